// navigate.cu -- a6-a15: phases 2+3 of the hot path as frontier sweeps over the index in HBM.
//
// Replaces navigate_one_bwt (/root/reference/ebwt2InDel.cpp:555-676) and navigate_two_bwts
// (:679-831): the two explicit-stack DFS loops over suffix-tree leaves and right-maximal internal
// nodes, with dna_bwt::LF(sa_node) / next_nodes / next_leaves (internal/dna_bwt.hpp:323-404),
// update_LCP_leaf (:344-355), update_lcp_threshold (internal/include.hpp:826-860),
// update_lcp_minima (:357-391), update_DA (:394-449) and find_leaves (:474-527).
//
// B200 design.  Every LCP / DA bit has exactly one writer, so traversal order is free.  The
// frontier is swept breadth-first, and every sweep keeps its nodes SORTED BY SUFFIX-ARRAY
// POSITION: the children cW of a sorted frontier are appended, in tile order, to four queues
// (one per c); A-queue ++ C-queue ++ G-queue ++ T-queue is again sorted because LF is monotone
// per symbol.  A sorted frontier turns the reference's random rank gathers into one
// near-sequential pass over the 64-byte index blocks per sweep (neighbouring nodes share blocks
// and DRAM pages) and makes the bit updates land in neighbouring words.  Ordered appends use a
// single-pass decoupled look-back (lookback.cuh).  When a sweep would not fit the frontier budget
// it is cut into position-contiguous chunks that are finished depth-first (bounded memory).
//
// Work mapping.  One thread per internal node (per node pair in mode -2): it walks the node's <= 6
// distinct boundaries, turns the rank differences into the five sub-interval sizes of each child
// cW, keeps the children with >= 2 non-empty sub-intervals (number_of_children >= 2) and appends
// them in order.  Leaves: one thread per leaf (two ranks per BWT).  All nodes of a sweep have the
// same depth, so the depth is a launch argument, not part of the records.
//
// Records.  Internal node, WIDE form (top of the tree): 48 bytes = {base, s0} {s1, s2} {s3, s4} as
// u64 (first position and the sizes of the TERM, A, C, G, T children).  SMALL form (every level
// whose nodes are all shorter than 2^16 -- all but the first ~10 levels): 16 bytes =
// { base[31:0], base[39:32] | s0 << 16, s1 | s2 << 16, s3 | s4 << 16 }.  Replaces the 56-byte
// sa_node (include.hpp:394-413).  Leaf: 16 bytes {first, second}.  A pair (mode -2) is two records.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <memory>

#include "common.cuh"
#include "lookback.cuh"

namespace e2i {

constexpr int kCompThreads = 256;                     // 8 compute warps: one node (pair) / leaf (pair) per thread and tile
constexpr int kPersistThreads = kCompThreads + 32;    // + the scan warp
constexpr int kStageBlocks = 512;                     // index blocks staged in shared memory per CTA (32 KB)
constexpr int kStripes = 128;                         // striped statistics counters (avoid single-address atomics)
enum { C_LCP = 0, C_NMIN, C_RANK, C_BITUPD, C_DA, C_NCOUNTERS = 8 };
#ifndef E2I_SMALL_LIMIT
#define E2I_SMALL_LIMIT 65536                         // node sizes below this use the 16-byte record (test builds lower it)
#endif
constexpr uint64_t kSmallLimit = E2I_SMALL_LIMIT;
static_assert(kSmallLimit <= 65536, "SMALL records hold 16-bit sizes");
constexpr uint32_t kExitTile = 0xffffffffu;

// Per-sweep control.  Device side: one zeroed 64-byte block per sweep (a ring, so no per-sweep
// memset): tile ticket, exit counter, the four child totals, the largest input node.  Host side: a
// page-locked, device-mapped block that the LAST CTA to leave the sweep fills in, followed by the
// sweep's sequence number; the host polls that word instead of a copy + stream synchronisation.
constexpr uint32_t kSweepSlots = 16384;
struct SweepDev {
    uint32_t ticket, done;
    unsigned long long counts[4];
    unsigned long long maxsz;
    unsigned long long pad[2];
};
static_assert(sizeof(SweepDev) == 64, "one sweep slot per 64 bytes");
struct HostCtl {
    unsigned long long out_count[4];
    unsigned long long maxsz;
    unsigned long long seq;
};

struct Segs {                 // a position-sorted run of records given as <= 4 segments
    const uint4 *p[4];
    uint32_t end[4];          // cumulative record counts
    uint32_t total;
};

struct NavArgs {
    DevIndex ix1, ix2;
    uint32_t *thr;            // 2 bits per merged position
    uint32_t *minima;         // 1 bit per merged position
    uint32_t *da;             // 1 bit per merged position (mode -2)
    unsigned long long *stripes;
    unsigned long long *desc;
    SweepDev *sweep;          // this sweep's device control block
    HostCtl *host;            // mapped page-locked result block
    unsigned long long seq;   // sequence number of this sweep
    uint4 *out[4];
    uint32_t epoch;
    uint32_t n_tiles;
    uint32_t bits;            // (depth >= K) | (depth >= k_right) << 1 for the records of this sweep
    int write;                // 0: expand only (redundant top of the tree on shards != 0)
};

__device__ __forceinline__ const uint4 *seg_record(const Segs &s, uint32_t g, int ru) {
    int k = 0;
    uint32_t start = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
        if (g >= s.end[i]) { k = i + 1; start = s.end[i]; }
    return s.p[k] + (size_t)(g - start) * ru;
}

// set bits [lo, hi) of a u32 bit array, keeping only those selected by the 32-bit periodic pattern
__device__ __forceinline__ void fill_bits(uint32_t *words, uint64_t lo, uint64_t hi, uint32_t pattern) {
    if (hi <= lo || pattern == 0) return;
    uint64_t w = lo >> 5;
    const uint64_t wl = (hi - 1) >> 5;
    for (; w <= wl; ++w) {
        uint32_t m = pattern;
        if (w == (lo >> 5)) m &= 0xffffffffu << (lo & 31);
        if (w == wl && (hi & 31)) m &= 0xffffffffu >> (32 - (hi & 31));
        if (m) atomicOr(words + w, m);
    }
}

// one atomicOr per touched word instead of one per bit
struct WordAcc {
    uint32_t *words;
    uint64_t w;
    uint32_t m;
    __device__ __forceinline__ void add(uint64_t word, uint32_t mask) {
        if (word != w) { flush(); w = word; m = mask; } else m |= mask;
    }
    __device__ __forceinline__ void flush() { if (m) atomicOr(words + w, m); m = 0; }
};

// named barriers: 1 = the compute warps among themselves; 2 = "prefix of the pending tile resolved"
// (scan warp arrives, compute warps wait); 3 = "a tile was posted" (compute warps arrive, scan warp
// waits).  Waiting warps are descheduled by the hardware instead of spinning on shared memory.
__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void bar_prefix_wait() { asm volatile("bar.sync 2, 288;" ::: "memory"); }
__device__ __forceinline__ void bar_prefix_arrive() { asm volatile("bar.arrive 2, 288;" ::: "memory"); }
__device__ __forceinline__ void bar_post_wait() { asm volatile("bar.sync 3, 288;" ::: "memory"); }
__device__ __forceinline__ void bar_post_arrive() { asm volatile("bar.arrive 3, 288;" ::: "memory"); }

// ---- the part of a sweep kernel that does not depend on what a record is ------------------------
struct SweepShared {
    unsigned long long base[4];                       // resolved exclusive prefix of the pending tile
    uint32_t agg[4];                                  // child counts of the pending tile
    uint32_t pend_tile;
    uint32_t tile;                                    // next ticket, handed from thread 0 to the CTA
    uint32_t wpk[kCompThreads / 32];                  // per-warp child counts, 4 x 8 bit
    uint32_t rng[4];
};

// Scan warp: publishes the counts of every posted tile at once, resolves its exclusive prefix by
// look-back while the compute warps already work on the next tile, and hands the four base slots
// back.  The last CTA to leave the sweep reports the totals to the host.
__device__ __forceinline__ void scan_warp_loop(const NavArgs &a, SweepShared &sh) {
    const int lane = threadIdx.x & 31;
    while (true) {
        bar_post_wait();
        const uint32_t tile = *(volatile uint32_t *)&sh.pend_tile;
        if (tile == kExitTile) break;
        uint32_t agg[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) agg[c] = ((volatile uint32_t *)sh.agg)[c];
        unsigned long long excl[4];
        lookback4(a.desc, a.epoch, tile, agg, excl);
        if (lane < 4) {
            unsigned long long e = 0, g2 = 0;
#pragma unroll
            for (int c = 0; c < 4; ++c) if (lane == c) { e = excl[c]; g2 = agg[c]; }
            ((volatile unsigned long long *)sh.base)[lane] = e;
            if (tile == a.n_tiles - 1) ((volatile unsigned long long *)a.sweep->counts)[lane] = e + g2;
        }
        __threadfence_block();
        __syncwarp();
        bar_prefix_arrive();
    }
    // every counter of this CTA is out (the compute warps flushed before posting the exit)
    if (lane == 0) {
        __threadfence();
        const uint32_t done = atomicAdd(&a.sweep->done, 1u);
        if (done == gridDim.x - 1) {
            __threadfence();
#pragma unroll
            for (int c = 0; c < 4; ++c)
                ((volatile unsigned long long *)a.host->out_count)[c] = ((volatile unsigned long long *)a.sweep->counts)[c];
            *(volatile unsigned long long *)&a.host->maxsz = *(volatile unsigned long long *)&a.sweep->maxsz;
            __threadfence_system();
            *(volatile unsigned long long *)&a.host->seq = a.seq;
        }
    }
}

// children of the pending tile: shared memory -> their resolved global slots (coalesced 16-byte stores)
template <int RU>
__device__ __forceinline__ void flush_pending(const NavArgs &a, SweepShared &sh, uint4 (*child)[kCompThreads * RU]) {
    bar_prefix_wait();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const uint32_t n16 = sh.agg[c] * RU;
        uint4 *dst = a.out[c] + sh.base[c] * RU;
        for (uint32_t i = threadIdx.x; i < n16; i += kCompThreads) dst[i] = child[c][i];
    }
}

// ballot the four validity flags of the warp's threads: position of this thread's children inside the
// warp (before[]), the validity mask, and the per-warp counts (4 x 8 bit) in shared memory
__device__ __forceinline__ uint32_t warp_child_slots(SweepShared &sh, const bool (&valid)[4], uint32_t (&before)[4]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t vm = 0, packed = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const uint32_t bal = __ballot_sync(0xffffffffu, valid[c]);
        before[c] = __popc(bal & ((1u << lane) - 1u));
        packed |= (uint32_t)__popc(bal) << (8 * c);
        if (valid[c]) vm |= 1u << c;
    }
    if (lane == 0) sh.wpk[warp] = packed;
    return vm;
}

// exclusive prefix of the per-warp counts for this warp (exw) and the tile totals (tot), by shuffles
__device__ __forceinline__ void tile_child_prefix(const SweepShared &sh, uint32_t (&exw)[4], uint32_t (&tot)[4]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t mine = lane < kCompThreads / 32 ? sh.wpk[lane] : 0u;
    uint32_t lo = (mine & 0xffu) | (((mine >> 8) & 0xffu) << 16);          // A, C as 16-bit fields (<= 256 per tile)
    uint32_t hi = ((mine >> 16) & 0xffu) | ((mine >> 24) << 16);           // G, T
    const uint32_t mlo = lo, mhi = hi;
#pragma unroll
    for (int s = 1; s < kCompThreads / 32; s <<= 1) {
        const uint32_t ylo = __shfl_up_sync(0xffffffffu, lo, s), yhi = __shfl_up_sync(0xffffffffu, hi, s);
        if (lane >= s) { lo += ylo; hi += yhi; }
    }
    const uint32_t elo = __shfl_sync(0xffffffffu, lo - mlo, warp), ehi = __shfl_sync(0xffffffffu, hi - mhi, warp);
    const uint32_t tlo = __shfl_sync(0xffffffffu, lo, kCompThreads / 32 - 1), thi = __shfl_sync(0xffffffffu, hi, kCompThreads / 32 - 1);
    exw[0] = elo & 0xffffu; exw[1] = elo >> 16; exw[2] = ehi & 0xffffu; exw[3] = ehi >> 16;
    tot[0] = tlo & 0xffffu; tot[1] = tlo >> 16; tot[2] = thi & 0xffffu; tot[3] = thi >> 16;
}

__device__ __forceinline__ void post_tile(SweepShared &sh, uint32_t tile, const uint32_t (&tot)[4]) {
    if (threadIdx.x == 0) {
#pragma unroll
        for (int c = 0; c < 4; ++c) sh.agg[c] = tot[c];
        sh.pend_tile = tile;
    }
    bar_compute();                                    // children, counts and tile id are in shared memory
    __threadfence_block();
    bar_post_arrive();
}

// end-of-kernel flush of the per-thread statistics (one atomic per warp and counter)
__device__ __forceinline__ void flush_stat(const NavArgs &a, int which, unsigned long long v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    if ((threadIdx.x & 31) == 0 && v)
        atomicAdd(a.stripes + (size_t)((blockIdx.x * 8 + (threadIdx.x >> 5)) & (kStripes - 1)) * C_NCOUNTERS + which, v);
}

// ---- staged index blocks ---------------------------------------------------------------------------
// Where the rank queries of a tile read their index blocks.  WINDOW: the tile's whole block range
// [origin, origin + n) sits in shared memory (dense tiles: nodes of one depth are disjoint and sorted).
// SLOTS: sparse tiles (a traversal shard of a multi-GPU run, the leaf frontier): every thread gets
// the first and the last block of its own interval, fetched by the whole CTA with 16-byte asynchronous
// copies in which 4 consecutive lanes take one 64-byte block (one L1 wavefront per block instead of
// one per 16 bytes).  GLOBAL: no staging (WIDE records at the top of the tree, sparse pairs).
enum { SRC_GLOBAL = 0, SRC_WINDOW = 1, SRC_SLOTS = 2 };

struct RankSrc {               // per thread and BWT: relative position 0 = start of the node's first block
    const uint4 *stage;        // this BWT's part of the staging buffer
    uint32_t origin_blk;       // index block of relative position 0
    uint32_t slot0;            // WINDOW: slot of block origin_blk (slot = slot0 + rel); SLOTS: slot0 holds block
                               // origin_blk and slot0 + 1 holds block origin_blk + d1
    uint32_t d1;               // SLOTS: relative index of the second staged block (0 = none)
};

__device__ __forceinline__ void load_block_smem(const uint4 *stage, uint32_t slot, uint4 &cnt, uint4 &a, uint4 &b, uint4 &t) {
    const uint32_t sw = (slot >> 1) & 3u;
    const uint4 *p = stage + slot * 4;
    cnt = p[sw]; a = p[1u ^ sw]; b = p[2u ^ sw]; t = p[3u ^ sw];
}

// #A,#C,#G,#T before relative position rpos, counted from the start of the superblock of origin_blk
// (mod 2^32: the base cancels in every difference, which is all a node shorter than 2^32 needs).
// multi_super (CTA-uniform): the tile reaches into a second superblock.
__device__ __forceinline__ void rank_rel(const DevIndex &ix, const RankSrc &r, int mode, bool multi_super, uint32_t rpos, uint32_t out[4]) {
    const uint32_t rel = rpos >> kBlockShift;
    uint4 cnt, a, b, t;
    if (mode == SRC_WINDOW) {
        load_block_smem(r.stage, r.slot0 + rel, cnt, a, b, t);
    } else if (mode == SRC_SLOTS && (rel == 0 || rel == r.d1)) {
        load_block_smem(r.stage, r.slot0 + (rel != 0), cnt, a, b, t);
    } else {
        const uint4 *p = ix.blocks + (size_t)(r.origin_blk + rel) * 4;
        cnt = __ldg(p); a = __ldg(p + 1); b = __ldg(p + 2); t = __ldg(p + 3);
    }
    uint32_t pc[4];
    block_popc(a, b, t, (int)(rpos & (kBlockSyms - 1)), pc);
    out[0] = cnt.x + pc[0];
    out[1] = cnt.y + pc[1];
    out[2] = cnt.z + pc[2];
    out[3] = cnt.w + pc[3];
    if (multi_super) {                                   // the interval may reach into the next superblock
        const uint32_t sb = (r.origin_blk + rel) >> (kSuperShift - kBlockShift), sb0 = r.origin_blk >> (kSuperShift - kBlockShift);
        if (sb != sb0) {
#pragma unroll
            for (int c = 0; c < 4; ++c) out[c] += (uint32_t)(ix.super[(size_t)sb * 4 + c] - ix.super[(size_t)sb0 * 4 + c]);
        }
    }
}

// children of one BWT side, kept in registers between the rank phase and the ordered append
template <typename W>
struct ChildSide {
    uint64_t base[4];       // F[c] + rank_c(first)
    W sz[5][4];             // the five sub-interval sizes of child c
};

template <bool OUT_S, typename W>
__device__ __forceinline__ void store_child(uint4 *dst, const ChildSide<W> &k, int c) {
    if (OUT_S) {
        dst[0] = make_uint4((uint32_t)k.base[c], (uint32_t)(k.base[c] >> 32) | ((uint32_t)k.sz[0][c] << 16),
                            (uint32_t)k.sz[1][c] | ((uint32_t)k.sz[2][c] << 16), (uint32_t)k.sz[3][c] | ((uint32_t)k.sz[4][c] << 16));
    } else {
        const uint64_t s0 = k.sz[0][c], s1 = k.sz[1][c], s2 = k.sz[2][c], s3 = k.sz[3][c], s4 = k.sz[4][c];
        dst[0] = make_uint4((uint32_t)k.base[c], (uint32_t)(k.base[c] >> 32), (uint32_t)s0, (uint32_t)(s0 >> 32));
        dst[1] = make_uint4((uint32_t)s1, (uint32_t)(s1 >> 32), (uint32_t)s2, (uint32_t)(s2 >> 32));
        dst[2] = make_uint4((uint32_t)s3, (uint32_t)(s3 >> 32), (uint32_t)s4, (uint32_t)(s4 >> 32));
    }
}

template <bool IN_S, typename W>
__device__ __forceinline__ void load_node(const uint4 *rec, uint64_t &base, W (&s)[5]) {
    if (IN_S) {
        const uint4 v = rec[0];
        base = v.x | ((uint64_t)(v.y & 0xffu) << 32);
        s[0] = v.y >> 16; s[1] = v.z & 0xffffu; s[2] = v.z >> 16; s[3] = v.w & 0xffffu; s[4] = v.w >> 16;
    } else {
        const uint4 v0 = rec[0], v1 = rec[1], v2 = rec[2];
        base = v0.x | ((uint64_t)v0.y << 32);
        s[0] = (W)(v0.z | ((uint64_t)v0.w << 32));
        s[1] = (W)(v1.x | ((uint64_t)v1.y << 32)); s[2] = (W)(v1.z | ((uint64_t)v1.w << 32));
        s[3] = (W)(v2.x | ((uint64_t)v2.y << 32)); s[4] = (W)(v2.z | ((uint64_t)v2.w << 32));
    }
}

struct NodeStat { uint32_t lcp = 0, nmin = 0, rank = 0, upd = 0, da = 0; };

// Bit updates of one (merged) node: merge_nodes (include.hpp:476-490), update_lcp_threshold
// (include.hpp:826-860: the border after child j is written iff child j is non-empty and the border
// is not the end of the node), update_lcp_minima (ebwt2InDel.cpp:357-391: after children A, C, G of
// size >= 2 whose end lies before last - 1), find_leaves (:474-527: children of summed size 1 get
// their DA bit, mode -2).
template <bool TWO, typename W>
__device__ __forceinline__ void node_bit_updates(const NavArgs &a, uint64_t mbase, const W (&s1)[5], const W (&s2)[5], NodeStat &st) {
    uint64_t ms[5], last = mbase;
#pragma unroll
    for (int j = 0; j < 5; ++j) { ms[j] = (uint64_t)s1[j] + (TWO ? (uint64_t)s2[j] : 0ull); last += ms[j]; }
    WordAcc thr{a.thr, ~0ull, 0u}, mn{a.minima, ~0ull, 0u};
    uint64_t mb = mbase;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        if (TWO) {
            if (ms[j] == 1) {
                st.da++;
                if (s2[j] == 1) atomicOr(a.da + (mb >> 5), 1u << (mb & 31));
            }
        }
        mb += ms[j];                           // border after child j = first position of child j+1
        if (j < 4 && mb != last) {
            if (ms[j] > 0) {
                st.lcp++;
                if (a.bits) { thr.add(mb >> 4, a.bits << ((mb & 15) * 2)); st.upd++; }
            }
            if (j >= 1 && ms[j] >= 2 && mb < last - 1) {
                st.nmin++;
                st.upd++;
                mn.add(mb >> 5, 1u << (mb & 31));
            }
        }
    }
    thr.flush();
    mn.flush();
}

// turn the ranks at one boundary into the sizes of sub-interval j of every child; nzp counts, per
// symbol, the non-empty sub-intervals (union over both BWTs, include.hpp:784-792)
template <bool TWO, typename W>
__device__ __forceinline__ void take_boundary(int j, const W (&cur1)[4], W (&prev1)[4], const W (&cur2)[4], W (&prev2)[4],
                                              ChildSide<W> &k1, ChildSide<W> &k2, uint32_t &nzp) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const W d1 = cur1[c] - prev1[c];
        W any = d1;
        k1.sz[j][c] = d1;
        prev1[c] = cur1[c];
        if (TWO) {
            const W d2 = cur2[c] - prev2[c];
            any |= d2;
            k2.sz[j][c] = d2;
            prev2[c] = cur2[c];
        }
        nzp += (any != 0 ? 1u : 0u) << (8 * c);
    }
}

// LF(sa_node) (dna_bwt.hpp:323-356) for one SMALL node (pair): ranks at the distinct boundaries (equal
// neighbours reuse the previous result, :334-347) in 32-bit arithmetic relative to the staged blocks.
template <bool TWO>
__device__ __forceinline__ void expand_small(const NavArgs &a, int mode, bool multi_super, const RankSrc &r1, const RankSrc &r2,
                                             uint64_t base1, uint32_t rpos1, const uint32_t (&s1)[5],
                                             uint64_t base2, uint32_t rpos2, const uint32_t (&s2)[5],
                                             ChildSide<uint32_t> &k1, ChildSide<uint32_t> &k2, uint32_t &nzp, uint32_t &st_rank) {
    uint32_t prev1[4], cur1[4], prev2[4] = {0, 0, 0, 0}, cur2[4] = {0, 0, 0, 0};
    rank_rel(a.ix1, r1, mode, multi_super, rpos1, prev1);
    st_rank++;
    if (TWO) { rank_rel(a.ix2, r2, mode, multi_super, rpos2, prev2); st_rank++; }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        // prev is relative to the superblock of the node's first block, which is the superblock of `base`
        uint64_t b1 = a.ix1.F[c] + prev1[c];
        if (a.ix1.n >> kSuperShift) b1 += a.ix1.super[(base1 >> kSuperShift) * 4 + c];
        k1.base[c] = b1;
        if (TWO) {
            uint64_t b2 = a.ix2.F[c] + prev2[c];
            if (a.ix2.n >> kSuperShift) b2 += a.ix2.super[(base2 >> kSuperShift) * 4 + c];
            k2.base[c] = b2;
        }
    }
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        rpos1 += s1[j];
        if (s1[j]) { rank_rel(a.ix1, r1, mode, multi_super, rpos1, cur1); st_rank++; }
        else { cur1[0] = prev1[0]; cur1[1] = prev1[1]; cur1[2] = prev1[2]; cur1[3] = prev1[3]; }
        if (TWO) {
            rpos2 += s2[j];
            if (s2[j]) { rank_rel(a.ix2, r2, mode, multi_super, rpos2, cur2); st_rank++; }
            else { cur2[0] = prev2[0]; cur2[1] = prev2[1]; cur2[2] = prev2[2]; cur2[3] = prev2[3]; }
        }
        take_boundary<TWO, uint32_t>(j, cur1, prev1, cur2, prev2, k1, k2, nzp);
    }
}

// the same for a WIDE node (pair): absolute 64-bit ranks read from HBM
template <bool TWO>
__device__ __forceinline__ void expand_wide(const NavArgs &a, uint64_t base1, const uint64_t (&s1)[5], uint64_t base2, const uint64_t (&s2)[5],
                                            ChildSide<uint64_t> &k1, ChildSide<uint64_t> &k2, uint32_t &nzp, uint32_t &st_rank) {
    uint64_t prev1[4], cur1[4], prev2[4] = {0, 0, 0, 0}, cur2[4] = {0, 0, 0, 0};
    rank4(a.ix1, base1, prev1);
    st_rank++;
    if (TWO) { rank4(a.ix2, base2, prev2); st_rank++; }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        k1.base[c] = a.ix1.F[c] + prev1[c];
        if (TWO) k2.base[c] = a.ix2.F[c] + prev2[c];
    }
    uint64_t b1 = base1, b2 = base2;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        b1 += s1[j];
        if (s1[j]) { rank4(a.ix1, b1, cur1); st_rank++; }
        else { cur1[0] = prev1[0]; cur1[1] = prev1[1]; cur1[2] = prev1[2]; cur1[3] = prev1[3]; }
        if (TWO) {
            b2 += s2[j];
            if (s2[j]) { rank4(a.ix2, b2, cur2); st_rank++; }
            else { cur2[0] = prev2[0]; cur2[1] = prev2[1]; cur2[2] = prev2[2]; cur2[3] = prev2[3]; }
        }
        take_boundary<TWO, uint64_t>(j, cur1, prev1, cur2, prev2, k1, k2, nzp);
    }
}

// SLOTS staging: every thread has announced up to two block ids in need[2 * tid], need[2 * tid + 1]
// (~0u = none); the CTA copies them with 4 consecutive lanes per 64-byte block
__device__ __forceinline__ void stage_slots(const DevIndex &ix, uint4 *stage, const uint32_t *need) {
#pragma unroll
    for (int it = 0; it < kStageBlocks * 4 / kCompThreads; ++it) {
        const uint32_t k = threadIdx.x + it * kCompThreads;
        const uint32_t slot = k >> 2, blk = need[slot];
        if (blk != ~0u) cp_async16(&stage[stage_slot(slot, k & 3)], ix.blocks + (size_t)blk * 4 + (k & 3));
    }
}

__device__ __forceinline__ void stage_window(const DevIndex &ix, uint4 *stage, uint32_t lo_blk, uint32_t n_blk) {
    const uint4 *src = ix.blocks + (size_t)lo_blk * 4;
    for (uint32_t k = threadIdx.x; k < n_blk * 4; k += kCompThreads) cp_async16(&stage[stage_slot(k >> 2, k & 3)], src + k);
}

// ---------------------------------------------------------------------------------------------
// Phase 3 sweep: internal nodes.  Persistent, warp-specialised: a CTA is 8 compute warps + 1 scan
// warp and loops over tiles taken by ticket (a tile's predecessors have always started: needed by the
// look-back).
//   compute warps  tile t+1: records -> staged index blocks -> bit updates -> ranks -> children in
//                  registers; THEN flush the children of tile t from shared memory to their final,
//                  by now resolved, global slots; park the children of t+1 in shared memory and post
//                  their counts to the scan warp;
//   scan warp      see scan_warp_loop.
// Publication is never delayed, so the look-back window stays short; nothing waits unless the
// prefix of tile t is still unresolved after the whole compute phase of tile t+1, and a warp that
// waits sits in a hardware barrier instead of polling.
// ---------------------------------------------------------------------------------------------
template <bool TWO, bool IN_S, bool OUT_S>
struct NodeSmem {
    static constexpr int RIN = (IN_S ? 1 : 3) * (TWO ? 2 : 1);    // uint4 per input record
    static constexpr int ROUT = (OUT_S ? 1 : 3) * (TWO ? 2 : 1);  // uint4 per output record
    uint4 stage[IN_S ? kStageBlocks * 4 : 4];         // staged index blocks
    uint4 child[4][kCompThreads * ROUT];              // parked children of the pending tile, per symbol
    uint4 recbuf[kCompThreads * RIN];                 // records of the NEXT tile, prefetched by LDGSTS (slot = thread)
    uint32_t need[IN_S ? kStageBlocks : 4];           // SLOTS staging: block ids wanted by the threads
    SweepShared sh;
};

template <bool TWO, bool IN_S, bool OUT_S>
__global__ void __launch_bounds__(kPersistThreads, IN_S ? (TWO ? 2 : 3) : 1)
expand_nodes_persistent(const NavArgs a, const Segs in) {
    using SM = NodeSmem<TWO, IN_S, OUT_S>;
    using W = typename std::conditional<IN_S, uint32_t, uint64_t>::type;
    constexpr int RIN = SM::RIN, ROUT = SM::ROUT, RSIDE_IN = IN_S ? 1 : 3, RSIDE_OUT = OUT_S ? 1 : 3;
    constexpr int STAGE = TWO ? kStageBlocks / 2 : kStageBlocks;           // blocks staged per BWT
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    SM &sm = *reinterpret_cast<SM *>(smem_raw);
    SweepShared &sh = sm.sh;

    if (threadIdx.x == 0) sh.pend_tile = 0;
    __syncthreads();
    if (threadIdx.x >= kCompThreads) { scan_warp_loop(a, sh); return; }

    // --------------------------------- compute warps ---------------------------------
    const int lane = threadIdx.x & 31;
    uint32_t my_seq = 0;                    // tiles this CTA has posted so far
    NodeStat st;
    uint64_t max_size = 0;
    uint4 *stage1 = sm.stage, *stage2 = sm.stage + (IN_S ? STAGE * 4 : 0);

    // software pipeline: the ticket and the records of tile t+1 are fetched while tile t is processed
    auto prefetch_records = [&](uint32_t tl) {
        const uint32_t g = tl * kCompThreads + threadIdx.x;
        if (tl < a.n_tiles && g < in.total) {
            const uint4 *rec = seg_record(in, g, RIN);
#pragma unroll
            for (int k = 0; k < RIN; ++k) cp_async16(&sm.recbuf[threadIdx.x * RIN + k], rec + k);
        }
    };
    if (threadIdx.x == 0) sh.tile = atomicAdd(&a.sweep->ticket, 1u);
    bar_compute();
    uint32_t tile = sh.tile;
    prefetch_records(tile);
    while (tile < a.n_tiles) {
        uint32_t nxt = 0;
        if (threadIdx.x == 0) nxt = atomicAdd(&a.sweep->ticket, 1u);       // its latency hides behind this tile
        const uint32_t g = tile * kCompThreads + threadIdx.x;
        const bool active = g < in.total;

        uint64_t base1 = 0, base2 = 0;
        W s1[5] = {0, 0, 0, 0, 0}, s2[5] = {0, 0, 0, 0, 0};
        cp_async_wait_all();                                                // this thread's own record has landed
        if (active) {
            load_node<IN_S, W>(&sm.recbuf[threadIdx.x * RIN], base1, s1);
            if (TWO) load_node<IN_S, W>(&sm.recbuf[threadIdx.x * RIN + RSIDE_IN], base2, s2);
        }
        const uint64_t size1 = (uint64_t)s1[0] + s1[1] + s1[2] + s1[3] + s1[4], size2 = (uint64_t)s2[0] + s2[1] + s2[2] + s2[3] + s2[4];
        max_size = max(max_size, max(size1, size2));

        // ---- stage the index blocks of the tile in shared memory (SMALL records only) ----
        int mode = SRC_GLOBAL;
        bool multi_super = false;
        const uint32_t fb1 = (uint32_t)(base1 >> kBlockShift), fb2 = (uint32_t)(base2 >> kBlockShift);
        RankSrc r1{stage1, fb1, 0u, 0u}, r2{stage2, fb2, 0u, 0u};
        if (IN_S) {
            const uint32_t last_active = min((uint32_t)kCompThreads, in.total - tile * kCompThreads) - 1;
            const uint32_t lb1 = (uint32_t)((base1 + size1) >> kBlockShift), lb2 = (uint32_t)((base2 + size2) >> kBlockShift);
            if (threadIdx.x == 0) { sh.rng[0] = fb1; if (TWO) sh.rng[2] = fb2; }
            if (threadIdx.x == last_active) { sh.rng[1] = lb1; if (TWO) sh.rng[3] = lb2; }
            if (!TWO) {                                                     // candidates for SLOTS staging
                sm.need[2 * threadIdx.x] = active ? fb1 : ~0u;
                sm.need[2 * threadIdx.x + 1] = (active && lb1 != fb1) ? lb1 : ~0u;
            }
            bar_compute();
            // nodes of one depth are disjoint and sorted: the tile touches the block range [lo, hi]
            const uint32_t lo1 = sh.rng[0], hi1 = sh.rng[1], lo2 = TWO ? sh.rng[2] : 0u, hi2 = TWO ? sh.rng[3] : 0u;
            const uint32_t span1 = hi1 - lo1 + 1, span2 = TWO ? hi2 - lo2 + 1 : 0u;
            constexpr uint32_t sbs = kSuperShift - kBlockShift;
            multi_super = (lo1 >> sbs) != (hi1 >> sbs) || (TWO && (lo2 >> sbs) != (hi2 >> sbs));
            if (span1 <= (uint32_t)STAGE && span2 <= (uint32_t)STAGE) {
                mode = SRC_WINDOW;
                stage_window(a.ix1, stage1, lo1, span1);
                if (TWO) stage_window(a.ix2, stage2, lo2, span2);
                r1.slot0 = fb1 - lo1;
                r2.slot0 = fb2 - lo2;
            } else if (!TWO) {
                mode = SRC_SLOTS;
                stage_slots(a.ix1, stage1, sm.need);
                r1.slot0 = 2 * threadIdx.x;
                r1.d1 = lb1 - fb1;
            }
        }
        const uint32_t rpos1 = (uint32_t)base1 & (kBlockSyms - 1), rpos2 = (uint32_t)base2 & (kBlockSyms - 1);

        // ---- bit updates on the merged node, while the copies are in flight ----
        if (active && a.write) node_bit_updates<TWO, W>(a, base1 + base2, s1, s2, st);
        if (threadIdx.x == 0) sh.tile = nxt;
        cp_async_wait_all();
        bar_compute();
        const uint32_t next_tile = sh.tile;
        prefetch_records(next_tile);                                         // overlaps with the rank phase below

        // ---- ranks -> children; child c is right-maximal iff >= 2 of its 5 gaps are non-empty ----
        ChildSide<W> k1, k2;
        uint32_t nzp = 0;
        if (active) {
            if constexpr (IN_S) expand_small<TWO>(a, mode, multi_super, r1, r2, base1, rpos1, s1, base2, rpos2, s2, k1, k2, nzp, st.rank);
            else expand_wide<TWO>(a, base1, s1, base2, s2, k1, k2, nzp, st.rank);
        }
        bool valid[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) valid[c] = ((nzp >> (8 * c)) & 0xffu) >= 2u;
        uint32_t before[4];
        const uint32_t vm = warp_child_slots(sh, valid, before);
        // the staged blocks and the per-warp counts are complete; the previous tile's children can leave
        if (my_seq) flush_pending<ROUT>(a, sh, sm.child);
        bar_compute();                                   // wpk visible; child buffer and agg free again
        uint32_t exw[4], tot[4];
        tile_child_prefix(sh, exw, tot);
        if (vm) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if ((vm >> c) & 1u) {
                    uint4 *dst = &sm.child[c][(exw[c] + before[c]) * ROUT];
                    store_child<OUT_S, W>(dst, k1, c);
                    if (TWO) store_child<OUT_S, W>(dst + RSIDE_OUT, k2, c);
                }
            }
        }
        post_tile(sh, tile, tot);
        ++my_seq;
        tile = next_tile;
    }
    if (my_seq) flush_pending<ROUT>(a, sh, sm.child);
    if (a.write) {
        flush_stat(a, C_LCP, st.lcp);
        flush_stat(a, C_NMIN, st.nmin);
        flush_stat(a, C_RANK, st.rank);
        flush_stat(a, C_BITUPD, st.upd);
        if (TWO) flush_stat(a, C_DA, st.da);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) max_size = max(max_size, (uint64_t)__shfl_xor_sync(0xffffffffu, (unsigned long long)max_size, s));
    if (lane == 0 && max_size) atomicMax(&a.sweep->maxsz, (unsigned long long)max_size);
    __threadfence();                                      // statistics and children are out before the CTA counts as done
    const uint32_t none[4] = {0, 0, 0, 0};
    post_tile(sh, kExitTile, none);
}

// ---------------------------------------------------------------------------------------------
// Phase 2 sweep: leaves (intervals of W#).  One thread per leaf (pair); record = 16 bytes
// {first, second} (a pair: 32 bytes).  Same persistent compute-warps + scan-warp structure.  Leaves
// are sparse in position space: the two blocks of a leaf are staged per thread (SLOTS).
// ---------------------------------------------------------------------------------------------
template <bool TWO>
struct LeafSmem {
    static constexpr int RU = TWO ? 2 : 1;            // uint4 per record
    uint4 stage[kStageBlocks * 4];
    uint4 child[4][kCompThreads * RU];
    uint4 recbuf[kCompThreads * RU];
    uint32_t need[kStageBlocks];
    SweepShared sh;
};

// rank at an absolute position whose block is staged in `slot` (SLOTS staging of the leaf kernel)
__device__ __forceinline__ void rank_slot(const DevIndex &ix, const uint4 *stage, uint32_t slot, uint64_t pos, uint64_t out[4]) {
    uint4 cnt, a, b, t;
    load_block_smem(stage, slot, cnt, a, b, t);
    uint32_t pc[4];
    block_popc(a, b, t, (int)((uint32_t)pos & (kBlockSyms - 1)), pc);
    out[0] = (uint64_t)cnt.x + pc[0];
    out[1] = (uint64_t)cnt.y + pc[1];
    out[2] = (uint64_t)cnt.z + pc[2];
    out[3] = (uint64_t)cnt.w + pc[3];
    if (ix.n >> kSuperShift) {
        const uint64_t *sb = ix.super + (pos >> kSuperShift) * 4;
        out[0] += sb[0]; out[1] += sb[1]; out[2] += sb[2]; out[3] += sb[3];
    }
}

template <bool TWO>
__global__ void __launch_bounds__(kPersistThreads, TWO ? 2 : 3)
expand_leaves_persistent(const NavArgs a, const Segs in) {
    constexpr int RU = TWO ? 2 : 1;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    LeafSmem<TWO> &sm = *reinterpret_cast<LeafSmem<TWO> *>(smem_raw);
    SweepShared &sh = sm.sh;

    if (threadIdx.x == 0) sh.pend_tile = 0;
    __syncthreads();
    if (threadIdx.x >= kCompThreads) { scan_warp_loop(a, sh); return; }

    uint32_t my_seq = 0;
    unsigned long long st_lcp = 0, st_da = 0;
    uint32_t st_rank = 0;
    auto prefetch_records = [&](uint32_t tl) {
        const uint32_t g = tl * kCompThreads + threadIdx.x;
        if (tl < a.n_tiles && g < in.total) {
            const uint4 *rec = seg_record(in, g, RU);
#pragma unroll
            for (int k = 0; k < RU; ++k) cp_async16(&sm.recbuf[threadIdx.x * RU + k], rec + k);
        }
    };
    if (threadIdx.x == 0) sh.tile = atomicAdd(&a.sweep->ticket, 1u);
    bar_compute();
    uint32_t tile = sh.tile;
    prefetch_records(tile);
    while (tile < a.n_tiles) {
        uint32_t nxt = 0;
        if (threadIdx.x == 0) nxt = atomicAdd(&a.sweep->ticket, 1u);
        const uint32_t g = tile * kCompThreads + threadIdx.x;
        const bool active = g < in.total;
        uint64_t f1 = 0, s1 = 0, f2 = 0, s2 = 0;
        cp_async_wait_all();                               // this thread's own record has landed
        if (active) {
            const ulonglong2 *rec = reinterpret_cast<const ulonglong2 *>(&sm.recbuf[threadIdx.x * RU]);
            const ulonglong2 x = rec[0];
            f1 = x.x; s1 = x.y;
            if (TWO) { const ulonglong2 z = rec[1]; f2 = z.x; s2 = z.y; }
        }
        // the (up to) two index blocks per BWT this leaf (pair) needs: slots 2t, 2t+1 (TWO: one BWT each,
        // the second boundary of a side reads HBM unless it shares the block of the first)
        const uint32_t fb1 = (uint32_t)(f1 >> kBlockShift), lb1 = (uint32_t)(s1 >> kBlockShift);
        const uint32_t fb2 = (uint32_t)(f2 >> kBlockShift), lb2 = (uint32_t)(s2 >> kBlockShift);
        if (!TWO) {
            sm.need[2 * threadIdx.x] = active ? fb1 : ~0u;
            sm.need[2 * threadIdx.x + 1] = (active && lb1 != fb1) ? lb1 : ~0u;
        } else {
            sm.need[2 * threadIdx.x] = active ? fb1 : ~0u;
            sm.need[2 * threadIdx.x + 1] = active ? fb2 : ~0u;
        }
        bar_compute();
        if (!TWO) {
            stage_slots(a.ix1, sm.stage, sm.need);
        } else {                                           // even slots come from BWT 1, odd slots from BWT 2
#pragma unroll
            for (int it = 0; it < kStageBlocks * 4 / kCompThreads; ++it) {
                const uint32_t k = threadIdx.x + it * kCompThreads;
                const uint32_t slot = k >> 2, blk = sm.need[slot];
                const uint4 *src = (slot & 1u) ? a.ix2.blocks : a.ix1.blocks;
                if (blk != ~0u) cp_async16(&sm.stage[stage_slot(slot, k & 3)], src + (size_t)blk * 4 + (k & 3));
            }
        }
        if (active && a.write) {
            // update_LCP_leaf (:344-355) / update_DA (:394-425) at merged coordinates
            const uint64_t start1 = f1 + f2, start2 = f2 + s1, end = s1 + s2;
            if (end > start1) st_lcp += end - start1 - 1;
            const uint32_t pat = ((a.bits & 1u) ? 0x55555555u : 0u) | ((a.bits & 2u) ? 0xaaaaaaaau : 0u);
            if (end > start1 + 1) fill_bits(a.thr, 2 * (start1 + 1), 2 * end, pat);
            if (TWO) {
                st_da += end - start1;
                fill_bits(a.da, start2, end, 0xffffffffu);
            }
        }
        if (threadIdx.x == 0) sh.tile = nxt;
        cp_async_wait_all();
        bar_compute();                                     // staged blocks and the next ticket are visible
        const uint32_t next_tile = sh.tile;
        prefetch_records(next_tile);
        // next_leaves (dna_bwt.hpp:358-379; two BWTs: ebwt2InDel.cpp:452-472): LF(range) = 2 ranks per BWT
        uint64_t lo1[4] = {0, 0, 0, 0}, hi1[4] = {0, 0, 0, 0}, lo2[4] = {0, 0, 0, 0}, hi2[4] = {0, 0, 0, 0};
        if (active) {
            rank_slot(a.ix1, sm.stage, 2 * threadIdx.x, f1, lo1);
            st_rank++;
            if (s1 > f1) {
                if (lb1 == fb1) rank_slot(a.ix1, sm.stage, 2 * threadIdx.x, s1, hi1);
                else if (!TWO) rank_slot(a.ix1, sm.stage, 2 * threadIdx.x + 1, s1, hi1);
                else rank4(a.ix1, s1, hi1);
                st_rank++;
            } else { hi1[0] = lo1[0]; hi1[1] = lo1[1]; hi1[2] = lo1[2]; hi1[3] = lo1[3]; }
            if (TWO) {
                rank_slot(a.ix2, sm.stage, 2 * threadIdx.x + 1, f2, lo2);
                st_rank++;
                if (s2 > f2) {
                    if (lb2 == fb2) rank_slot(a.ix2, sm.stage, 2 * threadIdx.x + 1, s2, hi2);
                    else rank4(a.ix2, s2, hi2);
                    st_rank++;
                } else { hi2[0] = lo2[0]; hi2[1] = lo2[1]; hi2[2] = lo2[2]; hi2[3] = lo2[3]; }
            }
        }
        bool valid[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) valid[c] = active && ((hi1[c] - lo1[c]) + (hi2[c] - lo2[c]) >= 2);
        uint32_t before[4];
        const uint32_t vm = warp_child_slots(sh, valid, before);
        if (my_seq) flush_pending<RU>(a, sh, sm.child);
        bar_compute();                                     // child buffer and agg free again
        uint32_t exw[4], tot[4];
        tile_child_prefix(sh, exw, tot);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if ((vm >> c) & 1u) {
                ulonglong2 *o = reinterpret_cast<ulonglong2 *>(&sm.child[c][(exw[c] + before[c]) * RU]);
                o[0] = make_ulonglong2(a.ix1.F[c] + lo1[c], a.ix1.F[c] + hi1[c]);
                if (TWO) o[1] = make_ulonglong2(a.ix2.F[c] + lo2[c], a.ix2.F[c] + hi2[c]);
            }
        }
        post_tile(sh, tile, tot);
        ++my_seq;
        tile = next_tile;
    }
    if (my_seq) flush_pending<RU>(a, sh, sm.child);
    if (a.write) {
        flush_stat(a, C_LCP, st_lcp);
        flush_stat(a, C_RANK, st_rank);
        if (TWO) flush_stat(a, C_DA, st_da);
    }
    __threadfence();
    const uint32_t none[4] = {0, 0, 0, 0};
    post_tile(sh, kExitTile, none);
}

// ---------------------------------------------------------------------------------------------
// Host-side frontier driver
// ---------------------------------------------------------------------------------------------
static int node_rec_u4(bool small, bool two) { return (small ? 1 : 3) * (two ? 2 : 1); }

static void pack_wide_host(uint64_t *rec, uint64_t base, const uint64_t F[4], uint64_t n) {
    rec[0] = base; rec[1] = F[0] - base; rec[2] = F[1] - F[0]; rec[3] = F[2] - F[1]; rec[4] = F[3] - F[2]; rec[5] = n - F[3];
}

// size of the node stored at `rec` (one side of a record)
static uint64_t node_size_host(const uint64_t *rec, bool small) {
    if (!small) return rec[1] + rec[2] + rec[3] + rec[4] + rec[5];
    uint32_t w[4];
    std::memcpy(w, rec, sizeof w);
    return (uint64_t)(w[1] >> 16) + (w[2] & 0xffffu) + (w[2] >> 16) + (w[3] & 0xffffu) + (w[3] >> 16);
}

struct Frame {
    Arena *arena;
    int side;
    void *p;
    ~Frame() { if (p) arena->free(side, p); }
};

struct Chunk {
    uint4 *p[4];
    uint64_t cnt[4];
    int level = 0;                      // tree depth of the records (selects the arena end of the next frame)
    bool small = false;                 // record form (internal nodes)
    uint64_t bound = ~0ull;             // upper bound on the size of any node of the chunk
    std::shared_ptr<Frame> frame;
    uint64_t total() const { return cnt[0] + cnt[1] + cnt[2] + cnt[3]; }
};

struct SweepStats {
    uint64_t items = 0, sweeps = 0, max_chunk = 0;
    double ms_alloc = 0, ms_sync = 0, ms_max_alloc = 0, ms_max_sync = 0;   // host wall time (E2I_DEBUG)
};

static inline double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Cut the first `take` records off a chunk (position-contiguous prefix); ru = uint4 per record.
static Chunk split_head(Chunk &c, uint64_t take, int ru) {
    Chunk head = c;
    uint64_t left = take;
    for (int s = 0; s < 4; ++s) {
        const uint64_t k = std::min<uint64_t>(left, c.cnt[s]);
        head.cnt[s] = k;
        c.p[s] += k * ru;
        c.cnt[s] -= k;
        left -= k;
    }
    return head;
}

struct PassCfg {
    bool leaves, two;
    uint64_t budget;          // bytes of the frame arena
    uint64_t depth_hint;      // expected depth of the traversal below a cut level
    uint32_t K, k_right;
};

template <typename Launch>
static int run_frontier(e2i_ctx *ctx, Chunk root, const PassCfg &cfg, NavArgs &args, Launch launch, SweepStats &ss,
                        uint64_t stop_at_items, std::vector<Chunk> *stopped) {
    std::vector<Chunk> stack;
    stack.push_back(std::move(root));
    HostCtl *hctl = reinterpret_cast<HostCtl *>(ctx->ctl_host);
    while (!stack.empty()) {
        Chunk cur = std::move(stack.back());
        stack.pop_back();
        if (cur.total() == 0) continue;
        if (stopped && cur.total() >= stop_at_items) {   // hand the frontier back to the caller (sharding)
            stopped->push_back(std::move(cur));
            continue;
        }
        // record forms: the children of nodes shorter than kSmallLimit are shorter than kSmallLimit
        const bool in_small = !cfg.leaves && cur.small, out_small = !cfg.leaves && cur.bound < kSmallLimit;
        const int ru_in = cfg.leaves ? (cfg.two ? 2 : 1) : node_rec_u4(in_small, cfg.two);
        const int ru_out = cfg.leaves ? (cfg.two ? 2 : 1) : node_rec_u4(out_small, cfg.two);
        const double out_bytes = 16.0 * ru_out * 4;      // four queues, each sized for every input record
        // max_chunk: largest chunk swept whole (two of its frames fit the arena: level-synchronous case).
        // split_chunk: chunk size once a level has to be cut (depth-first case): a path of such chunks down
        // to the deepest level must fit the arena next to the frame that is being cut.
        const uint64_t max_chunk = std::max<uint64_t>(65536, (uint64_t)((double)cfg.budget / (out_bytes * 2.5)));
        const uint64_t split_chunk = std::max<uint64_t>(256, std::min<uint64_t>(max_chunk, (uint64_t)((double)cfg.budget * 0.45 / (out_bytes * (double)cfg.depth_hint))));
        Chunk work;
        uint64_t take = cur.total() <= max_chunk ? cur.total() : std::min<uint64_t>(cur.total(), split_chunk);
        void *mem = nullptr;
        const double ta = now_ms();
        while (true) {   // shrink the chunk until its output frame fits the pool
            mem = ctx->arena.alloc((cur.level + 1) & 1, take * 4 * ru_out * sizeof(uint4));
            if (mem) break;
            if (take <= 256) { set_error("frontier memory exhausted (arena %llu bytes, %llu in use): raise the frontier budget",
                                          (unsigned long long)ctx->arena.size(), (unsigned long long)ctx->arena.in_use()); return E2I_ERR_MEMORY; }
            take /= 2;
        }
        { const double d = now_ms() - ta; ss.ms_alloc += d; ss.ms_max_alloc = std::max(ss.ms_max_alloc, d); }
        if (take < cur.total()) {
            work = split_head(cur, take, ru_in);
            stack.push_back(std::move(cur));
        } else {
            work = std::move(cur);
        }
        auto frame = std::make_shared<Frame>();
        frame->arena = &ctx->arena;
        frame->side = (work.level + 1) & 1;
        frame->p = mem;
        Segs segs;
        uint32_t acc = 0;
        for (int s = 0; s < 4; ++s) {
            segs.p[s] = work.p[s];
            acc += (uint32_t)work.cnt[s];
            segs.end[s] = acc;
        }
        segs.total = acc;
        const uint32_t n_tiles = (acc + kCompThreads - 1) / kCompThreads;
        if ((size_t)n_tiles * kLb4Words > ctx->desc_words) {
            dfree(ctx, ctx->desc);
            ctx->desc = nullptr;
            ctx->desc_words = (size_t)n_tiles * kLb4Words * 3 / 2 + 1024;
            E2I_CUDA_TRY(dmalloc(ctx, &ctx->desc, ctx->desc_words * 8));
            E2I_CUDA_TRY(cudaMemsetAsync(ctx->desc, 0, ctx->desc_words * 8, ctx->stream));
            ctx->epoch = 0;
        }
        if (++ctx->epoch >= 0xffffu) {
            E2I_CUDA_TRY(cudaMemsetAsync(ctx->desc, 0, ctx->desc_words * 8, ctx->stream));
            ctx->epoch = 1;
        }
        for (int c = 0; c < 4; ++c) args.out[c] = reinterpret_cast<uint4 *>(mem) + (size_t)c * take * ru_out;
        args.desc = ctx->desc;
        args.epoch = ctx->epoch;
        args.n_tiles = n_tiles;
        if (ctx->ticket_next == kSweepSlots) {           // ring of sweep control blocks used up: zero it again
            E2I_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            E2I_CUDA_TRY(cudaMemsetAsync(ctx->ctl, 0, kSweepSlots * sizeof(SweepDev), ctx->stream));
            ctx->ticket_next = 0;
        }
        args.sweep = reinterpret_cast<SweepDev *>(ctx->ctl) + ctx->ticket_next++;
        args.host = hctl;
        args.seq = ++ctx->sweep_seq;
        const uint64_t depth = (uint64_t)work.level;     // every record of a sweep has this depth
        args.bits = (depth >= cfg.K ? 1u : 0u) | (depth >= cfg.k_right ? 2u : 0u);
        launch(args, segs, n_tiles, in_small, out_small);
        E2I_CUDA_TRY(cudaGetLastError());
        ctx->n_launch++;
        ctx->n_d2h += sizeof(HostCtl);
        const double tsy = now_ms();
        {   // wait for the totals (written by the last CTA to leave): poll the mapped sequence word
            volatile unsigned long long *seqp = &hctl->seq;
            unsigned spins = 0;
            while (*seqp != args.seq) {
                if ((++spins & 0xfffu) == 0) {
                    const cudaError_t q = cudaStreamQuery(ctx->stream);
                    if (q == cudaSuccess) {
                        if (*seqp == args.seq) break;
                        set_error("traversal sweep finished without publishing its counts");
                        return E2I_ERR_CUDA;
                    }
                    if (q != cudaErrorNotReady) { set_error("CUDA error in a traversal sweep: %s", cudaGetErrorString(q)); return E2I_ERR_CUDA; }
                }
            }
            std::atomic_thread_fence(std::memory_order_acquire);
        }
        { const double d = now_ms() - tsy; ss.ms_sync += d; ss.ms_max_sync = std::max(ss.ms_max_sync, d); }
        ss.items += acc;
        ss.sweeps++;
        ss.max_chunk = std::max<uint64_t>(ss.max_chunk, acc);
        Chunk next;
        next.frame = frame;
        next.level = work.level + 1;
        next.small = out_small;
        next.bound = cfg.leaves ? ~0ull : std::min<uint64_t>(work.bound, ((volatile unsigned long long *)&hctl->maxsz)[0]);
        for (int c = 0; c < 4; ++c) { next.p[c] = args.out[c]; next.cnt[c] = ((volatile unsigned long long *)hctl->out_count)[c]; }
        work.frame.reset();
        if (next.total()) stack.push_back(std::move(next));
    }
    return E2I_OK;
}

}  // namespace e2i

using namespace e2i;

static uint64_t padded_words32(uint64_t bits) { return ((bits + 31) / 32 + 63) / 64 * 64 + 64; }

namespace {
template <bool TWO, bool IN_S, bool OUT_S>
cudaError_t launch_nodes(const NavArgs &a, const Segs &segs, uint32_t grid, cudaStream_t s) {
    constexpr size_t smem = sizeof(NodeSmem<TWO, IN_S, OUT_S>);
    const cudaError_t e = cudaFuncSetAttribute(expand_nodes_persistent<TWO, IN_S, OUT_S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    expand_nodes_persistent<TWO, IN_S, OUT_S><<<grid, kPersistThreads, smem, s>>>(a, segs);
    return cudaSuccess;
}
template <bool TWO>
cudaError_t launch_leaves(const NavArgs &a, const Segs &segs, uint32_t grid, cudaStream_t s) {
    constexpr size_t smem = sizeof(LeafSmem<TWO>);
    const cudaError_t e = cudaFuncSetAttribute(expand_leaves_persistent<TWO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    expand_leaves_persistent<TWO><<<grid, kPersistThreads, smem, s>>>(a, segs);
    return cudaSuccess;
}
}  // namespace

extern "C" int e2i_navigate_shard(e2i_ctx *ctx, const e2i_index *b1, const e2i_index *b2, const e2i_params *p,
                                  int shard, int n_shards, e2i_lcpbits **out, e2i_bits **da_out, e2i_stats *st) {
    if (!ctx || !b1 || !p || !out || !st) { set_error("e2i_navigate: null argument"); return E2I_ERR_ARG; }
    if (b2 && !da_out) { set_error("e2i_navigate: da_out is required with two BWTs"); return E2I_ERR_ARG; }
    if (n_shards < 1 || shard < 0 || shard >= n_shards) { set_error("e2i_navigate: bad shard %d/%d", shard, n_shards); return E2I_ERR_ARG; }
    if (p->K < 1 || p->k_right < 1) { set_error("e2i_navigate: K and k_right must be >= 1"); return E2I_ERR_ARG; }
    if ((b1->n >> 39) || (b2 && (b2->n >> 39))) { set_error("e2i_navigate: BWT longer than 2^39 symbols"); return E2I_ERR_ARG; }
    E2I_CUDA_TRY(cudaSetDevice(ctx->device));
    Accounting acct(ctx, st);
    cudaStream_t s = ctx->stream;
    const bool two = b2 != nullptr;
    const uint64_t n = b1->n + (two ? b2->n : 0);

    e2i_lcpbits *l = new e2i_lcpbits();
    l->ctx = ctx;
    l->n = n;
    l->thr_words32 = padded_words32(2 * n);
    l->min_words32 = padded_words32(n);
    e2i_bits *da = nullptr;
    unsigned long long *stripes = nullptr;
    auto fail = [&](int rc) { e2i_lcpbits_free(l); e2i_bits_free(da); dfree(ctx, stripes); return rc; };
#define TRYF(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__, cudaGetErrorString(_e)); return fail(E2I_ERR_CUDA); } } while (0)
    TRYF(dmalloc(ctx, &l->thr, l->thr_words32 * 4));
    TRYF(dmalloc(ctx, &l->minima, l->min_words32 * 4));
    TRYF(cudaMemsetAsync(l->thr, 0, l->thr_words32 * 4, s));
    TRYF(cudaMemsetAsync(l->minima, 0, l->min_words32 * 4, s));
    if (two) {
        da = new e2i_bits();
        da->ctx = ctx;
        da->n = n;
        da->n_words32 = padded_words32(n);
        TRYF(dmalloc(ctx, &da->words, da->n_words32 * 4));
        TRYF(cudaMemsetAsync(da->words, 0, da->n_words32 * 4, s));
    }
    const size_t stripe_bytes = (size_t)kStripes * C_NCOUNTERS * sizeof(unsigned long long);
    TRYF(dmalloc(ctx, &stripes, stripe_bytes));
    TRYF(cudaMemsetAsync(ctx->ctl, 0, kSweepSlots * sizeof(SweepDev), s));
    ctx->ticket_next = 0;

    // frontier budget: what is free now, minus head-room, unless the caller set one
    size_t free_b = 0, total_b = 0;
    TRYF(cudaMemGetInfo(&free_b, &total_b));
    {   // blocks cached by the context's stream-ordered pool are available to us as well
        uint64_t reserved = 0, used = 0;
        TRYF(cudaMemPoolGetAttribute(ctx->pool, cudaMemPoolAttrReservedMemCurrent, &reserved));
        TRYF(cudaMemPoolGetAttribute(ctx->pool, cudaMemPoolAttrUsedMemCurrent, &used));
        if (reserved > used) free_b += reserved - used;
    }
    uint64_t budget = ctx->frontier_budget ? ctx->frontier_budget : (uint64_t)(free_b * 0.85);
    {   // the frame arena: kept across calls, re-allocated only when this input needs a larger one
        const uint64_t want = std::min<uint64_t>(budget, std::max<uint64_t>(1ull << 30, 4 * n));
        if (ctx->arena_bytes < want && ctx->arena_bytes < budget) {
            dfree(ctx, ctx->arena_mem);
            ctx->arena_mem = nullptr;
            ctx->arena_bytes = 0;
            TRYF(dmalloc(ctx, &ctx->arena_mem, want));
            ctx->arena_bytes = want;
        }
        const uint64_t use = ctx->frontier_budget ? std::min<uint64_t>(ctx->arena_bytes, ctx->frontier_budget) : ctx->arena_bytes;
        ctx->arena.reset(static_cast<char *>(ctx->arena_mem), use);
        budget = use;
    }

    NavArgs args{};
    args.ix1 = b1->dev();
    args.ix2 = two ? b2->dev() : b1->dev();
    args.thr = l->thr;
    args.minima = l->minima;
    args.da = da ? da->words : nullptr;
    args.stripes = stripes;

    std::vector<unsigned long long> hstripes((size_t)kStripes * C_NCOUNTERS);
    auto sum_stripes = [&](unsigned long long tot[C_NCOUNTERS]) -> int {
        E2I_CUDA_TRY(cudaMemcpyAsync(hstripes.data(), stripes, stripe_bytes, cudaMemcpyDeviceToHost, s));
        ctx->n_d2h += stripe_bytes;
        E2I_CUDA_TRY(cudaStreamSynchronize(s));
        for (int k = 0; k < C_NCOUNTERS; ++k) tot[k] = 0;
        for (int i = 0; i < kStripes; ++i) for (int k = 0; k < C_NCOUNTERS; ++k) tot[k] += hstripes[(size_t)i * C_NCOUNTERS + k];
        return E2I_OK;
    };

    // Sharding: the top of the tree is expanded on every shard (only shard 0 writes
    // its bits); once a sweep holds >= kDealItems nodes it is dealt in position-contiguous slices of
    // equal cumulated interval length, and every shard finishes its slice independently.  The cut
    // depends only on the input and n_shards (never on a rank's free memory), so all shards agree on it.
    const uint64_t kDealItems = 4096ull * (uint64_t)n_shards;

    auto run_pass = [&](bool leaves, SweepStats &ss) -> int {
        PassCfg cfg;
        cfg.leaves = leaves;
        cfg.two = two;
        cfg.budget = budget;
        // depth of the traversal: the internal-node pass is never deeper than the leaf pass that ran before it
        cfg.depth_hint = leaves ? 1024 : st->levels_leaves + 16;
        cfg.K = (uint32_t)p->K;
        cfg.k_right = (uint32_t)p->k_right;
        const int ru_root = leaves ? (two ? 2 : 1) : node_rec_u4(false, two);
        void *rootmem = ctx->arena.alloc(0, (size_t)ru_root * 16);
        if (!rootmem) { set_error("frontier arena too small"); return E2I_ERR_MEMORY; }
        uint64_t rec[12] = {0};
        if (leaves) {                                   // first_leaf (dna_bwt.hpp:313-317)
            rec[0] = 0; rec[1] = b1->F[0];
            if (two) { rec[2] = 0; rec[3] = b2->F[0]; }
        } else {                                        // root (dna_bwt.hpp:296-308) as a WIDE record
            pack_wide_host(rec, 0, b1->F, b1->n);
            if (two) pack_wide_host(rec + 6, 0, b2->F, b2->n);
        }
        E2I_CUDA_TRY(cudaMemcpyAsync(rootmem, rec, (size_t)ru_root * 16, cudaMemcpyHostToDevice, s));
        Chunk root{};
        root.p[0] = reinterpret_cast<uint4 *>(rootmem);
        root.cnt[0] = 1;
        root.small = false;
        root.bound = std::max<uint64_t>(b1->n, two ? b2->n : 0);
        root.frame = std::make_shared<Frame>();
        root.frame->arena = &ctx->arena;
        root.frame->side = 0;
        root.frame->p = rootmem;
        cudaError_t launch_err = cudaSuccess;
        auto launch = [&](NavArgs &a, const Segs &segs, uint32_t n_tiles, bool in_small, bool out_small) {
            cudaError_t e;
            if (leaves) {
                const uint32_t grid = std::min<uint32_t>(n_tiles, (uint32_t)ctx->sm_count * (two ? 2u : 3u));
                e = two ? launch_leaves<true>(a, segs, grid, s) : launch_leaves<false>(a, segs, grid, s);
            } else if (in_small) {
                const uint32_t grid = std::min<uint32_t>(n_tiles, (uint32_t)ctx->sm_count * (two ? 2u : 3u));
                e = two ? launch_nodes<true, true, true>(a, segs, grid, s) : launch_nodes<false, true, true>(a, segs, grid, s);
            } else {
                const uint32_t grid = std::min<uint32_t>(n_tiles, (uint32_t)ctx->sm_count);
                if (out_small) e = two ? launch_nodes<true, false, true>(a, segs, grid, s) : launch_nodes<false, false, true>(a, segs, grid, s);
                else e = two ? launch_nodes<true, false, false>(a, segs, grid, s) : launch_nodes<false, false, false>(a, segs, grid, s);
            }
            if (e != cudaSuccess) launch_err = e;
        };
        auto run = [&](Chunk c, SweepStats &stats, uint64_t stop_at, std::vector<Chunk> *stopped) -> int {
            const int rc = run_frontier(ctx, std::move(c), cfg, args, launch, stats, stop_at, stopped);
            if (rc == E2I_OK && launch_err != cudaSuccess) { set_error("kernel configuration failed: %s", cudaGetErrorString(launch_err)); return E2I_ERR_CUDA; }
            return rc;
        };
        if (n_shards == 1) {
            args.write = 1;
            return run(std::move(root), ss, 0, nullptr);
        }
        // shared top of the tree
        std::vector<Chunk> dealt;
        args.write = shard == 0;
        SweepStats top;
        E2I_TRY(run(std::move(root), top, kDealItems, &dealt));
        if (shard == 0) { ss.items += top.items; ss.sweeps += top.sweeps; ss.max_chunk = std::max(ss.max_chunk, top.max_chunk); }
        args.write = 1;
        for (Chunk &c : dealt) {
            // deal by cumulated interval length: fetch every record
            const int ru = leaves ? (two ? 2 : 1) : node_rec_u4(c.small, two);
            const int words = ru * 2, side_words = words / (two ? 2 : 1);
            const uint64_t tot = c.total();
            std::vector<uint64_t> host((size_t)tot * words);
            uint64_t off = 0;
            for (int q = 0; q < 4; ++q) {
                if (!c.cnt[q]) continue;
                E2I_CUDA_TRY(cudaMemcpyAsync(host.data() + off * words, c.p[q], c.cnt[q] * words * 8, cudaMemcpyDeviceToHost, s));
                off += c.cnt[q];
            }
            E2I_CUDA_TRY(cudaStreamSynchronize(s));
            auto weight = [&](uint64_t i) -> uint64_t {
                const uint64_t *r = host.data() + i * words;
                if (leaves) return (r[1] - r[0]) + (two ? r[3] - r[2] : 0) + 1;
                return node_size_host(r, c.small) + (two ? node_size_host(r + side_words, c.small) : 0) + 1;
            };
            unsigned __int128 wsum = 0;
            for (uint64_t i = 0; i < tot; ++i) wsum += weight(i);
            unsigned __int128 acc = 0;
            uint64_t lo = tot, hi = tot;
            bool have_lo = false;
            for (uint64_t i = 0; i < tot; ++i) {   // record i belongs to shard floor(acc * n_shards / wsum)
                const int owner = (int)((acc * (unsigned)n_shards) / wsum);
                if (owner == shard && !have_lo) { lo = i; have_lo = true; }
                if (owner > shard) { hi = i; break; }
                acc += weight(i);
            }
            if (!have_lo || lo >= hi) continue;
            Chunk mine = c;
            (void)split_head(mine, lo, ru);             // drop [0, lo)
            Chunk part = split_head(mine, hi - lo, ru);
            part.frame = c.frame;
            E2I_TRY(run(std::move(part), ss, 0, nullptr));
        }
        return E2I_OK;
    };

    unsigned long long tot[C_NCOUNTERS];
    // ---- Phase 2: leaves ----
    TRYF(cudaMemsetAsync(stripes, 0, stripe_bytes, s));
    TRYF(cudaEventRecord(ctx->ev[0], s));
    SweepStats sl;
    int rc = run_pass(true, sl);
    if (rc != E2I_OK) return fail(rc);
    TRYF(cudaEventRecord(ctx->ev[1], s));
    rc = sum_stripes(tot);
    if (rc != E2I_OK) return fail(rc);
    const uint64_t first = shard == 0 ? 1 : 0;          // lcp_values starts at 1 (ebwt2InDel.cpp:575)
    st->leaves += sl.items;
    st->levels_leaves += sl.sweeps;
    st->rank_leaves += tot[C_RANK];
    st->lcp_values_leaves += first + tot[C_LCP];
    st->lcp_values += first + tot[C_LCP];
    st->da_values += tot[C_DA];
    st->da_values_leaves += tot[C_DA];
    st->max_frontier = std::max<uint64_t>(st->max_frontier, sl.max_chunk);
    // ---- Phase 3: internal nodes ----
    TRYF(cudaMemsetAsync(stripes, 0, stripe_bytes, s));
    TRYF(cudaEventRecord(ctx->ev[2], s));
    SweepStats sn;
    rc = run_pass(false, sn);
    if (rc != E2I_OK) return fail(rc);
    TRYF(cudaEventRecord(ctx->ev[3], s));
    rc = sum_stripes(tot);
    if (rc != E2I_OK) return fail(rc);
    st->nodes += sn.items;
    st->levels_nodes += sn.sweeps;
    st->rank_nodes += tot[C_RANK];
    st->lcp_values += tot[C_LCP];
    st->n_min += tot[C_NMIN];
    st->da_values += tot[C_DA];
    st->bit_updates += tot[C_BITUPD];
    st->max_frontier = std::max<uint64_t>(st->max_frontier, sn.max_chunk);
    if (std::getenv("E2I_DEBUG"))
        std::fprintf(stderr, "[e2i] leaves: %llu sweeps, alloc %.2f ms (max %.2f), sync %.2f ms (max %.2f) | nodes: %llu sweeps, alloc %.2f ms (max %.2f), sync %.2f ms (max %.2f)\n",
                     (unsigned long long)sl.sweeps, sl.ms_alloc, sl.ms_max_alloc, sl.ms_sync, sl.ms_max_sync,
                     (unsigned long long)sn.sweeps, sn.ms_alloc, sn.ms_max_alloc, sn.ms_sync, sn.ms_max_sync);
    float ms = 0;
    TRYF(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
    st->ms_leaves += ms;
    TRYF(cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]));
    st->ms_nodes += ms;
#undef TRYF
    dfree(ctx, stripes);
    *out = l;
    if (da_out) *da_out = da;
    return E2I_OK;
}

extern "C" int e2i_navigate(e2i_ctx *ctx, const e2i_index *b1, const e2i_index *b2, const e2i_params *p,
                            e2i_lcpbits **out, e2i_bits **da_out, e2i_stats *st) {
    return e2i_navigate_shard(ctx, b1, b2, p, 0, 1, out, da_out, st);
}

extern "C" int e2i_lcpbits_fetch(e2i_ctx *ctx, const e2i_lcpbits *l, uint64_t *host_thr_words, uint64_t *host_min_words) {
    if (!ctx || !l) { set_error("e2i_lcpbits_fetch: null argument"); return E2I_ERR_ARG; }
    if (host_thr_words) E2I_CUDA_TRY(cudaMemcpy(host_thr_words, l->thr, ((2 * l->n + 63) / 64) * 8, cudaMemcpyDeviceToHost));
    if (host_min_words) E2I_CUDA_TRY(cudaMemcpy(host_min_words, l->minima, ((l->n + 63) / 64) * 8, cudaMemcpyDeviceToHost));
    return E2I_OK;
}

extern "C" int e2i_lcpbits_device(const e2i_lcpbits *l, void **dev_thr, uint64_t *thr_words32, void **dev_min, uint64_t *min_words32) {
    if (!l) { set_error("e2i_lcpbits_device: null argument"); return E2I_ERR_ARG; }
    if (dev_thr) *dev_thr = l->thr;
    if (thr_words32) *thr_words32 = l->thr_words32;
    if (dev_min) *dev_min = l->minima;
    if (min_words32) *min_words32 = l->min_words32;
    return E2I_OK;
}

extern "C" void e2i_lcpbits_free(e2i_lcpbits *l) {
    if (!l) return;
    dfree(l->ctx, l->thr);
    dfree(l->ctx, l->minima);
    delete l;
}
