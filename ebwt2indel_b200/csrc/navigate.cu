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
// distinct boundaries (neighbouring boundaries of a small node share one 64-byte index block, so
// all but the first fetch hit L1), turns the rank differences into the five sub-interval sizes of
// each child cW, keeps the children with >= 2 non-empty sub-intervals (number_of_children >= 2) and
// appends them, in order, as 32-byte compact records.  Leaves: one thread per leaf (two ranks per BWT).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <memory>

#include "common.cuh"
#include "lookback.cuh"

namespace e2i {

constexpr int kNavThreads = 256;
constexpr int kStageBlocks = 512;       // index blocks staged in shared memory per CTA (32 KB)
constexpr int kStripes = 128;           // striped statistics counters (avoid single-address atomics)
enum { C_LCP = 0, C_NMIN, C_RANK, C_BITUPD, C_DA, C_NCOUNTERS = 8 };

// Per-sweep control.  Device side: one zeroed ticket counter per sweep (a ring, so no per-sweep
// memset).  Host side: a page-locked, device-mapped block that the last tile of the sweep writes
// the four child counts into, followed by the sweep's sequence number; the host polls that word
// instead of a copy + stream synchronisation and sizes the next sweep as soon as the counts exist
// (the next launch is stream-ordered behind the running one anyway).
constexpr uint32_t kTicketSlots = 16384;
struct HostCtl {
    unsigned long long out_count[4];
    unsigned long long seq;
};

struct Segs {                 // a position-sorted run of records given as <= 4 segments
    const uint64_t *p[4];
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
    uint32_t *ticket;          // this sweep's ticket counter
    HostCtl *host;             // mapped page-locked result block
    unsigned long long seq;    // sequence number of this sweep
    uint64_t *out[4];
    uint32_t epoch;
    uint32_t n_tiles;
    uint32_t K, k_right;
    int write;                // 0: expand only (redundant top of the tree on shards != 0)
};

__device__ __forceinline__ const uint64_t *seg_record(const Segs &s, uint32_t g, int words) {
    int k = 0;
    uint32_t start = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
        if (g >= s.end[i]) { k = i + 1; start = s.end[i]; }
    return s.p[k] + (size_t)(g - start) * words;
}

__device__ __forceinline__ void stripe_add(unsigned long long *stripes, uint32_t tile, int which, unsigned long long v) {
    if (v) atomicAdd(stripes + (size_t)(tile & (kStripes - 1)) * C_NCOUNTERS + which, v);
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

// ---------------------------------------------------------------------------------------------
// Phase 3 sweep: internal nodes.  One THREAD per node (per node pair with two BWTs).
//
// Compact node record, 32 bytes = 2 x uint4 (a pair is two records, BWT 1 then BWT 2):
//   lo = { s0, s1, s2, s3 }                        low 32 bits of the sizes of the TERM,A,C,G children
//   hi = { s4, base, hi8(s0..s3), hi8(s4) | hi8(base) << 8 | depth << 16 }
// i.e. 40-bit first position, five 40-bit child sizes (first_A = base + s0, ... last = base + sum)
// and a 16-bit saturating depth (only depth >= K and depth >= k_right are ever tested,
// include.hpp:836-837).  Replaces the 56-byte sa_node (include.hpp:394-413).
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kDepthMax = 0xffffu;

__device__ __forceinline__ void unpack_node(const uint4 lo, const uint4 hi, uint64_t &base, uint64_t (&s)[5], uint32_t &depth) {
    s[0] = lo.x | ((uint64_t)(hi.z & 0xffu) << 32);
    s[1] = lo.y | ((uint64_t)((hi.z >> 8) & 0xffu) << 32);
    s[2] = lo.z | ((uint64_t)((hi.z >> 16) & 0xffu) << 32);
    s[3] = lo.w | ((uint64_t)(hi.z >> 24) << 32);
    s[4] = hi.x | ((uint64_t)(hi.w & 0xffu) << 32);
    base = hi.y | ((uint64_t)((hi.w >> 8) & 0xffu) << 32);
    depth = hi.w >> 16;
}

// children of one BWT side, kept in registers between the rank phase and the ordered append
struct ChildSide {
    uint64_t base[4];       // F[c] + rank_c(first)
    uint32_t lo[5][4];      // low 32 bits of the five sub-interval sizes of child c
    uint32_t hz[4];         // high bytes of sizes 0..3 of child c
    uint32_t h4;            // high byte of size 4, one byte per c
};

__device__ __forceinline__ void store_child(uint4 *dst, const ChildSide &k, int c, uint32_t depth1) {
    dst[0] = make_uint4(k.lo[0][c], k.lo[1][c], k.lo[2][c], k.lo[3][c]);
    dst[1] = make_uint4(k.lo[4][c], (uint32_t)k.base[c], k.hz[c],
                        ((k.h4 >> (8 * c)) & 0xffu) | ((uint32_t)(k.base[c] >> 32) << 8) | (depth1 << 16));
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

// LF(sa_node) (dna_bwt.hpp:323-356) for one node (pair): ranks at the distinct boundaries, turned
// into the five sub-interval sizes of every child.  W = uint32_t when the whole node is shorter
// than 2^32 (all but the top of the tree), uint64_t otherwise.
template <bool TWO, typename W>
__device__ __forceinline__ void expand_core(const NavArgs &a, const uint4 *stage1, uint32_t lo1, uint32_t nst1,
                                            const uint4 *stage2, uint32_t lo2, uint32_t nst2,
                                            uint64_t base1, const uint64_t (&s1)[5], uint64_t base2, const uint64_t (&s2)[5],
                                            ChildSide &k1, ChildSide &k2, uint32_t &nzp, uint32_t &st_rank) {
    constexpr bool WIDE = sizeof(W) == 8;
    uint64_t abs1[4], abs2[4] = {0, 0, 0, 0};
    rank4w<uint64_t>(a.ix1, stage1, lo1, nst1, base1, abs1);
    st_rank++;
    if (TWO) { rank4w<uint64_t>(a.ix2, stage2, lo2, nst2, base2, abs2); st_rank++; }
    W prev1[4], cur1[4], prev2[4], cur2[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        k1.base[c] = a.ix1.F[c] + abs1[c];
        prev1[c] = (W)abs1[c];
        if (TWO) { k2.base[c] = a.ix2.F[c] + abs2[c]; prev2[c] = (W)abs2[c]; }
    }
    uint64_t b1 = base1, b2 = base2;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        b1 += s1[j];
        if (s1[j]) { rank4w<W>(a.ix1, stage1, lo1, nst1, b1, cur1); st_rank++; }
        else { cur1[0] = prev1[0]; cur1[1] = prev1[1]; cur1[2] = prev1[2]; cur1[3] = prev1[3]; }
        if (TWO) {
            b2 += s2[j];
            if (s2[j]) { rank4w<W>(a.ix2, stage2, lo2, nst2, b2, cur2); st_rank++; }
            else { cur2[0] = prev2[0]; cur2[1] = prev2[1]; cur2[2] = prev2[2]; cur2[3] = prev2[3]; }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const W d1 = cur1[c] - prev1[c];
            W any = d1;
            k1.lo[j][c] = (uint32_t)d1;
            if (WIDE) { if (j < 4) k1.hz[c] |= (uint32_t)((uint64_t)d1 >> 32) << (8 * j); else k1.h4 |= (uint32_t)((uint64_t)d1 >> 32) << (8 * c); }
            prev1[c] = cur1[c];
            if (TWO) {
                const W d2 = cur2[c] - prev2[c];
                any |= d2;
                k2.lo[j][c] = (uint32_t)d2;
                if (WIDE) { if (j < 4) k2.hz[c] |= (uint32_t)((uint64_t)d2 >> 32) << (8 * j); else k2.h4 |= (uint32_t)((uint64_t)d2 >> 32) << (8 * c); }
                prev2[c] = cur2[c];
            }
            nzp += (any != 0 ? 1u : 0u) << (8 * c);
        }
    }
}

// The common case, stripped of everything it does not need: the tile's whole block range is staged,
// lies inside one 2^32-symbol superblock, and the node is shorter than 2^32.  Positions are 32-bit
// offsets into the staged window and ranks are block counter + popcount (the superblock base
// cancels in every difference and is added once, to the child's first position).
__device__ __forceinline__ void rank4_window(const uint4 *stage, uint32_t rpos, uint32_t out[4]) {
    const uint32_t r = rpos >> kBlockShift;
    const uint32_t sw = (r >> 1) & 3u;
    const uint4 *p = stage + r * 4;
    const uint4 cnt = p[sw], a = p[1u ^ sw], b = p[2u ^ sw], t = p[3u ^ sw];
    uint32_t pc[4];
    block_popc(a, b, t, (int)(rpos & (kBlockSyms - 1)), pc);
    out[0] = cnt.x + pc[0];
    out[1] = cnt.y + pc[1];
    out[2] = cnt.z + pc[2];
    out[3] = cnt.w + pc[3];
}

template <bool TWO>
__device__ __forceinline__ void expand_core_window(const NavArgs &a, const uint4 *stage1, uint32_t r1, const uint64_t (&s1)[5],
                                                   const uint4 *stage2, uint32_t r2, const uint64_t (&s2)[5],
                                                   uint64_t sup_blk1, uint64_t sup_blk2,
                                                   ChildSide &k1, ChildSide &k2, uint32_t &nzp, uint32_t &st_rank) {
    uint32_t prev1[4], cur1[4], prev2[4] = {0, 0, 0, 0}, cur2[4] = {0, 0, 0, 0};
    rank4_window(stage1, r1, prev1);
    st_rank++;
    if (TWO) { rank4_window(stage2, r2, prev2); st_rank++; }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint64_t b1 = a.ix1.F[c] + prev1[c];
        if (a.ix1.n >> kSuperShift) b1 += a.ix1.super[sup_blk1 * 4 + c];
        k1.base[c] = b1;
        if (TWO) {
            uint64_t b2 = a.ix2.F[c] + prev2[c];
            if (a.ix2.n >> kSuperShift) b2 += a.ix2.super[sup_blk2 * 4 + c];
            k2.base[c] = b2;
        }
    }
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const uint32_t z1 = (uint32_t)s1[j];
        r1 += z1;
        if (z1) { rank4_window(stage1, r1, cur1); st_rank++; }
        else { cur1[0] = prev1[0]; cur1[1] = prev1[1]; cur1[2] = prev1[2]; cur1[3] = prev1[3]; }
        if (TWO) {
            const uint32_t z2 = (uint32_t)s2[j];
            r2 += z2;
            if (z2) { rank4_window(stage2, r2, cur2); st_rank++; }
            else { cur2[0] = prev2[0]; cur2[1] = prev2[1]; cur2[2] = prev2[2]; cur2[3] = prev2[3]; }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint32_t d1 = cur1[c] - prev1[c];
            uint32_t any = d1;
            k1.lo[j][c] = d1;
            prev1[c] = cur1[c];
            if (TWO) {
                const uint32_t d2 = cur2[c] - prev2[c];
                any |= d2;
                k2.lo[j][c] = d2;
                prev2[c] = cur2[c];
            }
            nzp += (any != 0 ? 1u : 0u) << (8 * c);
        }
    }
}

#ifndef E2I_NODE_MINBLOCKS
#define E2I_NODE_MINBLOCKS 4
#endif
template <bool TWO>
__global__ void __launch_bounds__(kNavThreads, TWO ? 2 : E2I_NODE_MINBLOCKS)
expand_nodes_kernel(const NavArgs a, const Segs in) {
    constexpr int WORDS = TWO ? 8 : 4;                 // u64 words per record
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_wpk[kNavThreads / 32];       // per-warp child counts, 4 x 8 bit
    __shared__ uint32_t s_wlo[kNavThreads / 32], s_whi[kNavThreads / 32];   // exclusive prefix, 2 x 16 bit each
    __shared__ unsigned long long s_base[4];
    __shared__ unsigned long long s_stat[C_NCOUNTERS];
    __shared__ uint32_t s_rng[4];                      // first / last index block touched by the tile, per BWT
    constexpr int STAGE = TWO ? kStageBlocks / 2 : kStageBlocks;   // blocks staged per BWT
    __shared__ uint4 s_stage[kStageBlocks * 4];

    if (threadIdx.x == 0) s_tile = atomicAdd(a.ticket, 1u);
    if (threadIdx.x < C_NCOUNTERS) s_stat[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t g = tile * kNavThreads + threadIdx.x;
    const bool active = g < in.total;

    uint64_t base1 = 0, s1[5] = {0, 0, 0, 0, 0}, base2 = 0, s2[5] = {0, 0, 0, 0, 0};
    uint32_t depth = 0;
    bool narrow = true;                                // every size of the node (pair) fits 32 bits, and so does their sum
    if (active) {
        const uint4 *rec = reinterpret_cast<const uint4 *>(seg_record(in, g, WORDS));
        const uint4 lo = __ldg(rec), hi = __ldg(rec + 1);
        unpack_node(lo, hi, base1, s1, depth);
        narrow = (hi.z | (hi.w & 0xffu)) == 0 && ((s1[0] + s1[1] + s1[2] + s1[3] + s1[4]) >> 32) == 0;
        if (TWO) {
            uint32_t d2;
            const uint4 lo2 = __ldg(rec + 2), hi2 = __ldg(rec + 3);
            unpack_node(lo2, hi2, base2, s2, d2);
            narrow = narrow && (hi2.z | (hi2.w & 0xffu)) == 0 && ((s2[0] + s2[1] + s2[2] + s2[3] + s2[4]) >> 32) == 0;
        }
    }
    uint32_t st_lcp = 0, st_min = 0, st_rank = 0, st_upd = 0, st_da = 0;

    // ---- stage the index blocks of the tile in shared memory ----
    // All nodes of a sweep have the same depth, so their intervals are disjoint and the frontier is
    // sorted: the tile touches the block range [block(first node), block(end of last node)].  When
    // that range is dense enough it is copied once with 16-byte asynchronous copies (LDGSTS) while
    // the threads do their bit updates, and the up to 6 rank queries per node read shared memory;
    // boundaries outside the window fall back to global loads.
    {
        const uint32_t last_active = min((uint32_t)kNavThreads, in.total - tile * kNavThreads) - 1;
        if (threadIdx.x == 0) { s_rng[0] = (uint32_t)(base1 >> kBlockShift); if (TWO) s_rng[2] = (uint32_t)(base2 >> kBlockShift); }
        if (threadIdx.x == last_active) {
            s_rng[1] = (uint32_t)((base1 + s1[0] + s1[1] + s1[2] + s1[3] + s1[4]) >> kBlockShift);
            if (TWO) s_rng[3] = (uint32_t)((base2 + s2[0] + s2[1] + s2[2] + s2[3] + s2[4]) >> kBlockShift);
        }
    }
    __syncthreads();
    const uint32_t lo1 = s_rng[0], lo2 = TWO ? s_rng[2] : 0u;
    uint32_t nst1 = 0, nst2 = 0;
    {
        constexpr int ITER = STAGE * 4 / kNavThreads;
        const uint32_t span1 = s_rng[1] >= lo1 ? s_rng[1] - lo1 + 1 : 0u;
        if (span1 <= 2u * STAGE) nst1 = min(span1, (uint32_t)STAGE);
        const uint4 *src1 = a.ix1.blocks + (size_t)lo1 * 4;
#pragma unroll
        for (int it = 0; it < ITER; ++it) {
            const uint32_t k = threadIdx.x + it * kNavThreads;
            if (k < nst1 * 4) cp_async16(&s_stage[stage_slot(k >> 2, k & 3)], src1 + k);
        }
        if (TWO) {
            const uint32_t end2 = s_rng[3] + 1;
            const uint32_t span2 = end2 > lo2 ? end2 - lo2 : 0u;
            if (span2 <= 2u * STAGE) nst2 = min(span2, (uint32_t)STAGE);
            const uint4 *src2 = a.ix2.blocks + (size_t)lo2 * 4;
#pragma unroll
            for (int it = 0; it < ITER; ++it) {
                const uint32_t k = threadIdx.x + it * kNavThreads;
                if (k < nst2 * 4) cp_async16(&s_stage[STAGE * 4 + stage_slot(k >> 2, k & 3)], src2 + k);
            }
        }
    }
    const uint4 *stage1 = s_stage, *stage2 = s_stage + STAGE * 4;

    // ---- bit updates on the merged node (merge_nodes, include.hpp:476-490), while the copies are in flight ----
    if (active && a.write) {
        const uint64_t mbase = base1 + base2;
        uint64_t ms[5], last = mbase;
#pragma unroll
        for (int j = 0; j < 5; ++j) { ms[j] = s1[j] + s2[j]; last += ms[j]; }
        const uint32_t bits = (depth >= a.K ? 1u : 0u) | (depth >= a.k_right ? 2u : 0u);
        WordAcc thr{a.thr, ~0ull, 0u}, mn{a.minima, ~0ull, 0u};
        uint64_t mb = mbase;
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            if (TWO) {
                // find_leaves (ebwt2InDel.cpp:474-527): children of summed size exactly 1
                if (ms[j] == 1) {
                    st_da++;
                    if (s2[j] == 1) atomicOr(a.da + (mb >> 5), 1u << (mb & 31));
                }
            }
            mb += ms[j];                           // border after child j = first position of child j+1
            if (j < 4 && mb != last) {
                // update_lcp_threshold (include.hpp:826-860): border written iff the child before it is non-empty
                if (ms[j] > 0) {
                    st_lcp++;
                    if (bits) { thr.add(mb >> 4, bits << ((mb & 15) * 2)); st_upd++; }
                }
                // update_lcp_minima (ebwt2InDel.cpp:357-391): after children A, C, G of size >= 2
                if (j >= 1 && ms[j] >= 2 && mb < last - 1) {
                    st_min++;
                    st_upd++;
                    mn.add(mb >> 5, 1u << (mb & 31));
                }
            }
        }
        thr.flush();
        mn.flush();
    }
    cp_async_wait_all();
    __syncthreads();

    // ---- ranks -> children ----
    ChildSide k1, k2;
    uint32_t nzp = 0;                                  // per symbol: number of non-empty gaps (union of both BWTs), 4 x 8 bit
    k1.h4 = 0; k2.h4 = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) { k1.hz[c] = 0; k2.hz[c] = 0; k1.base[c] = 0; k2.base[c] = 0; }
    if (active) {
        if (narrow) expand_core<TWO, uint32_t>(a, stage1, lo1, nst1, stage2, lo2, nst2, base1, s1, base2, s2, k1, k2, nzp, st_rank);
        else expand_core<TWO, uint64_t>(a, stage1, lo1, nst1, stage2, lo2, nst2, base1, s1, base2, s2, k1, k2, nzp, st_rank);
    }

    // ---- child c is right-maximal iff >= 2 of its 5 gaps are non-empty (number_of_children, include.hpp:760-792) ----
    uint32_t vm = 0, packed = 0, before[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const bool v = ((nzp >> (8 * c)) & 0xffu) >= 2u;
        const uint32_t bal = __ballot_sync(0xffffffffu, v);
        before[c] = __popc(bal & ((1u << lane) - 1u));
        packed |= (uint32_t)__popc(bal) << (8 * c);     // <= 32 per warp and symbol
        if (v) vm |= 1u << c;
    }
    if (lane == 0) s_wpk[warp] = packed;
    st_lcp = __reduce_add_sync(0xffffffffu, st_lcp);
    st_min = __reduce_add_sync(0xffffffffu, st_min);
    st_rank = __reduce_add_sync(0xffffffffu, st_rank);
    st_upd = __reduce_add_sync(0xffffffffu, st_upd);
    if (TWO) st_da = __reduce_add_sync(0xffffffffu, st_da);
    if (lane == 0) {
        if (st_lcp) atomicAdd(&s_stat[C_LCP], (unsigned long long)st_lcp);
        if (st_min) atomicAdd(&s_stat[C_NMIN], (unsigned long long)st_min);
        if (st_rank) atomicAdd(&s_stat[C_RANK], (unsigned long long)st_rank);
        if (st_upd) atomicAdd(&s_stat[C_BITUPD], (unsigned long long)st_upd);
        if (TWO && st_da) atomicAdd(&s_stat[C_DA], (unsigned long long)st_da);
    }
    __syncthreads();
    if (warp == 0) {
        // exclusive scan of the 8 per-warp counts (16-bit fields: up to 256 per symbol and tile), then the look-back
        const uint32_t mine = lane < kNavThreads / 32 ? s_wpk[lane] : 0u;
        uint32_t lo = (mine & 0xffu) | (((mine >> 8) & 0xffu) << 16);          // A, C
        uint32_t hi = ((mine >> 16) & 0xffu) | ((mine >> 24) << 16);           // G, T
        const uint32_t mlo = lo, mhi = hi;
#pragma unroll
        for (int s = 1; s < 8; s <<= 1) {
            const uint32_t ylo = __shfl_up_sync(0xffffffffu, lo, s), yhi = __shfl_up_sync(0xffffffffu, hi, s);
            if (lane >= s) { lo += ylo; hi += yhi; }
        }
        if (lane < kNavThreads / 32) { s_wlo[lane] = lo - mlo; s_whi[lane] = hi - mhi; }
        const uint32_t tlo = __shfl_sync(0xffffffffu, lo, 7), thi = __shfl_sync(0xffffffffu, hi, 7);
        const uint32_t agg[4] = {tlo & 0xffffu, tlo >> 16, thi & 0xffffu, thi >> 16};
        unsigned long long excl[4];
        lookback4(a.desc, a.epoch, tile, agg, excl);
        if (lane < 4) {
            unsigned long long e = 0, g2 = 0;
#pragma unroll
            for (int c = 0; c < 4; ++c) if (lane == c) { e = excl[c]; g2 = agg[c]; }
            s_base[lane] = e;
            if (tile == a.n_tiles - 1) ((volatile unsigned long long *)a.host->out_count)[lane] = e + g2;
        }
        if (tile == a.n_tiles - 1) {                       // tell the host that the counts of this sweep are final
            __threadfence_system();
            __syncwarp();
            if (lane == 0) *(volatile unsigned long long *)&a.host->seq = a.seq;
        }
        if (lane < C_NCOUNTERS && a.write) stripe_add(a.stripes, tile, lane, s_stat[lane]);
    }
    __syncthreads();
    // ---- ordered append of the surviving children ----
    if (vm) {
        const uint32_t exlo = s_wlo[warp], exhi = s_whi[warp];
        const uint32_t exw[4] = {exlo & 0xffffu, exlo >> 16, exhi & 0xffffu, exhi >> 16};
        const uint32_t depth1 = depth >= kDepthMax ? kDepthMax : depth + 1;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if ((vm >> c) & 1u) {
                const unsigned long long slot = s_base[c] + exw[c] + before[c];
                uint4 *dst = reinterpret_cast<uint4 *>(a.out[c] + slot * WORDS);
                store_child(dst, k1, c, depth1);
                if (TWO) store_child(dst + 2, k2, c, depth1);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Phase 3 sweep, persistent warp-specialised form (the default).
//
// The ordered append makes a tile wait until every earlier tile has published its child counts;
// with ~600 tiles in flight that wait was 40 % of the warp time of the one-tile-per-CTA kernel
// above (profiles/r01_ncu_nodes_c4s16_v4_raw.csv).  Here a CTA is 8 compute warps + 1 scan warp
// and loops over tiles taken by ticket:
//   compute warps  tile t+1: records -> staged index window -> bit updates -> ranks -> children in
//                  registers; THEN flush the children of tile t from shared memory to their final,
//                  now resolved, global slots (coalesced 16-byte stores); park the children of t+1
//                  in shared memory and post their counts to the scan warp;
//   scan warp      publishes the counts of a posted tile at once, resolves its exclusive prefix by
//                  look-back while the compute warps already work on the next tile, and hands the
//                  four base slots back.
// Publication is never delayed, so the look-back window stays short; nothing waits unless the
// prefix of tile t is still unresolved after the whole compute phase of tile t+1.
// ---------------------------------------------------------------------------------------------
constexpr int kCompThreads = 256;                     // 8 compute warps
constexpr int kPersistThreads = kCompThreads + 32;    // + the scan warp
constexpr uint32_t kExitSeq = 0xffffffffu;

__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

template <bool TWO>
struct PersistSmem {
    static constexpr int RU = TWO ? 4 : 2;            // uint4 per record
    uint4 stage[kStageBlocks * 4];                    // staged index window(s)
    uint4 child[4][kCompThreads * RU];                // parked children of the pending tile, per symbol
    uint4 recbuf[kCompThreads * RU];                  // records of the NEXT tile, prefetched by LDGSTS (slot = thread)
    unsigned long long base[4];                       // resolved exclusive prefix of the pending tile
    unsigned long long stat[C_NCOUNTERS];
    uint32_t agg[4];                                  // child counts of the pending tile
    uint32_t pend_tile;
    uint32_t seq_posted, seq_done;                    // handshake compute warps <-> scan warp
    uint32_t tile;
    uint32_t rng[4];
    uint32_t wpk[kCompThreads / 32];
};

template <bool TWO>
__global__ void __launch_bounds__(kPersistThreads, TWO ? 2 : 3)
expand_nodes_persistent(const NavArgs a, const Segs in) {
    constexpr int WORDS = TWO ? 8 : 4;
    constexpr int RU = TWO ? 4 : 2;
    constexpr int STAGE = TWO ? kStageBlocks / 2 : kStageBlocks;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PersistSmem<TWO> &sm = *reinterpret_cast<PersistSmem<TWO> *>(smem_raw);
    volatile uint32_t *v_posted = &sm.seq_posted, *v_done = &sm.seq_done;
    volatile unsigned long long *v_base = sm.base;

    if (threadIdx.x == 0) { sm.seq_posted = 0; sm.seq_done = 0; }
    __syncthreads();

    if (threadIdx.x >= kCompThreads) {
        // ------------------------------- scan warp -------------------------------
        const int lane = threadIdx.x & 31;
        uint32_t seen = 0;
        while (true) {
            uint32_t p = *v_posted;
            while (p == seen) { __nanosleep(40); p = *v_posted; }
            if (p == kExitSeq) break;
            seen = p;
            __threadfence_block();
            const uint32_t tile = *(volatile uint32_t *)&sm.pend_tile;
            uint32_t agg[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) agg[c] = ((volatile uint32_t *)sm.agg)[c];
            unsigned long long excl[4];
            lookback4(a.desc, a.epoch, tile, agg, excl);
            if (lane < 4) {
                unsigned long long e = 0, g2 = 0;
#pragma unroll
                for (int c = 0; c < 4; ++c) if (lane == c) { e = excl[c]; g2 = agg[c]; }
                v_base[lane] = e;
                if (tile == a.n_tiles - 1) ((volatile unsigned long long *)a.host->out_count)[lane] = e + g2;
            }
            if (tile == a.n_tiles - 1) {                   // tell the host that the counts of this sweep are final
                __threadfence_system();
                __syncwarp();
                if (lane == 0) *(volatile unsigned long long *)&a.host->seq = a.seq;
            }
            __threadfence_block();
            __syncwarp();
            if (lane == 0) *v_done = seen;
        }
        return;
    }

    // --------------------------------- compute warps ---------------------------------
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t my_seq = 0;                    // tiles this CTA has posted so far
    const uint4 *stage1 = sm.stage, *stage2 = sm.stage + STAGE * 4;

    auto flush_pending = [&]() {
        // children of the pending tile: shared memory -> their resolved global slots
        while (*v_done != my_seq) { }
        __threadfence_block();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint32_t n16 = ((volatile uint32_t *)sm.agg)[c] * RU;
            uint4 *dst = reinterpret_cast<uint4 *>(a.out[c]) + v_base[c] * RU;
            for (uint32_t i = threadIdx.x; i < n16; i += kCompThreads) dst[i] = sm.child[c][i];
        }
    };

    // software pipeline: the ticket and the records of tile t+1 are fetched while tile t is processed
    auto prefetch_records = [&](uint32_t tl) {
        const uint32_t g = tl * kCompThreads + threadIdx.x;
        if (tl < a.n_tiles && g < in.total) {
            const uint4 *rec = reinterpret_cast<const uint4 *>(seg_record(in, g, WORDS));
#pragma unroll
            for (int k = 0; k < RU; ++k) cp_async16(&sm.recbuf[threadIdx.x * RU + k], rec + k);
        }
    };
    if (threadIdx.x == 0) sm.tile = atomicAdd(a.ticket, 1u);
    bar_compute();
    uint32_t tile = sm.tile;
    prefetch_records(tile);
    while (tile < a.n_tiles) {
        uint32_t nxt = 0;
        if (threadIdx.x == 0) nxt = atomicAdd(a.ticket, 1u);          // its latency hides behind this tile
        if (threadIdx.x < C_NCOUNTERS) sm.stat[threadIdx.x] = 0;
        const uint32_t g = tile * kCompThreads + threadIdx.x;
        const bool active = g < in.total;

        uint64_t base1 = 0, s1[5] = {0, 0, 0, 0, 0}, base2 = 0, s2[5] = {0, 0, 0, 0, 0};
        uint32_t depth = 0;
        bool narrow = true;
        cp_async_wait_all();                                                // this thread's own record has landed
        if (active) {
            const uint4 *rec = &sm.recbuf[threadIdx.x * RU];
            const uint4 lo = rec[0], hi = rec[1];
            unpack_node(lo, hi, base1, s1, depth);
            narrow = (hi.z | (hi.w & 0xffu)) == 0 && ((s1[0] + s1[1] + s1[2] + s1[3] + s1[4]) >> 32) == 0;
            if (TWO) {
                uint32_t d2;
                const uint4 lo2 = rec[2], hi2 = rec[3];
                unpack_node(lo2, hi2, base2, s2, d2);
                narrow = narrow && (hi2.z | (hi2.w & 0xffu)) == 0 && ((s2[0] + s2[1] + s2[2] + s2[3] + s2[4]) >> 32) == 0;
            }
        }
        uint32_t st_lcp = 0, st_min = 0, st_rank = 0, st_upd = 0, st_da = 0;
        {
            const uint32_t last_active = min((uint32_t)kCompThreads, in.total - tile * kCompThreads) - 1;
            if (threadIdx.x == 0) { sm.rng[0] = (uint32_t)(base1 >> kBlockShift); if (TWO) sm.rng[2] = (uint32_t)(base2 >> kBlockShift); }
            if (threadIdx.x == last_active) {
                sm.rng[1] = (uint32_t)((base1 + s1[0] + s1[1] + s1[2] + s1[3] + s1[4]) >> kBlockShift);
                if (TWO) sm.rng[3] = (uint32_t)((base2 + s2[0] + s2[1] + s2[2] + s2[3] + s2[4]) >> kBlockShift);
            }
        }
        bar_compute();
        const uint32_t lo1 = sm.rng[0], lo2 = TWO ? sm.rng[2] : 0u;
        uint32_t nst1 = 0, nst2 = 0, span1_all = 0, span2_all = 0;
        {
            constexpr int ITER = STAGE * 4 / kCompThreads;
            const uint32_t span1 = sm.rng[1] >= lo1 ? sm.rng[1] - lo1 + 1 : 0u;
            span1_all = span1;
            if (span1 <= 2u * STAGE) nst1 = min(span1, (uint32_t)STAGE);
            const uint4 *src1 = a.ix1.blocks + (size_t)lo1 * 4;
#pragma unroll
            for (int it = 0; it < ITER; ++it) {
                const uint32_t k = threadIdx.x + it * kCompThreads;
                if (k < nst1 * 4) cp_async16(&sm.stage[stage_slot(k >> 2, k & 3)], src1 + k);
            }
            if (TWO) {
                const uint32_t end2 = sm.rng[3] + 1;
                const uint32_t span2 = end2 > lo2 ? end2 - lo2 : 0u;
                span2_all = span2;
                if (span2 <= 2u * STAGE) nst2 = min(span2, (uint32_t)STAGE);
                const uint4 *src2 = a.ix2.blocks + (size_t)lo2 * 4;
#pragma unroll
                for (int it = 0; it < ITER; ++it) {
                    const uint32_t k = threadIdx.x + it * kCompThreads;
                    if (k < nst2 * 4) cp_async16(&sm.stage[STAGE * 4 + stage_slot(k >> 2, k & 3)], src2 + k);
                }
            }
        }
        // bit updates on the merged node while the copies are in flight (same rules as expand_nodes_kernel)
        if (active && a.write) {
            const uint64_t mbase = base1 + base2;
            uint64_t ms[5], last = mbase;
#pragma unroll
            for (int j = 0; j < 5; ++j) { ms[j] = s1[j] + s2[j]; last += ms[j]; }
            const uint32_t bits = (depth >= a.K ? 1u : 0u) | (depth >= a.k_right ? 2u : 0u);
            WordAcc thr{a.thr, ~0ull, 0u}, mn{a.minima, ~0ull, 0u};
            uint64_t mb = mbase;
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                if (TWO) {
                    if (ms[j] == 1) {                        // find_leaves (ebwt2InDel.cpp:474-527)
                        st_da++;
                        if (s2[j] == 1) atomicOr(a.da + (mb >> 5), 1u << (mb & 31));
                    }
                }
                mb += ms[j];
                if (j < 4 && mb != last) {
                    if (ms[j] > 0) {                         // update_lcp_threshold (include.hpp:826-860)
                        st_lcp++;
                        if (bits) { thr.add(mb >> 4, bits << ((mb & 15) * 2)); st_upd++; }
                    }
                    if (j >= 1 && ms[j] >= 2 && mb < last - 1) {   // update_lcp_minima (ebwt2InDel.cpp:357-391)
                        st_min++;
                        st_upd++;
                        mn.add(mb >> 5, 1u << (mb & 31));
                    }
                }
            }
            thr.flush();
            mn.flush();
        }
        if (threadIdx.x == 0) sm.tile = nxt;
        cp_async_wait_all();
        bar_compute();
        const uint32_t next_tile = sm.tile;
        prefetch_records(next_tile);                                         // overlaps with the rank phase below

        ChildSide k1, k2;
        uint32_t nzp = 0;
        k1.h4 = 0; k2.h4 = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) { k1.hz[c] = 0; k2.hz[c] = 0; k1.base[c] = 0; k2.base[c] = 0; }
        // whole range staged and inside one superblock (CTA-uniform), for both BWTs
        bool window = nst1 == span1_all && nst1 > 0 && (lo1 >> (kSuperShift - kBlockShift)) == ((lo1 + nst1 - 1) >> (kSuperShift - kBlockShift));
        if (TWO) window = window && nst2 == span2_all && nst2 > 0 &&
                          (lo2 >> (kSuperShift - kBlockShift)) == ((lo2 + nst2 - 1) >> (kSuperShift - kBlockShift));
        if (active) {
            if (window && narrow)
                expand_core_window<TWO>(a, stage1, (uint32_t)(base1 - ((uint64_t)lo1 << kBlockShift)), s1,
                                        stage2, (uint32_t)(base2 - ((uint64_t)lo2 << kBlockShift)), s2,
                                        lo1 >> (kSuperShift - kBlockShift), lo2 >> (kSuperShift - kBlockShift), k1, k2, nzp, st_rank);
            else if (narrow) expand_core<TWO, uint32_t>(a, stage1, lo1, nst1, stage2, lo2, nst2, base1, s1, base2, s2, k1, k2, nzp, st_rank);
            else expand_core<TWO, uint64_t>(a, stage1, lo1, nst1, stage2, lo2, nst2, base1, s1, base2, s2, k1, k2, nzp, st_rank);
        }
        uint32_t vm = 0, packed = 0, before[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const bool v = ((nzp >> (8 * c)) & 0xffu) >= 2u;
            const uint32_t bal = __ballot_sync(0xffffffffu, v);
            before[c] = __popc(bal & ((1u << lane) - 1u));
            packed |= (uint32_t)__popc(bal) << (8 * c);
            if (v) vm |= 1u << c;
        }
        if (lane == 0) sm.wpk[warp] = packed;
        st_lcp = __reduce_add_sync(0xffffffffu, st_lcp);
        st_min = __reduce_add_sync(0xffffffffu, st_min);
        st_rank = __reduce_add_sync(0xffffffffu, st_rank);
        st_upd = __reduce_add_sync(0xffffffffu, st_upd);
        if (TWO) st_da = __reduce_add_sync(0xffffffffu, st_da);
        if (lane == 0) {
            if (st_lcp) atomicAdd(&sm.stat[C_LCP], (unsigned long long)st_lcp);
            if (st_min) atomicAdd(&sm.stat[C_NMIN], (unsigned long long)st_min);
            if (st_rank) atomicAdd(&sm.stat[C_RANK], (unsigned long long)st_rank);
            if (st_upd) atomicAdd(&sm.stat[C_BITUPD], (unsigned long long)st_upd);
            if (TWO && st_da) atomicAdd(&sm.stat[C_DA], (unsigned long long)st_da);
        }
        // the staged window and the per-warp counts are complete; the previous tile's children can leave
        if (my_seq) flush_pending();
        bar_compute();                                   // wpk visible; child buffer and agg free again
        uint32_t exw[4] = {0, 0, 0, 0}, tot[4] = {0, 0, 0, 0};
#pragma unroll
        for (int w = 0; w < kCompThreads / 32; ++w) {
            const uint32_t pk = sm.wpk[w];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint32_t n = (pk >> (8 * c)) & 0xffu;
                if (w < warp) exw[c] += n;
                tot[c] += n;
            }
        }
        if (vm) {
            const uint32_t depth1 = depth >= kDepthMax ? kDepthMax : depth + 1;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if ((vm >> c) & 1u) {
                    uint4 *dst = &sm.child[c][(exw[c] + before[c]) * RU];
                    store_child(dst, k1, c, depth1);
                    if (TWO) store_child(dst + 2, k2, c, depth1);
                }
            }
        }
        if (threadIdx.x == 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c) sm.agg[c] = tot[c];
            sm.pend_tile = tile;
        }
        if (threadIdx.x < C_NCOUNTERS && a.write) stripe_add(a.stripes, tile, threadIdx.x, sm.stat[threadIdx.x]);
        bar_compute();                                   // children, counts and tile id are in shared memory
        ++my_seq;
        if (threadIdx.x == 0) { __threadfence_block(); *v_posted = my_seq; }
        tile = next_tile;
    }
    if (my_seq) flush_pending();
    bar_compute();
    if (threadIdx.x == 0) *v_posted = kExitSeq;
}

// ---------------------------------------------------------------------------------------------
// Phase 2 sweep: leaves (intervals of W#).  One thread per leaf (pair).
// Record = 4 u64 {first, second, depth, 0}; mode -2: 8 u64 {f1, s1, depth, 0, f2, s2, 0, 0}.
// ---------------------------------------------------------------------------------------------
template <bool TWO>
__global__ void __launch_bounds__(kNavThreads)
expand_leaves_kernel(const NavArgs a, const Segs in) {
    constexpr int WORDS = TWO ? 8 : 4;
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_cnt[8];
    __shared__ uint32_t s_excl[8];
    __shared__ unsigned long long s_base[4];
    __shared__ unsigned long long s_stat[C_NCOUNTERS];
    if (threadIdx.x == 0) s_tile = atomicAdd(a.ticket, 1u);
    if (threadIdx.x < C_NCOUNTERS) s_stat[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t g = tile * kNavThreads + threadIdx.x;
    const bool active = g < in.total;
    uint64_t f1 = 0, s1 = 0, f2 = 0, s2 = 0, depth = 0;
    if (active) {
        const ulonglong2 *rec = reinterpret_cast<const ulonglong2 *>(seg_record(in, g, WORDS));
        const ulonglong2 x = rec[0], y = rec[1];
        f1 = x.x; s1 = x.y; depth = y.x;
        if (TWO) { const ulonglong2 z = rec[2]; f2 = z.x; s2 = z.y; }
    }
    unsigned long long st_lcp = 0, st_da = 0;
    uint32_t st_rank = 0;
    if (active && a.write) {
        // update_LCP_leaf (:344-355) / update_DA (:394-425) at merged coordinates
        const uint64_t start1 = f1 + f2, start2 = f2 + s1, end = s1 + s2;
        if (end > start1) st_lcp = end - start1 - 1;
        const uint32_t pat = (depth >= a.K ? 0x55555555u : 0u) | (depth >= a.k_right ? 0xaaaaaaaau : 0u);
        if (end > start1 + 1) fill_bits(a.thr, 2 * (start1 + 1), 2 * end, pat);
        if (TWO) {
            st_da = end - start1;
            fill_bits(a.da, start2, end, 0xffffffffu);
        }
    }
    // next_leaves (dna_bwt.hpp:358-379; two BWTs: ebwt2InDel.cpp:452-472): LF(range) = 2 ranks per BWT
    uint64_t lo1[4] = {0, 0, 0, 0}, hi1[4] = {0, 0, 0, 0}, lo2[4] = {0, 0, 0, 0}, hi2[4] = {0, 0, 0, 0};
    if (active) {
        rank4(a.ix1, f1, lo1);
        st_rank++;
        if (s1 > f1) { rank4(a.ix1, s1, hi1); st_rank++; }
        else { hi1[0] = lo1[0]; hi1[1] = lo1[1]; hi1[2] = lo1[2]; hi1[3] = lo1[3]; }
        if (TWO) {
            rank4(a.ix2, f2, lo2);
            st_rank++;
            if (s2 > f2) { rank4(a.ix2, s2, hi2); st_rank++; }
            else { hi2[0] = lo2[0]; hi2[1] = lo2[1]; hi2[2] = lo2[2]; hi2[3] = lo2[3]; }
        }
    }
    uint32_t vm = 0, packed = 0, before[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const bool v = active && ((hi1[c] - lo1[c]) + (hi2[c] - lo2[c]) >= 2);
        const uint32_t bal = __ballot_sync(0xffffffffu, v);
        before[c] = __popc(bal & ((1u << lane) - 1u));
        packed |= (uint32_t)__popc(bal) << (8 * c);
        if (v) vm |= 1u << c;
    }
    // warp-level totals can reach 32 per symbol: 8-bit fields hold up to 255 per tile (256 leaves: use 9+ bits)
    // -> keep per-warp counts in 8-bit fields but accumulate the tile scan in two u32 (16-bit fields).
    if (lane == 0) s_cnt[warp] = packed;
    st_lcp += __shfl_xor_sync(0xffffffffu, st_lcp, 16);
    st_lcp += __shfl_xor_sync(0xffffffffu, st_lcp, 8);
    st_lcp += __shfl_xor_sync(0xffffffffu, st_lcp, 4);
    st_lcp += __shfl_xor_sync(0xffffffffu, st_lcp, 2);
    st_lcp += __shfl_xor_sync(0xffffffffu, st_lcp, 1);
    if (TWO) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) st_da += __shfl_xor_sync(0xffffffffu, st_da, s);
    }
    st_rank = __reduce_add_sync(0xffffffffu, st_rank);
    if (lane == 0) {
        if (st_lcp) atomicAdd(&s_stat[C_LCP], st_lcp);
        if (st_rank) atomicAdd(&s_stat[C_RANK], (unsigned long long)st_rank);
        if (TWO && st_da) atomicAdd(&s_stat[C_DA], st_da);
    }
    __syncthreads();
    if (warp == 0) {
        // 8 per-warp entries; widen to 16-bit fields (two words) before the scan
        const uint32_t mine = lane < 8 ? s_cnt[lane] : 0u;
        uint32_t lo = (mine & 0xffu) | (((mine >> 8) & 0xffu) << 16);          // A, C
        uint32_t hi = ((mine >> 16) & 0xffu) | ((mine >> 24) << 16);           // G, T
        const uint32_t mlo = lo, mhi = hi;
#pragma unroll
        for (int s = 1; s < 8; s <<= 1) {
            const uint32_t ylo = __shfl_up_sync(0xffffffffu, lo, s), yhi = __shfl_up_sync(0xffffffffu, hi, s);
            if (lane >= s) { lo += ylo; hi += yhi; }
        }
        if (lane < 8) {
            s_excl[lane] = 0;  // unused
            // store exclusive prefix as two words in s_cnt/s_excl
            s_cnt[lane] = lo - mlo;
            s_excl[lane] = hi - mhi;
        }
        const uint32_t tlo = __shfl_sync(0xffffffffu, lo, 7), thi = __shfl_sync(0xffffffffu, hi, 7);
        const uint32_t agg[4] = {tlo & 0xffffu, tlo >> 16, thi & 0xffffu, thi >> 16};
        unsigned long long excl[4];
        lookback4(a.desc, a.epoch, tile, agg, excl);      // all four queues in one look-back round
        if (lane < 4) {
            unsigned long long e = 0, g2 = 0;
#pragma unroll
            for (int c = 0; c < 4; ++c) if (lane == c) { e = excl[c]; g2 = agg[c]; }
            s_base[lane] = e;
            if (tile == a.n_tiles - 1) ((volatile unsigned long long *)a.host->out_count)[lane] = e + g2;
        }
        if (tile == a.n_tiles - 1) {                       // tell the host that the counts of this sweep are final
            __threadfence_system();
            __syncwarp();
            if (lane == 0) *(volatile unsigned long long *)&a.host->seq = a.seq;
        }
        if (lane < C_NCOUNTERS && a.write) stripe_add(a.stripes, tile, lane, s_stat[lane]);
    }
    __syncthreads();
    const uint32_t exlo = s_cnt[warp], exhi = s_excl[warp];
    const uint32_t exw[4] = {exlo & 0xffffu, exlo >> 16, exhi & 0xffffu, exhi >> 16};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        if ((vm >> c) & 1u) {
            const unsigned long long slot = s_base[c] + exw[c] + before[c];
            ulonglong2 *o = reinterpret_cast<ulonglong2 *>(a.out[c] + slot * WORDS);
            o[0] = make_ulonglong2(a.ix1.F[c] + lo1[c], a.ix1.F[c] + hi1[c]);
            o[1] = make_ulonglong2(depth + 1, 0);
            if (TWO) {
                o[2] = make_ulonglong2(a.ix2.F[c] + lo2[c], a.ix2.F[c] + hi2[c]);
                o[3] = make_ulonglong2(0, 0);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Phase 2 sweep, persistent warp-specialised form (same structure as expand_nodes_persistent: the
// one-tile-per-CTA leaf kernel spent most of its warp time waiting behind the ordered look-back,
// profiles/r01_ncu_leaves_c4s16_raw.csv).  Leaves are sparse in position space, so there is no
// staged index window: the two (four) rank queries of a leaf read global memory.
// ---------------------------------------------------------------------------------------------
template <bool TWO>
struct LeafSmem {
    static constexpr int RU = TWO ? 4 : 2;            // uint4 per record
    uint4 child[4][kCompThreads * RU];                // parked children of the pending tile, per symbol
    uint4 recbuf[kCompThreads * RU];                  // records of the next tile (LDGSTS, slot = thread)
    unsigned long long base[4];
    unsigned long long stat[C_NCOUNTERS];
    uint32_t agg[4];
    uint32_t pend_tile;
    uint32_t seq_posted, seq_done;
    uint32_t tile;
    uint32_t wpk[kCompThreads / 32];
};

template <bool TWO>
__global__ void __launch_bounds__(kPersistThreads, TWO ? 2 : 3)
expand_leaves_persistent(const NavArgs a, const Segs in) {
    constexpr int WORDS = TWO ? 8 : 4;
    constexpr int RU = TWO ? 4 : 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LeafSmem<TWO> &sm = *reinterpret_cast<LeafSmem<TWO> *>(smem_raw);
    volatile uint32_t *v_posted = &sm.seq_posted, *v_done = &sm.seq_done;
    volatile unsigned long long *v_base = sm.base;

    if (threadIdx.x == 0) { sm.seq_posted = 0; sm.seq_done = 0; }
    __syncthreads();

    if (threadIdx.x >= kCompThreads) {
        // ------------------------------- scan warp (as in expand_nodes_persistent) -------------------------------
        const int lane = threadIdx.x & 31;
        uint32_t seen = 0;
        while (true) {
            uint32_t p = *v_posted;
            while (p == seen) { __nanosleep(40); p = *v_posted; }
            if (p == kExitSeq) break;
            seen = p;
            __threadfence_block();
            const uint32_t tile = *(volatile uint32_t *)&sm.pend_tile;
            uint32_t agg[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) agg[c] = ((volatile uint32_t *)sm.agg)[c];
            unsigned long long excl[4];
            lookback4(a.desc, a.epoch, tile, agg, excl);
            if (lane < 4) {
                unsigned long long e = 0, g2 = 0;
#pragma unroll
                for (int c = 0; c < 4; ++c) if (lane == c) { e = excl[c]; g2 = agg[c]; }
                v_base[lane] = e;
                if (tile == a.n_tiles - 1) ((volatile unsigned long long *)a.host->out_count)[lane] = e + g2;
            }
            if (tile == a.n_tiles - 1) {                   // tell the host that the counts of this sweep are final
                __threadfence_system();
                __syncwarp();
                if (lane == 0) *(volatile unsigned long long *)&a.host->seq = a.seq;
            }
            __threadfence_block();
            __syncwarp();
            if (lane == 0) *v_done = seen;
        }
        return;
    }

    // --------------------------------- compute warps ---------------------------------
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t my_seq = 0;

    auto flush_pending = [&]() {
        while (*v_done != my_seq) { }
        __threadfence_block();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint32_t n16 = ((volatile uint32_t *)sm.agg)[c] * RU;
            uint4 *dst = reinterpret_cast<uint4 *>(a.out[c]) + v_base[c] * RU;
            for (uint32_t i = threadIdx.x; i < n16; i += kCompThreads) dst[i] = sm.child[c][i];
        }
    };
    auto prefetch_records = [&](uint32_t tl) {
        const uint32_t g = tl * kCompThreads + threadIdx.x;
        if (tl < a.n_tiles && g < in.total) {
            const uint4 *rec = reinterpret_cast<const uint4 *>(seg_record(in, g, WORDS));
#pragma unroll
            for (int k = 0; k < RU; ++k) cp_async16(&sm.recbuf[threadIdx.x * RU + k], rec + k);
        }
    };
    if (threadIdx.x == 0) sm.tile = atomicAdd(a.ticket, 1u);
    bar_compute();
    uint32_t tile = sm.tile;
    prefetch_records(tile);
    while (tile < a.n_tiles) {
        uint32_t nxt = 0;
        if (threadIdx.x == 0) nxt = atomicAdd(a.ticket, 1u);
        if (threadIdx.x < C_NCOUNTERS) sm.stat[threadIdx.x] = 0;
        const uint32_t g = tile * kCompThreads + threadIdx.x;
        const bool active = g < in.total;
        uint64_t f1 = 0, s1 = 0, f2 = 0, s2 = 0, depth = 0;
        cp_async_wait_all();                               // this thread's own record has landed
        if (active) {
            const ulonglong2 *rec = reinterpret_cast<const ulonglong2 *>(&sm.recbuf[threadIdx.x * RU]);
            const ulonglong2 x = rec[0], y = rec[1];
            f1 = x.x; s1 = x.y; depth = y.x;
            if (TWO) { const ulonglong2 z = rec[2]; f2 = z.x; s2 = z.y; }
        }
        unsigned long long st_lcp = 0, st_da = 0;
        uint32_t st_rank = 0;
        if (active && a.write) {
            // update_LCP_leaf (:344-355) / update_DA (:394-425) at merged coordinates
            const uint64_t start1 = f1 + f2, start2 = f2 + s1, end = s1 + s2;
            if (end > start1) st_lcp = end - start1 - 1;
            const uint32_t pat = (depth >= a.K ? 0x55555555u : 0u) | (depth >= a.k_right ? 0xaaaaaaaau : 0u);
            if (end > start1 + 1) fill_bits(a.thr, 2 * (start1 + 1), 2 * end, pat);
            if (TWO) {
                st_da = end - start1;
                fill_bits(a.da, start2, end, 0xffffffffu);
            }
        }
        // next_leaves (dna_bwt.hpp:358-379; two BWTs: ebwt2InDel.cpp:452-472): LF(range) = 2 ranks per BWT
        uint64_t lo1[4] = {0, 0, 0, 0}, hi1[4] = {0, 0, 0, 0}, lo2[4] = {0, 0, 0, 0}, hi2[4] = {0, 0, 0, 0};
        if (active) {
            rank4(a.ix1, f1, lo1);
            st_rank++;
            if (s1 > f1) { rank4(a.ix1, s1, hi1); st_rank++; }
            else { hi1[0] = lo1[0]; hi1[1] = lo1[1]; hi1[2] = lo1[2]; hi1[3] = lo1[3]; }
            if (TWO) {
                rank4(a.ix2, f2, lo2);
                st_rank++;
                if (s2 > f2) { rank4(a.ix2, s2, hi2); st_rank++; }
                else { hi2[0] = lo2[0]; hi2[1] = lo2[1]; hi2[2] = lo2[2]; hi2[3] = lo2[3]; }
            }
        }
        uint32_t vm = 0, packed = 0, before[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const bool v = active && ((hi1[c] - lo1[c]) + (hi2[c] - lo2[c]) >= 2);
            const uint32_t bal = __ballot_sync(0xffffffffu, v);
            before[c] = __popc(bal & ((1u << lane) - 1u));
            packed |= (uint32_t)__popc(bal) << (8 * c);
            if (v) vm |= 1u << c;
        }
        if (lane == 0) sm.wpk[warp] = packed;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            st_lcp += __shfl_xor_sync(0xffffffffu, st_lcp, s);
            if (TWO) st_da += __shfl_xor_sync(0xffffffffu, st_da, s);
        }
        st_rank = __reduce_add_sync(0xffffffffu, st_rank);
        if (threadIdx.x == 0) sm.tile = nxt;
        bar_compute();                                     // counters reset, per-warp counts and the next ticket are visible
        const uint32_t next_tile = sm.tile;
        prefetch_records(next_tile);
        if (lane == 0) {
            if (st_lcp) atomicAdd(&sm.stat[C_LCP], st_lcp);
            if (st_rank) atomicAdd(&sm.stat[C_RANK], (unsigned long long)st_rank);
            if (TWO && st_da) atomicAdd(&sm.stat[C_DA], st_da);
        }
        if (my_seq) flush_pending();
        bar_compute();                                     // child buffer and agg free again; statistics complete
        uint32_t exw[4] = {0, 0, 0, 0}, tot[4] = {0, 0, 0, 0};
#pragma unroll
        for (int w = 0; w < kCompThreads / 32; ++w) {
            const uint32_t pk = sm.wpk[w];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint32_t n = (pk >> (8 * c)) & 0xffu;
                if (w < warp) exw[c] += n;
                tot[c] += n;
            }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if ((vm >> c) & 1u) {
                ulonglong2 *o = reinterpret_cast<ulonglong2 *>(&sm.child[c][(exw[c] + before[c]) * RU]);
                o[0] = make_ulonglong2(a.ix1.F[c] + lo1[c], a.ix1.F[c] + hi1[c]);
                o[1] = make_ulonglong2(depth + 1, 0);
                if (TWO) {
                    o[2] = make_ulonglong2(a.ix2.F[c] + lo2[c], a.ix2.F[c] + hi2[c]);
                    o[3] = make_ulonglong2(0, 0);
                }
            }
        }
        if (threadIdx.x == 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c) sm.agg[c] = tot[c];
            sm.pend_tile = tile;
        }
        if (threadIdx.x < C_NCOUNTERS && a.write) stripe_add(a.stripes, tile, threadIdx.x, sm.stat[threadIdx.x]);
        bar_compute();                                     // children, counts and tile id are in shared memory
        ++my_seq;
        if (threadIdx.x == 0) { __threadfence_block(); *v_posted = my_seq; }
        tile = next_tile;
    }
    if (my_seq) flush_pending();
    bar_compute();
    if (threadIdx.x == 0) *v_posted = kExitSeq;
}

// ---------------------------------------------------------------------------------------------
// Host-side frontier driver
// ---------------------------------------------------------------------------------------------
// host-side view of the compact node record (see unpack_node)
static void pack_node_host(uint64_t *rec, uint64_t base, const uint64_t F[4], uint64_t n, uint32_t depth) {
    const uint64_t s[5] = {F[0] - base, F[1] - F[0], F[2] - F[1], F[3] - F[2], n - F[3]};
    uint32_t w[8];
    for (int j = 0; j < 4; ++j) w[j] = (uint32_t)s[j];
    w[4] = (uint32_t)s[4];
    w[5] = (uint32_t)base;
    w[6] = (uint32_t)((s[0] >> 32) & 0xff) | (uint32_t)((s[1] >> 32) & 0xff) << 8 | (uint32_t)((s[2] >> 32) & 0xff) << 16 |
           (uint32_t)((s[3] >> 32) & 0xff) << 24;
    w[7] = (uint32_t)((s[4] >> 32) & 0xff) | (uint32_t)((base >> 32) & 0xff) << 8 | depth << 16;
    std::memcpy(rec, w, sizeof w);
}

static uint64_t node_size_host(const uint64_t *rec) {
    uint32_t w[8];
    std::memcpy(w, rec, sizeof w);
    uint64_t t = (uint64_t)w[0] + w[1] + w[2] + w[3] + w[4];
    t += ((uint64_t)(w[6] & 0xff) + ((w[6] >> 8) & 0xff) + ((w[6] >> 16) & 0xff) + (w[6] >> 24) + (w[7] & 0xff)) << 32;
    return t;
}

struct Frame {
    Arena *arena;
    int side;
    void *p;
    ~Frame() { if (p) arena->free(side, p); }
};

struct Chunk {
    uint64_t *p[4];
    uint64_t cnt[4];
    int level = 0;                      // tree depth of the records (selects the arena end of the next frame)
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

// Cut the first `take` records off a chunk (position-contiguous prefix).
static Chunk split_head(Chunk &c, uint64_t take, int words) {
    Chunk head = c;
    uint64_t left = take;
    for (int s = 0; s < 4; ++s) {
        const uint64_t k = std::min<uint64_t>(left, c.cnt[s]);
        head.cnt[s] = k;
        c.p[s] += k * words;
        c.cnt[s] -= k;
        left -= k;
    }
    return head;
}

template <typename Launch>
// max_chunk: largest chunk swept whole (two of its frames fit the arena: level-synchronous case).
// split_chunk: chunk size once a level has to be cut (depth-first case): a path of such chunks down
// to the deepest level must fit the arena next to the frame that is being cut.
static int run_frontier(e2i_ctx *ctx, Chunk root, int words, int tile_items, uint64_t max_chunk, uint64_t split_chunk,
                        NavArgs &args, Launch launch, SweepStats &ss, uint64_t stop_at_items, std::vector<Chunk> *stopped) {
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
        Chunk work;
        uint64_t take = cur.total() <= max_chunk ? cur.total() : std::min<uint64_t>(cur.total(), split_chunk);
        void *mem = nullptr;
        const double ta = now_ms();
        while (true) {   // shrink the chunk until its output frame fits the pool
            mem = ctx->arena.alloc((cur.level + 1) & 1, take * 4 * words * sizeof(uint64_t));
            if (mem) break;
            if (take <= 256) { set_error("frontier memory exhausted (arena %llu bytes, %llu in use): raise the frontier budget",
                                          (unsigned long long)ctx->arena.size(), (unsigned long long)ctx->arena.in_use()); return E2I_ERR_MEMORY; }
            take /= 2;
        }
        { const double d = now_ms() - ta; ss.ms_alloc += d; ss.ms_max_alloc = std::max(ss.ms_max_alloc, d); }
        if (take < cur.total()) {
            work = split_head(cur, take, words);
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
        const uint32_t n_tiles = (acc + tile_items - 1) / tile_items;
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
        for (int c = 0; c < 4; ++c) args.out[c] = reinterpret_cast<uint64_t *>(mem) + (size_t)c * take * words;
        args.desc = ctx->desc;
        args.epoch = ctx->epoch;
        args.n_tiles = n_tiles;
        if (ctx->ticket_next == kTicketSlots) {          // ring of ticket counters used up: zero it again
            E2I_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            E2I_CUDA_TRY(cudaMemsetAsync(ctx->ctl, 0, kTicketSlots * sizeof(uint32_t), ctx->stream));
            ctx->ticket_next = 0;
        }
        args.ticket = reinterpret_cast<uint32_t *>(ctx->ctl) + ctx->ticket_next++;
        args.host = hctl;
        args.seq = ++ctx->sweep_seq;
        launch(args, segs, n_tiles);
        E2I_CUDA_TRY(cudaGetLastError());
        ctx->n_launch++;
        ctx->n_d2h += sizeof(HostCtl);
        const double tsy = now_ms();
        {   // wait for the counts (not for the kernel): poll the mapped sequence word
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
        for (int c = 0; c < 4; ++c) { next.p[c] = args.out[c]; next.cnt[c] = ((volatile unsigned long long *)hctl->out_count)[c]; }
        work.frame.reset();
        if (next.total()) stack.push_back(std::move(next));
    }
    return E2I_OK;
}

}  // namespace e2i

using namespace e2i;

static uint64_t padded_words32(uint64_t bits) { return ((bits + 31) / 32 + 63) / 64 * 64 + 64; }

extern "C" int e2i_navigate_shard(e2i_ctx *ctx, const e2i_index *b1, const e2i_index *b2, const e2i_params *p,
                                  int shard, int n_shards, e2i_lcpbits **out, e2i_bits **da_out, e2i_stats *st) {
    if (!ctx || !b1 || !p || !out || !st) { set_error("e2i_navigate: null argument"); return E2I_ERR_ARG; }
    if (b2 && !da_out) { set_error("e2i_navigate: da_out is required with two BWTs"); return E2I_ERR_ARG; }
    if (n_shards < 1 || shard < 0 || shard >= n_shards) { set_error("e2i_navigate: bad shard %d/%d", shard, n_shards); return E2I_ERR_ARG; }
    if (p->K < 1 || p->k_right < 1 || p->K > 65535 || p->k_right > 65535) { set_error("e2i_navigate: K and k_right must be in [1, 65535]"); return E2I_ERR_ARG; }
    if ((b1->n >> 39) || (b2 && (b2->n >> 39))) { set_error("e2i_navigate: BWT longer than 2^39 symbols"); return E2I_ERR_ARG; }
    E2I_CUDA_TRY(cudaSetDevice(ctx->device));
    Accounting acct(ctx, st);
    cudaStream_t s = ctx->stream;
    const bool two = b2 != nullptr;
    // the node kernels stage 32 KB per CTA: ask for the large shared-memory carveout so that 4 CTAs fit an SM
    // E2I_NODE_KERNEL=tile selects the one-tile-per-CTA kernel (kept for A/B measurements)
    const char *nk = std::getenv("E2I_NODE_KERNEL");
    const bool persistent = !(nk && std::strcmp(nk, "tile") == 0);
    E2I_CUDA_TRY(cudaFuncSetAttribute(expand_nodes_persistent<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PersistSmem<false>)));
    E2I_CUDA_TRY(cudaFuncSetAttribute(expand_nodes_persistent<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PersistSmem<true>)));
    E2I_CUDA_TRY(cudaFuncSetAttribute(expand_leaves_persistent<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LeafSmem<false>)));
    E2I_CUDA_TRY(cudaFuncSetAttribute(expand_leaves_persistent<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LeafSmem<true>)));
    if (const char *cv = std::getenv("E2I_CARVEOUT")) {
        const int pct = atoi(cv);
        E2I_CUDA_TRY(cudaFuncSetAttribute(expand_nodes_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
        E2I_CUDA_TRY(cudaFuncSetAttribute(expand_nodes_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    }
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
    TRYF(cudaMemsetAsync(ctx->ctl, 0, kTicketSlots * sizeof(uint32_t), s));
    ctx->ticket_next = 0;

    // frontier budget: what is free now, minus head-room, unless the caller set one
    size_t free_b = 0, total_b = 0;
    TRYF(cudaMemGetInfo(&free_b, &total_b));
    {   // blocks cached by the stream-ordered pool are available to us as well
        cudaMemPool_t mp;
        uint64_t reserved = 0, used = 0;
        TRYF(cudaDeviceGetDefaultMemPool(&mp, ctx->device));
        TRYF(cudaMemPoolGetAttribute(mp, cudaMemPoolAttrReservedMemCurrent, &reserved));
        TRYF(cudaMemPoolGetAttribute(mp, cudaMemPoolAttrUsedMemCurrent, &used));
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
    args.K = (uint32_t)p->K;
    args.k_right = (uint32_t)p->k_right;

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
    // equal cumulated interval length, and every shard finishes its slice independently.
    const uint64_t kDealItems = 4096ull * (uint64_t)n_shards;

    auto run_pass = [&](bool leaves, SweepStats &ss) -> int {
        const int words = two ? 8 : 4;                  // u64 words per record (leaf: 32 B, compact node: 32 B)
        const int tile_items = kNavThreads;
        // two frames of the largest chunk (input's successor + its own output) plus slack must fit the arena
        const uint64_t max_chunk = std::max<uint64_t>(65536, (uint64_t)((double)budget / ((double)words * 8 * 4 * 2.5)));
        // depth of the traversal: the internal-node pass is never deeper than the leaf pass that ran before it
        const uint64_t depth_hint = leaves ? 1024 : st->levels_leaves + 16;
        const uint64_t split_chunk = std::max<uint64_t>(256, std::min<uint64_t>(max_chunk, (uint64_t)((double)budget * 0.45 / ((double)words * 8 * 4 * (double)depth_hint))));
        // root record
        void *rootmem = nullptr;
        rootmem = ctx->arena.alloc(0, (size_t)words * 8);
        if (!rootmem) { set_error("frontier arena too small"); return E2I_ERR_MEMORY; }
        uint64_t rec[16] = {0};
        if (leaves) {                                   // first_leaf (dna_bwt.hpp:313-317)
            rec[0] = 0; rec[1] = b1->F[0]; rec[2] = 0;
            if (two) { rec[4] = 0; rec[5] = b2->F[0]; }
        } else {                                        // root (dna_bwt.hpp:296-308) as a compact record
            pack_node_host(rec, 0, b1->F, b1->n, 0);
            if (two) pack_node_host(rec + 4, 0, b2->F, b2->n, 0);
        }
        E2I_CUDA_TRY(cudaMemcpyAsync(rootmem, rec, (size_t)words * 8, cudaMemcpyHostToDevice, s));
        Chunk root{};
        root.p[0] = reinterpret_cast<uint64_t *>(rootmem);
        root.cnt[0] = 1;
        root.frame = std::make_shared<Frame>();
        root.frame->arena = &ctx->arena;
        root.frame->side = 0;
        root.frame->p = rootmem;
        auto launch = [&](NavArgs &a, const Segs &segs, uint32_t n_tiles) {
            if (leaves && persistent) {
                const uint32_t grid = std::min<uint32_t>(n_tiles, (uint32_t)ctx->sm_count * (two ? 2u : 3u));
                if (two) expand_leaves_persistent<true><<<grid, kPersistThreads, sizeof(LeafSmem<true>), s>>>(a, segs);
                else expand_leaves_persistent<false><<<grid, kPersistThreads, sizeof(LeafSmem<false>), s>>>(a, segs);
            } else if (leaves) {
                if (two) expand_leaves_kernel<true><<<n_tiles, kNavThreads, 0, s>>>(a, segs);
                else expand_leaves_kernel<false><<<n_tiles, kNavThreads, 0, s>>>(a, segs);
            } else if (persistent) {
                const uint32_t grid = std::min<uint32_t>(n_tiles, (uint32_t)ctx->sm_count * (two ? 2u : 3u));
                if (two) expand_nodes_persistent<true><<<grid, kPersistThreads, sizeof(PersistSmem<true>), s>>>(a, segs);
                else expand_nodes_persistent<false><<<grid, kPersistThreads, sizeof(PersistSmem<false>), s>>>(a, segs);
            } else {
                if (two) expand_nodes_kernel<true><<<n_tiles, kNavThreads, 0, s>>>(a, segs);
                else expand_nodes_kernel<false><<<n_tiles, kNavThreads, 0, s>>>(a, segs);
            }
        };
        if (n_shards == 1) {
            args.write = 1;
            return run_frontier(ctx, std::move(root), words, tile_items, max_chunk, split_chunk, args, launch, ss, 0, nullptr);
        }
        // shared top of the tree
        std::vector<Chunk> dealt;
        args.write = shard == 0;
        SweepStats top;
        E2I_TRY(run_frontier(ctx, std::move(root), words, tile_items, max_chunk, split_chunk, args, launch, top, kDealItems, &dealt));
        if (shard == 0) { ss.items += top.items; ss.sweeps += top.sweeps; ss.max_chunk = std::max(ss.max_chunk, top.max_chunk); }
        args.write = 1;
        for (Chunk &c : dealt) {
            // deal by cumulated interval length: fetch (first, last) of every record
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
                if (leaves) return (r[1] - r[0]) + (two ? r[5] - r[4] : 0) + 1;
                return node_size_host(r) + (two ? node_size_host(r + 4) : 0) + 1;
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
            (void)split_head(mine, lo, words);          // drop [0, lo)
            Chunk part = split_head(mine, hi - lo, words);
            part.frame = c.frame;
            E2I_TRY(run_frontier(ctx, std::move(part), words, tile_items, max_chunk, split_chunk, args, launch, ss, 0, nullptr));
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
