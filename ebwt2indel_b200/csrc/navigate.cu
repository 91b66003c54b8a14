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
// POSITION: the children cW of a sorted frontier go to four queues (one per c), each in input
// order; A-queue ++ C-queue ++ G-queue ++ T-queue is again sorted because LF is monotone per
// symbol.  A sorted frontier turns the reference's random rank gathers into one near-sequential
// pass over the 64-byte index blocks per sweep (neighbouring nodes share blocks and DRAM pages)
// and makes the bit updates land in neighbouring words.
//
// No cross-CTA (or cross-warp) dependency inside a sweep.  The input of a sweep is cut into RUNS
// of consecutive records; a warp takes a run by ticket and writes the children of queue c to a
// region of the output frame that belongs to (c, run) alone -- a node has at most one child per
// symbol, so the region is as large as the run.  The frame is therefore "gappy": a tiny kernel
// (frame_index_kernel) scans the per-(queue, run) counts after the sweep and the next sweep
// addresses records by global index through that prefix array.  This replaces an ordered append by
// decoupled look-back, whose resolution latency (hundreds of tiles in flight) was the critical path
// of round 1's kernel; a warp now synchronises with nobody but itself.  When a sweep would not fit
// the frontier budget it is cut into index ranges that are finished depth-first (bounded memory).
//
// Work mapping.  One thread per internal node (per node pair in mode -2): it walks the node's <= 6
// distinct boundaries, turns the rank differences into the five sub-interval sizes of each child
// cW, keeps the children with >= 2 non-empty sub-intervals (number_of_children >= 2) and writes
// them in lane order.  Leaves: one thread per leaf (two ranks per BWT).  All nodes of a sweep have the
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

namespace e2i {

constexpr int kNavWarps = 4;                          // warps per CTA; every warp works on its own
constexpr int kNavThreads = kNavWarps * 32;
constexpr int kWarpStage = 128;                       // index blocks staged in shared memory per warp (4 KB)
#ifndef E2I_WINDOW_MAX
#define E2I_WINDOW_MAX 96
#endif
constexpr int kWindowMax = E2I_WINDOW_MAX;             // longest block range staged as one window (one BWT)
#ifndef E2I_WINDOW_TMA
#define E2I_WINDOW_TMA 0                              // 1: windows are staged by cp.async.bulk + mbarrier (measured slower:
                                                      // C4 nodes 468 vs 382 ms, profiles/README.md), 0: by LDGSTS
#endif
constexpr bool kWindowTma = E2I_WINDOW_TMA != 0;
constexpr int kMaxRun = 1024;                         // records per run (multiple of 32), upper bound
#ifndef E2I_NODE_CTAS
#define E2I_NODE_CTAS 7                               // resident CTAs per SM the one-BWT kernels are compiled for
#endif
constexpr int kNodeCtas = E2I_NODE_CTAS, kPairCtas = 4;
constexpr int kStripes = 128;                         // striped statistics counters (avoid single-address atomics)
enum { C_LCP = 0, C_NMIN, C_RANK, C_BITUPD, C_DA, C_NCOUNTERS = 8 };
#ifndef E2I_SMALL_LIMIT
#define E2I_SMALL_LIMIT 65536                         // node sizes below this use the 16-byte record (test builds lower it)
#endif
constexpr uint64_t kSmallLimit = E2I_SMALL_LIMIT;
static_assert(kSmallLimit <= 65536, "SMALL records hold 16-bit sizes");

// Per-sweep control.  Device side: one zeroed 64-byte block per sweep (a ring, so no per-sweep
// memset): run ticket and the largest input node.  Host side: a page-locked, device-mapped block
// that frame_index_kernel fills in after the sweep, followed by the sweep's sequence number; the
// host polls that word instead of a copy + stream synchronisation.
constexpr uint32_t kSweepSlots = 16384;
constexpr int kMaxDest = 16;                          // destinations of a frame = GPUs of one box
struct SweepDev {
    uint32_t ticket, done;    // run ticket; CTAs of frame_index_kernel that have finished
    unsigned long long maxsz;
    unsigned long long pad[6];
};
static_assert(sizeof(SweepDev) == 64, "one sweep slot per 64 bytes");
struct HostCtl {
    unsigned long long total;                         // records of the frame
    unsigned long long maxsz;                         // largest node the sweep read
    unsigned long long seq;
    uint32_t qstart[4 * kMaxDest + 1];                // P[(4h+c)K]: where destination h, queue c starts
};

// A frame: the records one sweep produced.  Queue c, run k owns the slots
// base + ((c * K + k) * run_cap ...) and fills them in input order.  With D destinations (position
// ranges of D GPUs, D = 1 on one GPU) the children of a region are counted per destination h -- they are
// sorted by position, so each destination's share is a contiguous piece -- and entry e = (h * 4 + c) * K + k
// holds cnt[e] records starting off[e] slots into the region.  P is the exclusive prefix of cnt in entry
// order (4 D K + 1 values), so global index g lives in the entry e with P[e] <= g < P[e+1], and the records
// of destination h, queue c are the global indices [P[(4h+c)K], P[(4h+c+1)K)).  hint[t] = the entry that
// holds global index 256 t.
struct FrameSrc {
    const uint4 *base;
    const uint32_t *P;
    const uint32_t *off;      // nullptr when the frame has one destination
    const uint32_t *hint;
    uint32_t K, run_cap;
};
constexpr int kMaxSeg = 4 * kMaxDest;
struct Segment {                                      // a contiguous piece of records (own or a peer's memory) of the sweep's input
    const uint4 *ptr;         // its first record
    uint32_t start, len;      // position in the concatenated input, records
};
template <bool MULTI> struct FrameInT;
template <> struct FrameInT<false> {                  // one local gappy frame, read by global index
    FrameSrc s;
    uint32_t g_lo, g_hi;      // this sweep reads the records [g_lo, g_hi)
};
template <> struct FrameInT<true> {                   // compacted pieces of several ranks' frames, concatenated in position order
    uint32_t g_lo, g_hi;
    uint32_t n_seg, pad;
    Segment seg[kMaxSeg];
};
using FrameIn = FrameInT<false>;
struct FrameOut {
    uint4 *base;
    uint32_t *cnt;
    uint32_t *gsum;           // sums of cnt over groups of 256 consecutive entries (group_sum_kernel)
    uint32_t K, run_cap;      // runs of this sweep's input, records per run (multiple of 32)
    uint32_t D, pad;          // destinations
    uint64_t range_len;       // destination h owns the positions [h * range_len, (h + 1) * range_len)
};

struct NavArgs {
    DevIndex ix1, ix2;
    uint32_t *thr;            // 2 bits per merged position
    uint32_t *minima;         // 1 bit per merged position
    uint32_t *da;             // 1 bit per merged position (mode -2)
    unsigned long long *stripes;
    SweepDev *sweep;          // this sweep's device control block
    uint32_t bits;            // (depth >= K) | (depth >= k_right) << 1 for the records of this sweep
    int write;                // 0: expand only (redundant top of the tree on shards != 0)
};

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

// ---- reading gappy frames by global index ----------------------------------------------------------
struct Cursor {               // per lane: the entry that holds the lane's current record
    uint32_t j, c, k, pj, pj1;
    uint32_t s, delta, seg_end, offj;                 // MULTI: segment, source index - position, end of the segment
};

// entry of a frame that holds global index x: the hints bracket it, a binary search on P finds it -- the
// entries in between may be thousands of empty ones (a queue without children for a destination)
__device__ __forceinline__ uint32_t locate_entry(const uint32_t *__restrict__ P, const uint32_t *__restrict__ hint, uint32_t x) {
    uint32_t lo = __ldg(hint + (x >> 8)), hi = __ldg(hint + (x >> 8) + 1);   // last j in [lo, hi] with P[j] <= x
    while (lo < hi) {
        const uint32_t mid = (lo + hi + 1) >> 1;
        if (__ldg(P + mid) <= x) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__device__ __forceinline__ void cursor_open(const FrameInT<false> &in, Cursor &cur, uint32_t g) {
    cur.j = locate_entry(in.s.P, in.s.hint, g);
    cur.c = cur.j / in.s.K;
    cur.k = cur.j - cur.c * in.s.K;
    cur.pj = __ldg(in.s.P + cur.j);
    cur.pj1 = __ldg(in.s.P + cur.j + 1);
}

// move to the entry that holds g (g never decreases, g < P[last])
__device__ __forceinline__ void cursor_seek(const FrameInT<false> &in, Cursor &cur, uint32_t g) {
    if (g >= cur.pj1) cursor_open(in, cur, g);
}

__device__ __forceinline__ const uint4 *cursor_record(const FrameInT<false> &in, const Cursor &cur, uint32_t g, int ru) {
    return in.s.base + (((size_t)cur.c * in.s.K + cur.k) * in.s.run_cap + (g - cur.pj)) * ru;
}

// MULTI: position g of the concatenated input -> segment (the pieces are plain arrays: frames are compacted
// per destination before the peers read them)
__device__ __forceinline__ void cursor_enter(const FrameInT<true> &in, Cursor &cur, uint32_t g) {
    uint32_t s = cur.s;
    while (g >= in.seg[s].start + in.seg[s].len) ++s;
    cur.s = s;
    cur.delta = in.seg[s].start;
    cur.seg_end = in.seg[s].start + in.seg[s].len;
}
__device__ __forceinline__ void cursor_open(const FrameInT<true> &in, Cursor &cur, uint32_t g) {
    cur.s = 0;
    cursor_enter(in, cur, g);
}
__device__ __forceinline__ void cursor_seek(const FrameInT<true> &in, Cursor &cur, uint32_t g) {
    if (g >= cur.seg_end) cursor_enter(in, cur, g);
}
__device__ __forceinline__ const uint4 *cursor_record(const FrameInT<true> &in, const Cursor &cur, uint32_t g, int ru) {
    return in.seg[cur.s].ptr + (size_t)(g - cur.delta) * ru;
}

// ---- MULTI: children of a run counted per destination (position range) ---------------------------------
// The children of queue c leave a run sorted by position, so the destination never decreases: dcur / dbound
// (warp-uniform) hold the current destination and the first position beyond it; only a step that crosses a
// range boundary takes the slow path.
__device__ __forceinline__ void count_dests(uint32_t *dcnt, uint32_t &dcur, uint64_t &dbound, const FrameOut &out, int lane,
                                            bool valid, uint64_t pos, uint32_t tot) {
    const uint32_t over = __ballot_sync(0xffffffffu, valid && pos >= dbound);
    if (!over) {
        if (lane == 0 && tot) dcnt[dcur] += tot;
        return;
    }
    uint32_t rest = __ballot_sync(0xffffffffu, valid);
    while (rest) {
        const uint32_t ov = __ballot_sync(0xffffffffu, ((rest >> lane) & 1u) && pos >= dbound);
        const uint32_t below = ov ? (rest & ((1u << (__ffs(ov) - 1)) - 1u)) : rest;
        if (lane == 0 && below) dcnt[dcur] += __popc(below);
        rest &= ~below;
        if (!rest) break;
        const uint64_t p0 = __shfl_sync(0xffffffffu, (unsigned long long)pos, __ffs(ov) - 1);
        dcur = (uint32_t)min((uint64_t)(out.D - 1), p0 / out.range_len);
        dbound = dcur == out.D - 1 ? ~0ull : (uint64_t)(dcur + 1) * out.range_len;
    }
}

// end-of-kernel flush of the per-thread statistics (one atomic per warp and counter)
__device__ __forceinline__ void flush_stat(const NavArgs &a, int which, unsigned long long v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    if ((threadIdx.x & 31) == 0 && v)
        atomicAdd(a.stripes + (size_t)((blockIdx.x * kNavWarps + (threadIdx.x >> 5)) & (kStripes - 1)) * C_NCOUNTERS + which, v);
}

// ---- staged index blocks ---------------------------------------------------------------------------
// Where the rank queries of a warp's 32 nodes read their index blocks.  WINDOW: their whole block range
// [origin, origin + n) sits in shared memory (dense tiles: nodes of one depth are disjoint and sorted).
// SLOTS: sparse tiles (a traversal shard of a multi-GPU run, the leaf frontier): every thread gets
// the first and the last block of its own interval, fetched by the whole warp with 16-byte asynchronous
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

__device__ __forceinline__ void load_block_smem(const uint4 *stage, uint32_t slot, uint4 &lo, uint4 &hi) {
    const uint32_t sw = (slot >> 2) & 1u;
    const uint4 *p = stage + slot * 2;
    lo = p[sw]; hi = p[1u ^ sw];
}
// a window staged by the bulk-copy engine is a plain image of the block range
__device__ __forceinline__ void load_block_linear(const uint4 *stage, uint32_t slot, uint4 &lo, uint4 &hi) {
    const uint4 *p = stage + slot * 2;
    lo = p[0]; hi = p[1];
}

// #A,#C,#G,#T before relative position rpos, counted from the start of the superblock of origin_blk
// (mod 2^32: the base cancels in every difference, which is all a node shorter than 2^32 needs).
// multi_super (warp-uniform): the 32 nodes reach into a second superblock.
template <int MODE>
__device__ __forceinline__ void rank_rel(const DevIndex &ix, const RankSrc &r, bool multi_super, uint32_t rpos, uint32_t out[4]) {
    const uint32_t rel = rpos >> kBlockShift;
    uint4 lo, hi;
    if (MODE == SRC_WINDOW) {
        if (kWindowTma) load_block_linear(r.stage, r.slot0 + rel, lo, hi); else load_block_smem(r.stage, r.slot0 + rel, lo, hi);
    } else if (MODE == SRC_SLOTS && (rel == 0 || rel == r.d1)) {
        load_block_smem(r.stage, r.slot0 + (rel != 0), lo, hi);
    } else {
        const uint4 *p = ix.blocks + (size_t)(r.origin_blk + rel) * kBlockU4;
        lo = __ldg(p); hi = __ldg(p + 1);
    }
    block_rank(lo, hi, (int)(rpos & (kBlockSyms - 1)), out);
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

// Turn the ranks at one boundary into the sizes of sub-interval j of every child.  Child c is
// right-maximal iff >= 2 of its 5 sub-intervals are non-empty in the union of both BWTs
// (number_of_children, include.hpp:760-792), i.e. iff the sum of the (summed) sizes exceeds their
// maximum: Gaps keeps both per symbol.
template <typename W>
struct Gaps {
    W sum[4] = {0, 0, 0, 0}, mx[4] = {0, 0, 0, 0};
    __device__ __forceinline__ bool valid(int c) const { return sum[c] > mx[c]; }
};

template <bool TWO, typename W>
__device__ __forceinline__ void take_boundary(int j, const W (&cur1)[4], W (&prev1)[4], const W (&cur2)[4], W (&prev2)[4],
                                              ChildSide<W> &k1, ChildSide<W> &k2, Gaps<W> &gp) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const W d1 = cur1[c] - prev1[c];
        W e = d1;
        k1.sz[j][c] = d1;
        prev1[c] = cur1[c];
        if (TWO) {
            const W d2 = cur2[c] - prev2[c];
            e += d2;
            k2.sz[j][c] = d2;
            prev2[c] = cur2[c];
        }
        gp.sum[c] += e;
        gp.mx[c] = max(gp.mx[c], e);
    }
}

// LF(sa_node) (dna_bwt.hpp:323-356) for one SMALL node (pair): ranks at the distinct boundaries (equal
// neighbours reuse the previous result, :334-347) in 32-bit arithmetic relative to the staged blocks.
template <bool TWO, int MODE>
__device__ __forceinline__ void expand_small(const NavArgs &a, bool multi_super, const RankSrc &r1, const RankSrc &r2,
                                             uint64_t base1, uint32_t rpos1, const uint32_t (&s1)[5],
                                             uint64_t base2, uint32_t rpos2, const uint32_t (&s2)[5],
                                             ChildSide<uint32_t> &k1, ChildSide<uint32_t> &k2, Gaps<uint32_t> &gp, uint32_t &st_rank) {
    uint32_t prev1[4], cur1[4], prev2[4] = {0, 0, 0, 0}, cur2[4] = {0, 0, 0, 0};
    rank_rel<MODE>(a.ix1, r1, multi_super, rpos1, prev1);
    st_rank++;
    if (TWO) { rank_rel<MODE>(a.ix2, r2, multi_super, rpos2, prev2); st_rank++; }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        // prev is relative to the superblock of the node's first block, which is the superblock of `base`
        k1.base[c] = a.ix1.F[c] + prev1[c] + __ldg(a.ix1.super + (base1 >> kSuperShift) * 4 + c);
        if (TWO) k2.base[c] = a.ix2.F[c] + prev2[c] + __ldg(a.ix2.super + (base2 >> kSuperShift) * 4 + c);
    }
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        rpos1 += s1[j];
        if (s1[j]) { rank_rel<MODE>(a.ix1, r1, multi_super, rpos1, cur1); st_rank++; }
        else { cur1[0] = prev1[0]; cur1[1] = prev1[1]; cur1[2] = prev1[2]; cur1[3] = prev1[3]; }
        if (TWO) {
            rpos2 += s2[j];
            if (s2[j]) { rank_rel<MODE>(a.ix2, r2, multi_super, rpos2, cur2); st_rank++; }
            else { cur2[0] = prev2[0]; cur2[1] = prev2[1]; cur2[2] = prev2[2]; cur2[3] = prev2[3]; }
        }
        take_boundary<TWO, uint32_t>(j, cur1, prev1, cur2, prev2, k1, k2, gp);
    }
}

// the same for a WIDE node (pair): absolute 64-bit ranks read from HBM
template <bool TWO>
__device__ __forceinline__ void expand_wide(const NavArgs &a, uint64_t base1, const uint64_t (&s1)[5], uint64_t base2, const uint64_t (&s2)[5],
                                            ChildSide<uint64_t> &k1, ChildSide<uint64_t> &k2, Gaps<uint64_t> &gp, uint32_t &st_rank) {
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
        take_boundary<TWO, uint64_t>(j, cur1, prev1, cur2, prev2, k1, k2, gp);
    }
}

// SLOTS staging: every lane has announced up to two block ids in need[2 * lane], need[2 * lane + 1]
// (~0u = none); the warp copies them with 2 consecutive lanes per 32-byte block
__device__ __forceinline__ void stage_slots(const uint4 *blocks, uint4 *stage, const uint32_t *need, int lane) {
#pragma unroll
    for (int it = 0; it < 64 * kBlockU4 / 32; ++it) {
        const uint32_t k = lane + it * 32;
        const uint32_t slot = k >> 1, blk = need[slot];
        if (blk != ~0u) cp_async16(&stage[stage_slot(slot, k & 1)], blocks + (size_t)blk * kBlockU4 + (k & 1));
    }
}

__device__ __forceinline__ void stage_window(const DevIndex &ix, uint4 *stage, uint32_t lo_blk, uint32_t n_blk, int lane) {
    const uint4 *src = ix.blocks + (size_t)lo_blk * kBlockU4;
    for (uint32_t k = lane; k < n_blk * kBlockU4; k += 32) cp_async16(&stage[stage_slot(k >> 1, k & 1)], src + k);
}

// ---- the same window staged by the TMA / bulk-copy engine: the block range is one contiguous piece of HBM, so
//      ONE elected lane issues ONE cp.async.bulk (UBLKCP) per BWT and the warp waits on its own mbarrier; replaces
//      up to 6 LDGSTS per lane with their address arithmetic ------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void *smem, const void *gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((uint32_t)__cvta_generic_to_shared(smem)),
                 "l"(gmem), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
// generic-proxy reads of the staging buffer (previous step) are ordered before the async-proxy writes that follow
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// MULTI (position-range sharding): the records of a step are a stretch of at most 32 consecutive entries of the
// concatenated input -- plain arrays in the frames of the ranks (most of them behind NVLink).  ONE lane fetches the
// step with bulk copies (one per segment the stretch touches, almost always one: 512 bytes for 32 SMALL records)
// that complete on an mbarrier of their own, so the records of the next kRecRing steps are in flight while the
// warp works -- independent of the ordered cp.async groups of the index staging, which capped the LDGSTS prefetch
// at one step (a round trip to a peer is several steps long).
constexpr int kRecRing = 4;
__device__ __forceinline__ void fetch_records_bulk(const FrameInT<true> &in, uint32_t &seg, uint32_t g0, uint32_t g_end, int ru,
                                                   uint4 *slot, uint64_t *bar) {
    uint32_t cnt = min(32u, g_end - g0), g = g0;
    fence_proxy_async();                                  // the warp's reads of this slot (generic proxy) are over
    mbar_expect_tx(bar, cnt * (uint32_t)ru * 16u);
    while (cnt) {
        while (g >= in.seg[seg].start + in.seg[seg].len) ++seg;
        const uint32_t take = min(cnt, in.seg[seg].start + in.seg[seg].len - g);
        bulk_copy_g2s(slot + (size_t)(g - g0) * ru, in.seg[seg].ptr + (size_t)(g - in.seg[seg].start) * ru, take * (uint32_t)ru * 16u, bar);
        g += take;
        cnt -= take;
    }
}

// position of this lane's children inside the step (before[c]) and the step's totals (tot[c])
__device__ __forceinline__ uint32_t warp_child_slots(const bool (&valid)[4], int lane, uint32_t (&before)[4], uint32_t (&tot)[4]) {
    uint32_t vm = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const uint32_t bal = __ballot_sync(0xffffffffu, valid[c]);
        before[c] = __popc(bal & ((1u << lane) - 1u));
        tot[c] = __popc(bal);
        if (valid[c]) vm |= 1u << c;
    }
    return vm;
}

// ---------------------------------------------------------------------------------------------
// Phase 3 sweep: internal nodes.  Persistent grid; every WARP loops over runs taken by ticket and,
// inside a run, over steps of 32 records:
//   records (prefetched by LDGSTS during the previous step) -> the warp's index blocks staged in
//   its 4 KB of shared memory by asynchronous copies, overlapped with the bit updates of the step
//   -> ranks -> children written straight to the run's own region of the output frame.
// ---------------------------------------------------------------------------------------------
template <bool TWO, bool IN_S, bool MULTI>
struct NodeSmem {
    static constexpr int RIN = (IN_S ? 1 : 3) * (TWO ? 2 : 1);    // uint4 per input record
    static constexpr int RING = MULTI && IN_S ? kRecRing : 2;     // steps of records in flight (WIDE records: 2, they are few)
    uint4 stage[kNavWarps][IN_S ? kWarpStage * kBlockU4 : 1];     // staged index blocks, per warp
    uint4 recbuf[kNavWarps][RING * 32 * RIN];                     // records of the next steps (ring, slot = lane)
    uint64_t rbar[kNavWarps][kRecRing];                           // MULTI: completion barriers of the record ring
    uint32_t need[kNavWarps][IN_S ? 64 : 1];                      // SLOTS staging: block ids wanted by the lanes
    uint64_t mbar[kNavWarps];                                     // completion barrier of the warp's bulk copies
    uint32_t dcnt[kNavWarps][4 * kMaxDest];                       // MULTI: children of the current run per (queue, destination)
};

template <bool TWO, bool IN_S, bool OUT_S, bool MULTI>
__global__ void __launch_bounds__(kNavThreads, IN_S ? (TWO ? kPairCtas : kNodeCtas) : 1)
expand_nodes_kernel(const NavArgs a, const __grid_constant__ FrameInT<MULTI> in, const FrameOut out) {
    using SM = NodeSmem<TWO, IN_S, MULTI>;
    using W = typename std::conditional<IN_S, uint32_t, uint64_t>::type;
    constexpr int RING = SM::RING;
    constexpr int RIN = SM::RIN, RSIDE_IN = IN_S ? 1 : 3, RSIDE_OUT = OUT_S ? 1 : 3, ROUT = RSIDE_OUT * (TWO ? 2 : 1);
    constexpr int STAGE = TWO ? kWarpStage / 2 : kWarpStage;               // blocks staged per BWT
    __shared__ __align__(1024) SM sm;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4 *stage1 = sm.stage[warp], *stage2 = sm.stage[warp] + (IN_S ? STAGE * kBlockU4 : 0);
    uint4 *recbuf = sm.recbuf[warp] + lane * RIN;
    uint32_t *need = sm.need[warp];

    uint64_t *mbar = &sm.mbar[warp];
    uint32_t mbar_phase = 0;
    if (IN_S && kWindowTma) {
        if (lane == 0) mbar_init(mbar, 1);
        __syncwarp();
    }
    uint32_t it = 0;                                      // MULTI: steps this warp has taken (ring slot it % RING, parity of its barrier)
    if (MULTI) {
        if (lane == 0) for (int d = 0; d < RING; ++d) mbar_init(&sm.rbar[warp][d], 1);
        __syncwarp();
    }

    uint32_t st_lcp = 0, st_min = 0, st_rank = 0, st_upd = 0, st_da = 0;
    uint64_t max_size = 0;
    // the first run of a warp is its own index (no atomic: small sweeps never touch the ticket), the
    // following ones are taken by ticket
    uint32_t next_run = blockIdx.x * kNavWarps + warp;
    while (true) {
        const uint32_t run = __shfl_sync(0xffffffffu, next_run, 0);
        if (run >= out.K) break;
        if (lane == 0) next_run = gridDim.x * kNavWarps + atomicAdd(&a.sweep->ticket, 1u);   // its latency hides behind this run
        const uint32_t g_begin = in.g_lo + run * out.run_cap, g_end = min(in.g_hi, g_begin + out.run_cap);
        Cursor cur;
        uint32_t seg = 0;                                 // MULTI: segment cursor of the lane that fetches
        if constexpr (MULTI) {
            __syncwarp();                                 // every lane is done with the previous run's records
            if (lane == 0) {
#pragma unroll
                for (int d = 0; d < RING; ++d)
                    if (g_begin + 32u * d < g_end)
                        fetch_records_bulk(in, seg, g_begin + 32u * d, g_end, RIN, sm.recbuf[warp] + ((it + d) % RING) * (32 * RIN), &sm.rbar[warp][(it + d) % RING]);
            }
        } else {
            cursor_open(in, cur, g_begin);
            // the record of the first step
            if (g_begin + lane < g_end) {
                cursor_seek(in, cur, g_begin + lane);
                const uint4 *rec = cursor_record(in, cur, g_begin + lane, RIN);
#pragma unroll
                for (int k = 0; k < RIN; ++k) cp_async16(recbuf + k, rec + k);
            }
            cp_async_commit();
            if (g_begin + 32 + lane < g_end) {            // ... and of the second
                cursor_seek(in, cur, g_begin + 32 + lane);
                const uint4 *rec = cursor_record(in, cur, g_begin + 32 + lane, RIN);
#pragma unroll
                for (int k = 0; k < RIN; ++k) cp_async16(recbuf + 32 * RIN + k, rec + k);
            }
            cp_async_commit();
        }
        uint32_t ring = 0;                                // ring slot of the current step
        uint32_t run_cnt[4] = {0, 0, 0, 0};
        uint4 *region[4];                                 // next free slot of the run's region, per queue
#pragma unroll
        for (int c = 0; c < 4; ++c) region[c] = out.base + (((size_t)c * out.K + run) * out.run_cap) * ROUT;
        uint32_t dcur[4] = {0, 0, 0, 0};                  // MULTI: current destination of every queue
        uint64_t dbound[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) dbound[c] = out.D > 1 ? out.range_len : ~0ull;
        if (MULTI) {
            for (int i = lane; i < 4 * kMaxDest; i += 32) sm.dcnt[warp][i] = 0;
            __syncwarp();
        }
        for (uint32_t g0 = g_begin; g0 < g_end; g0 += 32) {
            const uint32_t g = g0 + lane;
            const bool active = g < g_end;
            uint64_t base1 = 0, base2 = 0;
            W s1[5] = {0, 0, 0, 0, 0}, s2[5] = {0, 0, 0, 0, 0};
            // copy groups complete in order: all but the newest (the record of the step after this one) are
            // waited for, so a record has a whole step to arrive -- what remote frames (NVLink) need
            uint4 *rb;
            if constexpr (MULTI) {
                mbar_wait(&sm.rbar[warp][it % RING], (it / RING) & 1u);     // the step's records have landed
                rb = recbuf + (it % RING) * (32 * RIN);
            } else {
                cp_async_wait_group<1>();                                   // this lane's own record has landed
                rb = recbuf + ring * (32 * RIN);
            }
            if (active) {
                load_node<IN_S, W>(rb, base1, s1);
                if (TWO) load_node<IN_S, W>(rb + RSIDE_IN, base2, s2);
            }
            if constexpr (MULTI) {                                          // the slot just read takes the step RING ahead
                __syncwarp();
                if (lane == 0 && g0 + 32u * RING < g_end)
                    fetch_records_bulk(in, seg, g0 + 32u * RING, g_end, RIN, sm.recbuf[warp] + (it % RING) * (32 * RIN), &sm.rbar[warp][it % RING]);
                ++it;
            }
            const uint64_t size1 = (uint64_t)s1[0] + s1[1] + s1[2] + s1[3] + s1[4], size2 = (uint64_t)s2[0] + s2[1] + s2[2] + s2[3] + s2[4];
            max_size = max(max_size, max(size1, size2));

            // ---- stage the index blocks of the step in the warp's shared memory (SMALL records only) ----
            int mode = SRC_GLOBAL;
            bool multi_super = false;
            const uint32_t fb1 = (uint32_t)(base1 >> kBlockShift), fb2 = (uint32_t)(base2 >> kBlockShift);
            RankSrc r1{stage1, fb1, 0u, 0u}, r2{stage2, fb2, 0u, 0u};
            if (IN_S) {
                const int last_active = (int)min(32u, g_end - g0) - 1;
                const uint32_t lb1 = (uint32_t)((base1 + size1) >> kBlockShift), lb2 = (uint32_t)((base2 + size2) >> kBlockShift);
                // nodes of one depth are disjoint and sorted: the step touches the block range [lo, hi]
                const uint32_t lo1 = __shfl_sync(0xffffffffu, fb1, 0), hi1 = __shfl_sync(0xffffffffu, lb1, last_active);
                const uint32_t lo2 = TWO ? __shfl_sync(0xffffffffu, fb2, 0) : 0u, hi2 = TWO ? __shfl_sync(0xffffffffu, lb2, last_active) : 0u;
                const uint32_t span1 = hi1 - lo1 + 1, span2 = TWO ? hi2 - lo2 + 1 : 0u;
                constexpr uint32_t sbs = kSuperShift - kBlockShift;
                multi_super = (lo1 >> sbs) != (hi1 >> sbs) || (TWO && (lo2 >> sbs) != (hi2 >> sbs));
                __syncwarp();                                               // the previous step's reads of the staging buffer are over
                // WINDOW copies the whole block range, SLOTS at most two blocks per node: the window pays off while
                // the range is not longer than what SLOTS would fetch (pairs: up to the staging capacity)
                if (span1 <= (uint32_t)(TWO ? STAGE : kWindowMax) && span2 <= (uint32_t)STAGE) {
                    mode = SRC_WINDOW;
                    if (kWindowTma) {
                        if (lane == 0) {
                            fence_proxy_async();
                            mbar_expect_tx(mbar, (span1 + span2) * (uint32_t)(kBlockU4 * sizeof(uint4)));
                            bulk_copy_g2s(stage1, a.ix1.blocks + (size_t)lo1 * kBlockU4, span1 * (uint32_t)(kBlockU4 * sizeof(uint4)), mbar);
                            if (TWO) bulk_copy_g2s(stage2, a.ix2.blocks + (size_t)lo2 * kBlockU4, span2 * (uint32_t)(kBlockU4 * sizeof(uint4)), mbar);
                        }
                    } else {
                        stage_window(a.ix1, stage1, lo1, span1, lane);
                        if (TWO) stage_window(a.ix2, stage2, lo2, span2, lane);
                    }
                    r1.slot0 = fb1 - lo1;
                    r2.slot0 = fb2 - lo2;
                } else if (!TWO) {
                    mode = SRC_SLOTS;
                    need[2 * lane] = active ? fb1 : ~0u;
                    need[2 * lane + 1] = (active && lb1 != fb1) ? lb1 : ~0u;
                    __syncwarp();
                    stage_slots(a.ix1.blocks, stage1, need, lane);
                    r1.slot0 = 2 * lane;
                    r1.d1 = lb1 - fb1;
                }
            }
            cp_async_commit();
            // the record of the step after the next one goes into the ring slot just consumed
            if constexpr (!MULTI) {
                if (g + 64 < g_end) {
                    cursor_seek(in, cur, g + 64);
                    const uint4 *rec = cursor_record(in, cur, g + 64, RIN);
#pragma unroll
                    for (int k = 0; k < RIN; ++k) cp_async16(rb + k, rec + k);
                }
            }
            cp_async_commit();
            ring ^= 1u;
            const uint32_t rpos1 = (uint32_t)base1 & (kBlockSyms - 1), rpos2 = (uint32_t)base2 & (kBlockSyms - 1);

            // ---- bit updates on the merged node, while the copies are in flight ----
            if (active && a.write) {
                NodeStat st;
                node_bit_updates<TWO, W>(a, base1 + base2, s1, s2, st);
                st_lcp += st.lcp; st_min += st.nmin; st_upd += st.upd; st_da += st.da;
            }
            cp_async_wait_group<1>();                                       // the staged blocks (and the next step's record)
            if (IN_S && kWindowTma && mode == SRC_WINDOW) { mbar_wait(mbar, mbar_phase); mbar_phase ^= 1u; }
            __syncwarp();                                                   // every lane's copies are visible to the warp

            // ---- ranks -> children; child c is right-maximal iff >= 2 of its 5 gaps are non-empty ----
            ChildSide<W> k1, k2;
            Gaps<W> gp;
            if (active) {
                if constexpr (IN_S) {
                    if (mode == SRC_WINDOW) expand_small<TWO, SRC_WINDOW>(a, multi_super, r1, r2, base1, rpos1, s1, base2, rpos2, s2, k1, k2, gp, st_rank);
                    else if (mode == SRC_SLOTS) expand_small<TWO, SRC_SLOTS>(a, multi_super, r1, r2, base1, rpos1, s1, base2, rpos2, s2, k1, k2, gp, st_rank);
                    else expand_small<TWO, SRC_GLOBAL>(a, multi_super, r1, r2, base1, rpos1, s1, base2, rpos2, s2, k1, k2, gp, st_rank);
                } else {
                    expand_wide<TWO>(a, base1, s1, base2, s2, k1, k2, gp, st_rank);
                }
            }
            bool valid[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) valid[c] = gp.valid(c);
            uint32_t before[4], tot[4];
            const uint32_t vm = warp_child_slots(valid, lane, before, tot);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if ((vm >> c) & 1u) {
                    uint4 *dst = region[c] + before[c] * ROUT;
                    store_child<OUT_S, W>(dst, k1, c);
                    if (TWO) store_child<OUT_S, W>(dst + RSIDE_OUT, k2, c);
                }
                region[c] += tot[c] * ROUT;
                run_cnt[c] += tot[c];
                if (MULTI) count_dests(sm.dcnt[warp] + c * kMaxDest, dcur[c], dbound[c], out, lane, (vm >> c) & 1u,
                                       k1.base[c] + (TWO ? k2.base[c] : 0ull), tot[c]);
            }
        }
        if (!MULTI) {
            if (lane < 4) {
                uint32_t v = 0;
#pragma unroll
                for (int c = 0; c < 4; ++c) if (lane == c) v = run_cnt[c];
                out.cnt[(size_t)lane * out.K + run] = v;
            }
        } else {                                          // entry (h, c, run) for every destination h
            __syncwarp();
            for (int i = lane; i < 4 * kMaxDest; i += 32) {
                const uint32_t c = i / kMaxDest, h = i % kMaxDest;
                if (h < out.D) out.cnt[((size_t)h * 4 + c) * out.K + run] = sm.dcnt[warp][i];
            }
            __syncwarp();
        }
    }
    if (a.write) {
        flush_stat(a, C_LCP, st_lcp);
        flush_stat(a, C_NMIN, st_min);
        flush_stat(a, C_RANK, st_rank);
        flush_stat(a, C_BITUPD, st_upd);
        if (TWO) flush_stat(a, C_DA, st_da);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) max_size = max(max_size, (uint64_t)__shfl_xor_sync(0xffffffffu, (unsigned long long)max_size, s));
    if (lane == 0 && max_size) atomicMax(&a.sweep->maxsz, (unsigned long long)max_size);
}

// ---------------------------------------------------------------------------------------------
// Phase 2 sweep: leaves (intervals of W#).  One thread per leaf (pair); record = 16 bytes
// {first, second} (a pair: 32 bytes).  Same warp-per-run structure.  Leaves are sparse in position
// space: the two blocks of a leaf are staged per lane (SLOTS).
// ---------------------------------------------------------------------------------------------
template <bool TWO, bool MULTI>
struct LeafSmem {
    static constexpr int RU = TWO ? 2 : 1;            // uint4 per record
    static constexpr int RING = MULTI ? kRecRing : 2; // steps of records in flight
    uint4 stage[kNavWarps][64 * kBlockU4];            // two blocks per lane
    uint4 recbuf[kNavWarps][RING * 32 * RU];          // records of the next steps (ring, slot = lane)
    uint64_t rbar[kNavWarps][kRecRing];               // MULTI: completion barriers of the record ring
    uint32_t need[kNavWarps][64];
    uint32_t dcnt[kNavWarps][4 * kMaxDest];           // MULTI: children of the current run per (queue, destination)
};

// rank at an absolute position whose block is staged in `slot`
__device__ __forceinline__ void rank_slot(const DevIndex &ix, const uint4 *stage, uint32_t slot, uint64_t pos, uint64_t out[4]) {
    uint4 lo, hi;
    load_block_smem(stage, slot, lo, hi);
    uint32_t r[4];
    block_rank(lo, hi, (int)((uint32_t)pos & (kBlockSyms - 1)), r);
    const uint64_t *sb = ix.super + (pos >> kSuperShift) * 4;
    out[0] = sb[0] + r[0]; out[1] = sb[1] + r[1]; out[2] = sb[2] + r[2]; out[3] = sb[3] + r[3];
}

template <bool TWO, bool MULTI>
__global__ void __launch_bounds__(kNavThreads, TWO ? kPairCtas : kNodeCtas)
expand_leaves_kernel(const NavArgs a, const __grid_constant__ FrameInT<MULTI> in, const FrameOut out) {
    constexpr int RU = TWO ? 2 : 1, RING = LeafSmem<TWO, MULTI>::RING;
    __shared__ __align__(1024) LeafSmem<TWO, MULTI> sm;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4 *stage = sm.stage[warp];
    uint4 *recbuf = sm.recbuf[warp] + lane * RU;
    uint32_t *need = sm.need[warp];

    uint32_t it = 0;                                  // MULTI: steps this warp has taken (see the node sweep)
    if (MULTI) {
        if (lane == 0) for (int d = 0; d < RING; ++d) mbar_init(&sm.rbar[warp][d], 1);
        __syncwarp();
    }
    unsigned long long st_lcp = 0, st_da = 0;
    uint32_t st_rank = 0;
    // the first run of a warp is its own index (no atomic: small sweeps never touch the ticket), the
    // following ones are taken by ticket
    uint32_t next_run = blockIdx.x * kNavWarps + warp;
    while (true) {
        const uint32_t run = __shfl_sync(0xffffffffu, next_run, 0);
        if (run >= out.K) break;
        if (lane == 0) next_run = gridDim.x * kNavWarps + atomicAdd(&a.sweep->ticket, 1u);   // its latency hides behind this run
        const uint32_t g_begin = in.g_lo + run * out.run_cap, g_end = min(in.g_hi, g_begin + out.run_cap);
        Cursor cur;
        uint32_t seg = 0;
        if constexpr (MULTI) {
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int d = 0; d < RING; ++d)
                    if (g_begin + 32u * d < g_end)
                        fetch_records_bulk(in, seg, g_begin + 32u * d, g_end, RU, sm.recbuf[warp] + ((it + d) % RING) * (32 * RU), &sm.rbar[warp][(it + d) % RING]);
            }
        } else {
            cursor_open(in, cur, g_begin);
            if (g_begin + lane < g_end) {
                cursor_seek(in, cur, g_begin + lane);
                const uint4 *rec = cursor_record(in, cur, g_begin + lane, RU);
#pragma unroll
                for (int k = 0; k < RU; ++k) cp_async16(recbuf + k, rec + k);
            }
            cp_async_commit();
            if (g_begin + 32 + lane < g_end) {
                cursor_seek(in, cur, g_begin + 32 + lane);
                const uint4 *rec = cursor_record(in, cur, g_begin + 32 + lane, RU);
#pragma unroll
                for (int k = 0; k < RU; ++k) cp_async16(recbuf + 32 * RU + k, rec + k);
            }
            cp_async_commit();
        }
        uint32_t ring = 0;
        uint32_t run_cnt[4] = {0, 0, 0, 0};
        uint32_t dcur[4] = {0, 0, 0, 0};
        uint64_t dbound[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) dbound[c] = out.D > 1 ? out.range_len : ~0ull;
        if (MULTI) {
            for (int i = lane; i < 4 * kMaxDest; i += 32) sm.dcnt[warp][i] = 0;
            __syncwarp();
        }
        for (uint32_t g0 = g_begin; g0 < g_end; g0 += 32) {
            const uint32_t g = g0 + lane;
            const bool active = g < g_end;
            uint64_t f1 = 0, s1 = 0, f2 = 0, s2 = 0;
            uint4 *rb;
            if constexpr (MULTI) {
                mbar_wait(&sm.rbar[warp][it % RING], (it / RING) & 1u);
                rb = recbuf + (it % RING) * (32 * RU);
            } else {
                cp_async_wait_group<1>();                  // this lane's own record has landed (see the node sweep)
                rb = recbuf + ring * (32 * RU);
            }
            if (active) {
                const ulonglong2 *rec = reinterpret_cast<const ulonglong2 *>(rb);
                const ulonglong2 x = rec[0];
                f1 = x.x; s1 = x.y;
                if (TWO) { const ulonglong2 z = rec[1]; f2 = z.x; s2 = z.y; }
            }
            if constexpr (MULTI) {
                __syncwarp();
                if (lane == 0 && g0 + 32u * RING < g_end)
                    fetch_records_bulk(in, seg, g0 + 32u * RING, g_end, RU, sm.recbuf[warp] + (it % RING) * (32 * RU), &sm.rbar[warp][it % RING]);
                ++it;
            }
            // the (up to) two index blocks this leaf (pair) needs: slots 2 lane, 2 lane + 1 (TWO: one BWT each,
            // the second boundary of a side reads HBM unless it shares the block of the first)
            const uint32_t fb1 = (uint32_t)(f1 >> kBlockShift), lb1 = (uint32_t)(s1 >> kBlockShift);
            const uint32_t fb2 = (uint32_t)(f2 >> kBlockShift), lb2 = (uint32_t)(s2 >> kBlockShift);
            __syncwarp();                                  // the previous step's reads of the staging buffer are over
            need[2 * lane] = active ? fb1 : ~0u;
            if (!TWO) need[2 * lane + 1] = (active && lb1 != fb1) ? lb1 : ~0u;
            else need[2 * lane + 1] = active ? fb2 : ~0u;
            __syncwarp();
            if (!TWO) {
                stage_slots(a.ix1.blocks, stage, need, lane);
            } else {                                       // even slots come from BWT 1, odd slots from BWT 2
#pragma unroll
                for (int it = 0; it < 64 * kBlockU4 / 32; ++it) {
                    const uint32_t k = lane + it * 32;
                    const uint32_t slot = k >> 1, blk = need[slot];
                    const uint4 *src = (slot & 1u) ? a.ix2.blocks : a.ix1.blocks;
                    if (blk != ~0u) cp_async16(&stage[stage_slot(slot, k & 1)], src + (size_t)blk * kBlockU4 + (k & 1));
                }
            }
            cp_async_commit();
            if constexpr (!MULTI) {
                if (g + 64 < g_end) {                      // the record of the step after the next one
                    cursor_seek(in, cur, g + 64);
                    const uint4 *rec = cursor_record(in, cur, g + 64, RU);
#pragma unroll
                    for (int k = 0; k < RU; ++k) cp_async16(rb + k, rec + k);
                }
            }
            cp_async_commit();
            ring ^= 1u;
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
            cp_async_wait_group<1>();
            __syncwarp();
            // next_leaves (dna_bwt.hpp:358-379; two BWTs: ebwt2InDel.cpp:452-472): LF(range) = 2 ranks per BWT
            uint64_t lo1[4] = {0, 0, 0, 0}, hi1[4] = {0, 0, 0, 0}, lo2[4] = {0, 0, 0, 0}, hi2[4] = {0, 0, 0, 0};
            if (active) {
                rank_slot(a.ix1, stage, 2 * lane, f1, lo1);
                st_rank++;
                if (s1 > f1) {
                    if (lb1 == fb1) rank_slot(a.ix1, stage, 2 * lane, s1, hi1);
                    else if (!TWO) rank_slot(a.ix1, stage, 2 * lane + 1, s1, hi1);
                    else rank4(a.ix1, s1, hi1);
                    st_rank++;
                } else { hi1[0] = lo1[0]; hi1[1] = lo1[1]; hi1[2] = lo1[2]; hi1[3] = lo1[3]; }
                if (TWO) {
                    rank_slot(a.ix2, stage, 2 * lane + 1, f2, lo2);
                    st_rank++;
                    if (s2 > f2) {
                        if (lb2 == fb2) rank_slot(a.ix2, stage, 2 * lane + 1, s2, hi2);
                        else rank4(a.ix2, s2, hi2);
                        st_rank++;
                    } else { hi2[0] = lo2[0]; hi2[1] = lo2[1]; hi2[2] = lo2[2]; hi2[3] = lo2[3]; }
                }
            }
            bool valid[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) valid[c] = active && ((hi1[c] - lo1[c]) + (hi2[c] - lo2[c]) >= 2);
            uint32_t before[4], tot[4];
            const uint32_t vm = warp_child_slots(valid, lane, before, tot);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if ((vm >> c) & 1u) {
                    ulonglong2 *o = reinterpret_cast<ulonglong2 *>(out.base + (((size_t)c * out.K + run) * out.run_cap + run_cnt[c] + before[c]) * RU);
                    o[0] = make_ulonglong2(a.ix1.F[c] + lo1[c], a.ix1.F[c] + hi1[c]);
                    if (TWO) o[1] = make_ulonglong2(a.ix2.F[c] + lo2[c], a.ix2.F[c] + hi2[c]);
                }
                run_cnt[c] += tot[c];
                // a leaf (pair) lives at the merged position of its first suffix
                if (MULTI) count_dests(sm.dcnt[warp] + c * kMaxDest, dcur[c], dbound[c], out, lane, (vm >> c) & 1u,
                                       a.ix1.F[c] + lo1[c] + (TWO ? a.ix2.F[c] + lo2[c] : 0ull), tot[c]);
            }
        }
        if (!MULTI) {
            if (lane < 4) {
                uint32_t v = 0;
#pragma unroll
                for (int c = 0; c < 4; ++c) if (lane == c) v = run_cnt[c];
                out.cnt[(size_t)lane * out.K + run] = v;
            }
        } else {
            __syncwarp();
            for (int i = lane; i < 4 * kMaxDest; i += 32) {
                const uint32_t c = i / kMaxDest, h = i % kMaxDest;
                if (h < out.D) out.cnt[((size_t)h * 4 + c) * out.K + run] = sm.dcnt[warp][i];
            }
            __syncwarp();
        }
    }
    if (a.write) {
        flush_stat(a, C_LCP, st_lcp);
        flush_stat(a, C_RANK, st_rank);
        if (TWO) flush_stat(a, C_DA, st_da);
    }
}

// ---------------------------------------------------------------------------------------------
// After a sweep: exclusive prefix of the per-(queue, run) counts, the hints of the next sweep's reads
// and the total for the host.  One CTA per group of 256 entries: group_sum_kernel adds up every group,
// frame_index_kernel takes the sum of the groups before its own as offset and scans its 256 entries.
// ---------------------------------------------------------------------------------------------
constexpr int kIndexThreads = 256;
__global__ void __launch_bounds__(kIndexThreads)
group_sum_kernel(const uint32_t *__restrict__ cnt, uint32_t n_entries, uint32_t *__restrict__ gsum) {
    __shared__ uint32_t s_warp[kIndexThreads / 32];
    const uint32_t j = blockIdx.x * kIndexThreads + threadIdx.x;
    uint32_t v = j < n_entries ? cnt[j] : 0u;
    v = __reduce_add_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < kIndexThreads / 32; ++w) t += s_warp[w];
        gsum[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(kIndexThreads)
frame_index_kernel(const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ gsum, uint32_t n_entries, uint32_t K, uint32_t D,
                   uint32_t *__restrict__ P, uint32_t *__restrict__ off, uint32_t *__restrict__ hint, SweepDev *sweep, HostCtl *host,
                   unsigned long long seq) {
    __shared__ unsigned long long s_warp[kIndexThreads / 32];
    __shared__ unsigned long long s_off;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // offset of this group
    unsigned long long part = 0;
    for (uint32_t i = threadIdx.x; i < blockIdx.x; i += kIndexThreads) part += gsum[i];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) part += __shfl_xor_sync(0xffffffffu, part, s);
    if (lane == 0) s_warp[warp] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < kIndexThreads / 32; ++w) t += s_warp[w];
        s_off = t;
    }
    __syncthreads();
    const unsigned long long goff = s_off;
    __syncthreads();
    // scan of the group's entries
    const uint32_t j = blockIdx.x * kIndexThreads + threadIdx.x;
    const uint32_t c = j < n_entries ? cnt[j] : 0u;
    uint32_t incl = c;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, s);
        if (lane >= s) incl += y;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned long long p = goff + incl - c;
    for (int w = 0; w < warp; ++w) p += s_warp[w];
    if (j < n_entries) {
        P[j] = (uint32_t)p;
        // every multiple of 256 inside [p, p + c) starts a hint
        for (unsigned long long t = (p + 255) >> 8; (t << 8) < p + c; ++t) hint[t] = j;
        if (off) {                                        // ranged sweeps: slots of the region taken by the destinations before this one
            const uint32_t h = j / (4 * K), ck = j - h * 4 * K;
            uint32_t o = 0;
            for (uint32_t hh = 0; hh < h; ++hh) o += cnt[(size_t)hh * 4 * K + ck];
            off[j] = o;
        }
        if (j % K == 0) ((volatile uint32_t *)host->qstart)[j / K] = (uint32_t)p;
    }
    if (j == n_entries - 1) {                             // the last entry knows the total
        const unsigned long long total = p + c;
        P[n_entries] = (uint32_t)total;
        hint[(total + 255) >> 8] = n_entries - 1;         // sentinel: upper bound of the last partial group
        ((volatile uint32_t *)host->qstart)[4 * D] = (uint32_t)total;
        *(volatile unsigned long long *)&host->total = total;
        *(volatile unsigned long long *)&host->maxsz = sweep->maxsz;
    }
    // the sequence number goes out after every CTA's part of the table
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t done = atomicAdd(&sweep->done, 1u);
        if (done == gridDim.x - 1) {
            __threadfence_system();
            *(volatile unsigned long long *)&host->seq = seq;
        }
    }
}

// MULTI: the records of every entry copied to their global index (one warp per entry), so that destination h,
// queue c is ONE contiguous array [P[(4h+c)K], P[(4h+c+1)K)) that a peer reads with coalesced loads and no metadata
__global__ void __launch_bounds__(256)
compact_frame_kernel(const uint4 *__restrict__ base, const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ P,
                     const uint32_t *__restrict__ off, uint32_t n_entries, uint32_t K, uint32_t run_cap, int ru, uint4 *__restrict__ dst) {
    const uint32_t e = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (e >= n_entries) return;
    const uint32_t n = cnt[e] * ru;
    if (!n) return;
    const uint32_t h = e / (4 * K), ck = e - h * 4 * K;             // ck = c * K + k: the region
    const uint4 *src = base + ((size_t)ck * run_cap + off[e]) * ru;
    uint4 *out = dst + (size_t)P[e] * ru;
    for (uint32_t i = lane; i < n; i += 32) out[i] = src[i];
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

// layout of one frame inside its arena block: records, then cnt[4DK], gsum[], P[4DK+1], off[4DK], hint[]
struct FrameLayout {
    uint32_t K, run_cap, D;
    size_t rec_bytes, cnt_off, gsum_off, p_off, off_off, hint_off, total_bytes;
    FrameLayout(uint64_t n_in, uint32_t run, int ru_out, uint32_t dests = 1, bool ranged = false) {
        auto pad = [](size_t b) { return (b + 255) & ~(size_t)255; };
        run_cap = run;
        D = dests;
        K = (uint32_t)((n_in + run - 1) / run);
        const size_t ne = (size_t)4 * D * K;
        rec_bytes = (size_t)4 * K * run_cap * ru_out * sizeof(uint4);
        cnt_off = pad(rec_bytes);
        gsum_off = cnt_off + pad(ne * 4);
        p_off = gsum_off + pad((ne / 256 + 1) * 4);
        off_off = p_off + pad((ne + 1) * 4);
        hint_off = off_off + pad(D > 1 || ranged ? ne * 4 : 0);       // `off`: only the ranged sweeps compact their frames
        total_bytes = hint_off + pad(((size_t)4 * K * run_cap / 256 + 2) * 4);
    }
};

struct Chunk {                          // records [g_lo, g_hi) of a frame
    FrameIn in{};
    int level = 0;                      // tree depth of the records (selects the arena end of the next frame)
    bool small = false;                 // record form (internal nodes)
    uint64_t bound = ~0ull;             // upper bound on the size of any node of the chunk
    std::shared_ptr<Frame> frame;
    uint64_t total() const { return (uint64_t)in.g_hi - in.g_lo; }
};

struct SweepStats {
    uint64_t items = 0, sweeps = 0, max_chunk = 0;
    double ms_alloc = 0, ms_sync = 0, ms_max_alloc = 0, ms_max_sync = 0;   // host wall time (E2I_DEBUG)
    double ms_sweep = 0, ms_index = 0;                                    // device time of the two kernels (E2I_DEBUG)
};

static inline double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Cut the first `take` records off a chunk (position-contiguous prefix).
static Chunk split_head(Chunk &c, uint64_t take) {
    Chunk head = c;
    head.in.g_hi = c.in.g_lo + (uint32_t)take;
    c.in.g_lo = head.in.g_hi;
    return head;
}

struct PassCfg {
    bool leaves, two;
    uint64_t budget;          // bytes of the frame arena
    uint64_t depth_hint;      // expected depth of the traversal below a cut level
    uint32_t K, k_right;
    uint32_t resident_warps;  // warps the persistent grid keeps in flight
    uint32_t runs_per_warp;   // runs a warp should get in a large sweep
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
        const int ru_out = cfg.leaves ? (cfg.two ? 2 : 1) : node_rec_u4(out_small, cfg.two);
        const double out_bytes = 16.0 * ru_out * 4 + 1;  // four queues, each with room for every input record
        // max_chunk: largest chunk swept whole (two of its frames fit the arena: level-synchronous case).
        // split_chunk: chunk size once a level has to be cut (depth-first case): a path of such chunks down
        // to the deepest level must fit the arena next to the frame that is being cut.
        const uint64_t max_chunk = std::min<uint64_t>(1ull << 30, std::max<uint64_t>(65536, (uint64_t)((double)cfg.budget / (out_bytes * 2.5))));
        const uint64_t split_chunk = std::max<uint64_t>(256, std::min<uint64_t>(max_chunk, (uint64_t)((double)cfg.budget * 0.45 / (out_bytes * (double)cfg.depth_hint))));
        uint64_t take = cur.total() <= max_chunk ? cur.total() : std::min<uint64_t>(cur.total(), split_chunk);
        void *mem = nullptr;
        const double ta = now_ms();
        uint32_t run_len = 32;
        while (true) {   // shrink the chunk until its output frame fits the pool
            // runs: a few per resident warp (the first is the warp's own index, the others are taken by
            // ticket, which evens out the tail), between 32 and kMaxRun records each
            run_len = (uint32_t)std::min<uint64_t>(kMaxRun, std::max<uint64_t>(32, (take / ((uint64_t)cfg.resident_warps * cfg.runs_per_warp) + 31) / 32 * 32));
            const FrameLayout lay(take, run_len, ru_out);
            mem = ctx->arena.alloc((cur.level + 1) & 1, lay.total_bytes);
            if (mem) break;
            if (take <= 256) { set_error("frontier memory exhausted (arena %llu bytes, %llu in use): raise the frontier budget",
                                          (unsigned long long)ctx->arena.size(), (unsigned long long)ctx->arena.in_use()); return E2I_ERR_MEMORY; }
            take /= 2;
        }
        { const double d = now_ms() - ta; ss.ms_alloc += d; ss.ms_max_alloc = std::max(ss.ms_max_alloc, d); }
        Chunk work;
        if (take < cur.total()) {
            work = split_head(cur, take);
            stack.push_back(std::move(cur));
        } else {
            work = std::move(cur);
        }
        auto frame = std::make_shared<Frame>();
        frame->arena = &ctx->arena;
        frame->side = (work.level + 1) & 1;
        frame->p = mem;
        const FrameLayout lay(take, run_len, ru_out);
        char *fb = static_cast<char *>(mem);
        FrameOut fo{};
        fo.base = reinterpret_cast<uint4 *>(fb);
        fo.cnt = reinterpret_cast<uint32_t *>(fb + lay.cnt_off);
        fo.gsum = reinterpret_cast<uint32_t *>(fb + lay.gsum_off);
        fo.K = lay.K;
        fo.run_cap = lay.run_cap;
        fo.D = 1;
        fo.range_len = ~0ull;
        uint32_t *P = reinterpret_cast<uint32_t *>(fb + lay.p_off), *hint = reinterpret_cast<uint32_t *>(fb + lay.hint_off);
        if (ctx->ticket_next == kSweepSlots) {           // ring of sweep control blocks used up: zero it again
            E2I_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            E2I_CUDA_TRY(cudaMemsetAsync(ctx->ctl, 0, kSweepSlots * sizeof(SweepDev), ctx->stream));
            ctx->ticket_next = 0;
        }
        args.sweep = reinterpret_cast<SweepDev *>(ctx->ctl) + ctx->ticket_next++;
        const uint64_t depth = (uint64_t)work.level;     // every record of a sweep has this depth
        args.bits = (depth >= cfg.K ? 1u : 0u) | (depth >= cfg.k_right ? 2u : 0u);
        const unsigned long long seq = ++ctx->sweep_seq;
        static const bool debug = std::getenv("E2I_DEBUG") != nullptr;
        if (debug) cudaEventRecord(ctx->ev[4], ctx->stream);
        const uint32_t n_groups = (4 * fo.K + kIndexThreads - 1) / kIndexThreads;
        launch(args, work.in, fo, in_small, out_small);
        if (debug) cudaEventRecord(ctx->ev[5], ctx->stream);
        group_sum_kernel<<<n_groups, kIndexThreads, 0, ctx->stream>>>(fo.cnt, 4 * fo.K, fo.gsum);
        frame_index_kernel<<<n_groups, kIndexThreads, 0, ctx->stream>>>(fo.cnt, fo.gsum, 4 * fo.K, fo.K, 1, P, nullptr, hint, args.sweep, hctl, seq);
        if (debug) cudaEventRecord(ctx->ev[6], ctx->stream);
        E2I_CUDA_TRY(cudaGetLastError());
        ctx->n_launch += 3;
        ctx->n_d2h += sizeof(HostCtl);
        const double tsy = now_ms();
        {   // wait for the totals: poll the mapped sequence word
            volatile unsigned long long *seqp = &hctl->seq;
            unsigned spins = 0;
            while (*seqp != seq) {
                if ((++spins & 0xfffu) == 0) {
                    const cudaError_t q = cudaStreamQuery(ctx->stream);
                    if (q == cudaSuccess) {
                        if (*seqp == seq) break;
                        set_error("traversal sweep finished without publishing its counts");
                        return E2I_ERR_CUDA;
                    }
                    if (q != cudaErrorNotReady) { set_error("CUDA error in a traversal sweep: %s", cudaGetErrorString(q)); return E2I_ERR_CUDA; }
                }
            }
            std::atomic_thread_fence(std::memory_order_acquire);
        }
        { const double d = now_ms() - tsy; ss.ms_sync += d; ss.ms_max_sync = std::max(ss.ms_max_sync, d); }
        if (debug) {
            float m1 = 0, m2 = 0;
            cudaEventSynchronize(ctx->ev[6]);
            cudaEventElapsedTime(&m1, ctx->ev[4], ctx->ev[5]);
            cudaEventElapsedTime(&m2, ctx->ev[5], ctx->ev[6]);
            ss.ms_sweep += m1; ss.ms_index += m2;
        }
        ss.items += take;
        ss.sweeps++;
        ss.max_chunk = std::max<uint64_t>(ss.max_chunk, take);
        const uint64_t n_out = *(volatile unsigned long long *)&hctl->total;
        Chunk next;
        next.frame = frame;
        next.level = work.level + 1;
        next.small = out_small;
        next.bound = cfg.leaves ? ~0ull : std::min<uint64_t>(work.bound, ((volatile unsigned long long *)&hctl->maxsz)[0]);
        next.in.s.base = fo.base;
        next.in.s.P = P;
        next.in.s.off = nullptr;
        next.in.s.hint = hint;
        next.in.s.K = fo.K;
        next.in.s.run_cap = fo.run_cap;
        next.in.g_lo = 0;
        next.in.g_hi = (uint32_t)n_out;
        work.frame.reset();
        if (n_out) stack.push_back(std::move(next));
    }
    return E2I_OK;
}

// records [g_lo, g_hi) of a chunk copied to the host in order (sharding: the frames involved are small)
static int fetch_chunk(e2i_ctx *ctx, const Chunk &c, int ru, std::vector<uint64_t> &host) {
    const uint32_t n_entries = 4 * c.in.s.K;
    std::vector<uint32_t> P(n_entries + 1);
    E2I_CUDA_TRY(cudaMemcpyAsync(P.data(), c.in.s.P, P.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    E2I_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    host.resize((size_t)c.total() * ru * 2);
    for (uint32_t j = 0; j < n_entries; ++j) {
        const uint64_t lo = std::max<uint64_t>(P[j], c.in.g_lo), hi = std::min<uint64_t>(P[j + 1], c.in.g_hi);
        if (lo >= hi) continue;
        const uint4 *src = c.in.s.base + ((size_t)j * c.in.s.run_cap + (lo - P[j])) * ru;   // entry j = (queue, run) in order
        E2I_CUDA_TRY(cudaMemcpyAsync(host.data() + (lo - c.in.g_lo) * ru * 2, src, (hi - lo) * ru * 16, cudaMemcpyDeviceToHost, ctx->stream));
    }
    E2I_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return E2I_OK;
}

}  // namespace e2i

using namespace e2i;

static uint64_t padded_words32(uint64_t bits) { return ((bits + 31) / 32 + 63) / 64 * 64 + 64; }

// comm == nullptr: this rank traverses the subtrees dealt to `shard` (or everything when n_shards == 1).
// comm != nullptr: position-range sharding -- rank r processes the nodes (leaves) whose first position lies in
// its range of the suffix array and pulls its records from the frames of all ranks over peer memory, level
// by level (run_pass_ranged).
static int navigate_impl(e2i_ctx *ctx, const e2i_index *b1, const e2i_index *b2, const e2i_params *p, int shard, int n_shards, e2i_comm *comm,
                         e2i_lcpbits **out, e2i_bits **da_out, e2i_stats *st) {
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
    const bool dbg_time = std::getenv("E2I_DEBUG") != nullptr;
    const double t_enter = now_ms();

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
        const bool ipc = comm && comm->needs_ipc();
        if ((ctx->arena_bytes < want && ctx->arena_bytes < budget) || (ipc && !ctx->arena_ipc)) TRYF(arena_alloc(ctx, want, ipc));
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
    {   // the sweep kernels use ~20 KB of static shared memory per CTA and want 4-7 CTAs per SM: ask for the large carveout
        const int pct = (int)cudaSharedmemCarveoutMaxShared;
        TRYF(cudaFuncSetAttribute(expand_nodes_kernel<false, true, true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
        TRYF(cudaFuncSetAttribute(expand_nodes_kernel<true, true, true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
        TRYF(cudaFuncSetAttribute(expand_leaves_kernel<false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
        TRYF(cudaFuncSetAttribute(expand_leaves_kernel<true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    }
    const uint32_t ctas_per_sm = two ? (uint32_t)kPairCtas : (uint32_t)kNodeCtas;

    auto run_pass = [&](bool leaves, SweepStats &ss) -> int {
        PassCfg cfg;
        cfg.leaves = leaves;
        cfg.two = two;
        cfg.budget = budget;
        // depth of the traversal: the internal-node pass is never deeper than the leaf pass that ran before it
        cfg.depth_hint = leaves ? 1024 : st->levels_leaves + 16;
        cfg.K = (uint32_t)p->K;
        cfg.k_right = (uint32_t)p->k_right;
        cfg.resident_warps = (uint32_t)ctx->sm_count * ctas_per_sm * kNavWarps;
        cfg.runs_per_warp = leaves ? 1u : 4u;
        if (const char *e = std::getenv(leaves ? "E2I_RUNS_LEAVES" : "E2I_RUNS_NODES")) cfg.runs_per_warp = (uint32_t)std::max(1, atoi(e));
        // the root frame: one record, one run
        const int ru_root = leaves ? (two ? 2 : 1) : node_rec_u4(false, two);
        const size_t root_bytes = 1024;
        char *rootmem = static_cast<char *>(ctx->arena.alloc(0, root_bytes));
        if (!rootmem) { set_error("frontier arena too small"); return E2I_ERR_MEMORY; }
        uint64_t hostroot[root_bytes / 8] = {0};
        uint64_t *rec = hostroot;
        if (leaves) {                                   // first_leaf (dna_bwt.hpp:313-317)
            rec[0] = 0; rec[1] = b1->F[0];
            if (two) { rec[2] = 0; rec[3] = b2->F[0]; }
        } else {                                        // root (dna_bwt.hpp:296-308) as a WIDE record
            pack_wide_host(rec, 0, b1->F, b1->n);
            if (two) pack_wide_host(rec + 6, 0, b2->F, b2->n);
        }
        uint32_t *hp = reinterpret_cast<uint32_t *>(hostroot) + 128;     // byte 512: P[5] = {0,1,1,1,1}, byte 640: hint[2] = {0, 3}
        hp[0] = 0; hp[1] = hp[2] = hp[3] = hp[4] = 1;
        hp[32] = 0; hp[33] = 3;
        E2I_CUDA_TRY(cudaMemcpyAsync(rootmem, hostroot, root_bytes, cudaMemcpyHostToDevice, s));
        (void)ru_root;
        Chunk root{};
        root.in.s.base = reinterpret_cast<uint4 *>(rootmem);
        root.in.s.P = reinterpret_cast<uint32_t *>(rootmem + 512);
        root.in.s.off = nullptr;
        root.in.s.hint = reinterpret_cast<uint32_t *>(rootmem + 640);
        root.in.s.K = 1;
        root.in.s.run_cap = 1;
        root.in.g_lo = 0;
        root.in.g_hi = 1;
        root.small = false;
        root.bound = std::max<uint64_t>(b1->n, two ? b2->n : 0);
        root.frame = std::make_shared<Frame>();
        root.frame->arena = &ctx->arena;
        root.frame->side = 0;
        root.frame->p = rootmem;
        auto launch = [&](NavArgs &a, const FrameIn &fi, const FrameOut &fo, bool in_small, bool out_small) {
            const uint32_t want = (fo.K + kNavWarps - 1) / kNavWarps;
            if (leaves) {
                const uint32_t grid = std::min<uint32_t>(want, (uint32_t)ctx->sm_count * ctas_per_sm);
                if (two) expand_leaves_kernel<true, false><<<grid, kNavThreads, 0, s>>>(a, fi, fo);
                else expand_leaves_kernel<false, false><<<grid, kNavThreads, 0, s>>>(a, fi, fo);
            } else if (in_small) {
                const uint32_t grid = std::min<uint32_t>(want, (uint32_t)ctx->sm_count * ctas_per_sm);
                if (two) expand_nodes_kernel<true, true, true, false><<<grid, kNavThreads, 0, s>>>(a, fi, fo);
                else expand_nodes_kernel<false, true, true, false><<<grid, kNavThreads, 0, s>>>(a, fi, fo);
            } else {
                const uint32_t grid = std::min<uint32_t>(want, (uint32_t)ctx->sm_count * 2);
                if (out_small) { if (two) expand_nodes_kernel<true, false, true, false><<<grid, kNavThreads, 0, s>>>(a, fi, fo); else expand_nodes_kernel<false, false, true, false><<<grid, kNavThreads, 0, s>>>(a, fi, fo); }
                else { if (two) expand_nodes_kernel<true, false, false, false><<<grid, kNavThreads, 0, s>>>(a, fi, fo); else expand_nodes_kernel<false, false, false, false><<<grid, kNavThreads, 0, s>>>(a, fi, fo); }
            }
        };
        if (n_shards == 1) {
            args.write = 1;
            return run_frontier(ctx, std::move(root), cfg, args, launch, ss, 0, nullptr);
        }
        // shared top of the tree
        std::vector<Chunk> dealt;
        args.write = shard == 0;
        SweepStats top;
        E2I_TRY(run_frontier(ctx, std::move(root), cfg, args, launch, top, kDealItems, &dealt));
        if (shard == 0) { ss.items += top.items; ss.sweeps += top.sweeps; ss.max_chunk = std::max(ss.max_chunk, top.max_chunk); }
        args.write = 1;
        for (Chunk &c : dealt) {
            // deal by cumulated interval length: fetch every record
            const int ru = leaves ? (two ? 2 : 1) : node_rec_u4(c.small, two);
            const int words = ru * 2, side_words = words / (two ? 2 : 1);
            const uint64_t tot = c.total();
            std::vector<uint64_t> host;
            E2I_TRY(fetch_chunk(ctx, c, ru, host));
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
            Chunk part = c;
            part.in.g_lo = c.in.g_lo + (uint32_t)lo;
            part.in.g_hi = c.in.g_lo + (uint32_t)hi;
            E2I_TRY(run_frontier(ctx, std::move(part), cfg, args, launch, ss, 0, nullptr));
        }
        return E2I_OK;
    };

    // ---- position-range sharding (comm != nullptr) ------------------------------------------------------
    // Rank r owns the suffix-array positions [r * range_len, (r + 1) * range_len) and processes the records
    // whose (merged) first position lies there.  A sweep writes its children into the rank's OWN gappy frame,
    // counted per destination rank; after the sweep every rank publishes where each (destination, queue)
    // piece of its frame starts, and the next sweep PULLS its input -- the pieces addressed to it, in
    // (queue, source rank) order, which is position order -- straight out of the peers' frames over
    // NVLink (P2P loads in the kernel's record prefetch).  One barrier per level; no collective, no staging.
    // Every rank's frontier stays as dense as the single-GPU frontier, which is what the subtree deal loses.
    struct LevelInfo {                                  // what a rank publishes after a sweep (one comm slot)
        uint64_t total, maxsz;
        uint64_t rec_off;                               // the compacted records inside the rank's arena
        uint32_t small, error;
        uint32_t qstart[4 * kMaxDest + 1];              // record index where (destination h, queue c) starts, total at the end
        void *arena_base;
        cudaIpcMemHandle_t arena_handle;
    };
    static_assert(sizeof(LevelInfo) <= kCommSlotBytes, "LevelInfo must fit a comm slot");
    auto run_pass_ranged = [&](bool leaves, SweepStats &ss) -> int {
        const int world = comm->world, me = comm->rank;
        HostCtl *hctl = reinterpret_cast<HostCtl *>(ctx->ctl_host);
        const uint64_t range_len = (n + world - 1) / world;
        const uint32_t resident_warps = (uint32_t)ctx->sm_count * ctas_per_sm * kNavWarps;
        LevelInfo mine;
        std::memset(&mine, 0, sizeof mine);
        mine.arena_base = ctx->arena_mem;
        if (comm->needs_ipc()) E2I_CUDA_TRY(cudaIpcGetMemHandle(&mine.arena_handle, ctx->arena_mem));
        const cudaIpcMemHandle_t my_handle = mine.arena_handle;
        char *const abase = static_cast<char *>(ctx->arena_mem);
        std::shared_ptr<Frame> f_prev, f_old;           // compacted records of this level (read by everybody) / of the level before
        // level 0: the root lives on rank 0 (its first position is 0): one record, addressed to (destination 0, queue 0)
        if (me == 0) {
            const int ru_root = leaves ? (two ? 2 : 1) : node_rec_u4(false, two);
            char *mem = static_cast<char *>(ctx->arena.alloc(0, (size_t)ru_root * 16));
            if (!mem) { set_error("frontier arena too small"); return E2I_ERR_MEMORY; }
            uint64_t rec[12] = {0};
            if (leaves) { rec[0] = 0; rec[1] = b1->F[0]; if (two) { rec[2] = 0; rec[3] = b2->F[0]; } }
            else { pack_wide_host(rec, 0, b1->F, b1->n); if (two) pack_wide_host(rec + 6, 0, b2->F, b2->n); }
            E2I_CUDA_TRY(cudaMemcpyAsync(mem, rec, (size_t)ru_root * 16, cudaMemcpyHostToDevice, s));
            E2I_CUDA_TRY(cudaStreamSynchronize(s));
            f_prev = std::make_shared<Frame>();
            f_prev->arena = &ctx->arena; f_prev->side = 0; f_prev->p = mem;
            mine.total = 1;
            mine.rec_off = (uint64_t)(mem - abase);
            for (uint32_t i = 1; i <= 4 * (uint32_t)world; ++i) mine.qstart[i] = 1;
        }
        mine.maxsz = std::max<uint64_t>(b1->n, two ? b2->n : 0);
        mine.small = 0;
        std::vector<LevelInfo> info((size_t)world);
        std::vector<char *> peer_base((size_t)world, nullptr);
        double t_bar = 0, t_wait = 0, t_host = 0;
        static const bool debug = std::getenv("E2I_DEBUG") != nullptr;
        auto wait_seq = [&](unsigned long long seq) -> int {          // the mapped sequence word of the index kernel
            volatile unsigned long long *seqp = &hctl->seq;
            unsigned spins = 0;
            while (*seqp != seq) {
                if ((++spins & 0xfffu) == 0) {
                    const cudaError_t q = cudaStreamQuery(s);
                    if (q == cudaSuccess) { if (*seqp == seq) break; set_error("traversal sweep finished without publishing its counts"); return E2I_ERR_CUDA; }
                    if (q != cudaErrorNotReady) { set_error("CUDA error in a traversal sweep: %s", cudaGetErrorString(q)); return E2I_ERR_CUDA; }
                }
            }
            std::atomic_thread_fence(std::memory_order_acquire);
            return E2I_OK;
        };
        for (int level = 0;; ++level) {
            std::memcpy(comm->slot(me), &mine, sizeof mine);
            const double tb0 = now_ms();
            comm->barrier();                              // every rank has finished the previous sweep and published its records
            uint64_t all = 0, bound = 0;
            uint32_t err = 0;
            for (int r = 0; r < world; ++r) {
                std::memcpy(&info[r], comm->slot(r), sizeof(LevelInfo));
                all += info[r].total; bound = std::max(bound, info[r].maxsz); err |= info[r].error;
                if (!peer_base[r]) peer_base[r] = static_cast<char *>(r == me ? (void *)abase : comm->peer_ptr(r, info[r].arena_base, info[r].arena_handle));
                if (!peer_base[r]) { set_error("cannot map the frame arena of rank %d", r); err |= 1; }
            }
            comm->barrier();                              // everybody has read the slots: they may be rewritten
            t_bar += now_ms() - tb0;
            const double th0 = now_ms();
            f_old.reset();                                // the records of two levels ago are dead on every rank
            f_old = f_prev;
            f_prev.reset();
            if (err) { if (!mine.error) set_error("position-range traversal failed on another rank"); return E2I_ERR_MEMORY; }
            if (all == 0) break;
            // my input: for every queue, the pieces the ranks addressed to me, in rank order (= position order)
            FrameInT<true> in;
            std::memset(&in, 0, sizeof in);
            const bool in_small = !leaves && info[0].small, out_small = !leaves && bound < kSmallLimit;
            const int ru_in = leaves ? (two ? 2 : 1) : node_rec_u4(in_small, two);
            uint32_t n_in = 0;
            for (int c = 0; c < 4; ++c) {
                for (int r = 0; r < world; ++r) {
                    const uint32_t lo = info[r].qstart[me * 4 + c], hi = info[r].qstart[me * 4 + c + 1];
                    if (hi <= lo) continue;
                    Segment &sg = in.seg[in.n_seg++];
                    sg.ptr = reinterpret_cast<const uint4 *>(peer_base[r] + info[r].rec_off) + (size_t)lo * ru_in;
                    sg.start = n_in; sg.len = hi - lo;
                    n_in += hi - lo;
                }
            }
            in.g_lo = 0; in.g_hi = n_in;
            std::memset(&mine, 0, sizeof mine);
            mine.arena_base = ctx->arena_mem;
            mine.arena_handle = my_handle;
            mine.small = out_small;
            if (n_in == 0) continue;                      // nothing for me on this level (I still keep the barriers company)
            const int ru_out = leaves ? (two ? 2 : 1) : node_rec_u4(out_small, two);
            const uint32_t run_len = (uint32_t)std::min<uint64_t>(kMaxRun, std::max<uint64_t>(32, (n_in / ((uint64_t)resident_warps * (leaves ? 1 : 4)) + 31) / 32 * 32));
            const FrameLayout lay(n_in, run_len, ru_out, (uint32_t)world, true);
            // the gappy frame is a temporary on top of THIS level's records; the compacted children go to the other end
            const int side_tmp = level & 1, side_out = (level + 1) & 1;
            char *mem = static_cast<char *>(ctx->arena.alloc(side_tmp, lay.total_bytes));
            if (!mem) {
                set_error("frontier memory exhausted in the position-range traversal (arena %llu bytes, level of %u records): raise the frontier budget",
                          (unsigned long long)ctx->arena.size(), n_in);
                mine.error = 1;
                continue;                                 // the other ranks learn about it at the next barrier
            }
            Frame tmp{&ctx->arena, side_tmp, mem};
            FrameOut fo{};
            fo.base = reinterpret_cast<uint4 *>(mem);
            fo.cnt = reinterpret_cast<uint32_t *>(mem + lay.cnt_off);
            fo.gsum = reinterpret_cast<uint32_t *>(mem + lay.gsum_off);
            fo.K = lay.K; fo.run_cap = lay.run_cap; fo.D = (uint32_t)world; fo.range_len = range_len;
            uint32_t *P = reinterpret_cast<uint32_t *>(mem + lay.p_off), *off = reinterpret_cast<uint32_t *>(mem + lay.off_off),
                     *hint = reinterpret_cast<uint32_t *>(mem + lay.hint_off);
            if (ctx->ticket_next == kSweepSlots) {
                E2I_CUDA_TRY(cudaStreamSynchronize(s));
                E2I_CUDA_TRY(cudaMemsetAsync(ctx->ctl, 0, kSweepSlots * sizeof(SweepDev), s));
                ctx->ticket_next = 0;
            }
            args.sweep = reinterpret_cast<SweepDev *>(ctx->ctl) + ctx->ticket_next++;
            args.bits = ((uint64_t)level >= (uint64_t)p->K ? 1u : 0u) | ((uint64_t)level >= (uint64_t)p->k_right ? 2u : 0u);
            args.write = 1;
            const unsigned long long seq = ++ctx->sweep_seq;
            const uint32_t grid = std::min<uint32_t>((fo.K + kNavWarps - 1) / kNavWarps, (uint32_t)ctx->sm_count * ctas_per_sm);
            if (leaves) {
                if (two) expand_leaves_kernel<true, true><<<grid, kNavThreads, 0, s>>>(args, in, fo);
                else expand_leaves_kernel<false, true><<<grid, kNavThreads, 0, s>>>(args, in, fo);
            } else if (in_small) {
                if (two) expand_nodes_kernel<true, true, true, true><<<grid, kNavThreads, 0, s>>>(args, in, fo);
                else expand_nodes_kernel<false, true, true, true><<<grid, kNavThreads, 0, s>>>(args, in, fo);
            } else if (out_small) {
                if (two) expand_nodes_kernel<true, false, true, true><<<grid, kNavThreads, 0, s>>>(args, in, fo);
                else expand_nodes_kernel<false, false, true, true><<<grid, kNavThreads, 0, s>>>(args, in, fo);
            } else {
                if (two) expand_nodes_kernel<true, false, false, true><<<grid, kNavThreads, 0, s>>>(args, in, fo);
                else expand_nodes_kernel<false, false, false, true><<<grid, kNavThreads, 0, s>>>(args, in, fo);
            }
            const uint32_t n_entries = 4 * fo.D * fo.K, n_groups = (n_entries + kIndexThreads - 1) / kIndexThreads;
            group_sum_kernel<<<n_groups, kIndexThreads, 0, s>>>(fo.cnt, n_entries, fo.gsum);
            frame_index_kernel<<<n_groups, kIndexThreads, 0, s>>>(fo.cnt, fo.gsum, n_entries, fo.K, fo.D, P, off, hint, args.sweep, hctl, seq);
            E2I_CUDA_TRY(cudaGetLastError());
            ctx->n_launch += 3;
            ctx->n_d2h += sizeof(HostCtl);
            t_host += now_ms() - th0;
            const double tw0 = now_ms();
            E2I_TRY(wait_seq(seq));
            t_wait += now_ms() - tw0;
            ss.items += n_in;
            ss.sweeps++;
            ss.max_chunk = std::max<uint64_t>(ss.max_chunk, n_in);
            mine.total = *(volatile unsigned long long *)&hctl->total;
            mine.maxsz = leaves ? 0 : *(volatile unsigned long long *)&hctl->maxsz;
            for (uint32_t i = 0; i <= 4 * fo.D; ++i) mine.qstart[i] = ((volatile uint32_t *)hctl->qstart)[i];
            if (mine.total) {
                // compact: every (destination, queue) piece becomes one contiguous array that the peers read directly
                char *cmem = static_cast<char *>(ctx->arena.alloc(side_out, (size_t)mine.total * ru_out * 16));
                if (!cmem) {
                    set_error("frontier memory exhausted in the position-range traversal (arena %llu bytes): raise the frontier budget", (unsigned long long)ctx->arena.size());
                    mine.error = 1; mine.total = 0;
                    continue;
                }
                f_prev = std::make_shared<Frame>();
                f_prev->arena = &ctx->arena; f_prev->side = side_out; f_prev->p = cmem;
                compact_frame_kernel<<<(n_entries + 7) / 8, 256, 0, s>>>(fo.base, fo.cnt, P, off, n_entries, fo.K, fo.run_cap, ru_out, reinterpret_cast<uint4 *>(cmem));
                E2I_CUDA_TRY(cudaGetLastError());
                ctx->n_launch++;
                const double tc0 = now_ms();
                E2I_CUDA_TRY(cudaStreamSynchronize(s));   // the peers read these records right after the next barrier
                t_wait += now_ms() - tc0;
                mine.rec_off = (uint64_t)(cmem - abase);
            }
        }
        if (debug) std::fprintf(stderr, "[e2i] ranged %s pass, rank %d: %llu sweeps, barriers %.1f ms, host %.1f ms, waiting for the sweeps %.1f ms\n",
                                leaves ? "leaf" : "node", me, (unsigned long long)ss.sweeps, t_bar, t_host, t_wait);
        return E2I_OK;
    };

    unsigned long long tot[C_NCOUNTERS];
    double t_setup = 0;
    if (dbg_time) { cudaStreamSynchronize(s); t_setup = now_ms() - t_enter; }
    // ---- Phase 2: leaves ----
    TRYF(cudaMemsetAsync(stripes, 0, stripe_bytes, s));
    TRYF(cudaEventRecord(ctx->ev[0], s));
    SweepStats sl;
    // With a communicator the LEAF pass is position-range sharded (measured faster than the subtree deal at every N:
    // small records, no redundant top of the tree); the NODE pass is too only on request (E2I_RANGED_NODES=1): at N = 8
    // seven of eight records come out of a peer's HBM and the NVLink latency costs more than the sparse frontier
    // of a subtree shard does (C4, 8 GPUs: 158 ms against 119 ms per rank).
    const char *rn_env = std::getenv("E2I_RANGED_NODES");
    const bool ranged_nodes = comm && rn_env && std::strcmp(rn_env, "0") != 0;
    int rc = comm ? run_pass_ranged(true, sl) : run_pass(true, sl);
    if (rc != E2I_OK) return fail(rc);
    TRYF(cudaEventRecord(ctx->ev[1], s));
    rc = sum_stripes(tot);
    if (rc != E2I_OK) return fail(rc);
    const uint64_t first = (comm ? comm->rank : shard) == 0 ? 1 : 0;          // lcp_values starts at 1 (ebwt2InDel.cpp:575)
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
    rc = ranged_nodes ? run_pass_ranged(false, sn) : run_pass(false, sn);
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
        std::fprintf(stderr, "[e2i] leaves: %llu sweeps, alloc %.2f ms, sync %.2f ms (max %.2f), kernels %.2f + index %.2f ms | nodes: %llu sweeps, alloc %.2f ms, sync %.2f ms (max %.2f), kernels %.2f + index %.2f ms\n",
                     (unsigned long long)sl.sweeps, sl.ms_alloc, sl.ms_sync, sl.ms_max_sync, sl.ms_sweep, sl.ms_index,
                     (unsigned long long)sn.sweeps, sn.ms_alloc, sn.ms_sync, sn.ms_max_sync, sn.ms_sweep, sn.ms_index);
    if (dbg_time) std::fprintf(stderr, "[e2i] navigate: setup (bit vectors, arena) %.1f ms, %.1f ms of host wall time in all\n", t_setup, now_ms() - t_enter);
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

extern "C" int e2i_navigate_shard(e2i_ctx *ctx, const e2i_index *b1, const e2i_index *b2, const e2i_params *p,
                                  int shard, int n_shards, e2i_lcpbits **out, e2i_bits **da_out, e2i_stats *st) {
    return navigate_impl(ctx, b1, b2, p, shard, n_shards, nullptr, out, da_out, st);
}

extern "C" int e2i_navigate(e2i_ctx *ctx, const e2i_index *b1, const e2i_index *b2, const e2i_params *p,
                            e2i_lcpbits **out, e2i_bits **da_out, e2i_stats *st) {
    return navigate_impl(ctx, b1, b2, p, 0, 1, nullptr, out, da_out, st);
}

extern "C" int e2i_navigate_ranged(e2i_ctx *ctx, e2i_comm *comm, const e2i_index *b1, const e2i_index *b2, const e2i_params *p,
                                   e2i_lcpbits **out, e2i_bits **da_out, e2i_stats *st) {
    if (!comm || comm->world < 1 || comm->world > kMaxDest) { set_error("e2i_navigate_ranged: bad communicator"); return E2I_ERR_ARG; }
    // one rank: the plain traversal (E2I_RANGED_FORCE=1 keeps the ranged machinery, to measure what it costs by itself)
    if (comm->world == 1 && !std::getenv("E2I_RANGED_FORCE")) return navigate_impl(ctx, b1, b2, p, 0, 1, nullptr, out, da_out, st);
    return navigate_impl(ctx, b1, b2, p, comm->rank, comm->world, comm, out, da_out, st);
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
