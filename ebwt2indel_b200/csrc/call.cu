// call.cu -- a16-a20: phase 4 on the device.
//
// Replaces the cluster scans of run_one_dataset / run_two_datasets / run_two_datasets_da
// (/root/reference/ebwt2InDel.cpp:1609-1655, 1395-1445, 1510-1560) and the three find_variants
// (:840-934, 941-1005, 1013-1096) with extract_consensus (:265-319) and extract_dna (:325-342).
//
// Three kernels per slab of suffix-array positions:
//   1. scan_clusters: streams the 3n bits, finds maximal runs of  LCP_threshold[2i] & !LCP_minima[i],
//      keeps runs of length >= 2*mcov_out that are closed before n, builds the per-individual
//      symbol histograms (rank differences in modes -1/-2, a DA-guided walk in mode -d; TERM counts
//      as 'A', include.hpp:275-289) and appends the clusters that pass the frequent-allele filter
//      to a candidate list in SA order (ordered append by decoupled look-back).
//   2. consensus: one thread per (candidate, allele): LF(range, c), then k_left-1 steps of 4-way LF
//      following the largest child (ties -> A,C,G,T order).
//   3. right_context: first position of the cluster with LCP >= k_right, then k_right FL steps.
// Classification and text formatting (a21-a23) run on the host in snp_format.cpp.
#include <algorithm>
#include <chrono>

#include "common.cuh"
#include "lookback.cuh"

namespace e2i {

constexpr int kScanThreads = 256;
constexpr int kScanTile = kScanThreads * 64;     // positions per tile
constexpr int kMaxCand = 22;                     // clusters of length >= 2 that can start inside one 64-bit word

struct Candidate {
    uint64_t begin, end;        // merged SA range [begin, end)
    uint64_t b1, e1, b2, e2;    // range in BWT 1 / BWT 2 (mode -2); modes -1/-d: b1,e1 = begin,end
    uint32_t mask0, mask1;      // frequent alleles of individual 0 / 1 (bit c = A,C,G,T)
    uint32_t pad0, pad1;
};

struct CallCtl {
    uint32_t ticket;
    uint32_t pad;
    unsigned long long n_cand;
    unsigned long long n_clusters, clust_size, rank_q, n_pass;
    unsigned long long hist[201];
};

struct CallArgs {
    DevIndex ix1, ix2;
    const uint32_t *thr, *minima, *da;
    const uint64_t *da_rank512;
    uint64_t n;                 // merged length
    uint64_t pos_begin, pos_end;// clusters starting in [pos_begin, pos_end) belong to this launch
    uint64_t first_tile;        // pos_begin / kScanTile
    uint32_t n_tiles;
    uint32_t mcov, q;
    int mode;                   // 1, 2, 3
    int k_left, k_right;
    Candidate *cand;
    uint64_t cand_cap;
    unsigned long long *desc;
    uint32_t epoch;
    CallCtl *ctl;
};

__device__ __forceinline__ uint32_t even_bits(uint32_t x) {
    x &= 0x55555555u;
    x = (x | (x >> 1)) & 0x33333333u;
    x = (x | (x >> 2)) & 0x0f0f0f0fu;
    x = (x | (x >> 4)) & 0x00ff00ffu;
    x = (x | (x >> 8)) & 0x0000ffffu;
    return x;
}

// flagged positions [64w, 64w+64): LCP_threshold[2i] && !LCP_minima[i]  (ebwt2InDel.cpp:1611)
__device__ __forceinline__ uint64_t flag_word(const CallArgs &a, uint64_t w) {
    if (w * 64 >= a.n) return 0;
    const uint4 t = __ldg(reinterpret_cast<const uint4 *>(a.thr) + w);
    const uint2 m = __ldg(reinterpret_cast<const uint2 *>(a.minima) + w);
    const uint64_t lo = even_bits(t.x) | (even_bits(t.y) << 16), hi = even_bits(t.z) | (even_bits(t.w) << 16);
    uint64_t f = (lo | (hi << 32)) & ~((uint64_t)m.x | ((uint64_t)m.y << 32));
    const uint64_t rem = a.n - w * 64;
    if (rem < 64) f &= (1ull << rem) - 1;
    return f;
}

__device__ __forceinline__ int bit_at(const uint32_t *words, uint64_t i) { return (int)((__ldg(words + (i >> 5)) >> (i & 31)) & 1u); }

// number of 1s of the DA in [0, i)
__device__ __forceinline__ uint64_t da_rank1(const CallArgs &a, uint64_t i) {
    const uint64_t g = i >> 9;
    uint64_t r = __ldg(a.da_rank512 + g);
    const uint32_t *w = a.da + g * 16;
    const int full = (int)((i & 511) >> 5), rem = (int)(i & 31);
    for (int k = 0; k < full; ++k) r += __popc(__ldg(w + k));
    if (rem) r += __popc(__ldg(w + full) & ((1u << rem) - 1u));
    return r;
}

// Per-individual symbol histogram of BWT[begin, end) guided by the document array (mode -d,
// ebwt2InDel.cpp:1019), 32 positions per step: plane words and DA words are equally aligned.
// TERM has both low plane bits clear, so it lands in the A bin like base_to_int's default
// (include.hpp:275-289).
__device__ __forceinline__ void hist_da_words(const DevIndex &ix, const uint32_t *da, uint64_t begin, uint64_t end,
                                              uint64_t cnt0[4], uint64_t cnt1[4]) {
    const uint32_t *blocks32 = reinterpret_cast<const uint32_t *>(ix.blocks);
    const uint64_t w0 = begin >> 5, w1 = (end - 1) >> 5;
    uint32_t c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0};
    for (uint64_t w = w0; w <= w1; ++w) {
        uint32_t m = 0xffffffffu;
        if (w == w0) m &= 0xffffffffu << (begin & 31);
        if (w == w1 && (end & 31)) m &= 0xffffffffu >> (32 - (end & 31));
        const uint32_t *blk = blocks32 + (w >> 1) * 8;       // [0..1] counters, [2..3] plane a, [4..5] plane b, [6..7] TERM plane
        const int k = (int)(w & 1);
        const uint32_t pa = __ldg(blk + 2 + k), pb = __ldg(blk + 4 + k), d = __ldg(da + w);
        const uint32_t m1 = m & d, m0 = m & ~d;
        c0[0] += __popc(m0 & ~pa & ~pb); c0[1] += __popc(m0 & pa & ~pb); c0[2] += __popc(m0 & ~pa & pb); c0[3] += __popc(m0 & pa & pb);
        c1[0] += __popc(m1 & ~pa & ~pb); c1[1] += __popc(m1 & pa & ~pb); c1[2] += __popc(m1 & ~pa & pb); c1[3] += __popc(m1 & pa & pb);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) { cnt0[c] = c0[c]; cnt1[c] = c1[c]; }
}

__device__ __forceinline__ uint32_t frequent_mask(const uint64_t cnt[4], uint32_t mcov) {
    uint32_t m = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) if (cnt[c] >= mcov) m |= 1u << c;
    return m;
}

// Kernel 1.  Persistent CTAs take tiles by ticket (ordered append needs tile order).
__global__ void __launch_bounds__(kScanThreads)
scan_clusters_kernel(const CallArgs a) {
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_wcnt[kScanThreads / 32];
    __shared__ unsigned long long s_base;
    __shared__ unsigned int s_hist[201];
    __shared__ unsigned long long s_stat[4];
    for (int i = threadIdx.x; i < 201; i += kScanThreads) s_hist[i] = 0;
    if (threadIdx.x < 4) s_stat[threadIdx.x] = 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long t_clusters = 0, t_size = 0, t_rank = 0, t_cand = 0;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = atomicAdd(&a.ctl->ticket, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= a.n_tiles) break;
        const uint64_t w = (a.first_tile + tile) * kScanThreads + threadIdx.x;   // 64-bit word of flags
        const uint64_t f = flag_word(a, w);
        uint64_t prev = 0;
        if (w > 0 && f) prev = flag_word(a, w - 1) >> 63;
        uint64_t starts = f & ~((f << 1) | prev);
        // local candidate store
        uint64_t c_begin[kMaxCand], c_end[kMaxCand];
        uint64_t c_r[kMaxCand][4];
        uint32_t c_mask[kMaxCand];
        int nc = 0;
        while (starts) {
            const int b = __ffsll((long long)starts) - 1;
            starts &= starts - 1;
            const uint64_t begin = w * 64 + b;
            if (begin < a.pos_begin || begin >= a.pos_end) continue;
            // end = first unflagged position after begin
            uint64_t end;
            {
                uint64_t inv = ~f & (b == 63 ? 0ull : (~0ull << (b + 1)));
                uint64_t ww = w;
                while (!inv) { ++ww; inv = ~flag_word(a, ww); }
                end = ww * 64 + (__ffsll((long long)inv) - 1);
            }
            if (end >= a.n) continue;                      // a run still open at i = n is never closed (:1609-1655)
            const uint64_t len = end - begin;
            t_size += len;
            if (len <= 200) atomicAdd(&s_hist[len], 1u);
            if (len < 2ull * a.mcov) continue;             // :1629
            t_clusters++;
            uint64_t cnt0[4] = {0, 0, 0, 0}, cnt1[4] = {0, 0, 0, 0};
            uint64_t b1 = begin, e1 = end, b2 = 0, e2 = 0;
            if (a.mode == 1) {
                uint64_t rb[4], re[4];
                rank4(a.ix1, begin, rb);
                rank4(a.ix1, end, re);
                t_rank += 2;
                uint64_t acgt = 0;
#pragma unroll
                for (int c = 0; c < 4; ++c) { cnt0[c] = re[c] - rb[c]; acgt += cnt0[c]; }
                cnt0[0] += len - acgt;                     // TERM counted as 'A'
            } else if (a.mode == 3) {
                hist_da_words(a.ix1, a.da, begin, end, cnt0, cnt1);
            } else {
                const uint64_t ob = da_rank1(a, begin), oe = da_rank1(a, end);
                b2 = ob; e2 = oe; b1 = begin - ob; e1 = end - oe;
                uint64_t rb[4], re[4], acgt = 0;
                rank4(a.ix1, b1, rb); rank4(a.ix1, e1, re);
#pragma unroll
                for (int c = 0; c < 4; ++c) { cnt0[c] = re[c] - rb[c]; acgt += cnt0[c]; }
                cnt0[0] += (e1 - b1) - acgt;
                rank4(a.ix2, b2, rb); rank4(a.ix2, e2, re);
                acgt = 0;
#pragma unroll
                for (int c = 0; c < 4; ++c) { cnt1[c] = re[c] - rb[c]; acgt += cnt1[c]; }
                cnt1[0] += (e2 - b2) - acgt;
                t_rank += 4;
            }
            const uint32_t m0 = frequent_mask(cnt0, a.mcov), m1 = frequent_mask(cnt1, a.mcov);
            const int n0 = __popc(m0), n1 = __popc(m1);
            bool pass;
            if (a.mode == 1) pass = n0 >= 2 && !(a.q > 0 && (uint32_t)n0 > a.q);                         // :961-966
            else pass = n0 > 0 && n1 > 0 && !(a.q > 0 && ((uint32_t)n0 > a.q || (uint32_t)n1 > a.q));    // :870-880
            if (pass) t_cand++;
            // Two samples: a pair is only ever emitted when the last characters of the two left contexts --
            // the alleles themselves -- differ (:921, :1083).  A cluster where both individuals have the same
            // single frequent allele (the overwhelming majority) can produce no output and changes no
            // counter, so its consensus walks and right context are never computed.
            if (pass && a.mode != 1 && n0 == 1 && m0 == m1) pass = false;
            if (pass && nc < kMaxCand) {
                c_begin[nc] = begin; c_end[nc] = end;
                c_r[nc][0] = b1; c_r[nc][1] = e1; c_r[nc][2] = b2; c_r[nc][3] = e2;
                c_mask[nc] = m0 | (m1 << 8);
                nc++;
            }
        }
        // ordered append: thread order inside the tile, tile order across tiles
        const uint32_t bal_incl = [&] {
            uint32_t x = (uint32_t)nc;
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, s); if (lane >= s) x += y; }
            return x;
        }();
        if (lane == 31) s_wcnt[warp] = bal_incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t x = lane < kScanThreads / 32 ? s_wcnt[lane] : 0u;
            const uint32_t own = x;
#pragma unroll
            for (int s = 1; s < 8; s <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, s); if (lane >= s) x += y; }
            const uint32_t tot = __shfl_sync(0xffffffffu, x, 7);
            if (lane < kScanThreads / 32) s_wcnt[lane] = x - own;
            unsigned long long agg[1] = {tot}, excl[1];
            lookback_exclusive<1>(a.desc, a.epoch, tile, agg, excl);
            if (lane == 0) {
                s_base = excl[0];
                if (tile == a.n_tiles - 1) a.ctl->n_cand = excl[0] + tot;
            }
        }
        __syncthreads();
        unsigned long long slot = s_base + s_wcnt[warp] + (bal_incl - (uint32_t)nc);
        for (int k = 0; k < nc; ++k, ++slot) {
            if (slot < a.cand_cap) {
                Candidate cd;
                cd.begin = c_begin[k]; cd.end = c_end[k];
                cd.b1 = c_r[k][0]; cd.e1 = c_r[k][1]; cd.b2 = c_r[k][2]; cd.e2 = c_r[k][3];
                cd.mask0 = c_mask[k] & 0xffu; cd.mask1 = c_mask[k] >> 8; cd.pad0 = cd.pad1 = 0;
                a.cand[slot] = cd;
            }
        }
    }
    // flush statistics once per CTA
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        t_clusters += __shfl_xor_sync(0xffffffffu, t_clusters, s);
        t_size += __shfl_xor_sync(0xffffffffu, t_size, s);
        t_rank += __shfl_xor_sync(0xffffffffu, t_rank, s);
        t_cand += __shfl_xor_sync(0xffffffffu, t_cand, s);
    }
    if (lane == 0) {
        atomicAdd(&s_stat[0], t_clusters);
        atomicAdd(&s_stat[1], t_size);
        atomicAdd(&s_stat[2], t_rank);
        atomicAdd(&s_stat[3], t_cand);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_stat[0]) atomicAdd(&a.ctl->n_clusters, s_stat[0]);
        if (s_stat[1]) atomicAdd(&a.ctl->clust_size, s_stat[1]);
        if (s_stat[2]) atomicAdd(&a.ctl->rank_q, s_stat[2]);
        if (s_stat[3]) atomicAdd(&a.ctl->n_pass, s_stat[3]);
    }
    for (int i = threadIdx.x; i < 201; i += kScanThreads)
        if (s_hist[i]) atomicAdd(&a.ctl->hist[i], (unsigned long long)s_hist[i] * (unsigned long long)i);
}

// Kernel 2: consensus left contexts (extract_consensus, ebwt2InDel.cpp:265-319; consensus_letter :243-261).
// One thread per (candidate, slot); slots 0-3 = alleles of individual 0, 4-7 = individual 1.
__global__ void consensus_kernel(const CallArgs a, uint64_t n_cand, char *__restrict__ left,
                                 int32_t *__restrict__ support, uint8_t *__restrict__ reached,
                                 unsigned long long *rank_q) {
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n_cand * 8) return;
    const uint64_t ci = gid >> 3;
    const int slot = (int)(gid & 7), ind = slot >> 2, c = slot & 3;
    const Candidate cd = a.cand[ci];
    const uint32_t mask = ind ? cd.mask1 : cd.mask0;
    reached[gid] = 0;
    if (!((mask >> c) & 1u)) return;
    if (a.mode != 1 && !((ind ? cd.mask0 : cd.mask1) & ~(1u << c))) return;   // no partner allele differs: never emitted
    const bool second = a.mode == 2 && ind == 1;
    const DevIndex &ix = second ? a.ix2 : a.ix1;
    uint64_t first = second ? cd.b2 : cd.b1, last = second ? cd.e2 : cd.e1;
    if (a.mode == 3) { first = cd.begin; last = cd.end; }
    char *out = left + gid * (uint64_t)a.k_left;
    uint64_t rb[4], re[4];
    unsigned nq = 0;
    // LF(range, c) (dna_bwt.hpp:168-192)
    rank4(ix, first, rb);
    nq++;
    if (last > first) { rank4(ix, last, re); nq++; } else { re[0] = rb[0]; re[1] = rb[1]; re[2] = rb[2]; re[3] = rb[3]; }
    uint64_t lo = ix.F[c] + rb[c], hi = ix.F[c] + re[c];
    support[gid] = (int32_t)(hi - lo);
    int k = 0;
    out[a.k_left - 1 - k] = "ACGT"[c];
    k++;
    for (int rem = a.k_left - 1; rem > 0; --rem) {
        rank4(ix, lo, rb);
        nq++;
        if (hi > lo) { rank4(ix, hi, re); nq++; } else { re[0] = rb[0]; re[1] = rb[1]; re[2] = rb[2]; re[3] = rb[3]; }
        int best = 0;
        uint64_t bs = re[0] - rb[0];
#pragma unroll
        for (int x = 1; x < 4; ++x) { const uint64_t sz = re[x] - rb[x]; if (sz > bs) { bs = sz; best = x; } }
        if (bs == 0) break;
        out[a.k_left - 1 - k] = "ACGT"[best];
        k++;
        lo = ix.F[best] + rb[best];
        hi = ix.F[best] + re[best];
    }
    reached[gid] = (k == a.k_left);
    atomicAdd(rank_q + (gid & 63), (unsigned long long)nq);
}

// Kernel 3: right context (find_variants' scan for LCP >= k_right + extract_dna, :325-342, 979-988, 901-912)
__global__ void right_context_kernel(const CallArgs a, uint64_t n_cand, char *__restrict__ right,
                                     uint8_t *__restrict__ right_len, uint8_t *__restrict__ has_right,
                                     unsigned long long *rank_q) {
    const uint64_t ci = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ci >= n_cand) return;
    const Candidate cd = a.cand[ci];
    uint64_t i = cd.begin, i0 = cd.b1, i1 = cd.b2;
    while (i < cd.end && !bit_at(a.thr, 2 * i + 1)) {
        if (a.mode == 2) { if (bit_at(a.da, i)) i1++; else i0++; }
        ++i;
    }
    has_right[ci] = 0;
    right_len[ci] = 0;
    if (!(i < cd.end)) return;
    const bool second = a.mode == 2 && bit_at(a.da, i);
    const DevIndex &ix = second ? a.ix2 : a.ix1;
    uint64_t pos = a.mode == 2 ? (second ? i1 : i0) : i;
    char *out = right + ci * (uint64_t)a.k_right;
    int k = 0, len = a.k_right;
    int c = f_code(ix, pos);
    unsigned nq = 0;
    while (c != 4 && len > 0) {
        out[k++] = "ACGT"[c];
        pos = fl_map(ix, pos, c);
        nq++;
        c = f_code(ix, pos);
        len--;
    }
    has_right[ci] = 1;
    right_len[ci] = (uint8_t)k;
    atomicAdd(rank_q + (ci & 63), (unsigned long long)nq);
}

// Kernel 4: final record layout on the device.  Per candidate: the reached left contexts of each
// individual move to the front of its group of four slots (A,C,G,T order kept; a destination slot
// never lies after its source), supports alike, and the e2i_call_rec is filled in.
__global__ void pack_calls_kernel(const CallArgs a, uint64_t n_cand, char *__restrict__ left, const int32_t *__restrict__ support,
                                  const uint8_t *__restrict__ reached, const uint8_t *__restrict__ right_len,
                                  const uint8_t *__restrict__ has_right, e2i_call_rec *__restrict__ recs) {
    const uint64_t ci = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ci >= n_cand) return;
    const Candidate cd = a.cand[ci];
    e2i_call_rec rec = {};                                 // padding bytes included: the records are copied out as bytes
    rec.begin = cd.begin;
    rec.end = cd.end;
    rec.right_len = right_len[ci];
    rec.has_right = has_right[ci];
#pragma unroll
    for (int k = 0; k < 8; ++k) rec.support[k] = 0;
    char *L = left + ci * 8 * (uint64_t)a.k_left;
    for (int ind = 0; ind < 2; ++ind) {
        int k = 0;
        for (int c = 0; c < 4; ++c) {
            const uint64_t src = ci * 8 + ind * 4 + c;
            if (!reached[src]) continue;
            if (k != c) {
                const char *from = L + (ind * 4 + c) * a.k_left;
                char *to = L + (ind * 4 + k) * a.k_left;
                for (int i = 0; i < a.k_left; ++i) to[i] = from[i];
            }
            rec.support[ind * 4 + k] = support[src];
            k++;
        }
        if (ind == 0) rec.n0 = (uint8_t)k; else rec.n1 = (uint8_t)k;
    }
    recs[ci] = rec;
}

// popcount of every 512-bit group of the DA, then an in-place exclusive scan (one CTA)
__global__ void da_group_popc_kernel(const uint32_t *__restrict__ words, uint64_t n_groups, uint64_t *__restrict__ out) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const uint4 *p = reinterpret_cast<const uint4 *>(words) + g * 4;
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { const uint4 v = __ldg(p + k); c += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w); }
    out[g] = c;
}

__global__ void __launch_bounds__(1024) scan_u64_kernel(uint64_t *data, uint64_t n) {
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint64_t base = 0; base < n; base += 1024) {
        const uint64_t i = base + threadIdx.x;
        const unsigned long long v = i < n ? data[i] : 0ull;
        unsigned long long x = v;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) { const unsigned long long y = __shfl_up_sync(0xffffffffu, x, s); if (lane >= s) x += y; }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            unsigned long long y = s_warp[lane];
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) { const unsigned long long z = __shfl_up_sync(0xffffffffu, y, s); if (lane >= s) y += z; }
            s_warp[lane] = y;
        }
        __syncthreads();
        const unsigned long long ex = s_carry + (warp ? s_warp[warp - 1] : 0ull) + x - v;
        if (i < n) data[i] = ex;
        __syncthreads();
        if (threadIdx.x == 0) s_carry += s_warp[31];
        __syncthreads();
    }
}

}  // namespace e2i

using namespace e2i;

// The directory is rebuilt on every e2i_call (two tiny kernels): the DA words are exposed through
// e2i_bits_device and are modified in place by the cross-GPU OR-combine, so a cached copy could be stale.
static int build_da_rank(e2i_ctx *ctx, e2i_bits *da) {
    const uint64_t n_groups = da->n_words32 / 16;
    if (!da->rank512) E2I_CUDA_TRY(dmalloc(ctx, &da->rank512, (n_groups + 1) * 8));
    da_group_popc_kernel<<<(unsigned)((n_groups + 255) / 256), 256, 0, ctx->stream>>>(da->words, n_groups, da->rank512);
    scan_u64_kernel<<<1, 1024, 0, ctx->stream>>>(da->rank512, n_groups);
    E2I_CUDA_TRY(cudaGetLastError());
    ctx->n_launch += 2;
    return E2I_OK;
}

extern "C" void e2i_calls_free(e2i_calls *c);

// Where the text goes when phase 4 formats its records on the device (e2i_call_snp): a malloc'ed host buffer that
// grows batch by batch; the records themselves then never leave the device.
struct TextSink {
    char *buf = nullptr;
    size_t len = 0, cap = 0;
    uint64_t next_cluster = 1, events = 0, clusters = 0;
    double ms = 0;                                      // host wall time spent formatting + copying the text
};

static int call_impl(e2i_ctx *ctx, const e2i_index *b1, const e2i_index *b2, const e2i_bits *da,
                     const e2i_lcpbits *l, const e2i_params *p, uint64_t pos_begin, uint64_t pos_end,
                     e2i_calls **out, TextSink *sink, bool keep_device, e2i_stats *st) {
    if (!ctx || !b1 || !l || !p || (!out && !sink) || !st) { set_error("e2i_call: null argument"); return E2I_ERR_ARG; }
    if (b2 && !da) { set_error("e2i_call: two BWTs need the document array produced by e2i_navigate"); return E2I_ERR_ARG; }
    if (p->k_left < 1 || p->k_left > 255 || p->k_right < 1 || p->k_right > 255) { set_error("e2i_call: k_left and k_right must be in [1,255]"); return E2I_ERR_ARG; }
    if (p->mcov_out < 1) { set_error("e2i_call: mcov_out must be >= 1"); return E2I_ERR_ARG; }
    const int mode = b2 ? 2 : (da ? 3 : 1);
    const uint64_t n = b1->n + (b2 ? b2->n : 0);
    if (l->n != n || (da && da->n != n)) { set_error("e2i_call: bitvector length does not match the BWT length"); return E2I_ERR_ARG; }
    E2I_CUDA_TRY(cudaSetDevice(ctx->device));
    Accounting acct(ctx, st);
    cudaStream_t s = ctx->stream;
    if (mode == 2) E2I_TRY(build_da_rank(ctx, const_cast<e2i_bits *>(da)));
    pos_end = std::min<uint64_t>(pos_end, n);

    e2i_calls *calls = new e2i_calls();
    calls->ctx = ctx;
    calls->k_left = p->k_left;
    calls->k_right = p->k_right;
    calls->two_samples = mode != 1;

    CallArgs a{};
    a.ix1 = b1->dev();
    a.ix2 = b2 ? b2->dev() : b1->dev();
    a.thr = l->thr;
    a.minima = l->minima;
    a.da = da ? da->words : nullptr;
    a.da_rank512 = da ? da->rank512 : nullptr;
    a.n = n;
    a.mcov = (uint32_t)p->mcov_out;
    a.q = (uint32_t)std::max(0, p->max_variants_per_position);
    a.mode = mode;
    a.k_left = p->k_left;
    a.k_right = p->k_right;

    // Phase 4 runs after the traversal, so the frame arena is idle: the control block, the candidate
    // list and the per-candidate outputs are carved out of it (no allocation on this path).
    if (ctx->arena_bytes < (64ull << 20)) {
        if (arena_alloc(ctx, 1ull << 30, false) != cudaSuccess) { cudaGetLastError(); delete calls; set_error("e2i_call: out of device memory"); return E2I_ERR_MEMORY; }
    }
    char *const abase = static_cast<char *>(ctx->arena_mem);
    CallCtl *dctl = reinterpret_cast<CallCtl *>(abase);
    unsigned long long *rank_q = reinterpret_cast<unsigned long long *>(abase + 4096);
    Candidate *cand = reinterpret_cast<Candidate *>(abase + 8192);
    const size_t kl = (size_t)p->k_left, kr = (size_t)p->k_right;
    const size_t out_per = 8 * kl + kr + sizeof(e2i_call_rec) + 8 * sizeof(int32_t) + 8 + 2 + 64;   // output bytes per candidate (+ slack)
    const uint64_t cand_cap = (ctx->arena_bytes / 4) / sizeof(Candidate);
    char *const obase = abase + 8192 + ((cand_cap * sizeof(Candidate) + 255) & ~(size_t)255);
    const uint64_t batch_cap = (uint64_t)((abase + ctx->arena_bytes - obase) / out_per);
    auto fail = [&](int rc) { e2i_calls_free(calls); return rc; };
#define TRYF(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__, cudaGetErrorString(_e)); return fail(E2I_ERR_CUDA); } } while (0)
    TRYF(cudaMemsetAsync(rank_q, 0, 64 * 8, s));
    TRYF(cudaEventRecord(ctx->ev[4], s));
    CallCtl hctl;
    auto now_ms = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_scan = 0, t_walk = 0, t_mark = now_ms();
    auto lap = [&](double &acc) { const double t = now_ms(); acc += t - t_mark; t_mark = t; };

    // page-locked result arrays [recs | left | right] with room for `cap` records; grown by moving
    const uint64_t gen = ++ctx->pinned_gen;
    uint64_t cap = 0, n_out = 0;
    const size_t rec_bytes = sizeof(e2i_call_rec) + 8 * kl + kr;
    auto layout = [&](void *basep, uint64_t c, e2i_call_rec *&r, char *&lf, char *&rt) {
        char *bp = static_cast<char *>(basep);
        r = reinterpret_cast<e2i_call_rec *>(bp);
        lf = bp + c * sizeof(e2i_call_rec);
        rt = lf + c * 8 * kl;
    };
    auto ensure = [&](uint64_t want) -> int {
        if (want <= cap) return E2I_OK;
        const uint64_t ncap = std::max<uint64_t>(want, std::max<uint64_t>(1024, 2 * cap));
        if (ncap * rec_bytes <= ctx->pinned_bytes && n_out == 0) {       // the cached buffer is large enough
            cap = ctx->pinned_bytes / rec_bytes;
            layout(ctx->pinned, cap, calls->recs, calls->left, calls->right);
            return E2I_OK;
        }
        void *np = nullptr;
        if (cudaMallocHost(&np, ncap * rec_bytes) != cudaSuccess) { cudaGetLastError(); set_error("e2i_call: cannot page-lock %llu bytes", (unsigned long long)(ncap * rec_bytes)); return E2I_ERR_MEMORY; }
        e2i_call_rec *r; char *lf, *rt;
        layout(np, ncap, r, lf, rt);
        if (n_out) {
            std::memcpy(r, calls->recs, n_out * sizeof(e2i_call_rec));
            std::memcpy(lf, calls->left, n_out * 8 * kl);
            std::memcpy(rt, calls->right, n_out * kr);
        }
        if (ctx->pinned) cudaFreeHost(ctx->pinned);
        ctx->pinned = np;
        ctx->pinned_bytes = ncap * rec_bytes;
        cap = ncap;
        calls->recs = r; calls->left = lf; calls->right = rt;
        return E2I_OK;
    };

    // consensus walks + right contexts + record packing for the candidates gathered so far
    uint64_t n_acc = 0;
    auto flush = [&]() -> int {
        if (!sink && !keep_device) E2I_TRY(ensure(n_out + n_acc));
        if (keep_device && n_out + n_acc > calls->d_cap) {           // device arrays, grown by moving (one batch is the rule)
            const uint64_t ncap = std::max<uint64_t>(n_out + n_acc, 2 * calls->d_cap);
            const size_t o_left = (ncap * sizeof(e2i_call_rec) + 255) & ~(size_t)255, o_right = o_left + ((ncap * 8 * kl + 255) & ~(size_t)255);
            char *nb = nullptr;
            if (dmalloc(ctx, &nb, o_right + ncap * kr + 256) != cudaSuccess) { cudaGetLastError(); set_error("e2i_call_device: out of device memory for %llu records", (unsigned long long)ncap); return E2I_ERR_MEMORY; }
            if (n_out) {
                E2I_CUDA_TRY(cudaMemcpyAsync(nb, calls->d_recs, n_out * sizeof(e2i_call_rec), cudaMemcpyDeviceToDevice, s));
                E2I_CUDA_TRY(cudaMemcpyAsync(nb + o_left, calls->d_left, n_out * 8 * kl, cudaMemcpyDeviceToDevice, s));
                E2I_CUDA_TRY(cudaMemcpyAsync(nb + o_right, calls->d_right, n_out * kr, cudaMemcpyDeviceToDevice, s));
            }
            dfree(ctx, calls->d_block);
            calls->d_block = nb;
            calls->d_recs = reinterpret_cast<e2i_call_rec *>(nb);
            calls->d_left = nb + o_left;
            calls->d_right = nb + o_right;
            calls->d_cap = ncap;
        }
        for (uint64_t c0 = 0; c0 < n_acc; c0 += batch_cap) {
            const uint64_t nb = std::min<uint64_t>(batch_cap, n_acc - c0);
            char *o = obase;
            auto carve = [&](size_t bytes) { char *r = o; o += (bytes + 63) & ~(size_t)63; return r; };
            char *d_left = carve(nb * 8 * kl);
            char *d_right = carve(nb * kr);
            e2i_call_rec *d_recs = reinterpret_cast<e2i_call_rec *>(carve(nb * sizeof(e2i_call_rec)));
            int32_t *d_support = reinterpret_cast<int32_t *>(carve(nb * 8 * sizeof(int32_t)));
            uint8_t *d_reached = reinterpret_cast<uint8_t *>(carve(nb * 8));
            uint8_t *d_rlen = reinterpret_cast<uint8_t *>(carve(nb));
            uint8_t *d_has = reinterpret_cast<uint8_t *>(carve(nb));
            CallArgs ab = a;
            ab.cand = cand + c0;
            E2I_CUDA_TRY(cudaMemsetAsync(d_support, 0, nb * 8 * sizeof(int32_t), s));
            E2I_CUDA_TRY(cudaMemsetAsync(d_left, 0, nb * 8 * kl, s));
            E2I_CUDA_TRY(cudaMemsetAsync(d_right, 0, nb * kr, s));
            consensus_kernel<<<(unsigned)((nb * 8 + 127) / 128), 128, 0, s>>>(ab, nb, d_left, d_support, d_reached, rank_q);
            right_context_kernel<<<(unsigned)((nb + 127) / 128), 128, 0, s>>>(ab, nb, d_right, d_rlen, d_has, rank_q);
            pack_calls_kernel<<<(unsigned)((nb + 127) / 128), 128, 0, s>>>(ab, nb, d_left, d_support, d_reached, d_rlen, d_has, d_recs);
            E2I_CUDA_TRY(cudaGetLastError());
            ctx->n_launch += 3;
            if (sink) {                                   // text on the device: only the characters go to the host
                E2I_CUDA_TRY(cudaStreamSynchronize(s));   // the walks are phase 4 time, not formatting time
                const double t0 = now_ms();
                char *d_text = nullptr;
                uint64_t tl = 0, cl = 0, ev = 0;
                E2I_TRY(format_device(ctx, d_recs, d_left, d_right, nb, p, mode != 1, sink->next_cluster, &d_text, &tl, &cl, &ev));
                if (sink->len + tl + 1 > sink->cap) {     // page-locked buffer (text_alloc); a second batch moves what is there
                    const size_t want = std::max<size_t>(sink->len + tl + 1, sink->cap + sink->cap / 2);
                    char *nb2 = text_alloc(want);
                    if (!nb2) { dfree(ctx, d_text); set_error("e2i_call_snp: cannot page-lock %llu bytes for the text", (unsigned long long)want); return E2I_ERR_MEMORY; }
                    if (sink->len) std::memcpy(nb2, sink->buf, sink->len);
                    e2i_buffer_free(sink->buf);
                    sink->buf = nb2;
                    sink->cap = want;
                }
                if (tl) E2I_CUDA_TRY(cudaMemcpyAsync(sink->buf + sink->len, d_text, tl, cudaMemcpyDeviceToHost, s));
                E2I_CUDA_TRY(cudaStreamSynchronize(s));
                dfree(ctx, d_text);
                ctx->n_d2h += tl;
                sink->len += tl;
                sink->next_cluster += cl;
                sink->clusters += cl;
                sink->events += ev;
                sink->ms += now_ms() - t0;
                n_out += nb;
                continue;
            }
            if (keep_device) {                            // the records stay in HBM for e2i_calls_snp
                E2I_CUDA_TRY(cudaMemcpyAsync(calls->d_recs + n_out, d_recs, nb * sizeof(e2i_call_rec), cudaMemcpyDeviceToDevice, s));
                E2I_CUDA_TRY(cudaMemcpyAsync(calls->d_left + n_out * 8 * kl, d_left, nb * 8 * kl, cudaMemcpyDeviceToDevice, s));
                E2I_CUDA_TRY(cudaMemcpyAsync(calls->d_right + n_out * kr, d_right, nb * kr, cudaMemcpyDeviceToDevice, s));
                E2I_CUDA_TRY(cudaStreamSynchronize(s));
                n_out += nb;
                continue;
            }
            // the device layout is the final one: three copies straight into the page-locked result arrays
            E2I_CUDA_TRY(cudaMemcpyAsync(calls->recs + n_out, d_recs, nb * sizeof(e2i_call_rec), cudaMemcpyDeviceToHost, s));
            E2I_CUDA_TRY(cudaMemcpyAsync(calls->left + n_out * 8 * kl, d_left, nb * 8 * kl, cudaMemcpyDeviceToHost, s));
            E2I_CUDA_TRY(cudaMemcpyAsync(calls->right + n_out * kr, d_right, nb * kr, cudaMemcpyDeviceToHost, s));
            ctx->n_d2h += nb * rec_bytes;
            E2I_CUDA_TRY(cudaStreamSynchronize(s));
            n_out += nb;
        }
        n_acc = 0;
        return E2I_OK;
    };

    // scan [pos_begin, pos_end) slab by slab; the candidates of all slabs accumulate in one list
    uint64_t slab = 1ull << 30;
    const uint64_t first = pos_begin / kScanTile * kScanTile, last = (pos_end + kScanTile - 1) / kScanTile * kScanTile;
    for (uint64_t sb = first; sb < pos_end;) {
        const uint64_t se = std::min<uint64_t>(sb + slab, last);
        a.pos_begin = std::max(sb, pos_begin);
        a.pos_end = std::min(se, pos_end);
        a.first_tile = sb / kScanTile;
        a.n_tiles = (uint32_t)((se - sb + kScanTile - 1) / kScanTile);
        a.cand = cand + n_acc;
        a.cand_cap = cand_cap - n_acc;
        if ((size_t)a.n_tiles > ctx->desc_words) {
            dfree(ctx, ctx->desc);
            ctx->desc = nullptr;
            ctx->desc_words = (size_t)a.n_tiles + 1024;
            TRYF(dmalloc(ctx, &ctx->desc, ctx->desc_words * 8));
            TRYF(cudaMemsetAsync(ctx->desc, 0, ctx->desc_words * 8, s));
            ctx->epoch = 0;
        }
        if (++ctx->epoch >= 0xffffu) { TRYF(cudaMemsetAsync(ctx->desc, 0, ctx->desc_words * 8, s)); ctx->epoch = 1; }
        a.desc = ctx->desc;
        a.epoch = ctx->epoch;
        a.ctl = dctl;
        TRYF(cudaMemsetAsync(dctl, 0, sizeof(CallCtl), s));
        const int grid = (int)std::min<uint64_t>(a.n_tiles, (uint64_t)ctx->sm_count * 8);
        scan_clusters_kernel<<<grid, kScanThreads, 0, s>>>(a);
        TRYF(cudaGetLastError());
        ctx->n_launch++;
        ctx->n_d2h += sizeof(CallCtl);
        TRYF(cudaMemcpyAsync(&hctl, dctl, sizeof(CallCtl), cudaMemcpyDeviceToHost, s));
        TRYF(cudaStreamSynchronize(s));
        lap(t_scan);
        const uint64_t nc = hctl.n_cand;
        if (nc > a.cand_cap) {           // the list is full: empty it, or (if it was empty) redo the slab in halves
            if (n_acc) { const int rc = flush(); if (rc != E2I_OK) return fail(rc); lap(t_walk); continue; }
            if (slab <= (uint64_t)kScanTile) { set_error("e2i_call: candidate list overflow (%llu > %llu)", (unsigned long long)nc, (unsigned long long)cand_cap); return fail(E2I_ERR_MEMORY); }
            slab = std::max<uint64_t>((uint64_t)kScanTile, (slab / 2) / kScanTile * kScanTile);
            continue;
        }
        sb = se;
        n_acc += nc;
        st->n_clusters += hctl.n_clusters;
        st->clust_size += hctl.clust_size;
        st->rank_call += hctl.rank_q;
        for (int i = 0; i <= 200; ++i) st->clust_sizes[i] += hctl.hist[i];
        st->candidates += hctl.n_pass;
    }
    { const int rc = flush(); if (rc != E2I_OK) return fail(rc); lap(t_walk); }
    calls->n = n_out;
    calls->gen = gen;
    if (std::getenv("E2I_DEBUG"))
        std::fprintf(stderr, "[e2i] call: scan+sync %.1f ms, walks+d2h %.1f ms, %llu records\n", t_scan, t_walk, (unsigned long long)n_out);
    unsigned long long hq[64];
    TRYF(cudaMemcpyAsync(hq, rank_q, sizeof hq, cudaMemcpyDeviceToHost, s));
    TRYF(cudaEventRecord(ctx->ev[5], s));
    TRYF(cudaStreamSynchronize(s));
    for (int i = 0; i < 64; ++i) st->rank_call += hq[i];
    float ms = 0;
    TRYF(cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]));
    st->ms_call += ms;
#undef TRYF
    if (sink) {                                         // the formatting time is reported on its own
        st->ms_call -= std::min<double>(ms, sink->ms);
        st->ms_format += sink->ms;
        st->events += sink->events;
        st->clusters_out += sink->clusters;
        delete calls;
        return E2I_OK;
    }
    *out = calls;
    return E2I_OK;
}

extern "C" int e2i_call(e2i_ctx *ctx, const e2i_index *b1, const e2i_index *b2, const e2i_bits *da,
                        const e2i_lcpbits *l, const e2i_params *p, uint64_t pos_begin, uint64_t pos_end,
                        e2i_calls **out, e2i_stats *st) {
    if (!out) { set_error("e2i_call: null argument"); return E2I_ERR_ARG; }
    return call_impl(ctx, b1, b2, da, l, p, pos_begin, pos_end, out, nullptr, false, st);
}

extern "C" int e2i_call_device(e2i_ctx *ctx, const e2i_index *b1, const e2i_index *b2, const e2i_bits *da,
                               const e2i_lcpbits *l, const e2i_params *p, uint64_t pos_begin, uint64_t pos_end,
                               e2i_calls **out, e2i_stats *st) {
    if (!out) { set_error("e2i_call_device: null argument"); return E2I_ERR_ARG; }
    return call_impl(ctx, b1, b2, da, l, p, pos_begin, pos_end, out, nullptr, true, st);
}

static int calls_format(const e2i_calls *c, const e2i_params *p, uint64_t first, bool want_text, char **d_text, uint64_t *len,
                        uint64_t *clusters, uint64_t *events) {
    if (!c || !p) { set_error("e2i_calls_snp: null argument"); return E2I_ERR_ARG; }
    if (c->n && !c->d_recs) { set_error("e2i_calls_snp: the records are not on the device (use e2i_call_device)"); return E2I_ERR_ARG; }
    if (p->k_left != c->k_left || p->k_right != c->k_right) { set_error("e2i_calls_snp: k_left / k_right differ from the ones the records were made with"); return E2I_ERR_ARG; }
    E2I_CUDA_TRY(cudaSetDevice(c->ctx->device));
    return format_device(c->ctx, c->d_recs, c->d_left, c->d_right, c->n, p, c->two_samples, first ? first : 1, d_text, len, clusters, events, want_text);
}

extern "C" int e2i_calls_clusters(const e2i_calls *c, const e2i_params *p, uint64_t *clusters) {
    if (!clusters) { set_error("e2i_calls_clusters: null argument"); return E2I_ERR_ARG; }
    char *d_text = nullptr;
    uint64_t len = 0, ev = 0;
    return calls_format(c, p, 1, false, &d_text, &len, clusters, &ev);
}

extern "C" int e2i_calls_snp_device(const e2i_calls *c, const e2i_params *p, uint64_t first_cluster_nr, void **dev_text, uint64_t *len,
                                    e2i_stats *st) {
    if (!c || !dev_text || !len) { set_error("e2i_calls_snp_device: null argument"); return E2I_ERR_ARG; }
    char *d_text = nullptr;
    uint64_t cl = 0, ev = 0;
    Accounting acct(c->ctx, st);
    E2I_TRY(calls_format(c, p, first_cluster_nr, true, &d_text, len, &cl, &ev));
    E2I_CUDA_TRY(cudaStreamSynchronize(c->ctx->stream));       // the caller reads the text on a stream of its own
    *dev_text = d_text;
    if (st) { st->events += ev; st->clusters_out += cl; }
    return E2I_OK;
}

extern "C" void e2i_device_free(e2i_ctx *ctx, void *p) {
    if (ctx && p) { cudaSetDevice(ctx->device); dfree(ctx, p); }
}

extern "C" int e2i_calls_snp(const e2i_calls *c, const e2i_params *p, uint64_t first_cluster_nr, char **snp, size_t *snp_len, e2i_stats *st) {
    if (!snp || !snp_len) { set_error("e2i_calls_snp: null argument"); return E2I_ERR_ARG; }
    void *d_text = nullptr;
    uint64_t len = 0;
    E2I_TRY(e2i_calls_snp_device(c, p, first_cluster_nr, &d_text, &len, st));
    e2i_ctx *ctx = c->ctx;
    char *buf = text_alloc(len + 1);
    if (!buf) { dfree(ctx, d_text); set_error("e2i_calls_snp: cannot page-lock %llu bytes for the text", (unsigned long long)len); return E2I_ERR_MEMORY; }
    if (len) {
        cudaError_t e = cudaMemcpyAsync(buf, d_text, len, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { e2i_buffer_free(buf); dfree(ctx, d_text); set_error("CUDA error copying the .snp text: %s", cudaGetErrorString(e)); return E2I_ERR_CUDA; }
        ctx->n_d2h += len;
        if (st) st->d2h_bytes += len;
    }
    dfree(ctx, d_text);
    buf[len] = 0;
    *snp = buf;
    *snp_len = (size_t)len;
    return E2I_OK;
}

extern "C" int e2i_call_snp(e2i_ctx *ctx, const e2i_index *b1, const e2i_index *b2, const e2i_bits *da,
                            const e2i_lcpbits *l, const e2i_params *p, uint64_t pos_begin, uint64_t pos_end,
                            uint64_t first_cluster_nr, char **snp, size_t *snp_len, e2i_stats *st) {
    if (!snp || !snp_len) { set_error("e2i_call_snp: null argument"); return E2I_ERR_ARG; }
    TextSink sink;
    sink.next_cluster = first_cluster_nr ? first_cluster_nr : 1;
    const int rc = call_impl(ctx, b1, b2, da, l, p, pos_begin, pos_end, nullptr, &sink, false, st);
    if (rc != E2I_OK) { e2i_buffer_free(sink.buf); return rc; }
    if (!sink.buf) sink.buf = text_alloc(1);
    if (!sink.buf) { set_error("e2i_call_snp: out of host memory"); return E2I_ERR_MEMORY; }
    sink.buf[sink.len] = 0;
    *snp = sink.buf;
    *snp_len = sink.len;
    return E2I_OK;
}

extern "C" uint64_t e2i_calls_count(const e2i_calls *c) { return c ? c->n : 0; }

extern "C" int e2i_calls_fetch(const e2i_calls *c, e2i_call_rec *host_recs, char *host_left, char *host_right,
                               uint64_t cap, uint64_t *n) {
    if (!c || !n) { set_error("e2i_calls_fetch: null argument"); return E2I_ERR_ARG; }
    if (c->n && !c->recs) { set_error("e2i_calls_fetch: the records are on the device (e2i_call_device): format them with e2i_calls_snp"); return E2I_ERR_ARG; }
    if (c->gen != c->ctx->pinned_gen) { set_error("e2i_calls_fetch: stale handle (a later e2i_call on this context reused the staging buffer)"); return E2I_ERR_ARG; }
    const uint64_t k = std::min<uint64_t>(cap, c->n);
    if (k && (!host_recs || !host_left || !host_right)) { set_error("e2i_calls_fetch: null buffer"); return E2I_ERR_ARG; }
    if (k) {
        std::memcpy(host_recs, c->recs, k * sizeof(e2i_call_rec));
        std::memcpy(host_left, c->left, k * 8 * (size_t)c->k_left);
        std::memcpy(host_right, c->right, k * (size_t)c->k_right);
    }
    *n = k;
    return E2I_OK;
}

extern "C" int e2i_calls_view(const e2i_calls *c, const e2i_call_rec **recs, const char **left, const char **right, uint64_t *n) {
    if (!c || !recs || !left || !right || !n) { set_error("e2i_calls_view: null argument"); return E2I_ERR_ARG; }
    if (c->n && !c->recs) { set_error("e2i_calls_view: the records are on the device (e2i_call_device): format them with e2i_calls_snp"); return E2I_ERR_ARG; }
    if (c->gen != c->ctx->pinned_gen) { set_error("e2i_calls_view: stale handle (a later e2i_call on this context reused the staging buffer)"); return E2I_ERR_ARG; }
    *recs = c->recs; *left = c->left; *right = c->right; *n = c->n;
    return E2I_OK;
}

extern "C" void e2i_calls_free(e2i_calls *c) {
    if (c && c->d_block) { cudaSetDevice(c->ctx->device); dfree(c->ctx, c->d_block); }
    delete c;
}
