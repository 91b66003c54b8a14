// common.cuh -- shared host/device definitions of libe2i (B200 / sm_100a).
//
// HBM layout of the rank-indexed BWT (replaces dna_string, /root/reference/internal/dna_string.hpp):
//   one 64-byte block per 128 symbols = 4 x uint4:
//     [0] four u32 counts of A,C,G,T before the block, relative to the 2^32-symbol superblock
//     [1] plane 0 (bit 0 of the code)   [2] plane 1 (bit 1 of the code)   [3] TERM plane
//   symbol j of the block sits at bit j%32 of word j/32 of each plane (A=00, C=01, G=10, T=11;
//   TERM has its TERM-plane bit set and 00 in the others).  A superblock table (4 x u64 absolute
//   counts per 2^32 symbols) lifts the u32 counters to 64-bit positions.  One rank query =
//   one aligned 64-byte fetch (two 32-byte sectors) + 16 popcounts.
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "e2i.h"

namespace e2i {

constexpr int kBlockSyms = 128;          // symbols per index block
constexpr int kBlockShift = 7;
constexpr int kSuperShift = 32;          // symbols per superblock = 2^32
constexpr int kTileSyms = 16384;         // symbols per index-build tile (128 blocks)
constexpr int kTileShift = 14;
constexpr int kSuperTileShift = kSuperShift - kTileShift;

struct DevIndex {
    const uint4 *blocks;      // 4 x uint4 per block
    const uint64_t *super;    // 4 x u64 per superblock
    uint64_t n;
    uint64_t F[4];            // F_A, F_C, F_G, F_T (dna_bwt.hpp:412-415)
};

void set_error(const char *fmt, ...);

#define E2I_CUDA_TRY(expr)                                                                   \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            e2i::set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__,     \
                           __LINE__, cudaGetErrorString(_e));                                \
            return E2I_ERR_CUDA;                                                             \
        }                                                                                    \
    } while (0)

#define E2I_TRY(expr)                 \
    do {                              \
        int _rc = (expr);             \
        if (_rc != E2I_OK) return _rc; \
    } while (0)

// Frontier frames live in ONE device arena used as a double-ended stack: frames of even tree
// depth grow up from the bottom, frames of odd depth grow down from the top.  A frame of depth d is
// released either right after the sweep that produced depth d+1 (level-synchronous case) or after
// the whole subtree below it is finished (chunked depth-first case); both are LIFO per end, so
// allocation is a pointer bump: no driver call and no fragmentation on the hot path.
class Arena {
  public:
    void reset(char *base, size_t bytes) {
        bytes &= ~(size_t)255;                          // both ends hand out 256-byte aligned frames
        base_ = base; size_ = bytes; lo_ = 0; hi_ = bytes; live_[0].clear(); live_[1].clear();
    }
    char *base() const { return base_; }
    size_t size() const { return size_; }
    size_t in_use() const { return lo_ + (size_ - hi_); }
    void *alloc(int side, size_t bytes);            // nullptr when it does not fit
    void free(int side, void *p);

  private:
    struct Blk { size_t off, bytes; bool freed; };
    char *base_ = nullptr;
    size_t size_ = 0, lo_ = 0, hi_ = 0;
    std::vector<Blk> live_[2];
};

}  // namespace e2i

struct e2i_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaMemPool_t pool = nullptr;  // the context's own stream-ordered pool (freed blocks stay cached until e2i_trim)
    cudaEvent_t ev[8] = {};
    int sm_count = 148;
    uint64_t frontier_budget = 0;
    e2i::Arena arena;           // frontier frames (device memory of `arena_mem`)
    void *arena_mem = nullptr;
    size_t arena_bytes = 0;
    // look-back descriptors shared by all ordered-compaction kernels
    unsigned long long *desc = nullptr;
    size_t desc_words = 0;
    uint32_t epoch = 0;
    void *ctl = nullptr;        // device ring of per-sweep control blocks (64 bytes each), see navigate.cu
    void *ctl_host = nullptr;   // page-locked, device-mapped block the sweeps report their counts to
    uint32_t ticket_next = 0;
    unsigned long long sweep_seq = 0;
    void *pinned = nullptr;     // page-locked staging of the call records
    size_t pinned_bytes = 0;
    uint64_t pinned_gen = 0;
    // accounting (kernels launched, bytes copied) since the context was created
    uint64_t n_launch = 0, n_h2d = 0, n_d2h = 0;
};

namespace e2i {
template <typename T>
inline cudaError_t dmalloc(e2i_ctx *ctx, T **p, size_t bytes) {
    return cudaMallocFromPoolAsync(reinterpret_cast<void **>(p), bytes ? bytes : 16, ctx->pool, ctx->stream);
}
inline void dfree(e2i_ctx *ctx, void *p) {
    if (p) cudaFreeAsync(p, ctx->stream);
}
struct Accounting {            // adds what a call launched / copied to its e2i_stats on scope exit
    e2i_ctx *ctx; e2i_stats *st; uint64_t l0, h0, d0;
    Accounting(e2i_ctx *c, e2i_stats *s) : ctx(c), st(s), l0(c->n_launch), h0(c->n_h2d), d0(c->n_d2h) {}
    ~Accounting() {
        if (!st) return;
        st->kernel_launches += ctx->n_launch - l0;
        st->h2d_bytes += ctx->n_h2d - h0;
        st->d2h_bytes += ctx->n_d2h - d0;
    }
};
}  // namespace e2i

struct e2i_index {
    e2i_ctx *ctx = nullptr;
    uint4 *blocks = nullptr;
    uint64_t *super = nullptr;
    uint64_t n = 0, n_blocks = 0, n_super = 0, bytes = 0;
    uint64_t F[4] = {0, 0, 0, 0};
    uint8_t term = '#';
    // slice-wise construction state (e2i_index_slice_*)
    void *slice_cnt = nullptr, *slice_prefix = nullptr;
    uint64_t slice_begin = 0, slice_len = 0, slice_tiles = 0;
    e2i::DevIndex dev() const {
        e2i::DevIndex d;
        d.blocks = blocks;
        d.super = super;
        d.n = n;
        for (int i = 0; i < 4; ++i) d.F[i] = F[i];
        return d;
    }
};

struct e2i_bits {
    e2i_ctx *ctx = nullptr;
    uint32_t *words = nullptr;    // packed bits, zero-padded to a multiple of 64 u32 words
    uint64_t n = 0, n_words32 = 0;
    uint64_t *rank512 = nullptr;  // popcount before every 512-bit group (built on demand, mode -2)
};

struct e2i_lcpbits {
    e2i_ctx *ctx = nullptr;
    uint32_t *thr = nullptr;      // 2n bits: bit 2i = LCP[i] >= K, bit 2i+1 = LCP[i] >= k_right
    uint32_t *minima = nullptr;   // n bits
    uint64_t n = 0, thr_words32 = 0, min_words32 = 0;
};

// Call records live in the context's page-locked staging buffer (one D2H per array, no pageable
// bounce): a handle is valid until the next e2i_call on the same context (checked by `gen`).
struct e2i_calls {
    e2i_ctx *ctx = nullptr;
    e2i_call_rec *recs = nullptr;
    char *left = nullptr, *right = nullptr;
    uint64_t n = 0, gen = 0;
    int k_left = 0, k_right = 0;
};

#ifdef __CUDACC__
namespace e2i {

// lowest t bits set, t clamped to [0, 32]
__device__ __forceinline__ uint32_t low_mask(int t) {
    uint32_t m;
    const int tt = t < 0 ? 0 : t;
    asm("bmsk.clamp.b32 %0, 0, %1;" : "=r"(m) : "r"(tt));
    return m;
}

__device__ __forceinline__ uint32_t prefix_mask32(int off, int word) {
    // bits of `word` (32 symbols) that lie before block offset `off`
    return low_mask(off - 32 * word);
}

// #A,#C,#G,#T among the first `off` symbols of a block given its three planes
__device__ __forceinline__ void block_popc(const uint4 a, const uint4 b, const uint4 t, int off, uint32_t out[4]) {
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w}, tw[4] = {t.x, t.y, t.z, t.w};
    uint32_t nN = 0, nC = 0, nG = 0, nT = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t nt = ~tw[k] & low_mask(off - 32 * k);
        nN += __popc(nt);
        nC += __popc(nt & aw[k]);
        nG += __popc(nt & bw[k]);
        nT += __popc(nt & aw[k] & bw[k]);
    }
    nC -= nT;
    nG -= nT;
    out[0] = nN - nC - nG - nT;
    out[1] = nC;
    out[2] = nG;
    out[3] = nT;
}

// Index blocks staged in shared memory: block r of the window lives at stage[4r .. 4r+3] with its
// four 16-byte quarters XOR-swizzled by (r >> 1) & 3, so that lanes reading the same quarter of
// eight consecutive blocks hit eight different bank groups.
__device__ __forceinline__ int stage_slot(uint32_t r, int q) { return (int)(r * 4 + (q ^ ((r >> 1) & 3))); }

// a2: parallel_rank (dna_string.hpp:140-152): #A,#C,#G,#T in [0, i), 0 <= i <= n, in W arithmetic:
// W = uint64_t gives absolute counts; W = uint32_t gives them modulo 2^32, which is all that a
// DIFFERENCE of two ranks less than 2^32 apart needs (half the integer work of the 64-bit form).
// Blocks [blk_lo, blk_lo + n_staged) are read from the shared-memory window, the others from HBM.
template <typename W>
__device__ __forceinline__ void rank4w(const DevIndex &ix, const uint4 *stage, uint32_t blk_lo, uint32_t n_staged,
                                       uint64_t i, W out[4]) {
    const uint32_t blk = (uint32_t)(i >> kBlockShift);      // n < 2^39
    const uint32_t rel = blk - blk_lo;                       // wraps to a huge value below the window
    uint4 cnt, a, b, t;
    if (rel < n_staged) {
        cnt = stage[stage_slot(rel, 0)]; a = stage[stage_slot(rel, 1)]; b = stage[stage_slot(rel, 2)]; t = stage[stage_slot(rel, 3)];
    } else {
        const uint4 *p = ix.blocks + (size_t)blk * 4;
        cnt = __ldg(p); a = __ldg(p + 1); b = __ldg(p + 2); t = __ldg(p + 3);
    }
    uint32_t pc[4];
    block_popc(a, b, t, (int)((uint32_t)i & (kBlockSyms - 1)), pc);
    out[0] = (W)cnt.x + pc[0];
    out[1] = (W)cnt.y + pc[1];
    out[2] = (W)cnt.z + pc[2];
    out[3] = (W)cnt.w + pc[3];
    if (ix.n >> kSuperShift) {                               // more than one superblock: add its base counts
        const uint64_t *sb = ix.super + (i >> kSuperShift) * 4;
        out[0] += (W)sb[0]; out[1] += (W)sb[1]; out[2] += (W)sb[2]; out[3] += (W)sb[3];
    }
}

__device__ __forceinline__ void rank4(const DevIndex &ix, uint64_t i, uint64_t out[4]) {
    rank4w<uint64_t>(ix, nullptr, 0u, 0u, i, out);
}

// 16-byte asynchronous global -> shared copy (LDGSTS), L2-only caching
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// single-symbol count before block `blk` (absolute)
__device__ __forceinline__ uint64_t block_count(const DevIndex &ix, uint64_t blk, int c) {
    const uint32_t *cnt = reinterpret_cast<const uint32_t *>(ix.blocks + blk * 4);
    return ix.super[(blk >> (kSuperShift - kBlockShift)) * 4 + c] + __ldg(cnt + c);
}

// a3: operator[] (dna_string.hpp:113-135): code 0..3 = A,C,G,T, 4 = TERM
__device__ __forceinline__ int access_code(const DevIndex &ix, uint64_t i) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(ix.blocks + (i >> kBlockShift) * 4);
    const int off = (int)(i & (kBlockSyms - 1)), k = off >> 5, sh = off & 31;
    const uint32_t a = (__ldg(w + 4 + k) >> sh) & 1u, b = (__ldg(w + 8 + k) >> sh) & 1u, t = (__ldg(w + 12 + k) >> sh) & 1u;
    return t ? 4 : (int)(a | (b << 1));
}

// F column (dna_bwt.hpp:100-110): code of the first symbol of suffix i
__device__ __forceinline__ int f_code(const DevIndex &ix, uint64_t i) {
    return i < ix.F[0] ? 4 : i < ix.F[1] ? 0 : i < ix.F[2] ? 1 : i < ix.F[3] ? 2 : 3;
}

// a4: select (dna_string.hpp:182-188, 254-272): position of the r-th (0-based) symbol c.
// Binary search on the interleaved block counters, then a bit select inside the block.
__device__ __forceinline__ uint64_t select_sym(const DevIndex &ix, uint64_t r, int c) {
    uint64_t lo = 0, hi = ix.n >> kBlockShift;  // last block whose count <= r lies in [lo, hi]
    while (lo < hi) {
        const uint64_t mid = (lo + hi + 1) >> 1;
        if (block_count(ix, mid, c) <= r) lo = mid; else hi = mid - 1;
    }
    uint32_t k = (uint32_t)(r - block_count(ix, lo, c));  // k-th occurrence inside block lo
    const uint4 *p = ix.blocks + lo * 4;
    const uint4 a = __ldg(p + 1), b = __ldg(p + 2), t = __ldg(p + 3);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w}, tw[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        const uint32_t m = ~tw[w] & ((c & 1) ? aw[w] : ~aw[w]) & ((c & 2) ? bw[w] : ~bw[w]);
        const uint32_t pc = __popc(m);
        if (k < pc) return (lo << kBlockShift) + 32 * w + __fns(m, 0, (int)k + 1);
        k -= pc;
    }
    return ~0ull;  // unreachable for r < #c
}

// FL (dna_bwt.hpp:115-133); the caller guarantees F(i) != TERM
__device__ __forceinline__ uint64_t fl_map(const DevIndex &ix, uint64_t i, int c) {
    return select_sym(ix, i - ix.F[c], c);
}

__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src, int width = 32) {
    return __shfl_sync(0xffffffffu, v, src, width);
}
__device__ __forceinline__ uint64_t shfl_down_u64(uint64_t v, int d, int width = 32) {
    return __shfl_down_sync(0xffffffffu, v, d, width);
}
__device__ __forceinline__ uint64_t shfl_up_u64(uint64_t v, int d, int width = 32) {
    return __shfl_up_sync(0xffffffffu, v, d, width);
}

}  // namespace e2i
#endif
