// common.cuh -- shared host/device definitions of libe2i (B200 / sm_100a).
//
// HBM layout of the rank-indexed BWT (replaces dna_string, /root/reference/internal/dna_string.hpp):
//   one 32-byte block per 64 symbols = 2 x uint4 -- exactly one DRAM sector:
//     [0] = { #A | #C << 16, #G | #T << 16, plane a bits 0-31, plane a bits 32-63 }
//     [1] = { plane b bits 0-31, plane b bits 32-63, TERM plane bits 0-31, TERM plane bits 32-63 }
//   symbol j of the block sits at bit j of each plane (A=00, C=01, G=10, T=11 as (b,a); TERM has its
//   TERM-plane bit set and 00 in the others).  The four 16-bit counters hold the occurrences BEFORE
//   THE MIDDLE of the block (position 32), relative to the start of its 2^16-symbol superblock, so a
//   rank query popcounts one masked 32-bit word per plane -- upwards or downwards from the middle --
//   instead of up to four: 4 POPC per query (POPC is a quarter-rate instruction and was the limiter
//   of the 64-byte / 128-symbol layout of round 1).  A superblock table (4 x u64 absolute counts
//   per 2^16 symbols, 0.05 % of the index) lifts the counters to 64-bit positions.  Same 4 bits per
//   symbol as the reference's dna_string; one rank query = one aligned 32-byte sector + 4 popcounts.
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

constexpr int kBlockSyms = 64;           // symbols per index block
constexpr int kBlockShift = 6;
constexpr int kBlockU4 = 2;              // uint4 per block
constexpr int kSuperShift = 16;          // symbols per superblock = 2^16
constexpr int kTileSyms = 16384;         // symbols per index-build tile (256 blocks)
constexpr int kTileShift = 14;
constexpr int kTileBlocks = kTileSyms / kBlockSyms;
constexpr int kSuperTileShift = kSuperShift - kTileShift;

struct DevIndex {
    const uint4 *blocks;      // 2 x uint4 per block
    const uint64_t *super;    // 4 x u64 per superblock (absolute counts at its start)
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
    bool arena_ipc = false;     // the arena is a plain cudaMalloc block (exportable by CUDA IPC), not a pool block
    // look-back descriptors shared by all ordered-compaction kernels
    unsigned long long *desc = nullptr;
    size_t desc_words = 0;
    uint32_t epoch = 0;
    void *ctl = nullptr;        // device ring of per-sweep control blocks (64 bytes each), see navigate.cu
    void *ctl_host = nullptr;   // page-locked, device-mapped block the sweeps report their counts to
    uint32_t ticket_next = 0;
    unsigned long long sweep_seq = 0;
    void *ring[4] = {};         // page-locked ring of the streaming file ingest (index.cu)
    float last_h2d_ms = 0;      // copy-stream time of the last streamed upload
    void *pinned = nullptr;     // page-locked staging of the call records
    size_t pinned_bytes = 0;
    uint64_t pinned_gen = 0;
    // accounting (kernels launched, bytes copied) since the context was created
    uint64_t n_launch = 0, n_h2d = 0, n_d2h = 0;
};

// Ranks of a multi-GPU run (threads of one process, or processes of one box): a barrier and one 4 KB publish
// slot per rank that every rank can read (multi.cu).  Device memory of a peer is addressed through
// peer_ptr(): the pointer itself between threads of one process, a CUDA IPC mapping between processes.
struct e2i_comm {
    int rank = 0, world = 1;
    virtual ~e2i_comm() {}
    virtual void barrier() = 0;
    virtual unsigned char *slot(int r) = 0;                     // 4 KB, written by rank r only, read after a barrier
    // this process's view of `base`, a device allocation of rank r whose IPC handle is `handle`
    virtual void *peer_ptr(int r, void *base, const cudaIpcMemHandle_t &handle) = 0;
    virtual bool needs_ipc() const = 0;
};
constexpr size_t kCommSlotBytes = 4096;

namespace e2i {
template <typename T>
inline cudaError_t dmalloc(e2i_ctx *ctx, T **p, size_t bytes) {
    return cudaMallocFromPoolAsync(reinterpret_cast<void **>(p), bytes ? bytes : 16, ctx->pool, ctx->stream);
}
inline void dfree(e2i_ctx *ctx, void *p) {
    if (p) cudaFreeAsync(p, ctx->stream);
}
// the frame arena: a pool block normally, a plain cudaMalloc block when peers of other processes must map it
inline void arena_release(e2i_ctx *ctx) {
    if (ctx->arena_mem) {
        if (ctx->arena_ipc) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->arena_mem); }
        else dfree(ctx, ctx->arena_mem);
    }
    ctx->arena_mem = nullptr;
    ctx->arena_bytes = 0;
    ctx->arena_ipc = false;
}
inline cudaError_t arena_alloc(e2i_ctx *ctx, size_t bytes, bool ipc) {
    arena_release(ctx);
    const cudaError_t e = ipc ? cudaMalloc(&ctx->arena_mem, bytes) : cudaMallocFromPoolAsync(&ctx->arena_mem, bytes, ctx->pool, ctx->stream);
    if (e == cudaSuccess) { ctx->arena_bytes = bytes; ctx->arena_ipc = ipc; }
    return e;
}
// Host buffers for the .snp text handed to the caller: page-locked (the text arrives by one DMA copy, no page
// faults, no staging) and cached process-wide between calls; e2i_buffer_free gives them back (snp_format.cpp).
char *text_alloc(size_t bytes);          // ordinary malloc memory when nothing can be page-locked any more; nullptr: out of memory
bool text_release(void *p);              // false: not one of ours
void text_cache_trim();                  // frees the cached buffers that are not handed out
// .snp text of call records that are still in device memory (snp_format.cpp); *d_text is released with dfree
int format_device(e2i_ctx *ctx, const e2i_call_rec *d_recs, const char *d_left, const char *d_right, uint64_t n_recs,
                  const e2i_params *p, int two_samples, uint64_t first_cluster_nr,
                  char **d_text, uint64_t *text_len, uint64_t *clusters, uint64_t *events, bool want_text = true);
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
    uint4 *blocks = nullptr;      // kBlockU4 x uint4 per block
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
    e2i_call_rec *recs = nullptr;                       // page-locked host arrays (e2i_call) ...
    char *left = nullptr, *right = nullptr;
    uint64_t n = 0, gen = 0;
    int k_left = 0, k_right = 0;
    // ... or device arrays (e2i_call_device): [recs | left | right] in one pool block with room for d_cap records
    char *d_block = nullptr;
    e2i_call_rec *d_recs = nullptr;
    char *d_left = nullptr, *d_right = nullptr;
    uint64_t d_cap = 0;
    int two_samples = 0;
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

// #A,#C,#G,#T before offset `off` (0..63) of a block, relative to the block's superblock: the counters
// hold the counts before the middle, the masked half-word is added (off >= 32) or subtracted
__device__ __forceinline__ void block_rank(const uint4 lo, const uint4 hi, int off, uint32_t out[4]) {
    const bool up = off >= 32;
    const uint32_t a = up ? lo.w : lo.z, b = up ? hi.y : hi.x, t = up ? hi.w : hi.z;
    const uint32_t lm = low_mask(off & 31);               // the bits of the half that lie before `off`
    const uint32_t nt = ~t & (up ? lm : ~lm);             // non-terminators between the middle and `off`
    const uint32_t pn = __popc(nt), pa = __popc(nt & a), pb = __popc(nt & b), pab = __popc(nt & a & b);
    const uint32_t v[4] = {pn - pa - pb + pab, pa - pab, pb - pab, pab};
    const uint32_t c[4] = {lo.x & 0xffffu, lo.x >> 16, lo.y & 0xffffu, lo.y >> 16};
#pragma unroll
    for (int k = 0; k < 4; ++k) out[k] = up ? c[k] + v[k] : c[k] - v[k];
}

// Index blocks staged in shared memory: slot r lives at stage[2r], stage[2r+1] with its two 16-byte
// halves swapped when bit 2 of r is set, so that lanes reading the same half of random slots spread
// over all eight 16-byte bank groups.
__device__ __forceinline__ int stage_slot(uint32_t r, int h) { return (int)(r * 2 + (h ^ ((r >> 2) & 1))); }

// a2: parallel_rank (dna_string.hpp:140-152): #A,#C,#G,#T in [0, i), 0 <= i <= n, read from HBM
__device__ __forceinline__ void rank4(const DevIndex &ix, uint64_t i, uint64_t out[4]) {
    const uint4 *p = ix.blocks + (i >> kBlockShift) * kBlockU4;
    const uint4 lo = __ldg(p), hi = __ldg(p + 1);
    uint32_t r[4];
    block_rank(lo, hi, (int)((uint32_t)i & (kBlockSyms - 1)), r);
    const ulonglong2 *sb = reinterpret_cast<const ulonglong2 *>(ix.super + (i >> kSuperShift) * 4);
    const ulonglong2 s0 = __ldg(sb), s1 = __ldg(sb + 1);
    out[0] = s0.x + r[0]; out[1] = s0.y + r[1]; out[2] = s1.x + r[2]; out[3] = s1.y + r[3];
}

// 16-byte asynchronous global -> shared copy (LDGSTS), L2-only caching
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// occurrences of symbol c before the MIDDLE of block `blk` (absolute)
__device__ __forceinline__ uint64_t mid_count(const DevIndex &ix, uint64_t blk, int c) {
    const uint16_t *cnt = reinterpret_cast<const uint16_t *>(ix.blocks + blk * kBlockU4);
    return ix.super[(blk >> (kSuperShift - kBlockShift)) * 4 + c] + __ldg(cnt + c);
}

// a3: operator[] (dna_string.hpp:113-135): code 0..3 = A,C,G,T, 4 = TERM
__device__ __forceinline__ int access_code(const DevIndex &ix, uint64_t i) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(ix.blocks + (i >> kBlockShift) * kBlockU4);
    const int off = (int)(i & (kBlockSyms - 1)), k = off >> 5, sh = off & 31;
    const uint32_t a = (__ldg(w + 2 + k) >> sh) & 1u, b = (__ldg(w + 4 + k) >> sh) & 1u, t = (__ldg(w + 6 + k) >> sh) & 1u;
    return t ? 4 : (int)(a | (b << 1));
}

// F column (dna_bwt.hpp:100-110): code of the first symbol of suffix i
__device__ __forceinline__ int f_code(const DevIndex &ix, uint64_t i) {
    return i < ix.F[0] ? 4 : i < ix.F[1] ? 0 : i < ix.F[2] ? 1 : i < ix.F[3] ? 2 : 3;
}

// the 32 symbols of half h (0 = first, 1 = second) of block blk that equal c, as a bit mask
__device__ __forceinline__ uint32_t half_mask(const DevIndex &ix, uint64_t blk, int h, int c) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(ix.blocks + blk * kBlockU4);
    const uint32_t a = __ldg(w + 2 + h), b = __ldg(w + 4 + h), t = __ldg(w + 6 + h);
    return ~t & ((c & 1) ? a : ~a) & ((c & 2) ? b : ~b);
}

// a4: select (dna_string.hpp:182-188, 254-272): position of the r-th (0-based) symbol c.
// Binary search on the mid-block counters (last block whose middle has at most r occurrences before
// it), then a bit select in the 64 symbols from that middle to the next one.
__device__ __forceinline__ uint64_t select_sym(const DevIndex &ix, uint64_t r, int c) {
    if (mid_count(ix, 0, c) > r) {                        // inside the first half of block 0
        const uint32_t m = half_mask(ix, 0, 0, c);
        return (uint64_t)__fns(m, 0, (int)r + 1);
    }
    uint64_t lo = 0, hi = ix.n >> kBlockShift;            // last block with mid_count <= r lies in [lo, hi]
    while (lo < hi) {
        const uint64_t mid = (lo + hi + 1) >> 1;
        if (mid_count(ix, mid, c) <= r) lo = mid; else hi = mid - 1;
    }
    uint32_t k = (uint32_t)(r - mid_count(ix, lo, c));    // k-th occurrence at or after the middle of block lo
    const uint32_t m1 = half_mask(ix, lo, 1, c);
    const uint32_t p1 = __popc(m1);
    if (k < p1) return (lo << kBlockShift) + 32 + __fns(m1, 0, (int)k + 1);
    k -= p1;
    const uint32_t m2 = half_mask(ix, lo + 1, 0, c);
    return ((lo + 1) << kBlockShift) + __fns(m2, 0, (int)k + 1);   // exists for r < #c
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
