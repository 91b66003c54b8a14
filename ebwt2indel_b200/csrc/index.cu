// index.cu -- a1/a5: ASCII BWT -> rank-indexed blocks in HBM, and the batched test hooks.
//
// Replaces dna_string(path, TERM) + build_rank_support (/root/reference/internal/dna_string.hpp:
// 55-110, 275-315, 320-369) and the F-array scan of dna_bwt(path, TERM) (dna_bwt.hpp:36-62).
// Three streaming kernels: count symbols per 16384-symbol tile, scan the tile totals, pack
// (second read of the ASCII, one write of the 32-byte blocks).  HBM traffic: 2n read + n/2 write.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <thread>

#include "common.cuh"

namespace e2i {

// ---- symbol classification -------------------------------------------------------------------
// code: 0..3 = A,C,G,T; 4 = TERM; 5 = forbidden
__device__ __forceinline__ int classify(uint32_t ch, uint32_t term) {
    if (ch == term) return 4;
    if (ch == 'A') return 0;
    if (ch == 'C') return 1;
    if (ch == 'G') return 2;
    if (ch == 'T') return 3;
    return 5;
}

struct Piece {               // 16 symbols
    uint32_t p0, p1, pt;     // 16 plane bits each
    uint32_t cnt;            // 4 x 8-bit counts A,C,G,T
    int bad;                 // index (0..15) of the first forbidden symbol, or -1
};

// bit k of the result = least significant bit of byte k of x (x holds 0/1 per byte)
__device__ __forceinline__ uint32_t gather4(uint32_t x) { return ((x & 0x01010101u) * 0x01020408u) >> 24; }

// 0x80 in every byte of the result whose byte of v is zero (exact, no borrow between bytes)
__device__ __forceinline__ uint32_t zero_bytes(uint32_t v) { return ~(((v & 0x7f7f7f7fu) + 0x7f7f7f7fu) | v) & 0x80808080u; }

// 16 symbols at once, four per 32-bit word, without per-byte compares: the 2-bit code of A,C,G,T is read off bits
// 1-2 of the ASCII byte (A 0x41, C 0x43, G 0x47, T 0x54: y = bits 2..1 = 0,1,3,2; code = y ^ (y >> 1)), the byte that
// code stands for is rebuilt and compared with the input in one XOR, and the terminator test (which comes first, as
// in the reference: a terminator that is one of A,C,G,T stays a terminator) is one more.  ~35 integer operations per
// word against ~100 for five __vcmpeq4 (emulated on this architecture): the build kernels were issue-bound on them.
template <bool TAIL>
__device__ __forceinline__ Piece encode_words(const uint32_t (&w)[4], int valid, uint32_t t4) {
    Piece pc{0, 0, 0, 0, -1};
    uint32_t notok[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint32_t live = 0x80808080u;                                   // 0x80 per in-range symbol of this word
        if (TAIL) {
            const int k = valid - 4 * q;
            if (k < 4) live = k <= 0 ? 0u : (0x80808080u >> (8 * (4 - k)));
        }
        const uint32_t y = (w[q] >> 1) & 0x03030303u;
        const uint32_t hi = (y >> 1) & 0x01010101u, lo = (y ^ hi) & 0x01010101u;       // code = lo | hi << 1
        const uint32_t expect = 0x41414141u + lo * 2u + hi * 6u + (lo & hi) * 11u;      // +2 C, +6 G, +19 T
        const uint32_t isX = zero_bytes(w[q] ^ t4) & live;
        const uint32_t base = zero_bytes(w[q] ^ expect) & live & ~isX;                  // A,C,G,T that are not the terminator
        notok[q] = live & ~(base | isX);                                                // 0x80 per forbidden symbol
        const uint32_t b1 = base >> 7;                                                  // 0x01 per byte
        pc.p0 |= ((( lo & b1) * 0x01020408u) >> 24) << (4 * q);
        pc.p1 |= ((( hi & b1) * 0x01020408u) >> 24) << (4 * q);
        // positions past the end count as terminators: never as A
        pc.pt |= ((((isX | (TAIL ? ~live & 0x80808080u : 0u)) >> 7) * 0x01020408u) >> 24) << (4 * q);
    }
    if (notok[0] | notok[1] | notok[2] | notok[3]) {                   // rare: where is the first forbidden symbol
        for (int q = 3; q >= 0; --q) if (notok[q]) pc.bad = 4 * q + ((__ffs((int)notok[q]) - 1) >> 3);
    }
    // counts from the planes: outside A,C,G,T both plane bits are zero and the terminator bit tells them from A
    const uint32_t a = ~(pc.p0 | pc.p1 | pc.pt) & 0xffffu;
    pc.cnt = (uint32_t)__popc(a) | ((uint32_t)__popc(pc.p0 & ~pc.p1) << 8) | ((uint32_t)__popc(pc.p1 & ~pc.p0) << 16) |
             ((uint32_t)__popc(pc.p0 & pc.p1) << 24);
    return pc;
}

__device__ __noinline__ Piece encode_tail(uint4 v, int valid, uint32_t t4) {      // the last piece of the string: out of line
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    return encode_words<true>(w, valid, t4);
}

__device__ __forceinline__ Piece encode_piece(uint4 v, uint64_t pos0, uint64_t n, uint32_t term) {
    const uint32_t t4 = term * 0x01010101u;
    if (pos0 + 16 > n) return encode_tail(v, (int)(n - pos0), t4);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    return encode_words<false>(w, 16, t4);
}

__device__ __noinline__ uint4 load_tail(const uint8_t *ascii, uint64_t pos0, uint64_t n) {
    uint32_t w[4] = {0, 0, 0, 0};
    for (int j = 0; j < 16; ++j)
        if (pos0 + j < n) w[j >> 2] |= (uint32_t)ascii[pos0 + j] << (8 * (j & 3));
    return make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ uint4 load_piece(const uint8_t *ascii, uint64_t pos0, uint64_t n) {
    // 16-byte aligned vector load when the whole piece is inside the buffer, bytewise tail otherwise
    if (pos0 + 16 <= n) return __ldg(reinterpret_cast<const uint4 *>(ascii + pos0));
    return load_tail(ascii, pos0, n);
}

constexpr int kBuildThreads = 256;
constexpr int kPiecesPerTile = kTileSyms / 16;                       // 1024
constexpr int kPiecesPerThread = kPiecesPerTile / kBuildThreads;     // 4

// Kernel 1: per-tile symbol totals.  Grid-stride over tiles; coalesced 16-byte loads.
__global__ void __launch_bounds__(kBuildThreads)
count_tiles_kernel(const uint8_t *__restrict__ ascii, uint64_t n, uint32_t term, uint64_t tile0, uint64_t n_tiles,
                   uint4 *__restrict__ tile_cnt, unsigned long long *bad_pos) {
    // `ascii` points at global position tile0 << kTileShift; tiles tile0 .. tile0 + n_tiles - 1 are counted
    ascii -= tile0 << kTileShift;
    __shared__ unsigned long long s_part[kBuildThreads / 32];
    for (uint64_t lt = blockIdx.x; lt < n_tiles; lt += gridDim.x) {
        const uint64_t tile = tile0 + lt;
        unsigned long long acc = 0;  // 4 x 16-bit fields
#pragma unroll
        for (int it = 0; it < kPiecesPerThread; ++it) {
            const uint64_t pos0 = (tile << kTileShift) + (uint64_t)(it * kBuildThreads + threadIdx.x) * 16;
            if (pos0 < n) {
                const Piece pc = encode_piece(load_piece(ascii, pos0, n), pos0, n, term);
                if (pc.bad >= 0) atomicMin(bad_pos, (unsigned long long)(pos0 + pc.bad));
                acc += (unsigned long long)(pc.cnt & 0xffu) | ((unsigned long long)((pc.cnt >> 8) & 0xffu) << 16) |
                       ((unsigned long long)((pc.cnt >> 16) & 0xffu) << 32) | ((unsigned long long)(pc.cnt >> 24) << 48);
            }
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long t = 0;
            for (int w = 0; w < kBuildThreads / 32; ++w) t += s_part[w];
            tile_cnt[lt] = make_uint4((uint32_t)(t & 0xffff), (uint32_t)((t >> 16) & 0xffff),
                                        (uint32_t)((t >> 32) & 0xffff), (uint32_t)(t >> 48));
        }
        __syncthreads();
    }
}

// Kernel 2: exclusive scan of the tile totals (one CTA, coalesced chunks of 1024 tiles).
constexpr int kScanThreads = 1024;
__global__ void __launch_bounds__(kScanThreads)
scan_tiles_kernel(const uint4 *__restrict__ tile_cnt, uint64_t n_tiles, ulonglong4 *__restrict__ tile_prefix,
                  unsigned long long *__restrict__ totals) {
    __shared__ unsigned long long s_warp[32][4];
    __shared__ unsigned long long s_carry[4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 4) s_carry[threadIdx.x] = 0;
    __syncthreads();
    for (uint64_t base = 0; base < n_tiles; base += kScanThreads) {
        const uint64_t t = base + threadIdx.x;
        uint4 c = t < n_tiles ? tile_cnt[t] : make_uint4(0, 0, 0, 0);
        unsigned long long v[4] = {c.x, c.y, c.z, c.w}, incl[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            unsigned long long x = v[k];
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                const unsigned long long y = __shfl_up_sync(0xffffffffu, x, s);
                if (lane >= s) x += y;
            }
            incl[k] = x;
            if (lane == 31) s_warp[warp][k] = x;
        }
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                unsigned long long x = s_warp[lane][k];
#pragma unroll
                for (int s = 1; s < 32; s <<= 1) {
                    const unsigned long long y = __shfl_up_sync(0xffffffffu, x, s);
                    if (lane >= s) x += y;
                }
                s_warp[lane][k] = x;  // inclusive over warps
            }
        }
        __syncthreads();
        unsigned long long ex[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            ex[k] = s_carry[k] + (warp ? s_warp[warp - 1][k] : 0ull) + incl[k] - v[k];
        if (t < n_tiles) tile_prefix[t] = make_ulonglong4(ex[0], ex[1], ex[2], ex[3]);
        __syncthreads();
        if (threadIdx.x < 4) s_carry[threadIdx.x] += s_warp[31][threadIdx.x];
        __syncthreads();
    }
    if (threadIdx.x < 4) totals[threadIdx.x] = s_carry[threadIdx.x];
}

// Packing of ONE tile (256 blocks of 64 symbols) by a CTA of 256 threads: encode 4 pieces per thread into shared
// memory, scan the per-block totals, write the 512 uint4 of the tile coalesced.  `rel` = counts of A,C,G,T before the
// tile relative to its superblock start (< 2^16 each).  Returns the tile's totals (4 x 16 bit, every thread).
struct TileSmem {
    uint16_t plane[3][kPiecesPerTile];
    uint32_t cnt[kPiecesPerTile];
    unsigned long long mid[kTileBlocks];                  // counts before the middle of each block, 4 x 16 bit, tile-relative
    unsigned long long wsum[kBuildThreads / 32];
};

__device__ __forceinline__ unsigned long long widen_counts(uint32_t c) {
    return (unsigned long long)(c & 0xffu) | ((unsigned long long)((c >> 8) & 0xffu) << 16) |
           ((unsigned long long)((c >> 16) & 0xffu) << 32) | ((unsigned long long)(c >> 24) << 48);
}

__device__ __forceinline__ unsigned long long pack_one_tile(TileSmem &sm, const uint8_t *__restrict__ ascii, uint64_t n, uint32_t term, uint64_t tile,
                                                            const uint32_t (&rel)[4], uint4 *__restrict__ blocks, unsigned long long *bad_pos) {
    static_assert(kTileBlocks == kBuildThreads, "one thread per block in the per-block scan");
#pragma unroll
    for (int it = 0; it < kPiecesPerThread; ++it) {
        const int piece = it * kBuildThreads + threadIdx.x;
        const uint64_t pos0 = (tile << kTileShift) + (uint64_t)piece * 16;
        Piece pc{0, 0, 0xffffu, 0, -1};                                   // past the end: terminator bits, no counts
        if (pos0 < n) pc = encode_piece(load_piece(ascii, pos0, n), pos0, n, term);
        if (bad_pos && pc.bad >= 0) atomicMin(bad_pos, (unsigned long long)(pos0 + pc.bad));
        sm.plane[0][piece] = (uint16_t)pc.p0;
        sm.plane[1][piece] = (uint16_t)pc.p1;
        sm.plane[2][piece] = (uint16_t)pc.pt;
        sm.cnt[piece] = pc.cnt;
    }
    __syncthreads();
    // per-block totals (4 pieces each) and their exclusive scan over the 256 blocks of the tile
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint4 c4 = reinterpret_cast<const uint4 *>(sm.cnt)[threadIdx.x];
    const unsigned long long half = widen_counts(c4.x) + widen_counts(c4.y);
    const unsigned long long mine = half + widen_counts(c4.z) + widen_counts(c4.w);
    unsigned long long incl = mine;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const unsigned long long y = __shfl_up_sync(0xffffffffu, incl, s);
        if (lane >= s) incl += y;
    }
    if (lane == 31) sm.wsum[warp] = incl;
    __syncthreads();
    unsigned long long off = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kBuildThreads / 32; ++w) {
        const unsigned long long v = sm.wsum[w];
        if (w < warp) off += v;
        total += v;                                                   // <= 16384 per field
    }
    sm.mid[threadIdx.x] = off + incl - mine + half;                   // 16-bit fields: a tile has 16384 symbols
    __syncthreads();
    // 256 blocks x 2 uint4 = 512 uint4 per tile, written coalesced
    uint4 *out = blocks + (tile << (kTileShift - kBlockShift)) * kBlockU4;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const int q = it * kBuildThreads + threadIdx.x;
        const int blk = q >> 1, part = q & 1;
        uint4 v;
        if (part == 0) {
            const unsigned long long e = sm.mid[blk];
            const uint16_t *pl = sm.plane[0] + blk * 4;
            // counts before the middle of the block, relative to the superblock start (< 2^16)
            v.x = (rel[0] + (uint32_t)(e & 0xffff)) | ((rel[1] + (uint32_t)((e >> 16) & 0xffff)) << 16);
            v.y = (rel[2] + (uint32_t)((e >> 32) & 0xffff)) | ((rel[3] + (uint32_t)(e >> 48)) << 16);
            v.z = pl[0] | ((uint32_t)pl[1] << 16);
            v.w = pl[2] | ((uint32_t)pl[3] << 16);
        } else {
            const uint16_t *pb = sm.plane[1] + blk * 4, *pt = sm.plane[2] + blk * 4;
            v.x = pb[0] | ((uint32_t)pb[1] << 16);
            v.y = pb[2] | ((uint32_t)pb[3] << 16);
            v.z = pt[0] | ((uint32_t)pt[1] << 16);
            v.w = pt[2] | ((uint32_t)pt[3] << 16);
        }
        out[q] = v;
    }
    __syncthreads();
    return total;
}

// Kernel 3 (slice-wise build): pack tiles whose prefix counts are known.
__global__ void __launch_bounds__(kBuildThreads)
pack_tiles_kernel(const uint8_t *__restrict__ ascii, uint64_t n, uint32_t term, uint64_t tile0, uint64_t n_tiles,
                  const ulonglong4 *__restrict__ tile_prefix, const ulonglong4 base,
                  const unsigned long long *__restrict__ super_in, uint4 *__restrict__ blocks,
                  unsigned long long *__restrict__ super) {
    // Packs global tiles tile0 .. tile0 + n_tiles - 1 (`ascii` points at position tile0 << kTileShift);
    // tile_prefix is local to the slice and `base` holds the counts before it.  super_in = the complete table of
    // absolute counts at the superblock starts (nullptr: the slice starts at 0 and the entries are written here).
    ascii -= tile0 << kTileShift;
    __shared__ __align__(16) TileSmem sm;
    for (uint64_t lt = blockIdx.x; lt < n_tiles; lt += gridDim.x) {
        const uint64_t tile = tile0 + lt;
        ulonglong4 tp = tile_prefix[lt], sp;
        tp.x += base.x; tp.y += base.y; tp.z += base.z; tp.w += base.w;       // absolute counts before the tile
        if (super_in) {
            const unsigned long long *q = super_in + (tile >> kSuperTileShift) * 4;
            sp = make_ulonglong4(q[0], q[1], q[2], q[3]);
        } else {
            sp = tile_prefix[((tile >> kSuperTileShift) << kSuperTileShift) - tile0];
            if (threadIdx.x == 0 && (tile & ((1ull << kSuperTileShift) - 1)) == 0) {
                unsigned long long *s = super + (tile >> kSuperTileShift) * 4;
                s[0] = tp.x; s[1] = tp.y; s[2] = tp.z; s[3] = tp.w;
            }
        }
        const uint32_t rel[4] = {(uint32_t)(tp.x - sp.x), (uint32_t)(tp.y - sp.y), (uint32_t)(tp.z - sp.z), (uint32_t)(tp.w - sp.w)};
        pack_one_tile(sm, ascii, n, term, tile, rel, blocks, nullptr);
    }
}

// One-pass build (whole-string and streamed builds): a CTA packs a whole SUPERBLOCK (4 tiles), so the block
// counters -- which are relative to the superblock start -- need no prefix from anywhere else; the totals of the
// superblocks are scanned afterwards into the (tiny) table of absolute counts.  The ASCII text is read ONCE and
// the forbidden-symbol check rides along.  `ascii` is addressed with global positions and is valid in
// [first byte of superblock sb0, end): a streamed build passes the device copy of ONE chunk, shifted.
__global__ void __launch_bounds__(kBuildThreads)
pack_super_kernel(const uint8_t *__restrict__ ascii, uint64_t end, uint32_t term, uint64_t sb0, uint64_t n_sb, uint64_t n_tiles_total,
                  uint4 *__restrict__ blocks, uint4 *__restrict__ sb_tot, unsigned long long *bad_pos) {
    __shared__ __align__(16) TileSmem sm;
    for (uint64_t sb = sb0 + blockIdx.x; sb < sb0 + n_sb; sb += gridDim.x) {
        uint32_t rel[4] = {0, 0, 0, 0};
        for (uint64_t t = 0; t < (1ull << kSuperTileShift); ++t) {
            const uint64_t tile = (sb << kSuperTileShift) + t;
            if (tile >= n_tiles_total) break;
            const unsigned long long tot = pack_one_tile(sm, ascii, end, term, tile, rel, blocks, bad_pos);
            rel[0] += (uint32_t)(tot & 0xffff);
            rel[1] += (uint32_t)((tot >> 16) & 0xffff);
            rel[2] += (uint32_t)((tot >> 32) & 0xffff);
            rel[3] += (uint32_t)(tot >> 48);
        }
        if (threadIdx.x == 0) sb_tot[sb] = make_uint4(rel[0], rel[1], rel[2], rel[3]);
    }
}

// ---- batched test hooks ------------------------------------------------------------------------
__global__ void rank_batch_kernel(DevIndex ix, const uint64_t *__restrict__ pos, uint64_t m, uint64_t *__restrict__ out4) {
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= m) return;
    uint64_t r[4];
    rank4(ix, pos[q], r);
    reinterpret_cast<ulonglong2 *>(out4)[2 * q] = make_ulonglong2(r[0], r[1]);
    reinterpret_cast<ulonglong2 *>(out4)[2 * q + 1] = make_ulonglong2(r[2], r[3]);
}

__global__ void access_batch_kernel(DevIndex ix, const uint64_t *__restrict__ pos, uint64_t m, uint8_t *__restrict__ out, uint8_t term) {
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= m) return;
    const int c = access_code(ix, pos[q]);
    out[q] = c == 4 ? term : (uint8_t)"ACGT"[c];
}

__global__ void fl_batch_kernel(DevIndex ix, const uint64_t *__restrict__ pos, uint64_t m, uint64_t *__restrict__ out) {
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= m) return;
    const uint64_t i = pos[q];
    const int c = f_code(ix, i);
    out[q] = c == 4 ? ~0ull : fl_map(ix, i, c);
}

// document array: ASCII '1' -> 1, anything else -> 0 (ebwt2InDel.cpp:1503-1508); 32 positions per thread
__global__ void da_pack_kernel(const uint8_t *__restrict__ ascii, uint64_t n, uint32_t *__restrict__ words, uint64_t n_words, bool aligned16) {
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    uint32_t bits = 0;
    const uint64_t base = w * 32;
    if (base + 32 <= n && aligned16) {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(ascii + base));
        const uint4 b = __ldg(reinterpret_cast<const uint4 *>(ascii + base + 16));
        const uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int j = 0; j < 32; ++j) bits |= (uint32_t)(((v[j >> 2] >> (8 * (j & 3))) & 0xffu) == '1') << j;
    } else {
        for (int j = 0; j < 32; ++j) if (base + j < n) bits |= (uint32_t)(ascii[base + j] == '1') << j;
    }
    words[w] = bits;
}

}  // namespace e2i

using namespace e2i;

// =================================================================================================
// C ABI
// =================================================================================================
// ---- whole-string build: superblocks are packed as their bytes become available (all at once for a resident
//      string, chunk by chunk behind the copies of a streamed one), then the superblock totals are scanned ----
namespace {
struct IndexBuild {
    e2i_index *ix = nullptr;
    uint4 *sb_tot = nullptr;               // A,C,G,T totals of every superblock
    unsigned long long *scal = nullptr;    // [0..3] totals, [4] position of the first forbidden symbol
    uint64_t n_tiles = 0, n_sb = 0, packed_sb = 0;
};

void index_abort(e2i_ctx *ctx, IndexBuild &b) {
    dfree(ctx, b.sb_tot); dfree(ctx, b.scal);
    e2i_index_free(b.ix);
    b = IndexBuild();
}

int index_begin(e2i_ctx *ctx, uint64_t n, uint8_t term, IndexBuild &b) {
    cudaStream_t s = ctx->stream;
    e2i_index *ix = new e2i_index();
    b.ix = ix;
    ix->ctx = ctx;
    ix->n = n;
    ix->term = term;
    ix->n_blocks = n / kBlockSyms + 1;                       // rank(n) must be addressable (dna_string.hpp:62)
    b.n_tiles = (ix->n_blocks + kTileBlocks - 1) / kTileBlocks;
    ix->n_super = (n >> kSuperShift) + 1;
    b.n_sb = ix->n_super;                                    // = ceil(n_tiles / tiles per superblock)
    const size_t blk_bytes = b.n_tiles * kTileBlocks * kBlockU4 * sizeof(uint4);
    ix->bytes = blk_bytes + ix->n_super * 32;
    const unsigned long long init[5] = {0, 0, 0, 0, ~0ull};
    cudaError_t e = dmalloc(ctx, &ix->blocks, blk_bytes);
    if (e == cudaSuccess) e = dmalloc(ctx, &ix->super, ix->n_super * 32);
    if (e == cudaSuccess) e = dmalloc(ctx, &b.sb_tot, b.n_sb * sizeof(uint4));
    if (e == cudaSuccess) e = dmalloc(ctx, &b.scal, 5 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemcpyAsync(b.scal, init, sizeof init, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);      // `init` lives on this stack frame
    if (e != cudaSuccess) { set_error("index build: %s", cudaGetErrorString(e)); cudaGetLastError(); index_abort(ctx, b); return E2I_ERR_CUDA; }
    ctx->n_h2d += sizeof init;
    return E2I_OK;
}

// pack superblocks [sb0, sb0 + nsb); `ascii` is addressed with global positions and holds the bytes of these
// superblocks up to position `end` (the superblocks of a chunk: the chunk's device copy, shifted)
int index_pack(e2i_ctx *ctx, IndexBuild &b, const uint8_t *ascii, uint64_t sb0, uint64_t nsb, uint64_t end) {
    if (!nsb) return E2I_OK;
    const int grid = (int)std::min<uint64_t>(nsb, (uint64_t)ctx->sm_count * 8);
    pack_super_kernel<<<grid, kBuildThreads, 0, ctx->stream>>>(ascii, end, b.ix->term, sb0, nsb, b.n_tiles, b.ix->blocks, b.sb_tot, b.scal + 4);
    E2I_CUDA_TRY(cudaGetLastError());
    ctx->n_launch++;
    b.packed_sb = sb0 + nsb;
    return E2I_OK;
}

int index_finish(e2i_ctx *ctx, IndexBuild &b, e2i_index **out, uint64_t *bad_pos) {
    cudaStream_t s = ctx->stream;
    e2i_index *ix = b.ix;
    if (b.packed_sb < b.n_sb) {                              // only superblocks without a byte can be left (empty string)
        if ((b.packed_sb << kSuperShift) < ix->n) { set_error("index build: the input ended early"); index_abort(ctx, b); return E2I_ERR_ARG; }
        const int rc = index_pack(ctx, b, nullptr, b.packed_sb, b.n_sb - b.packed_sb, ix->n);
        if (rc != E2I_OK) { index_abort(ctx, b); return rc; }
    }
    scan_tiles_kernel<<<1, kScanThreads, 0, s>>>(b.sb_tot, b.n_sb, reinterpret_cast<ulonglong4 *>(ix->super), b.scal);
    unsigned long long res[5];
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(res, b.scal, sizeof res, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { set_error("index build: %s", cudaGetErrorString(e)); index_abort(ctx, b); return E2I_ERR_CUDA; }
    ctx->n_launch += 1;
    ctx->n_d2h += sizeof res;
    if (res[4] != ~0ull) {
        if (bad_pos) *bad_pos = res[4];
        set_error("forbidden character at position %llu: only A,C,G,T and the terminator (ASCII %d) are admitted in the input BWT",
                  res[4], (int)ix->term);
        index_abort(ctx, b);
        return E2I_ERR_SYMBOL;
    }
    const uint64_t acgt = res[0] + res[1] + res[2] + res[3];
    ix->F[0] = ix->n - acgt;                                  // dna_bwt.hpp:51-60
    ix->F[1] = ix->F[0] + res[0];
    ix->F[2] = ix->F[1] + res[1];
    ix->F[3] = ix->F[2] + res[2];
    dfree(ctx, b.sb_tot); dfree(ctx, b.scal);
    *out = ix;
    b = IndexBuild();
    return E2I_OK;
}
}  // namespace

extern "C" int e2i_index_build_device(e2i_ctx *ctx, const uint8_t *dev_ascii, uint64_t n, uint8_t term,
                                      e2i_index **out, uint64_t *bad_pos) {
    if (!ctx || !out || (n && !dev_ascii)) { set_error("e2i_index_build_device: null argument"); return E2I_ERR_ARG; }
    if (reinterpret_cast<uintptr_t>(dev_ascii) & 15) { set_error("e2i_index_build_device: input must be 16-byte aligned"); return E2I_ERR_ARG; }
    E2I_CUDA_TRY(cudaSetDevice(ctx->device));
    IndexBuild b;
    E2I_TRY(index_begin(ctx, n, term, b));
    const int rc = index_pack(ctx, b, dev_ascii, 0, b.n_sb, n);
    if (rc != E2I_OK) { index_abort(ctx, b); return rc; }
    return index_finish(ctx, b, out, bad_pos);
}

// ---- slice-wise construction (multi-GPU): every rank packs the blocks of one tile-aligned slice ----
extern "C" uint64_t e2i_index_slice_align(void) { return (uint64_t)kTileSyms; }

extern "C" int e2i_index_alloc(e2i_ctx *ctx, uint64_t n, uint8_t term, uint64_t tile_multiple, e2i_index **out) {
    if (!ctx || !out || tile_multiple == 0) { set_error("e2i_index_alloc: bad argument"); return E2I_ERR_ARG; }
    E2I_CUDA_TRY(cudaSetDevice(ctx->device));
    e2i_index *ix = new e2i_index();
    ix->ctx = ctx;
    ix->n = n;
    ix->term = term;
    ix->n_blocks = n / kBlockSyms + 1;
    uint64_t n_tiles = (ix->n_blocks + kTileBlocks - 1) / kTileBlocks;
    n_tiles = (n_tiles + tile_multiple - 1) / tile_multiple * tile_multiple;   // equal slices for the all-gather
    ix->n_super = (n >> kSuperShift) + 1;
    const size_t blk_bytes = n_tiles * kTileBlocks * kBlockU4 * sizeof(uint4);
    ix->bytes = blk_bytes + ix->n_super * 32;
    if (dmalloc(ctx, &ix->blocks, blk_bytes) != cudaSuccess || dmalloc(ctx, &ix->super, ix->n_super * 32) != cudaSuccess) {
        cudaGetLastError();
        set_error("e2i_index_alloc: out of device memory");
        e2i_index_free(ix);
        return E2I_ERR_MEMORY;
    }
    *out = ix;
    return E2I_OK;
}

extern "C" int e2i_index_slice_count(e2i_ctx *ctx, e2i_index *ix, const uint8_t *dev_slice, uint64_t begin, uint64_t len,
                                     uint64_t n_tiles, uint64_t counts[4], uint64_t *bad_pos) {
    if (!ctx || !ix || !counts || (len && !dev_slice)) { set_error("e2i_index_slice_count: null argument"); return E2I_ERR_ARG; }
    const uint64_t all_tiles = (ix->n_blocks + kTileBlocks - 1) / kTileBlocks;
    if (n_tiles && ((begin & (kTileSyms - 1)) || begin + len > ix->n || (reinterpret_cast<uintptr_t>(dev_slice) & 15) ||
                    (begin >> kTileShift) + n_tiles > all_tiles || len > n_tiles * kTileSyms)) {
        set_error("e2i_index_slice_count: slice must start on a multiple of %d symbols, lie inside the string and be 16-byte aligned", kTileSyms);
        return E2I_ERR_ARG;
    }
    E2I_CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    dfree(ctx, ix->slice_cnt); dfree(ctx, ix->slice_prefix);
    ix->slice_cnt = nullptr; ix->slice_prefix = nullptr;
    ix->slice_begin = n_tiles ? begin : 0;
    ix->slice_len = n_tiles ? len : 0;
    // n_tiles: the tiles this slice owns; the owner of the last tile also packs the block that makes rank(n)
    // addressable (that tile may hold no symbol at all).  An empty slice (n_tiles == 0) does nothing.
    ix->slice_tiles = n_tiles;
    const uint64_t end = ix->slice_begin + ix->slice_len;
    const uint64_t nt = ix->slice_tiles;
    for (int k = 0; k < 4; ++k) counts[k] = 0;
    if (nt == 0) return E2I_OK;
    unsigned long long *scal = nullptr;
    E2I_CUDA_TRY(dmalloc(ctx, &ix->slice_cnt, nt * sizeof(uint4)));
    E2I_CUDA_TRY(dmalloc(ctx, &ix->slice_prefix, nt * sizeof(ulonglong4)));
    E2I_CUDA_TRY(dmalloc(ctx, &scal, 5 * sizeof(unsigned long long)));
    unsigned long long init[5] = {0, 0, 0, 0, ~0ull}, res[5];
    E2I_CUDA_TRY(cudaMemcpyAsync(scal, init, sizeof init, cudaMemcpyHostToDevice, s));
    const int grid = (int)std::min<uint64_t>(nt, (uint64_t)ctx->sm_count * 16);
    count_tiles_kernel<<<grid, kBuildThreads, 0, s>>>(dev_slice, end, ix->term, begin >> kTileShift, nt,
                                                      static_cast<uint4 *>(ix->slice_cnt), scal + 4);
    scan_tiles_kernel<<<1, kScanThreads, 0, s>>>(static_cast<uint4 *>(ix->slice_cnt), nt, static_cast<ulonglong4 *>(ix->slice_prefix), scal);
    E2I_CUDA_TRY(cudaGetLastError());
    ctx->n_launch += 2;
    E2I_CUDA_TRY(cudaMemcpyAsync(res, scal, sizeof res, cudaMemcpyDeviceToHost, s));
    E2I_CUDA_TRY(cudaStreamSynchronize(s));
    dfree(ctx, scal);
    if (res[4] != ~0ull) {
        if (bad_pos) *bad_pos = res[4];
        set_error("forbidden character at position %llu: only A,C,G,T and the terminator (ASCII %d) are admitted in the input BWT", res[4], (int)ix->term);
        return E2I_ERR_SYMBOL;
    }
    for (int k = 0; k < 4; ++k) counts[k] = res[k];
    return E2I_OK;
}

extern "C" uint64_t e2i_index_super_count(const e2i_index *ix) { return ix ? ix->n_super : 0; }

extern "C" int e2i_index_slice_super(e2i_ctx *ctx, e2i_index *ix, const uint64_t before[4], uint64_t *host_super) {
    if (!ctx || !ix || !before || !host_super) { set_error("e2i_index_slice_super: null argument"); return E2I_ERR_ARG; }
    E2I_CUDA_TRY(cudaSetDevice(ctx->device));
    std::memset(host_super, 0, ix->n_super * 32);
    const uint64_t tile0 = ix->slice_begin >> kTileShift, step = 1ull << kSuperTileShift;
    // superblocks that start inside this slice: tiles tile0 <= sb * step < tile0 + slice_tiles
    const uint64_t sb_lo = (tile0 + step - 1) / step, sb_hi = std::min<uint64_t>(ix->n_super, (tile0 + ix->slice_tiles + step - 1) / step);
    if (sb_lo >= sb_hi || ix->slice_tiles == 0) return E2I_OK;
    // one strided copy: the prefix entry (32 bytes) of every `step`-th tile
    E2I_CUDA_TRY(cudaMemcpy2DAsync(host_super + sb_lo * 4, 32, static_cast<ulonglong4 *>(ix->slice_prefix) + (sb_lo * step - tile0), 32 * step, 32,
                                   sb_hi - sb_lo, cudaMemcpyDeviceToHost, ctx->stream));
    E2I_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    for (uint64_t sb = sb_lo; sb < sb_hi; ++sb)
        for (int k = 0; k < 4; ++k) host_super[sb * 4 + k] += before[k];
    return E2I_OK;
}

extern "C" int e2i_index_slice_pack(e2i_ctx *ctx, e2i_index *ix, const uint8_t *dev_slice, const uint64_t before[4],
                                    const uint64_t *host_super) {
    if (!ctx || !ix || !before || !host_super) { set_error("e2i_index_slice_pack: null argument"); return E2I_ERR_ARG; }
    E2I_CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    E2I_CUDA_TRY(cudaMemcpyAsync(ix->super, host_super, ix->n_super * 32, cudaMemcpyHostToDevice, s));
    const uint64_t nt = ix->slice_tiles;
    if (nt) {
        const int grid = (int)std::min<uint64_t>(nt, (uint64_t)ctx->sm_count * 16);
        pack_tiles_kernel<<<grid, kBuildThreads, 0, s>>>(dev_slice, ix->slice_begin + ix->slice_len, ix->term, ix->slice_begin >> kTileShift, nt,
                                                         static_cast<ulonglong4 *>(ix->slice_prefix),
                                                         make_ulonglong4(before[0], before[1], before[2], before[3]),
                                                         reinterpret_cast<const unsigned long long *>(ix->super), ix->blocks, nullptr);
        E2I_CUDA_TRY(cudaGetLastError());
        ctx->n_launch++;
    }
    E2I_CUDA_TRY(cudaStreamSynchronize(s));
    dfree(ctx, ix->slice_cnt); dfree(ctx, ix->slice_prefix);
    ix->slice_cnt = nullptr; ix->slice_prefix = nullptr;
    return E2I_OK;
}

extern "C" int e2i_index_finish(e2i_index *ix, const uint64_t totals[4]) {
    if (!ix || !totals) { set_error("e2i_index_finish: null argument"); return E2I_ERR_ARG; }
    const uint64_t acgt = totals[0] + totals[1] + totals[2] + totals[3];
    if (acgt > ix->n) { set_error("e2i_index_finish: symbol totals exceed the string length"); return E2I_ERR_ARG; }
    ix->F[0] = ix->n - acgt;
    ix->F[1] = ix->F[0] + totals[0];
    ix->F[2] = ix->F[1] + totals[1];
    ix->F[3] = ix->F[2] + totals[2];
    return E2I_OK;
}

extern "C" int e2i_index_device(const e2i_index *ix, void **dev_blocks, uint64_t *block_bytes) {
    if (!ix) { set_error("e2i_index_device: null argument"); return E2I_ERR_ARG; }
    if (dev_blocks) *dev_blocks = ix->blocks;
    if (block_bytes) *block_bytes = ix->bytes - ix->n_super * 32;
    return E2I_OK;
}

// ---- streaming ingest: the eBWT goes up in chunks on the copy stream while the counting
//      pass follows it on the compute stream; with a file as the source a reader thread fills a ring of
//      page-locked buffers, so disk reads, PCIe copies and the counting kernels overlap.  Replaces the
//      byte-at-a-time loop of dna_string.hpp:82-101. ---------------------------------------------------
namespace {
constexpr uint64_t kChunk = 64ull << 20;                  // multiple of the superblock size
static_assert(kChunk % (1ull << kSuperShift) == 0, "chunks must end on superblock boundaries");
constexpr int kRing = 4;

struct Uploader {                                           // H2D of consecutive chunks into a small device ring + packing behind them
    e2i_ctx *ctx;
    IndexBuild *b;
    uint64_t n, off = 0;
    uint8_t *ring = nullptr;                                // kRing chunk buffers in device memory: the ASCII text never sits in HBM as a whole
    cudaEvent_t ev[kRing] = {}, packed[kRing] = {};
    int k = 0;
    int init() {
        for (auto &e : ev) E2I_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto &e : packed) E2I_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        E2I_CUDA_TRY(dmalloc(ctx, &ring, (size_t)kRing * kChunk));
        // the copy stream must not run ahead of the allocations made on the compute stream
        E2I_CUDA_TRY(cudaEventRecord(ev[0], ctx->stream));
        E2I_CUDA_TRY(cudaStreamWaitEvent(ctx->copy_stream, ev[0], 0));
        return E2I_OK;
    }
    ~Uploader() {
        for (auto &e : ev) if (e) cudaEventDestroy(e);
        for (auto &e : packed) if (e) cudaEventDestroy(e);
        dfree(ctx, ring);                                   // stream-ordered: after the last packing kernel
    }
    // host must stay valid until event `slot` (returned) has completed
    int push(const uint8_t *host, uint64_t len, int *slot) {
        const int sl = k % kRing;
        uint8_t *dst = ring + (size_t)sl * kChunk;
        if (k >= kRing) E2I_CUDA_TRY(cudaStreamWaitEvent(ctx->copy_stream, packed[sl], 0));   // the kernel that read this buffer is done
        ++k;
        E2I_CUDA_TRY(cudaMemcpyAsync(dst, host, len, cudaMemcpyHostToDevice, ctx->copy_stream));
        E2I_CUDA_TRY(cudaEventRecord(ev[sl], ctx->copy_stream));
        E2I_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ev[sl], 0));
        ctx->n_h2d += len;
        // chunks are superblock-aligned except the last, which also owns what follows the string's end
        const bool last = off + len >= n;
        const uint64_t sb0 = off >> kSuperShift, sb1 = last ? b->n_sb : (off + len) >> kSuperShift;
        E2I_TRY(index_pack(ctx, *b, dst - off, sb0, sb1 - sb0, off + len));
        E2I_CUDA_TRY(cudaEventRecord(packed[sl], ctx->stream));
        off += len;
        if (slot) *slot = sl;
        return E2I_OK;
    }
};
}  // namespace

extern "C" int e2i_index_build(e2i_ctx *ctx, const uint8_t *host_ascii, uint64_t n, uint8_t term,
                               e2i_index **out, uint64_t *bad_pos) {
    if (!ctx || !out || (n && !host_ascii)) { set_error("e2i_index_build: null argument"); return E2I_ERR_ARG; }
    E2I_CUDA_TRY(cudaSetDevice(ctx->device));
    IndexBuild b;
    int rc = index_begin(ctx, n, term, b);
    if (rc != E2I_OK) return rc;
    {
        Uploader up{ctx, &b, n};
        rc = up.init();
        cudaEvent_t t0 = nullptr, t1 = nullptr;
        if (rc == E2I_OK && (cudaEventCreate(&t0) != cudaSuccess || cudaEventCreate(&t1) != cudaSuccess)) rc = E2I_ERR_CUDA;
        if (rc == E2I_OK) cudaEventRecord(t0, ctx->copy_stream);
        for (uint64_t off = 0; rc == E2I_OK && off < n; off += kChunk) rc = up.push(host_ascii + off, std::min(kChunk, n - off), nullptr);
        if (rc == E2I_OK) cudaEventRecord(t1, ctx->copy_stream);
        if (rc == E2I_OK) rc = index_finish(ctx, b, out, bad_pos); else index_abort(ctx, b);
        if (t0 && t1 && rc == E2I_OK) { float ms = 0; if (cudaEventElapsedTime(&ms, t0, t1) == cudaSuccess) ctx->last_h2d_ms = ms; }
        if (t0) cudaEventDestroy(t0);
        if (t1) cudaEventDestroy(t1);
        cudaStreamSynchronize(ctx->copy_stream);
    }
    return rc;
}

namespace {
// A file streamed through a ring of page-locked buffers by a reader thread.  `total` bytes are delivered:
// the file's bytes, then (DA semantics, ebwt2InDel.cpp:1503-1508) copies of its last byte.
struct FileStream {
    int fd = -1;
    uint64_t size = 0, total = 0;
    uint8_t *buf[kRing] = {};
    uint64_t len[kRing] = {};
    std::mutex m;
    std::condition_variable cv;
    uint64_t filled = 0, released = 0;      // chunks produced / chunks whose buffer may be reused
    bool failed = false;
    std::thread th;
    void reader() {
        const uint64_t n_chunks = (total + kChunk - 1) / kChunk;
        uint8_t last = 0;
        for (uint64_t c = 0; c < n_chunks; ++c) {
            {
                std::unique_lock<std::mutex> lk(m);
                cv.wait(lk, [&] { return c < released + kRing; });
            }
            const int sl = (int)(c % kRing);
            const uint64_t off = c * kChunk, want = std::min(kChunk, total - off);
            uint64_t got = 0;
            while (off + got < size && got < want) {
                const ssize_t r = ::pread(fd, buf[sl] + got, (size_t)std::min<uint64_t>(want - got, size - off - got), (off_t)(off + got));
                if (r <= 0) { std::lock_guard<std::mutex> lk(m); failed = true; cv.notify_all(); return; }
                got += (uint64_t)r;
            }
            if (got) last = buf[sl][got - 1];
            if (got < want) std::memset(buf[sl] + got, last, want - got);
            {
                std::lock_guard<std::mutex> lk(m);
                len[sl] = want;
                ++filled;
            }
            cv.notify_all();
        }
    }
    int open(e2i_ctx *ctx, const char *path, uint64_t want_total) {
        fd = ::open(path, O_RDONLY);
        struct stat sb;
        if (fd < 0 || ::fstat(fd, &sb) != 0) { set_error("could not read file %s", path); return E2I_ERR_IO; }
        size = (uint64_t)sb.st_size;
        total = want_total ? want_total : size;
        if (!ctx->ring[0]) {
            for (int i = 0; i < kRing; ++i)
                if (cudaMallocHost(&ctx->ring[i], kChunk) != cudaSuccess) { cudaGetLastError(); set_error("cannot page-lock the ingest ring"); return E2I_ERR_MEMORY; }
        }
        for (int i = 0; i < kRing; ++i) buf[i] = static_cast<uint8_t *>(ctx->ring[i]);
#ifdef POSIX_FADV_SEQUENTIAL
        ::posix_fadvise(fd, 0, 0, POSIX_FADV_SEQUENTIAL);
#endif
        th = std::thread([this] { reader(); });
        return E2I_OK;
    }
    // next filled chunk (blocks); false on a read error
    bool next(uint64_t c, const uint8_t **p, uint64_t *l) {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return failed || filled > c; });
        if (failed) return false;
        *p = buf[c % kRing];
        *l = len[c % kRing];
        return true;
    }
    void release() { { std::lock_guard<std::mutex> lk(m); ++released; } cv.notify_all(); }
    ~FileStream() {
        { std::lock_guard<std::mutex> lk(m); released = ~0ull >> 1; }
        cv.notify_all();
        if (th.joinable()) th.join();
        if (fd >= 0) ::close(fd);
    }
};
}  // namespace

extern "C" int e2i_index_build_file(e2i_ctx *ctx, const char *path, uint8_t term, e2i_index **out, uint64_t *bad_pos) {
    if (!ctx || !path || !out) { set_error("e2i_index_build_file: null argument"); return E2I_ERR_ARG; }
    E2I_CUDA_TRY(cudaSetDevice(ctx->device));
    FileStream fs;
    E2I_TRY(fs.open(ctx, path, 0));
    const uint64_t n = fs.total;
    IndexBuild b;
    int rc = index_begin(ctx, n, term, b);
    if (rc != E2I_OK) return rc;
    {
        Uploader up{ctx, &b, n};
        rc = up.init();
        const uint64_t n_chunks = (n + kChunk - 1) / kChunk;
        std::vector<int> slot_of(n_chunks, 0);
        for (uint64_t c = 0; rc == E2I_OK && c < n_chunks; ++c) {
            const uint8_t *p;
            uint64_t l;
            if (!fs.next(c, &p, &l)) { set_error("read error on %s", path); rc = E2I_ERR_IO; break; }
            int sl = 0;
            rc = up.push(p, l, &sl);
            slot_of[c] = sl;
            // the buffer of chunk c - (kRing - 2) is free once its copy has completed
            if (rc == E2I_OK && c + 2 >= (uint64_t)kRing) { cudaEventSynchronize(up.ev[slot_of[c + 2 - kRing]]); fs.release(); }
        }
        cudaStreamSynchronize(ctx->copy_stream);
        if (rc == E2I_OK) rc = index_finish(ctx, b, out, bad_pos); else index_abort(ctx, b);
    }
    return rc;
}

// document array from a file, streamed the same way; a file shorter than n repeats its last byte
extern "C" int e2i_da_load_file(e2i_ctx *ctx, const char *path, uint64_t n, e2i_bits **out) {
    if (!ctx || !path || !out) { set_error("e2i_da_load_file: null argument"); return E2I_ERR_ARG; }
    E2I_CUDA_TRY(cudaSetDevice(ctx->device));
    FileStream fs;
    E2I_TRY(fs.open(ctx, path, n));
    uint8_t *d = nullptr;
    E2I_CUDA_TRY(dmalloc(ctx, &d, n + 16));
    cudaEvent_t ev[kRing] = {};
    int rc = E2I_OK;
    for (auto &e : ev) if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) rc = E2I_ERR_CUDA;
    if (rc == E2I_OK && cudaEventRecord(ev[0], ctx->stream) == cudaSuccess) cudaStreamWaitEvent(ctx->copy_stream, ev[0], 0);
    const uint64_t n_chunks = (n + kChunk - 1) / kChunk;
    for (uint64_t c = 0; rc == E2I_OK && c < n_chunks; ++c) {
        const uint8_t *p;
        uint64_t l;
        if (!fs.next(c, &p, &l)) { set_error("read error on %s", path); rc = E2I_ERR_IO; break; }
        if (cudaMemcpyAsync(d + c * kChunk, p, l, cudaMemcpyHostToDevice, ctx->copy_stream) != cudaSuccess) { rc = E2I_ERR_CUDA; break; }
        cudaEventRecord(ev[c % kRing], ctx->copy_stream);
        ctx->n_h2d += l;
        if (c + 2 >= (uint64_t)kRing) { cudaEventSynchronize(ev[(c + 2 - kRing) % kRing]); fs.release(); }
    }
    cudaStreamSynchronize(ctx->copy_stream);
    for (auto &e : ev) if (e) cudaEventDestroy(e);
    if (rc == E2I_OK) rc = e2i_da_load_device(ctx, d, n, out);
    dfree(ctx, d);
    return rc;
}

// ---- packed-index sidecar: the index as it lies in HBM, so that a later run uploads n/2 bytes and skips the
//      build.  Replaces the reference's (unused) serialize / load, dna_bwt.hpp:238-289, dna_string.hpp:205-243.
namespace {
struct SidecarHeader {
    char magic[8];              // "E2IIDX02"
    uint64_t n, n_blocks, n_super, blk_bytes, F[4];
    uint32_t term, block_syms, super_shift, pad;
};
}  // namespace

extern "C" int e2i_index_save(const e2i_index *ix, const char *path) {
    if (!ix || !path) { set_error("e2i_index_save: null argument"); return E2I_ERR_ARG; }
    E2I_CUDA_TRY(cudaSetDevice(ix->ctx->device));
    SidecarHeader h = {};
    std::memcpy(h.magic, "E2IIDX02", 8);
    h.n = ix->n; h.n_blocks = ix->n_blocks; h.n_super = ix->n_super; h.blk_bytes = ix->bytes - ix->n_super * 32;
    for (int i = 0; i < 4; ++i) h.F[i] = ix->F[i];
    h.term = ix->term; h.block_syms = kBlockSyms; h.super_shift = kSuperShift;
    FILE *f = std::fopen(path, "wb");
    if (!f) { set_error("could not write %s", path); return E2I_ERR_IO; }
    bool ok = std::fwrite(&h, sizeof h, 1, f) == 1;
    std::vector<uint8_t> host((size_t)kChunk);
    auto dump = [&](const void *dev, uint64_t bytes) {
        for (uint64_t off = 0; ok && off < bytes; off += kChunk) {
            const uint64_t l = std::min(kChunk, bytes - off);
            ok = cudaMemcpy(host.data(), static_cast<const char *>(dev) + off, l, cudaMemcpyDeviceToHost) == cudaSuccess &&
                 std::fwrite(host.data(), 1, l, f) == l;
        }
    };
    dump(ix->super, ix->n_super * 32);
    dump(ix->blocks, h.blk_bytes);
    ok = (std::fclose(f) == 0) && ok;
    if (!ok) { set_error("could not write %s", path); return E2I_ERR_IO; }
    return E2I_OK;
}

extern "C" int e2i_index_load(e2i_ctx *ctx, const char *path, e2i_index **out) {
    if (!ctx || !path || !out) { set_error("e2i_index_load: null argument"); return E2I_ERR_ARG; }
    E2I_CUDA_TRY(cudaSetDevice(ctx->device));
    FileStream fs;
    E2I_TRY(fs.open(ctx, path, 0));
    const uint8_t *p;
    uint64_t l;
    SidecarHeader h;
    if (fs.total < sizeof h || !fs.next(0, &p, &l)) { set_error("%s is not a packed index", path); return E2I_ERR_IO; }
    std::memcpy(&h, p, sizeof h);
    if (std::memcmp(h.magic, "E2IIDX02", 8) != 0 || h.block_syms != (uint32_t)kBlockSyms || h.super_shift != (uint32_t)kSuperShift ||
        fs.total != sizeof h + h.n_super * 32 + h.blk_bytes || h.n_blocks != h.n / kBlockSyms + 1) {
        set_error("%s is not a packed index of this library version", path);
        return E2I_ERR_IO;
    }
    e2i_index *ix = new e2i_index();
    ix->ctx = ctx; ix->n = h.n; ix->n_blocks = h.n_blocks; ix->n_super = h.n_super; ix->term = (uint8_t)h.term;
    ix->bytes = h.blk_bytes + h.n_super * 32;
    for (int i = 0; i < 4; ++i) ix->F[i] = h.F[i];
    if (dmalloc(ctx, &ix->blocks, h.blk_bytes) != cudaSuccess || dmalloc(ctx, &ix->super, h.n_super * 32) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        cudaGetLastError(); set_error("e2i_index_load: out of device memory"); e2i_index_free(ix); return E2I_ERR_MEMORY;
    }
    // the payload [super | blocks] follows the header; chunks go up as they arrive
    cudaEvent_t ev[kRing] = {};
    int rc = E2I_OK;
    for (auto &e : ev) if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) rc = E2I_ERR_CUDA;
    const uint64_t s_bytes = h.n_super * 32, n_chunks = (fs.total + kChunk - 1) / kChunk;
    for (uint64_t c = 0; rc == E2I_OK && c < n_chunks; ++c) {
        if (c && !fs.next(c, &p, &l)) { set_error("read error on %s", path); rc = E2I_ERR_IO; break; }
        uint64_t lo = c * kChunk, hi = lo + l;                       // file range of this chunk
        const uint64_t pay0 = sizeof h;
        auto copy = [&](uint64_t f0, uint64_t f1, char *dst_base, uint64_t dst_f0) {   // file range [f0,f1) ∩ chunk -> device
            const uint64_t a = std::max(lo, f0), b2 = std::min(hi, f1);
            if (a >= b2 || rc != E2I_OK) return;
            if (cudaMemcpyAsync(dst_base + (a - dst_f0), p + (a - lo), b2 - a, cudaMemcpyHostToDevice, ctx->copy_stream) != cudaSuccess) rc = E2I_ERR_CUDA;
            ctx->n_h2d += b2 - a;
        };
        copy(pay0, pay0 + s_bytes, reinterpret_cast<char *>(ix->super), pay0);
        copy(pay0 + s_bytes, fs.total, reinterpret_cast<char *>(ix->blocks), pay0 + s_bytes);
        cudaEventRecord(ev[c % kRing], ctx->copy_stream);
        if (c + 2 >= (uint64_t)kRing) { cudaEventSynchronize(ev[(c + 2 - kRing) % kRing]); fs.release(); }
    }
    cudaStreamSynchronize(ctx->copy_stream);
    for (auto &e : ev) if (e) cudaEventDestroy(e);
    if (rc != E2I_OK) { if (rc == E2I_ERR_CUDA) set_error("e2i_index_load: copy failed"); e2i_index_free(ix); return rc; }
    *out = ix;
    return E2I_OK;
}

extern "C" void e2i_index_free(e2i_index *ix) {
    if (!ix) return;
    dfree(ix->ctx, ix->blocks);
    dfree(ix->ctx, ix->super);
    dfree(ix->ctx, ix->slice_cnt);
    dfree(ix->ctx, ix->slice_prefix);
    delete ix;
}

extern "C" uint64_t e2i_index_size(const e2i_index *ix) { return ix ? ix->n : 0; }
extern "C" uint64_t e2i_index_bytes(const e2i_index *ix) { return ix ? ix->bytes : 0; }
extern "C" int e2i_index_F(const e2i_index *ix, uint64_t F[4]) {
    if (!ix || !F) { set_error("e2i_index_F: null argument"); return E2I_ERR_ARG; }
    for (int i = 0; i < 4; ++i) F[i] = ix->F[i];
    return E2I_OK;
}

namespace {
template <typename TOut, typename Launch>
int batch_hook(e2i_ctx *ctx, const uint64_t *host_pos, uint64_t m, TOut *host_out, size_t out_per, Launch launch) {
    if (!ctx || (m && (!host_pos || !host_out))) { set_error("batch hook: null argument"); return E2I_ERR_ARG; }
    if (m == 0) return E2I_OK;
    E2I_CUDA_TRY(cudaSetDevice(ctx->device));
    uint64_t *dpos = nullptr;
    TOut *dout = nullptr;
    E2I_CUDA_TRY(dmalloc(ctx, &dpos, m * sizeof(uint64_t)));
    cudaError_t e = dmalloc(ctx, &dout, m * out_per * sizeof(TOut));
    if (e == cudaSuccess) e = cudaMemcpyAsync(dpos, host_pos, m * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) { launch(dpos, dout); e = cudaGetLastError(); }
    if (e == cudaSuccess) e = cudaMemcpyAsync(host_out, dout, m * out_per * sizeof(TOut), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    dfree(ctx, dpos);
    dfree(ctx, dout);
    if (e != cudaSuccess) { set_error("batch hook failed: %s", cudaGetErrorString(e)); return E2I_ERR_CUDA; }
    return E2I_OK;
}
}  // namespace

extern "C" int e2i_rank_batch(e2i_ctx *ctx, const e2i_index *ix, const uint64_t *host_pos, uint64_t m, uint64_t *host_out4) {
    if (!ix) { set_error("e2i_rank_batch: null index"); return E2I_ERR_ARG; }
    for (uint64_t k = 0; k < m; ++k) if (host_pos[k] > ix->n) { set_error("e2i_rank_batch: position %llu > n", (unsigned long long)host_pos[k]); return E2I_ERR_ARG; }
    return batch_hook<uint64_t>(ctx, host_pos, m, host_out4, 4, [&](uint64_t *dp, uint64_t *dout) {
        rank_batch_kernel<<<(unsigned)((m + 255) / 256), 256, 0, ctx->stream>>>(ix->dev(), dp, m, dout);
    });
}

extern "C" int e2i_rank_batch_device(e2i_ctx *ctx, const e2i_index *ix, const uint64_t *dev_pos, uint64_t m,
                                     uint64_t *dev_out4, float *ms) {
    if (!ctx || !ix || !dev_pos || !dev_out4) { set_error("e2i_rank_batch_device: null argument"); return E2I_ERR_ARG; }
    E2I_CUDA_TRY(cudaSetDevice(ctx->device));
    E2I_CUDA_TRY(cudaEventRecord(ctx->ev[0], ctx->stream));
    rank_batch_kernel<<<(unsigned)((m + 255) / 256), 256, 0, ctx->stream>>>(ix->dev(), dev_pos, m, dev_out4);
    E2I_CUDA_TRY(cudaGetLastError());
    E2I_CUDA_TRY(cudaEventRecord(ctx->ev[1], ctx->stream));
    E2I_CUDA_TRY(cudaEventSynchronize(ctx->ev[1]));
    if (ms) E2I_CUDA_TRY(cudaEventElapsedTime(ms, ctx->ev[0], ctx->ev[1]));
    return E2I_OK;
}

extern "C" int e2i_access_batch(e2i_ctx *ctx, const e2i_index *ix, const uint64_t *host_pos, uint64_t m, uint8_t *host_out) {
    if (!ix) { set_error("e2i_access_batch: null index"); return E2I_ERR_ARG; }
    for (uint64_t k = 0; k < m; ++k) if (host_pos[k] >= ix->n) { set_error("e2i_access_batch: position out of range"); return E2I_ERR_ARG; }
    return batch_hook<uint8_t>(ctx, host_pos, m, host_out, 1, [&](uint64_t *dp, uint8_t *dout) {
        access_batch_kernel<<<(unsigned)((m + 255) / 256), 256, 0, ctx->stream>>>(ix->dev(), dp, m, dout, ix->term);
    });
}

extern "C" int e2i_fl_batch(e2i_ctx *ctx, const e2i_index *ix, const uint64_t *host_pos, uint64_t m, uint64_t *host_out) {
    if (!ix) { set_error("e2i_fl_batch: null index"); return E2I_ERR_ARG; }
    for (uint64_t k = 0; k < m; ++k) if (host_pos[k] >= ix->n) { set_error("e2i_fl_batch: position out of range"); return E2I_ERR_ARG; }
    return batch_hook<uint64_t>(ctx, host_pos, m, host_out, 1, [&](uint64_t *dp, uint64_t *dout) {
        fl_batch_kernel<<<(unsigned)((m + 255) / 256), 256, 0, ctx->stream>>>(ix->dev(), dp, m, dout);
    });
}

// ---- document array ----------------------------------------------------------------------------
static uint64_t padded_words32(uint64_t bits) { return ((bits + 31) / 32 + 63) / 64 * 64 + 64; }

extern "C" int e2i_da_load_device(e2i_ctx *ctx, const uint8_t *dev_ascii01, uint64_t n, e2i_bits **out) {
    if (!ctx || !out || (n && !dev_ascii01)) { set_error("e2i_da_load_device: null argument"); return E2I_ERR_ARG; }
    E2I_CUDA_TRY(cudaSetDevice(ctx->device));
    e2i_bits *b = new e2i_bits();
    b->ctx = ctx;
    b->n = n;
    b->n_words32 = padded_words32(n);
    cudaError_t e = dmalloc(ctx, &b->words, b->n_words32 * 4);
    if (e == cudaSuccess) e = cudaMemsetAsync(b->words, 0, b->n_words32 * 4, ctx->stream);
    const uint64_t nw = (n + 31) / 32;
    if (e == cudaSuccess && nw) {
        da_pack_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, ctx->stream>>>(dev_ascii01, n, b->words, nw,
                                                                              (reinterpret_cast<uintptr_t>(dev_ascii01) & 15) == 0);   // a sliced tensor may be unaligned: byte loads then
        e = cudaGetLastError();
        ctx->n_launch++;
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { set_error("e2i_da_load_device: %s", cudaGetErrorString(e)); e2i_bits_free(b); return E2I_ERR_CUDA; }
    *out = b;
    return E2I_OK;
}

extern "C" int e2i_da_load(e2i_ctx *ctx, const uint8_t *host_ascii01, uint64_t n, e2i_bits **out) {
    if (!ctx || !out || (n && !host_ascii01)) { set_error("e2i_da_load: null argument"); return E2I_ERR_ARG; }
    E2I_CUDA_TRY(cudaSetDevice(ctx->device));
    uint8_t *d = nullptr;
    E2I_CUDA_TRY(dmalloc(ctx, &d, n + 16));
    cudaError_t e = cudaMemcpyAsync(d, host_ascii01, n, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) { dfree(ctx, d); set_error("H2D copy failed: %s", cudaGetErrorString(e)); return E2I_ERR_CUDA; }
    ctx->n_h2d += n;
    const int rc = e2i_da_load_device(ctx, d, n, out);
    dfree(ctx, d);
    return rc;
}

extern "C" int e2i_bits_fetch(e2i_ctx *ctx, const e2i_bits *b, uint64_t *host_words, uint64_t n_words) {
    if (!ctx || !b || !host_words) { set_error("e2i_bits_fetch: null argument"); return E2I_ERR_ARG; }
    const uint64_t have = b->n_words32 / 2;
    const uint64_t k = n_words < have ? n_words : have;
    E2I_CUDA_TRY(cudaMemcpy(host_words, b->words, k * 8, cudaMemcpyDeviceToHost));
    for (uint64_t i = k; i < n_words; ++i) host_words[i] = 0;
    return E2I_OK;
}

extern "C" uint64_t e2i_bits_size(const e2i_bits *b) { return b ? b->n : 0; }

extern "C" int e2i_bits_device(const e2i_bits *b, void **dev_words, uint64_t *words32) {
    if (!b) { set_error("e2i_bits_device: null argument"); return E2I_ERR_ARG; }
    if (dev_words) *dev_words = b->words;
    if (words32) *words32 = b->n_words32;
    return E2I_OK;
}

extern "C" void e2i_bits_free(e2i_bits *b) {
    if (!b) return;
    dfree(b->ctx, b->words);
    dfree(b->ctx, b->rank512);
    delete b;
}
