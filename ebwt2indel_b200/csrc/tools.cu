// tools.cu -- libe2i_tools.so: GPU helpers for building SYNTHETIC eBWT inputs (bench / test tooling).
//
// Not part of the product path (the reference takes its eBWT from external tools, README.md:38,
// 91-92; SURVEY.md §8f lists construction as a "next" item).  BCR-style column insertion: iteration
// k merges the m symbols preceding the length-(k+1) suffixes into BWT_k at their (sorted) target
// positions.  The LF step uses the product's own index + batched-rank kernels (libe2i.so); this
// file only holds the merge: mark the target positions in a bitvector, count marks per 4096-bit
// group, and copy old / inserted symbols to their final places.
#include "bcr_kernels.cuh"

using namespace e2i_bcr;

// bits: zeroed, n_words (a multiple of 128) 32-bit words covering S + m positions; groups: n_words / 128 entries.
extern "C" int e2i_tool_bcr_mark(const long long *dev_pos, unsigned long long m, unsigned int *dev_bits,
                                 unsigned long long n_words, long long *dev_groups) {
    if (n_words % kGroupWords) return 3;
    if (m) mark_kernel<<<(unsigned)((m + 255) / 256), 256>>>(dev_pos, m, dev_bits);
    group_popc_kernel<<<(unsigned)(n_words / kGroupWords), kGroupWords>>>(dev_bits, dev_groups);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// out[p] = marked(p) ? new_sym[#marks before p] : old_sym[p - #marks before p]; optional aux byte arrays move alike.
extern "C" int e2i_tool_bcr_merge(const uint8_t *dev_old, const unsigned int *dev_bits, const long long *dev_group_excl,
                                  const uint8_t *dev_new, unsigned long long n_out, unsigned long long n_words, uint8_t *dev_out,
                                  const uint8_t *dev_old_aux, const uint8_t *dev_new_aux, uint8_t *dev_out_aux) {
    if (n_words % kGroupWords) return 3;
    merge_kernel<<<(unsigned)(n_words / kGroupWords), kGroupWords>>>(dev_old, dev_bits, dev_group_excl, dev_new, n_out, dev_out,
                                                                    dev_old_aux, dev_new_aux, dev_out_aux);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// ---- random-sector gather micro-benchmark (the roofline BASELINE.json asks to report against) ----
// Every thread reads `sector` bytes (32, 64 or 128) at independent pseudo-random, sector-aligned
// offsets of a large buffer: the access pattern of an UNSORTED rank-query stream.
namespace {
__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

template <int SECTOR>
__global__ void gather_bench_kernel(const uint4 *__restrict__ buf, unsigned long long n_sectors, unsigned long long per_thread,
                                    unsigned long long seed, unsigned int *__restrict__ sink) {
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int acc = 0;
    // K independent sectors per thread (<= 128 data registers), all in flight before the first use;
    // n_sectors < 2^32, so a multiply-shift replaces the modulo
    constexpr int K = SECTOR == 128 ? 4 : 8;
    const uint4 *p[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const unsigned long long h = mix64(seed + t * K + k) >> 32;
        p[k] = buf + ((h * n_sectors) >> 32) * (SECTOR / 16);
    }
    uint4 v[K][SECTOR / 16];
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int q = 0; q < SECTOR / 16; ++q) v[k][q] = __ldg(p[k] + q);
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int q = 0; q < SECTOR / 16; ++q) acc ^= v[k][q].x ^ v[k][q].y ^ v[k][q].z ^ v[k][q].w;
    (void)per_thread;
    if (acc == 0x12345678u) sink[0] = acc;     // keeps the loads alive
}
}  // namespace

extern "C" int e2i_tool_gather_bench(const void *dev_buf, unsigned long long n_bytes, int sector, unsigned long long n_access,
                                     unsigned int *dev_sink, float *ms) {
    const unsigned long long per_thread = sector == 128 ? 4 : 8, threads = (n_access + per_thread - 1) / per_thread;
    const unsigned grid = (unsigned)((threads + 255) / 256);
    const unsigned long long n_sectors = n_bytes / (unsigned)sector;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {            // first pass warms up
        cudaEventRecord(e0);
        const uint4 *b = static_cast<const uint4 *>(dev_buf);
        if (sector == 32) gather_bench_kernel<32><<<grid, 256>>>(b, n_sectors, per_thread, 1234567ull + rep, dev_sink);
        else if (sector == 64) gather_bench_kernel<64><<<grid, 256>>>(b, n_sectors, per_thread, 1234567ull + rep, dev_sink);
        else if (sector == 128) gather_bench_kernel<128><<<grid, 256>>>(b, n_sectors, per_thread, 1234567ull + rep, dev_sink);
        else return 3;
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
    }
    cudaEventElapsedTime(ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
