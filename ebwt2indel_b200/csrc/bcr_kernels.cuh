// bcr_kernels.cuh -- the merge step of BCR-style column insertion (Bauer, Cox, Rosone 2013): mark the (sorted)
// target positions of the m new symbols in a bitvector, count marks per 4096-bit group, and copy old /
// inserted symbols to their final places.  Shared by the product's eBWT builder (ebwt_build.cu) and the
// synthetic-input tooling (tools.cu).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

namespace e2i_bcr {

constexpr int kGroupWords = 128;   // 4096 positions per group

__global__ void mark_kernel(const long long *__restrict__ pos, unsigned long long m, unsigned int *__restrict__ bits) {
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    const unsigned long long q = (unsigned long long)pos[t];
    atomicOr(bits + (q >> 5), 1u << (q & 31));
}

__global__ void __launch_bounds__(kGroupWords)
group_popc_kernel(const unsigned int *__restrict__ bits, long long *__restrict__ groups) {
    __shared__ unsigned int s_w[kGroupWords / 32];
    unsigned int c = __popc(bits[(size_t)blockIdx.x * kGroupWords + threadIdx.x]);
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) groups[blockIdx.x] = (long long)s_w[0] + s_w[1] + s_w[2] + s_w[3];
}

// one thread per 32 output positions
__global__ void __launch_bounds__(kGroupWords)
merge_kernel(const uint8_t *__restrict__ old_sym, const unsigned int *__restrict__ bits, const long long *__restrict__ group_excl,
             const uint8_t *__restrict__ new_sym, unsigned long long n_out, uint8_t *__restrict__ out,
             const uint8_t *__restrict__ old_aux, const uint8_t *__restrict__ new_aux, uint8_t *__restrict__ out_aux) {
    __shared__ unsigned int s_w[kGroupWords / 32];
    const size_t w = (size_t)blockIdx.x * kGroupWords + threadIdx.x;
    const unsigned int word = bits[w];
    const unsigned int pc = __popc(word);
    unsigned int incl = pc;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const unsigned int y = __shfl_up_sync(0xffffffffu, incl, s);
        if (lane >= s) incl += y;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    unsigned int before = incl - pc;
    for (int k = 0; k < warp; ++k) before += s_w[k];
    unsigned long long ins = (unsigned long long)group_excl[blockIdx.x] + before;   // inserted symbols before this word
    const unsigned long long p0 = (unsigned long long)w * 32;
    if (p0 >= n_out) return;
    unsigned long long oi = p0 - ins;                                                // old symbols before this word
    const int lim = n_out - p0 < 32 ? (int)(n_out - p0) : 32;
    for (int b = 0; b < lim; ++b) {
        const bool is_new = (word >> b) & 1u;
        out[p0 + b] = is_new ? new_sym[ins] : old_sym[oi];
        if (out_aux) out_aux[p0 + b] = is_new ? new_aux[ins] : old_aux[oi];
        if (is_new) ++ins; else ++oi;
    }
}

}  // namespace e2i_bcr
