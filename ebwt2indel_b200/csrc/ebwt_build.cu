// ebwt_build.cu -- SURVEY.md §8 (f2): the eBWT (and document array) of a read collection, built on the GPU.
//
// The reference does not build its input: README.md:38, 91-92 sends the user to BCR_LCP_GSA / egap / eGSA
// first.  e2i_ebwt_build makes the tool self-contained from the reads: same convention as those tools
// ('#'_i < '#'_j for i < j, '#' < A < C < G < T; raw ASCII, one byte per symbol; DA = ASCII '0' / '1').
//
// BCR-style column insertion (Bauer, Cox, Rosone 2013).  After iteration k the array holds the symbols
// preceding all read suffixes of length <= k, in suffix order, and P[i] is the position of the newest suffix
// of read order[i] (P ascending).  Iteration k: the LF step of all m reads is ONE batched rank query on the
// product's own index of the current array (index.cu); a stable 4-way partition by symbol keeps the new
// positions sorted, so the insertion is one monotone merge (bcr_kernels.cuh).  L iterations of O(current
// length) streaming work; reads of one fixed length L.
#include <cub/cub.cuh>

#include "bcr_kernels.cuh"
#include "common.cuh"

namespace e2i {

__device__ __forceinline__ int base_code(uint8_t ch) { return ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : 4; }

// iteration "-1": the m suffixes '#'_i, preceded by the last symbol of read i; every symbol is checked once here
__global__ void bcr_init_kernel(const uint8_t *__restrict__ reads, uint64_t m, uint32_t L, uint64_t second_from, uint8_t *__restrict__ bwt,
                                uint8_t *__restrict__ owner, long long *__restrict__ P, uint32_t *__restrict__ order,
                                unsigned long long *bad) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const uint8_t *r = reads + i * L;
    for (uint32_t j = 0; j < L; ++j)
        if (base_code(r[j]) > 3) atomicMin(bad, (unsigned long long)(i * L + j));
    bwt[i] = r[L - 1];
    if (owner) owner[i] = i >= second_from;
    P[i] = (long long)i;
    order[i] = (uint32_t)i;
}

// LF: position of the length-(k+1) suffix of read order[i] = m + #(symbols < c) + rank_c(P[i])
__global__ void bcr_newpos_kernel(const uint8_t *__restrict__ reads, uint64_t m, uint32_t L, uint32_t col, const uint32_t *__restrict__ order,
                                  const uint64_t *__restrict__ ranks4, ulonglong4 F, uint8_t *__restrict__ code, long long *__restrict__ newP,
                                  uint32_t *__restrict__ iota) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int c = base_code(reads[(uint64_t)order[i] * L + col]) & 3;
    const uint64_t f = c == 0 ? F.x : c == 1 ? F.y : c == 2 ? F.z : F.w;
    code[i] = (uint8_t)c;
    newP[i] = (long long)(m + f + ranks4[i * 4 + c]);
    iota[i] = (uint32_t)i;
}

// after the stable partition: permute order / newP and fetch the symbols (and owners) to insert
__global__ void bcr_gather_kernel(const uint8_t *__restrict__ reads, uint64_t m, uint32_t L, int col, uint8_t term, uint64_t second_from,
                                  const uint32_t *__restrict__ perm, const uint32_t *__restrict__ order, const long long *__restrict__ newP,
                                  uint32_t *__restrict__ order2, long long *__restrict__ newP2, uint8_t *__restrict__ new_sym,
                                  uint8_t *__restrict__ new_own) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const uint32_t i = perm[j], r = order[i];
    order2[j] = r;
    newP2[j] = newP[i];
    new_sym[j] = col >= 0 ? reads[(uint64_t)r * L + col] : term;
    if (new_own) new_own[j] = r >= second_from;
}

__global__ void bcr_da_ascii_kernel(uint8_t *__restrict__ owner, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) owner[i] = owner[i] ? '1' : '0';
}

}  // namespace e2i

using namespace e2i;

extern "C" int e2i_ebwt_build(e2i_ctx *ctx, const uint8_t *host_reads, uint64_t m, uint32_t L, uint64_t second_from, uint8_t term,
                              uint8_t *host_bwt, uint8_t *host_da) {
    if (!ctx || !host_reads || !host_bwt || m == 0 || L == 0) { set_error("e2i_ebwt_build: bad argument"); return E2I_ERR_ARG; }
    if (m >= 0x7fffffffull) { set_error("e2i_ebwt_build: more than 2^31 - 1 reads"); return E2I_ERR_ARG; }
    E2I_CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const uint64_t total = m * (uint64_t)(L + 1);
    const bool want_da = host_da != nullptr;
    uint8_t *reads = nullptr, *bwt[2] = {nullptr, nullptr}, *own[2] = {nullptr, nullptr}, *code = nullptr, *code2 = nullptr, *new_sym = nullptr, *new_own = nullptr;
    long long *P = nullptr, *newP = nullptr, *groups = nullptr, *gex = nullptr;
    uint32_t *order = nullptr, *order2 = nullptr, *iota = nullptr, *perm = nullptr, *bits = nullptr;
    uint64_t *ranks = nullptr;
    unsigned long long *bad = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    auto cleanup = [&] {
        for (void *p : {(void *)reads, (void *)bwt[0], (void *)bwt[1], (void *)own[0], (void *)own[1], (void *)code, (void *)code2, (void *)new_sym, (void *)new_own,
                        (void *)P, (void *)newP, (void *)groups, (void *)gex, (void *)order, (void *)order2, (void *)iota, (void *)perm, (void *)bits,
                        (void *)ranks, (void *)bad, tmp})
            dfree(ctx, p);
    };
#define TRYB(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { set_error("e2i_ebwt_build: CUDA error %s at line %d: %s", cudaGetErrorName(_e), __LINE__, cudaGetErrorString(_e)); cleanup(); return E2I_ERR_CUDA; } } while (0)
    const uint64_t n_words_max = (((total + 31) / 32 + e2i_bcr::kGroupWords - 1) / e2i_bcr::kGroupWords) * e2i_bcr::kGroupWords;
    TRYB(dmalloc(ctx, &reads, m * L + 16));
    TRYB(dmalloc(ctx, &bwt[0], total + 16));
    TRYB(dmalloc(ctx, &bwt[1], total + 16));
    if (want_da) { TRYB(dmalloc(ctx, &own[0], total + 16)); TRYB(dmalloc(ctx, &own[1], total + 16)); TRYB(dmalloc(ctx, &new_own, m)); }
    TRYB(dmalloc(ctx, &code, m)); TRYB(dmalloc(ctx, &code2, m)); TRYB(dmalloc(ctx, &new_sym, m));
    TRYB(dmalloc(ctx, &P, m * 8)); TRYB(dmalloc(ctx, &newP, m * 8));
    TRYB(dmalloc(ctx, &order, m * 4)); TRYB(dmalloc(ctx, &order2, m * 4)); TRYB(dmalloc(ctx, &iota, m * 4)); TRYB(dmalloc(ctx, &perm, m * 4));
    TRYB(dmalloc(ctx, &ranks, m * 32));
    TRYB(dmalloc(ctx, &bits, n_words_max * 4));
    TRYB(dmalloc(ctx, &groups, (n_words_max / e2i_bcr::kGroupWords) * 8));
    TRYB(dmalloc(ctx, &gex, (n_words_max / e2i_bcr::kGroupWords) * 8));
    TRYB(dmalloc(ctx, &bad, 8));
    {   // scratch of the two CUB primitives
        size_t a = 0, b = 0;
        TRYB(cub::DeviceRadixSort::SortPairs(nullptr, a, code, code2, iota, perm, (int)m, 0, 2, s));
        TRYB(cub::DeviceScan::ExclusiveSum(nullptr, b, groups, gex, (int)(n_words_max / e2i_bcr::kGroupWords), s));
        tmp_bytes = std::max(a, b);
        TRYB(dmalloc(ctx, &tmp, tmp_bytes));
    }
    TRYB(cudaMemcpyAsync(reads, host_reads, m * L, cudaMemcpyHostToDevice, s));
    TRYB(cudaMemsetAsync(bad, 0xff, 8, s));
    const unsigned gm = (unsigned)((m + 255) / 256);
    bcr_init_kernel<<<gm, 256, 0, s>>>(reads, m, L, second_from, bwt[0], want_da ? own[0] : nullptr, P, order, bad);
    unsigned long long hbad = 0;
    TRYB(cudaMemcpyAsync(&hbad, bad, 8, cudaMemcpyDeviceToHost, s));
    TRYB(cudaStreamSynchronize(s));
    if (hbad != ~0ull) {
        set_error("forbidden character in read %llu at offset %llu: only A,C,G,T are admitted", hbad / L, hbad % L);
        cleanup();
        return E2I_ERR_SYMBOL;
    }
    int cur = 0;
    uint64_t S = m;
    for (uint32_t k = 0; k < L; ++k) {
        // LF of every read's newest suffix on the index of the current array
        e2i_index *ix = nullptr;
        uint64_t badpos = 0;
        int rc = e2i_index_build_device(ctx, bwt[cur], S, term, &ix, &badpos);
        if (rc != E2I_OK) { cleanup(); return rc; }
        float ms = 0;
        rc = e2i_rank_batch_device(ctx, ix, reinterpret_cast<const uint64_t *>(P), m, ranks, &ms);
        uint64_t F[4];
        if (rc == E2I_OK) rc = e2i_index_F(ix, F);
        e2i_index_free(ix);
        if (rc != E2I_OK) { cleanup(); return rc; }
        bcr_newpos_kernel<<<gm, 256, 0, s>>>(reads, m, L, L - 1 - k, order, ranks, make_ulonglong4(F[0], F[1], F[2], F[3]), code, newP, iota);
        // stable 4-way partition by symbol (a 2-bit radix sort): the new positions come out ascending
        TRYB(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, code, code2, iota, perm, (int)m, 0, 2, s));
        bcr_gather_kernel<<<gm, 256, 0, s>>>(reads, m, L, k + 1 < L ? (int)(L - 2 - k) : -1, term, second_from, perm, order, newP, order2, P,
                                             new_sym, want_da ? new_own : nullptr);
        std::swap(order, order2);
        // merge: old symbols keep their order, the m new ones go to P (ascending)
        const uint64_t n_out = S + m;
        const uint64_t n_words = (((n_out + 31) / 32 + e2i_bcr::kGroupWords - 1) / e2i_bcr::kGroupWords) * e2i_bcr::kGroupWords;
        const unsigned n_groups = (unsigned)(n_words / e2i_bcr::kGroupWords);
        TRYB(cudaMemsetAsync(bits, 0, n_words * 4, s));
        e2i_bcr::mark_kernel<<<gm, 256, 0, s>>>(P, m, bits);
        e2i_bcr::group_popc_kernel<<<n_groups, e2i_bcr::kGroupWords, 0, s>>>(bits, groups);
        TRYB(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, groups, gex, (int)n_groups, s));
        e2i_bcr::merge_kernel<<<n_groups, e2i_bcr::kGroupWords, 0, s>>>(bwt[cur], bits, gex, new_sym, n_out, bwt[cur ^ 1],
                                                                      want_da ? own[cur] : nullptr, want_da ? new_own : nullptr,
                                                                      want_da ? own[cur ^ 1] : nullptr);
        TRYB(cudaGetLastError());
        ctx->n_launch += 8;
        cur ^= 1;
        S = n_out;
    }
    TRYB(cudaMemcpyAsync(host_bwt, bwt[cur], total, cudaMemcpyDeviceToHost, s));
    if (want_da) {
        bcr_da_ascii_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(own[cur], total);
        TRYB(cudaMemcpyAsync(host_da, own[cur], total, cudaMemcpyDeviceToHost, s));
    }
    TRYB(cudaStreamSynchronize(s));
#undef TRYB
    cleanup();
    return E2I_OK;
}
