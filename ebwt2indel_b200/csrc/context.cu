// context.cu -- device context, error reporting, parameter defaults, the frame allocator and the
// whole-path entry points e2i_run / e2i_run_device.
//
// e2i_run replaces run_one_dataset / run_two_datasets / run_two_datasets_da
// (/root/reference/ebwt2InDel.cpp:1584-1674, 1344-1465, 1471-1579): load + index the eBWT(s),
// traverse (phases 2-3), scan clusters and extract contexts (phase 4), format the .snp text.
#include <algorithm>
#include <chrono>

#include "common.cuh"

namespace e2i {

static thread_local std::string g_last_error;

void set_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

// ---- Arena -----------------------------------------------------------------------------------
void *Arena::alloc(int side, size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (bytes > hi_ - lo_) return nullptr;
    size_t off;
    if (side == 0) { off = lo_; lo_ += bytes; }
    else { hi_ -= bytes; off = hi_; }
    live_[side].push_back({off, bytes, false});
    return base_ + off;
}

void Arena::free(int side, void *p) {
    const size_t off = (size_t)(static_cast<char *>(p) - base_);
    std::vector<Blk> &v = live_[side];
    for (size_t i = v.size(); i-- > 0;)
        if (v[i].off == off) { v[i].freed = true; break; }
    while (!v.empty() && v.back().freed) {           // pop every released frame at the top of this end
        if (side == 0) lo_ -= v.back().bytes; else hi_ += v.back().bytes;
        v.pop_back();
    }
}

}  // namespace e2i

using namespace e2i;

extern "C" const char *e2i_last_error(void) { return g_last_error.c_str(); }
extern "C" const char *e2i_version(void) { return "ebwt2indel_b200 0.1.0 (sm_100a)"; }

// globals of ebwt2InDel.cpp:20-74
extern "C" void e2i_params_default(e2i_params *p) {
    if (!p) return;
    p->k_left = 31; p->k_right = 30; p->K = 16; p->max_gap = 10; p->max_snvs = 2; p->mcov_out = 3;
    p->complexity = 20;          // complexity_def = k_right_def - 10, computed from the DEFAULT -R (:64)
    p->max_variants_per_position = 0;
    p->term = '#';
}

// "0 means default" (ebwt2InDel.cpp:1740-1746); -q and -t are taken as given
extern "C" void e2i_params_resolve(e2i_params *p) {
    if (!p) return;
    e2i_params d;
    e2i_params_default(&d);
    if (p->complexity == 0) p->complexity = d.complexity;
    if (p->K == 0) p->K = d.K;
    if (p->max_gap == 0) p->max_gap = d.max_gap;
    if (p->k_left == 0) p->k_left = d.k_left;
    if (p->k_right == 0) p->k_right = d.k_right;
    if (p->max_snvs == 0) p->max_snvs = d.max_snvs;
    if (p->mcov_out == 0) p->mcov_out = d.mcov_out;
}

extern "C" int e2i_create(int device, e2i_ctx **out) {
    if (!out) { set_error("e2i_create: null argument"); return E2I_ERR_ARG; }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        set_error("e2i_create: no CUDA device available (%s); this library has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return E2I_ERR_CUDA;
    }
    if (device < 0 || device >= count) { set_error("e2i_create: device %d out of range [0,%d)", device, count); return E2I_ERR_ARG; }
    E2I_CUDA_TRY(cudaSetDevice(device));
    e2i_ctx *ctx = new e2i_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    E2I_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    E2I_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    E2I_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    {   // a pool of our own (the process-wide default pool is left alone); freed blocks stay cached in it
        // until e2i_trim / e2i_destroy
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        E2I_CUDA_TRY(cudaMemPoolCreate(&ctx->pool, &props));
        uint64_t thr = UINT64_MAX;
        E2I_CUDA_TRY(cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &thr));
    }
    for (auto &ev : ctx->ev) E2I_CUDA_TRY(cudaEventCreate(&ev));
    E2I_CUDA_TRY(cudaMalloc(&ctx->ctl, 16384 * 64));   // kSweepSlots x sizeof(SweepDev), navigate.cu
    E2I_CUDA_TRY(cudaHostAlloc(&ctx->ctl_host, 4096, cudaHostAllocMapped | cudaHostAllocPortable));
    std::memset(ctx->ctl_host, 0, 4096);
    *out = ctx;
    return E2I_OK;
}

extern "C" void e2i_destroy(e2i_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    dfree(ctx, ctx->desc);
    ctx->desc = nullptr;
    e2i_trim(ctx);               // frees the frame arena and returns the cached pool memory
    cudaFree(ctx->ctl);
    cudaFreeHost(ctx->ctl_host);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    for (auto &ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
    delete ctx;
}

extern "C" int e2i_trim(e2i_ctx *ctx) {
    if (!ctx) { set_error("e2i_trim: null context"); return E2I_ERR_ARG; }
    E2I_CUDA_TRY(cudaSetDevice(ctx->device));
    dfree(ctx, ctx->arena_mem);
    ctx->arena_mem = nullptr;
    ctx->arena_bytes = 0;
    ctx->arena.reset(nullptr, 0);
    E2I_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    E2I_CUDA_TRY(cudaMemPoolTrimTo(ctx->pool, 0));
    return E2I_OK;
}

extern "C" void *e2i_stream(const e2i_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

extern "C" int e2i_set_frontier_budget(e2i_ctx *ctx, uint64_t bytes) {
    if (!ctx) { set_error("e2i_set_frontier_budget: null context"); return E2I_ERR_ARG; }
    ctx->frontier_budget = bytes;
    return E2I_OK;
}

// ---- whole path ------------------------------------------------------------------------------
static int run_on_device(e2i_ctx *ctx, const uint8_t *dev_bwt1, uint64_t n1, const uint8_t *dev_bwt2, uint64_t n2,
                         const uint8_t *dev_da, const e2i_params *p, char **snp, size_t *snp_len, e2i_stats *st,
                         void (*release_inputs)(void *), void *release_arg) {
    e2i_index *b1 = nullptr, *b2 = nullptr;
    e2i_bits *da = nullptr, *da_nav = nullptr;
    e2i_lcpbits *lcp = nullptr;
    e2i_calls *calls = nullptr;
    auto cleanup = [&] {
        e2i_calls_free(calls); e2i_lcpbits_free(lcp); e2i_bits_free(da); e2i_bits_free(da_nav);
        e2i_index_free(b1); e2i_index_free(b2);
    };
    cudaStream_t s = ctx->stream;
    const auto w0 = std::chrono::steady_clock::now();
    cudaEventRecord(ctx->ev[6], s);
    uint64_t bad = 0;
    Accounting *acct = new Accounting(ctx, st);
    int rc = e2i_index_build_device(ctx, dev_bwt1, n1, (uint8_t)p->term, &b1, &bad);
    if (rc == E2I_OK && dev_bwt2) rc = e2i_index_build_device(ctx, dev_bwt2, n2, (uint8_t)p->term, &b2, &bad);
    if (rc == E2I_OK && dev_da) rc = e2i_da_load_device(ctx, dev_da, n1, &da);
    delete acct;   // index + DA build only; navigate and call account for themselves
    cudaEventRecord(ctx->ev[7], s);
    if (rc == E2I_OK) {
        cudaEventSynchronize(ctx->ev[7]);
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7]);
        st->ms_index += ms;
    }
    if (release_inputs) release_inputs(release_arg);   // the ASCII copies are dead once the index exists
    if (rc == E2I_OK) rc = e2i_navigate(ctx, b1, b2, p, &lcp, b2 ? &da_nav : nullptr, st);
    if (rc == E2I_OK) rc = e2i_call(ctx, b1, b2, b2 ? da_nav : da, lcp, p, 0, UINT64_MAX, &calls, st);
    if (std::getenv("E2I_DEBUG")) {
        cudaStreamSynchronize(s);
        std::fprintf(stderr, "[e2i] run: %.1f ms of host wall time before formatting (phases: index %.1f leaves %.1f nodes %.1f call %.1f)\n",
                     std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count(),
                     st->ms_index, st->ms_leaves, st->ms_nodes, st->ms_call);
    }
    const auto w1 = std::chrono::steady_clock::now();
    if (rc == E2I_OK)
        rc = e2i_snp_format(calls->recs, calls->left, calls->right, calls->n, p,
                            (b2 || da) ? 1 : 0, 1, snp, snp_len, st);
    const auto w2 = std::chrono::steady_clock::now();
    cleanup();
    st->ms_format += std::chrono::duration<double, std::milli>(w2 - w1).count();
    st->ms_wall += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
    return rc;
}

extern "C" int e2i_run_device(e2i_ctx *ctx, const uint8_t *dev_bwt1, uint64_t n1, const uint8_t *dev_bwt2, uint64_t n2,
                              const uint8_t *dev_da, const e2i_params *p, char **snp, size_t *snp_len, e2i_stats *st) {
    if (!ctx || !dev_bwt1 || !p || !snp || !snp_len || !st) { set_error("e2i_run_device: null argument"); return E2I_ERR_ARG; }
    if (dev_bwt2 && dev_da) { set_error("Document array (-d) can only be used with one input BWT file (-1)"); return E2I_ERR_ARG; }
    E2I_CUDA_TRY(cudaSetDevice(ctx->device));
    return run_on_device(ctx, dev_bwt1, n1, dev_bwt2, n2, dev_da, p, snp, snp_len, st, nullptr, nullptr);
}

namespace {
struct HostInputs {
    uint8_t *d1 = nullptr, *d2 = nullptr, *dd = nullptr;
};
struct HostInputsRef { HostInputs *in; e2i_ctx *ctx; };
void free_inputs(void *arg) {
    HostInputsRef *r = static_cast<HostInputsRef *>(arg);
    HostInputs *h = r->in;
    dfree(r->ctx, h->d1); dfree(r->ctx, h->d2); dfree(r->ctx, h->dd);
    h->d1 = h->d2 = h->dd = nullptr;
}
}  // namespace

extern "C" int e2i_run(e2i_ctx *ctx, const uint8_t *host_bwt1, uint64_t n1, const uint8_t *host_bwt2, uint64_t n2,
                       const uint8_t *host_da, const e2i_params *p, char **snp, size_t *snp_len, e2i_stats *st) {
    if (!ctx || !host_bwt1 || !p || !snp || !snp_len || !st) { set_error("e2i_run: null argument"); return E2I_ERR_ARG; }
    if (host_bwt2 && host_da) { set_error("Document array (-d) can only be used with one input BWT file (-1)"); return E2I_ERR_ARG; }
    E2I_CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    HostInputs in;
    HostInputsRef ref{&in, ctx};
    auto fail = [&](cudaError_t e) { set_error("e2i_run: %s", cudaGetErrorString(e)); free_inputs(&ref); return E2I_ERR_CUDA; };
    // One stream by default (a single copy engine saturates the link on an idle host).  With
    // E2I_H2D_STREAMS=2 large inputs go up as 256 MB pieces alternating between the context's two
    // streams; measurements on a shared host were too noisy to prefer either (profiles/e2e_time.py).
    const char *hs = std::getenv("E2I_H2D_STREAMS");
    const int n_streams = hs ? std::max(1, std::min(2, atoi(hs))) : 1;
    cudaEvent_t ev_copy = nullptr;
    auto upload = [&](uint8_t *dst, const uint8_t *src, uint64_t n) -> cudaError_t {
        const uint64_t piece = 256ull << 20;
        if (n_streams == 1 || n <= piece) return cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, s);
        int k = 0;
        for (uint64_t off = 0; off < n; off += piece, ++k) {
            const cudaError_t ce = cudaMemcpyAsync(dst + off, src + off, std::min(piece, n - off), cudaMemcpyHostToDevice,
                                                   (k & 1) ? ctx->copy_stream : s);
            if (ce != cudaSuccess) return ce;
        }
        return cudaSuccess;
    };
    cudaError_t e = cudaEventRecord(ctx->ev[6], s);
    if (e == cudaSuccess) e = dmalloc(ctx, &in.d1, n1 + 16);
    if (e == cudaSuccess && host_bwt2) e = dmalloc(ctx, &in.d2, n2 + 16);
    if (e == cudaSuccess && host_da) e = dmalloc(ctx, &in.dd, n1 + 16);
    // the second stream must not start before the allocations (stream-ordered) are done on the first
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev_copy, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(ev_copy, s);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->copy_stream, ev_copy, 0);
    if (e == cudaSuccess) e = upload(in.d1, host_bwt1, n1);
    if (e == cudaSuccess && host_bwt2) e = upload(in.d2, host_bwt2, n2);
    if (e == cudaSuccess && host_da) e = upload(in.dd, host_da, n1);
    if (e == cudaSuccess) e = cudaEventRecord(ev_copy, ctx->copy_stream);       // join the second stream
    if (e == cudaSuccess) e = cudaStreamWaitEvent(s, ev_copy, 0);
    if (e == cudaSuccess) e = cudaEventRecord(ctx->ev[7], s);
    if (e == cudaSuccess) e = cudaEventSynchronize(ctx->ev[7]);
    if (e != cudaSuccess) return fail(e);
    if (ev_copy) cudaEventDestroy(ev_copy);
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7]);
    st->ms_h2d += ms;
    st->h2d_bytes += n1 + (host_bwt2 ? n2 : 0) + (host_da ? n1 : 0);
    const int rc = run_on_device(ctx, in.d1, n1, in.d2, n2, in.dd, p, snp, snp_len, st, free_inputs, &ref);
    free_inputs(&ref);
    return rc;
}

extern "C" int e2i_host_alloc(uint64_t bytes, void **out) {
    if (!out) { set_error("e2i_host_alloc: null argument"); return E2I_ERR_ARG; }
    cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1);
    if (e != cudaSuccess) { cudaGetLastError(); set_error("e2i_host_alloc(%llu bytes): %s", (unsigned long long)bytes, cudaGetErrorString(e)); return E2I_ERR_CUDA; }
    return E2I_OK;
}

extern "C" void e2i_host_free(void *p) { if (p) cudaFreeHost(p); }
