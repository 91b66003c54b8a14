// context.cu -- device context, error reporting, parameter defaults, the frame allocator and the
// whole-path entry points e2i_run / e2i_run_device.
//
// e2i_run replaces run_one_dataset / run_two_datasets / run_two_datasets_da
// (/root/reference/ebwt2InDel.cpp:1584-1674, 1344-1465, 1471-1579): load + index the eBWT(s),
// traverse (phases 2-3), scan clusters and extract contexts (phase 4), format the .snp text.
#include <sys/stat.h>

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
    text_cache_trim();
    for (void *r : ctx->ring) if (r) cudaFreeHost(r);
    for (auto &ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
    delete ctx;
}

extern "C" int e2i_trim(e2i_ctx *ctx) {
    if (!ctx) { set_error("e2i_trim: null context"); return E2I_ERR_ARG; }
    E2I_CUDA_TRY(cudaSetDevice(ctx->device));
    arena_release(ctx);
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
namespace {
// phases 2-4 + text on indexes that already exist
int run_with_indexes(e2i_ctx *ctx, e2i_index *b1, e2i_index *b2, e2i_bits *da, const e2i_params *p, char **snp, size_t *snp_len, e2i_stats *st) {
    e2i_bits *da_nav = nullptr;
    e2i_lcpbits *lcp = nullptr;
    const auto w0 = std::chrono::steady_clock::now();
    int rc = e2i_navigate(ctx, b1, b2, p, &lcp, b2 ? &da_nav : nullptr, st);
    // phase 4 and the text in one go: the records are classified and printed where they are, in HBM
    // (st->ms_format: the formatting kernels + the copy of the text)
    if (rc == E2I_OK) rc = e2i_call_snp(ctx, b1, b2, b2 ? da_nav : da, lcp, p, 0, UINT64_MAX, 1, snp, snp_len, st);
    if (std::getenv("E2I_DEBUG"))
        std::fprintf(stderr, "[e2i] run: %.1f ms of host wall time (phases: index %.1f leaves %.1f nodes %.1f call %.1f format %.1f)\n",
                     std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count(),
                     st->ms_index, st->ms_leaves, st->ms_nodes, st->ms_call, st->ms_format);
    e2i_lcpbits_free(lcp);
    e2i_bits_free(da_nav);
    return rc;
}

// the index of one input: from the packed sidecar when E2I_INDEX_CACHE is set and a valid one exists, else
// streamed from the ASCII file (and saved as a sidecar when the cache is on)
int index_from_file(e2i_ctx *ctx, const char *path, uint8_t term, e2i_index **out, uint64_t *bad) {
    const char *cache = std::getenv("E2I_INDEX_CACHE");
    const bool use_cache = cache && *cache && std::strcmp(cache, "0") != 0;
    const std::string side = std::string(path) + ".e2ix";
    if (use_cache) {
        struct stat a, b;
        if (::stat(path, &a) == 0 && ::stat(side.c_str(), &b) == 0 && b.st_mtime >= a.st_mtime) {
            e2i_index *ix = nullptr;
            if (e2i_index_load(ctx, side.c_str(), &ix) == E2I_OK) {
                if (e2i_index_size(ix) == (uint64_t)a.st_size && ix->term == term) { *out = ix; return E2I_OK; }
                e2i_index_free(ix);
            }
        }
    }
    E2I_TRY(e2i_index_build_file(ctx, path, term, out, bad));
    if (use_cache && e2i_index_save(*out, side.c_str()) != E2I_OK)
        std::fprintf(stderr, "[e2i] warning: %s\n", e2i_last_error());
    return E2I_OK;
}

struct IndexTimer {             // device time of the index phase (uploads overlap with the counting pass)
    e2i_ctx *ctx; e2i_stats *st;
    IndexTimer(e2i_ctx *c, e2i_stats *s) : ctx(c), st(s) { cudaEventRecord(ctx->ev[6], ctx->stream); }
    void stop() {
        cudaEventRecord(ctx->ev[7], ctx->stream);
        if (cudaEventSynchronize(ctx->ev[7]) == cudaSuccess) { float ms = 0; if (cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7]) == cudaSuccess) st->ms_index += ms; }
    }
};
}  // namespace

extern "C" int e2i_run_device(e2i_ctx *ctx, const uint8_t *dev_bwt1, uint64_t n1, const uint8_t *dev_bwt2, uint64_t n2,
                              const uint8_t *dev_da, const e2i_params *p, char **snp, size_t *snp_len, e2i_stats *st) {
    if (!ctx || !dev_bwt1 || !p || !snp || !snp_len || !st) { set_error("e2i_run_device: null argument"); return E2I_ERR_ARG; }
    if (dev_bwt2 && dev_da) { set_error("Document array (-d) can only be used with one input BWT file (-1)"); return E2I_ERR_ARG; }
    E2I_CUDA_TRY(cudaSetDevice(ctx->device));
    const auto w0 = std::chrono::steady_clock::now();
    e2i_index *b1 = nullptr, *b2 = nullptr;
    e2i_bits *da = nullptr;
    uint64_t bad = 0;
    int rc;
    {
        Accounting acct(ctx, st);
        IndexTimer t(ctx, st);
        rc = e2i_index_build_device(ctx, dev_bwt1, n1, (uint8_t)p->term, &b1, &bad);
        if (rc == E2I_OK && dev_bwt2) rc = e2i_index_build_device(ctx, dev_bwt2, n2, (uint8_t)p->term, &b2, &bad);
        if (rc == E2I_OK && dev_da) rc = e2i_da_load_device(ctx, dev_da, n1, &da);
        t.stop();
    }
    if (rc == E2I_OK) rc = run_with_indexes(ctx, b1, b2, da, p, snp, snp_len, st);
    e2i_bits_free(da); e2i_index_free(b1); e2i_index_free(b2);
    st->ms_wall += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
    return rc;
}

extern "C" int e2i_run(e2i_ctx *ctx, const uint8_t *host_bwt1, uint64_t n1, const uint8_t *host_bwt2, uint64_t n2,
                       const uint8_t *host_da, const e2i_params *p, char **snp, size_t *snp_len, e2i_stats *st) {
    if (!ctx || !host_bwt1 || !p || !snp || !snp_len || !st) { set_error("e2i_run: null argument"); return E2I_ERR_ARG; }
    if (host_bwt2 && host_da) { set_error("Document array (-d) can only be used with one input BWT file (-1)"); return E2I_ERR_ARG; }
    E2I_CUDA_TRY(cudaSetDevice(ctx->device));
    const auto w0 = std::chrono::steady_clock::now();
    e2i_index *b1 = nullptr, *b2 = nullptr;
    e2i_bits *da = nullptr;
    uint64_t bad = 0;
    int rc;
    {   // the eBWT goes up in chunks on the copy stream while the counting pass of the index build follows it
        Accounting acct(ctx, st);
        IndexTimer t(ctx, st);
        rc = e2i_index_build(ctx, host_bwt1, n1, (uint8_t)p->term, &b1, &bad);
        st->ms_h2d += ctx->last_h2d_ms;
        if (rc == E2I_OK && host_bwt2) { rc = e2i_index_build(ctx, host_bwt2, n2, (uint8_t)p->term, &b2, &bad); st->ms_h2d += ctx->last_h2d_ms; }
        if (rc == E2I_OK && host_da) rc = e2i_da_load(ctx, host_da, n1, &da);
        t.stop();
    }
    if (rc == E2I_OK) rc = run_with_indexes(ctx, b1, b2, da, p, snp, snp_len, st);
    e2i_bits_free(da); e2i_index_free(b1); e2i_index_free(b2);
    st->ms_wall += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
    return rc;
}

// The whole path from files: what bin/ebwt2InDel calls.  Disk reads, PCIe copies and the counting pass of the
// index build overlap (e2i_index_build_file); with E2I_INDEX_CACHE=1 the packed index is kept next to the
// input as <file>.e2ix and later runs upload that instead (half the bytes, no build).
extern "C" int e2i_run_files(e2i_ctx *ctx, const char *path_bwt1, const char *path_bwt2, const char *path_da, const e2i_params *p,
                             char **snp, size_t *snp_len, e2i_stats *st, uint64_t *n1_out, uint64_t *n2_out, uint64_t *bad_pos) {
    if (!ctx || !path_bwt1 || !p || !snp || !snp_len || !st) { set_error("e2i_run_files: null argument"); return E2I_ERR_ARG; }
    if (path_bwt2 && path_da) { set_error("Document array (-d) can only be used with one input BWT file (-1)"); return E2I_ERR_ARG; }
    E2I_CUDA_TRY(cudaSetDevice(ctx->device));
    const auto w0 = std::chrono::steady_clock::now();
    e2i_index *b1 = nullptr, *b2 = nullptr;
    e2i_bits *da = nullptr;
    uint64_t bad = 0;
    int rc, which = 1;
    {
        Accounting acct(ctx, st);
        IndexTimer t(ctx, st);
        rc = index_from_file(ctx, path_bwt1, (uint8_t)p->term, &b1, &bad);
        if (rc == E2I_OK && path_bwt2) { which = 2; rc = index_from_file(ctx, path_bwt2, (uint8_t)p->term, &b2, &bad); }
        if (rc == E2I_OK && path_da) rc = e2i_da_load_file(ctx, path_da, b1->n, &da);
        t.stop();
    }
    if (rc == E2I_ERR_SYMBOL && bad_pos) { bad_pos[0] = bad; bad_pos[1] = (uint64_t)which; }
    if (n1_out) *n1_out = b1 ? b1->n : 0;
    if (n2_out) *n2_out = b2 ? b2->n : 0;
    const auto w1 = std::chrono::steady_clock::now();
    if (rc == E2I_OK) rc = run_with_indexes(ctx, b1, b2, da, p, snp, snp_len, st);
    const auto w2 = std::chrono::steady_clock::now();
    e2i_bits_free(da); e2i_index_free(b1); e2i_index_free(b2);
    st->ms_wall += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
    if (std::getenv("E2I_DEBUG"))
        std::fprintf(stderr, "[e2i] run_files: ingest + index %.1f ms, traversal + calls + text %.1f ms, frees %.1f ms (host wall)\n",
                     std::chrono::duration<double, std::milli>(w1 - w0).count(), std::chrono::duration<double, std::milli>(w2 - w1).count(),
                     std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w2).count());
    return rc;
}

extern "C" int e2i_host_alloc(uint64_t bytes, void **out) {
    if (!out) { set_error("e2i_host_alloc: null argument"); return E2I_ERR_ARG; }
    cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1);
    if (e != cudaSuccess) { cudaGetLastError(); set_error("e2i_host_alloc(%llu bytes): %s", (unsigned long long)bytes, cudaGetErrorString(e)); return E2I_ERR_CUDA; }
    return E2I_OK;
}

extern "C" void e2i_host_free(void *p) { if (p) cudaFreeHost(p); }
