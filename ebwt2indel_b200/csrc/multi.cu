// multi.cu -- the whole path on several GPUs of one box, driven from ONE process (no torch, no NCCL).
//
// Replaces the reference's parallel wrapper pebwt2InDel.sh (/root/reference/pebwt2InDel.sh:45-88:
// split the reads into pieces, one ebwt2InDel process per piece, concatenate), without its loss of
// cross-piece coverage: every GPU works on the SAME eBWT, so the output equals the single-GPU run's.
//
// One host thread and one e2i_ctx per GPU; the GPUs talk through peer memory over NVLink:
//   1. index     every GPU uploads, counts and packs ONE tile-aligned slice of the eBWT, then pulls the
//                other slices' blocks from its peers (cudaMemcpyPeerAsync): a replicated index
//   2. traverse  e2i_navigate_shard: subtrees dealt by position-contiguous slices of a shallow frontier
//   3. combine   the one exchange of the path: OR of the LCP (and DA) bit vectors.  One
//                kernel per GPU reads its 1/N slice of the words from every peer, ORs them and stores the
//                result back into every peer's copy -- reduce-scatter and all-gather fused, over P2P loads
//                and stores, no staging buffer
//   4. call      phase 4 on the GPU's own suffix-array range; .snp text per range, concatenated in order
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <map>
#include <memory>
#include <mutex>
#include <thread>

#include "common.cuh"

namespace e2i {

constexpr int kMaxGpus = 16;

struct PeerWords {
    uint32_t *p[kMaxGpus];
    int n;
};

// OR of words [w0, w1) over all ranks, written back to every rank (16-byte vectors; w0, w1 multiples of 4)
__global__ void __launch_bounds__(256)
or_allreduce_kernel(const PeerWords bufs, uint64_t w0, uint64_t w1) {
    const uint64_t v0 = w0 >> 2, v1 = w1 >> 2;
    for (uint64_t v = v0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < v1; v += (uint64_t)gridDim.x * blockDim.x) {
        uint4 acc = make_uint4(0, 0, 0, 0);
#pragma unroll 4
        for (int r = 0; r < bufs.n; ++r) {
            const uint4 x = reinterpret_cast<const uint4 *>(bufs.p[r])[v];
            acc.x |= x.x; acc.y |= x.y; acc.z |= x.z; acc.w |= x.w;
        }
        for (int r = 0; r < bufs.n; ++r) reinterpret_cast<uint4 *>(bufs.p[r])[v] = acc;
    }
}

class HostBarrier {
  public:
    explicit HostBarrier(int n) : n_(n) {}
    void wait() {
        std::unique_lock<std::mutex> lk(m_);
        const uint64_t gen = gen_;
        if (++count_ == n_) { count_ = 0; ++gen_; cv_.notify_all(); }
        else cv_.wait(lk, [&] { return gen_ != gen; });
    }
  private:
    std::mutex m_;
    std::condition_variable cv_;
    int n_, count_ = 0;
    uint64_t gen_ = 0;
};

// ---- communicators ------------------------------------------------------------------------------------
// ranks = threads of this process: a shared barrier, slots in ordinary memory, peer pointers used as they are
struct LocalShared {
    explicit LocalShared(int n) : bar(n), slots((size_t)n * kCommSlotBytes, 0) {}
    HostBarrier bar;
    std::vector<unsigned char> slots;
};
struct LocalComm : e2i_comm {
    std::shared_ptr<LocalShared> sh;
    void barrier() override { sh->bar.wait(); }
    unsigned char *slot(int r) override { return sh->slots.data() + (size_t)r * kCommSlotBytes; }
    void *peer_ptr(int, void *base, const cudaIpcMemHandle_t &) override { return base; }
    bool needs_ipc() const override { return false; }
};

// ranks = processes of one box (torchrun): barrier and slots in a POSIX shared-memory segment, peer device
// memory mapped through CUDA IPC handles (NVLink P2P between the processes' GPUs)
struct ShmHeader {
    std::atomic<uint32_t> ready, arrived, generation, pad;
};
struct ShmComm : e2i_comm {
    std::string name;
    void *map = nullptr;
    size_t bytes = 0;
    std::map<std::string, void *> opened;             // IPC handle bytes -> mapping
    ShmHeader *hdr() { return static_cast<ShmHeader *>(map); }
    void barrier() override {
        ShmHeader *h = hdr();
        const uint32_t gen = h->generation.load(std::memory_order_acquire);
        if (h->arrived.fetch_add(1, std::memory_order_acq_rel) + 1 == (uint32_t)world) {
            h->arrived.store(0, std::memory_order_relaxed);
            h->generation.store(gen + 1, std::memory_order_release);
        } else {
            unsigned spins = 0;
            while (h->generation.load(std::memory_order_acquire) == gen)
                if (++spins > 2000) { sched_yield(); spins = 0; }
        }
    }
    unsigned char *slot(int r) override { return static_cast<unsigned char *>(map) + 4096 + (size_t)r * kCommSlotBytes; }
    void *peer_ptr(int, void *, const cudaIpcMemHandle_t &handle) override {
        const std::string key(reinterpret_cast<const char *>(&handle), sizeof handle);
        auto it = opened.find(key);
        if (it != opened.end()) return it->second;
        void *p = nullptr;
        if (cudaIpcOpenMemHandle(&p, handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        opened[key] = p;
        return p;
    }
    bool needs_ipc() const override { return true; }
    ~ShmComm() override {
        for (auto &kv : opened) cudaIpcCloseMemHandle(kv.second);
        if (map) munmap(map, bytes);
        if (rank == 0 && !name.empty()) shm_unlink(name.c_str());
    }
};

// tile-aligned slices of [0, n): every tile (the last one may hold no symbol, only the block that makes
// rank(n) addressable) has exactly one owner; trailing ranks may be empty
struct Slice { uint64_t begin, len, tiles; };
static void index_slices(uint64_t n, int world, std::vector<Slice> &out, uint64_t &per) {
    const uint64_t tiles = (n / kBlockSyms + 1 + kTileBlocks - 1) / kTileBlocks;
    per = (tiles + world - 1) / world;
    out.resize(world);
    for (int r = 0; r < world; ++r) {
        const uint64_t lo = std::min<uint64_t>(tiles, (uint64_t)r * per), hi = std::min<uint64_t>(tiles, (uint64_t)(r + 1) * per);
        if (hi <= lo) { out[r] = {0, 0, 0}; continue; }
        const uint64_t b = std::min<uint64_t>(n, lo << kTileShift), e = std::min<uint64_t>(n, hi << kTileShift);
        out[r] = {b, e - b, hi - lo};
    }
}

struct MultiShared {
    int world = 0;
    std::vector<e2i_ctx *> ctx;
    std::vector<int> rc;
    std::vector<std::string> err;
    // index build
    std::vector<e2i_index *> ix[2];
    std::vector<uint64_t> counts[2];            // [rank * 4 + k]
    std::vector<std::vector<uint64_t>> super[2];
    // navigate
    std::vector<e2i_lcpbits *> lcp;
    std::vector<e2i_bits *> da_nav, da;
    // call + format
    std::vector<uint64_t> clusters;
    char *final_text = nullptr;          // the whole .snp text (page-locked, text_alloc): every rank copies its piece in
    std::vector<size_t> text_len;
    std::vector<e2i_stats> st;
    bool any_failed() const { for (int r : rc) if (r != E2I_OK) return true; return false; }
};

}  // namespace e2i

using namespace e2i;

extern "C" int e2i_comm_local(int world, e2i_comm **out) {
    if (!out || world < 1 || world > kMaxGpus) { set_error("e2i_comm_local: bad argument"); return E2I_ERR_ARG; }
    auto sh = std::make_shared<LocalShared>(world);
    for (int r = 0; r < world; ++r) {
        LocalComm *c = new LocalComm();
        c->rank = r; c->world = world; c->sh = sh;
        out[r] = c;
    }
    return E2I_OK;
}

extern "C" int e2i_comm_shm(const char *name, int rank, int world, e2i_comm **out) {
    if (!out || !name || world < 1 || world > kMaxGpus || rank < 0 || rank >= world) { set_error("e2i_comm_shm: bad argument"); return E2I_ERR_ARG; }
    ShmComm *c = new ShmComm();
    c->rank = rank; c->world = world; c->name = name;
    c->bytes = 4096 + (size_t)world * kCommSlotBytes;
    int fd = -1;
    if (rank == 0) {
        shm_unlink(name);
        fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd >= 0 && ftruncate(fd, (off_t)c->bytes) != 0) { close(fd); fd = -1; }
    } else {
        for (int tries = 0; tries < 60000 && fd < 0; ++tries) {       // rank 0 creates the segment
            fd = shm_open(name, O_RDWR, 0600);
            if (fd >= 0) { struct stat sb; if (fstat(fd, &sb) != 0 || (size_t)sb.st_size < c->bytes) { close(fd); fd = -1; } }
            if (fd < 0) usleep(1000);
        }
    }
    if (fd < 0) { set_error("e2i_comm_shm: cannot open shared memory segment %s", name); c->name.clear(); delete c; return E2I_ERR_IO; }
    c->map = mmap(nullptr, c->bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (c->map == MAP_FAILED) { c->map = nullptr; set_error("e2i_comm_shm: mmap failed"); delete c; return E2I_ERR_IO; }
    if (rank == 0) {
        std::memset(c->map, 0, c->bytes);
        c->hdr()->ready.store(0x600dc0deu, std::memory_order_release);
    } else {
        for (int tries = 0; c->hdr()->ready.load(std::memory_order_acquire) != 0x600dc0deu; ++tries) {
            if (tries > 60000) { set_error("e2i_comm_shm: rank 0 never initialised %s", name); delete c; return E2I_ERR_IO; }
            usleep(1000);
        }
    }
    *out = c;
    return E2I_OK;
}

extern "C" void e2i_comm_barrier(e2i_comm *c) { if (c) c->barrier(); }
extern "C" void e2i_comm_free(e2i_comm *c) { delete c; }

extern "C" int e2i_enable_peers(e2i_ctx **ctxs, int n) {
    if (!ctxs || n < 1 || n > kMaxGpus) { set_error("e2i_enable_peers: bad argument"); return E2I_ERR_ARG; }
    for (int a = 0; a < n; ++a) {
        E2I_CUDA_TRY(cudaSetDevice(ctxs[a]->device));
        std::vector<cudaMemAccessDesc> desc;
        for (int b = 0; b < n; ++b) {
            if (a == b || ctxs[a]->device == ctxs[b]->device) continue;
            int can = 0;
            E2I_CUDA_TRY(cudaDeviceCanAccessPeer(&can, ctxs[a]->device, ctxs[b]->device));
            if (!can) { set_error("GPU %d cannot access GPU %d as a peer", ctxs[a]->device, ctxs[b]->device); return E2I_ERR_CUDA; }
            const cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[b]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { set_error("cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e)); return E2I_ERR_CUDA; }
            cudaGetLastError();
            // the peers may read and write what this context allocates from its pool
            cudaMemAccessDesc d = {};
            d.location.type = cudaMemLocationTypeDevice;
            d.location.id = ctxs[b]->device;
            d.flags = cudaMemAccessFlagsProtReadWrite;
            desc.push_back(d);
        }
        if (!desc.empty()) E2I_CUDA_TRY(cudaMemPoolSetAccess(ctxs[a]->pool, desc.data(), desc.size()));
    }
    return E2I_OK;
}

// One GPU's part of the OR-combine: words of its 1/N slice from every peer, result to every peer.
extern "C" int e2i_or_allreduce(e2i_ctx *ctx, void *const *dev_words, int n_ranks, int rank, uint64_t words32) {
    if (!ctx || !dev_words || n_ranks < 1 || n_ranks > kMaxGpus || rank < 0 || rank >= n_ranks || (words32 & 3)) {
        set_error("e2i_or_allreduce: bad argument");
        return E2I_ERR_ARG;
    }
    E2I_CUDA_TRY(cudaSetDevice(ctx->device));
    PeerWords pw;
    pw.n = n_ranks;
    for (int r = 0; r < n_ranks; ++r) pw.p[r] = static_cast<uint32_t *>(dev_words[r]);
    const uint64_t vecs = words32 / 4, per = (vecs + n_ranks - 1) / n_ranks;
    const uint64_t w0 = std::min(vecs, per * rank) * 4, w1 = std::min(vecs, per * (rank + 1)) * 4;
    if (w1 > w0) {
        const unsigned grid = (unsigned)std::min<uint64_t>(((w1 - w0) / 4 + 255) / 256, (uint64_t)ctx->sm_count * 8);
        or_allreduce_kernel<<<grid, 256, 0, ctx->stream>>>(pw, w0, w1);
        E2I_CUDA_TRY(cudaGetLastError());
        ctx->n_launch++;
    }
    return E2I_OK;
}

extern "C" int e2i_run_multi(const int *devices, int n_devices, const uint8_t *host_bwt1, uint64_t n1, const uint8_t *host_bwt2,
                             uint64_t n2, const uint8_t *host_da, const e2i_params *p, uint64_t frontier_budget,
                             char **snp, size_t *snp_len, e2i_stats *st_out) {
    if (!devices || n_devices < 1 || n_devices > kMaxGpus || !host_bwt1 || !p || !snp || !snp_len || !st_out) { set_error("e2i_run_multi: bad argument"); return E2I_ERR_ARG; }
    if (host_bwt2 && host_da) { set_error("Document array (-d) can only be used with one input BWT file (-1)"); return E2I_ERR_ARG; }
    const int world = n_devices;
    const bool two = host_bwt2 != nullptr;
    const auto w_start = std::chrono::steady_clock::now();
    MultiShared sh;
    sh.world = world;
    sh.ctx.assign(world, nullptr);
    sh.rc.assign(world, E2I_OK);
    sh.err.assign(world, "");
    for (int b = 0; b < 2; ++b) { sh.ix[b].assign(world, nullptr); sh.counts[b].assign((size_t)world * 4, 0); sh.super[b].resize(world); }
    sh.lcp.assign(world, nullptr);
    sh.da_nav.assign(world, nullptr);
    sh.da.assign(world, nullptr);
    sh.clusters.assign(world, 0);
    sh.text_len.assign(world, 0);
    sh.st.resize(world);
    for (auto &s : sh.st) std::memset(&s, 0, sizeof s);
    auto destroy_all = [&] { for (e2i_ctx *c : sh.ctx) e2i_destroy(c); };
    for (int r = 0; r < world; ++r) {
        const int rc = e2i_create(devices[r], &sh.ctx[r]);
        if (rc != E2I_OK) { destroy_all(); return rc; }
        if (frontier_budget) e2i_set_frontier_budget(sh.ctx[r], frontier_budget);
    }
    if (world > 1) { const int rc = e2i_enable_peers(sh.ctx.data(), world); if (rc != E2I_OK) { destroy_all(); return rc; } }

    HostBarrier bar(world);
    // traversal: position-range sharding with peer pulls by default, E2I_SHARDING=subtree selects the independent
    // subtree shards of round 1
    const char *shard_env = std::getenv("E2I_SHARDING");
    const bool ranged = world > 1 && !(shard_env && std::strcmp(shard_env, "subtree") == 0);
    std::vector<e2i_comm *> comms((size_t)world, nullptr);
    if (ranged) { const int rc = e2i_comm_local(world, comms.data()); if (rc != E2I_OK) { destroy_all(); return rc; } }
    const uint64_t n = n1 + (two ? n2 : 0);
    std::vector<Slice> sl[2];
    uint64_t per[2] = {0, 0};
    index_slices(n1, world, sl[0], per[0]);
    if (two) index_slices(n2, world, sl[1], per[1]);
    const uint8_t *host_bwt[2] = {host_bwt1, host_bwt2};
    const uint64_t nn[2] = {n1, n2};
    const int n_bwt = two ? 2 : 1;

    auto worker = [&](int rank) {
        e2i_ctx *ctx = sh.ctx[rank];
        e2i_stats &st = sh.st[rank];
        cudaSetDevice(ctx->device);
        cudaStream_t s = ctx->stream;
        auto fail = [&](int rc) { sh.rc[rank] = rc; sh.err[rank] = e2i_last_error(); };
        // every collective step is entered by all ranks; a rank that failed earlier just keeps the barriers company
#define STEP(expr) do { if (sh.rc[rank] == E2I_OK) { const int _rc = (expr); if (_rc != E2I_OK) fail(_rc); } } while (0)
#define CUDA_STEP(expr) do { if (sh.rc[rank] == E2I_OK) { const cudaError_t _e = (expr); if (_e != cudaSuccess) { set_error("CUDA error %s: %s", cudaGetErrorName(_e), cudaGetErrorString(_e)); fail(E2I_ERR_CUDA); } } } while (0)
        // ---- 1. index: slice-wise build + pull of the peers' blocks ----
        cudaEventRecord(ctx->ev[6], s);
        uint8_t *dslice[2] = {nullptr, nullptr};
        for (int b = 0; b < n_bwt; ++b) {
            const Slice &my = sl[b][rank];
            STEP(e2i_index_alloc(ctx, nn[b], (uint8_t)p->term, (uint64_t)world, &sh.ix[b][rank]));
            if (my.len) {
                CUDA_STEP(dmalloc(ctx, &dslice[b], my.len + 16));
                CUDA_STEP(cudaMemcpyAsync(dslice[b], host_bwt[b] + my.begin, my.len, cudaMemcpyHostToDevice, s));
                st.h2d_bytes += my.len;
            }
            uint64_t bad = 0;
            if (sh.rc[rank] == E2I_OK) {
                const int rc = e2i_index_slice_count(ctx, sh.ix[b][rank], dslice[b], my.begin, my.len, my.tiles, &sh.counts[b][(size_t)rank * 4], &bad);
                if (rc != E2I_OK) fail(rc);
            }
        }
        bar.wait();                                      // all counts are in
        for (int b = 0; b < n_bwt; ++b) {
            uint64_t before[4] = {0, 0, 0, 0};
            for (int r = 0; r < rank; ++r) for (int k = 0; k < 4; ++k) before[k] += sh.counts[b][(size_t)r * 4 + k];
            sh.super[b][rank].assign(e2i_index_super_count(sh.ix[b][rank]) * 4, 0);
            if (!sh.any_failed()) STEP(e2i_index_slice_super(ctx, sh.ix[b][rank], before, sh.super[b][rank].data()));
        }
        bar.wait();                                      // all partial superblock tables are in
        for (int b = 0; b < n_bwt; ++b) {
            if (sh.any_failed()) break;
            uint64_t before[4] = {0, 0, 0, 0}, totals[4] = {0, 0, 0, 0};
            for (int r = 0; r < world; ++r) for (int k = 0; k < 4; ++k) { if (r < rank) before[k] += sh.counts[b][(size_t)r * 4 + k]; totals[k] += sh.counts[b][(size_t)r * 4 + k]; }
            std::vector<uint64_t> table(sh.super[b][rank].size(), 0);
            for (int r = 0; r < world; ++r) for (size_t i = 0; i < table.size(); ++i) table[i] += sh.super[b][r][i];
            STEP(e2i_index_slice_pack(ctx, sh.ix[b][rank], dslice[b], before, table.data()));
            STEP(e2i_index_finish(sh.ix[b][rank], totals));
        }
        for (int b = 0; b < n_bwt; ++b) { dfree(ctx, dslice[b]); dslice[b] = nullptr; }
        CUDA_STEP(cudaStreamSynchronize(s));
        bar.wait();                                      // every slice is packed
        if (world > 1 && !sh.any_failed()) {
            for (int b = 0; b < n_bwt; ++b) {
                const size_t seg = (size_t)per[b] * kTileBlocks * kBlockU4 * sizeof(uint4);
                for (int k = 1; k < world; ++k) {        // start with the next rank: spreads the load over the links
                    const int r = (rank + k) % world;
                    if (!sl[b][r].tiles) continue;
                    char *mine = reinterpret_cast<char *>(sh.ix[b][rank]->blocks) + seg * r;
                    const char *theirs = reinterpret_cast<const char *>(sh.ix[b][r]->blocks) + seg * r;
                    const size_t bytes = (size_t)sl[b][r].tiles * kTileBlocks * kBlockU4 * sizeof(uint4);
                    CUDA_STEP(cudaMemcpyPeerAsync(mine, ctx->device, theirs, sh.ctx[r]->device, bytes, s));
                }
            }
        }
        if (host_da) {
            if (sh.rc[rank] == E2I_OK) { const int rc = e2i_da_load(ctx, host_da, n1, &sh.da[rank]); if (rc != E2I_OK) fail(rc); st.h2d_bytes += n1; }
        }
        cudaEventRecord(ctx->ev[7], s);
        CUDA_STEP(cudaStreamSynchronize(s));
        { float ms = 0; if (cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7]) == cudaSuccess) st.ms_index += ms; }
        bar.wait();                                      // nobody reads a peer's blocks any more; the indexes are complete
        // ---- 2. traversal shard ----
        if (ranged) {                                    // every rank enters: the traversal has barriers of its own
            if (sh.any_failed()) { /* all ranks see the same flags here (after the barrier): nobody enters */ }
            else STEP(e2i_navigate_ranged(ctx, comms[rank], sh.ix[0][rank], two ? sh.ix[1][rank] : nullptr, p, &sh.lcp[rank], two ? &sh.da_nav[rank] : nullptr, &st));
        } else if (!sh.any_failed()) {
            STEP(e2i_navigate_shard(ctx, sh.ix[0][rank], two ? sh.ix[1][rank] : nullptr, p, rank, world, &sh.lcp[rank], two ? &sh.da_nav[rank] : nullptr, &st));
        }
        CUDA_STEP(cudaStreamSynchronize(s));
        bar.wait();                                      // every shard's bits are written
        // ---- 3. OR-combine over peer memory ----
        if (world > 1 && !sh.any_failed()) {
            void *ptrs[kMaxGpus];
            uint64_t words = 0;
            for (int r = 0; r < world; ++r) ptrs[r] = sh.lcp[r]->thr;
            words = sh.lcp[rank]->thr_words32;
            STEP(e2i_or_allreduce(ctx, ptrs, world, rank, words));
            for (int r = 0; r < world; ++r) ptrs[r] = sh.lcp[r]->minima;
            STEP(e2i_or_allreduce(ctx, ptrs, world, rank, sh.lcp[rank]->min_words32));
            if (two) {
                for (int r = 0; r < world; ++r) ptrs[r] = sh.da_nav[r]->words;
                STEP(e2i_or_allreduce(ctx, ptrs, world, rank, sh.da_nav[rank]->n_words32));
            }
            CUDA_STEP(cudaStreamSynchronize(s));
        }
        bar.wait();                                      // the combined vectors are everywhere
        // ---- 4. phase 4 on this GPU's suffix-array range, text per range ----
        e2i_calls *calls = nullptr;
        if (!sh.any_failed()) {
            const uint64_t lo = (uint64_t)((unsigned __int128)n * rank / world), hi = (uint64_t)((unsigned __int128)n * (rank + 1) / world);
            // the records stay in HBM: counted and printed there once the first cluster number of the range is known
            STEP(e2i_call_device(ctx, sh.ix[0][rank], two ? sh.ix[1][rank] : nullptr, two ? sh.da_nav[rank] : sh.da[rank], sh.lcp[rank], p, lo, hi, &calls, &st));
            STEP(e2i_calls_clusters(calls, p, &sh.clusters[rank]));
        }
        bar.wait();                                      // every range knows how many cluster numbers it consumes
        void *d_text = nullptr;
        if (!sh.any_failed()) {
            uint64_t first = 1, len = 0;
            for (int r = 0; r < rank; ++r) first += sh.clusters[r];
            const auto t0 = std::chrono::steady_clock::now();
            STEP(e2i_calls_snp_device(calls, p, first, &d_text, &len, &st));
            sh.text_len[rank] = (size_t)len;
            st.ms_format += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        }
        bar.wait();                                      // every range knows the length of its text
        if (rank == 0 && !sh.any_failed()) {             // one page-locked buffer for the whole text
            size_t total = 0;
            for (int r = 0; r < world; ++r) total += sh.text_len[r];
            sh.final_text = text_alloc(total + 1);
            if (!sh.final_text) { set_error("cannot page-lock %llu bytes for the .snp text", (unsigned long long)total); fail(E2I_ERR_MEMORY); }
            else sh.final_text[total] = 0;
        }
        bar.wait();
        if (!sh.any_failed() && sh.text_len[rank]) {     // every GPU copies its piece to its place
            size_t off = 0;
            for (int r = 0; r < rank; ++r) off += sh.text_len[r];
            CUDA_STEP(cudaMemcpy(sh.final_text + off, d_text, sh.text_len[rank], cudaMemcpyDeviceToHost));
            st.d2h_bytes += sh.text_len[rank];
        }
        e2i_device_free(ctx, d_text);
        e2i_calls_free(calls);
        e2i_lcpbits_free(sh.lcp[rank]);
        e2i_bits_free(sh.da_nav[rank]);
        e2i_bits_free(sh.da[rank]);
        bar.wait();                                      // peers are done with this rank's memory
        for (int b = 0; b < 2; ++b) e2i_index_free(sh.ix[b][rank]);
#undef STEP
#undef CUDA_STEP
    };
    std::vector<std::thread> th;
    for (int r = 1; r < world; ++r) th.emplace_back(worker, r);
    worker(0);
    for (auto &t : th) t.join();
    for (e2i_comm *c : comms) e2i_comm_free(c);

    int rc = E2I_OK;
    for (int r = 0; r < world; ++r)
        if (sh.rc[r] != E2I_OK) { rc = sh.rc[r]; set_error("GPU %d: %s", devices[r], sh.err[r].c_str()); break; }
    if (rc == E2I_OK) {
        size_t total = 0;
        for (int r = 0; r < world; ++r) total += sh.text_len[r];
        *snp = sh.final_text;
        *snp_len = total;
    } else if (sh.final_text) {
        e2i_buffer_free(sh.final_text);
    }
    // counters: every unit of work is done by exactly one rank -> sums; phase times -> max over ranks
    e2i_stats &o = *st_out;
    for (int r = 0; r < world; ++r) {
        const e2i_stats &s = sh.st[r];
        o.leaves += s.leaves; o.nodes += s.nodes; o.lcp_values += s.lcp_values; o.lcp_values_leaves += s.lcp_values_leaves;
        o.n_min += s.n_min; o.da_values += s.da_values; o.da_values_leaves += s.da_values_leaves; o.n_clusters += s.n_clusters;
        o.clust_size += s.clust_size; o.events += s.events; o.clusters_out += s.clusters_out;
        for (int i = 0; i <= 200; ++i) o.clust_sizes[i] += s.clust_sizes[i];
        o.rank_leaves += s.rank_leaves; o.rank_nodes += s.rank_nodes; o.rank_call += s.rank_call; o.bit_updates += s.bit_updates;
        o.candidates += s.candidates; o.kernel_launches += s.kernel_launches; o.h2d_bytes += s.h2d_bytes; o.d2h_bytes += s.d2h_bytes;
        o.levels_leaves = std::max(o.levels_leaves, s.levels_leaves); o.levels_nodes = std::max(o.levels_nodes, s.levels_nodes);
        o.max_frontier = std::max(o.max_frontier, s.max_frontier);
        o.ms_index = std::max(o.ms_index, s.ms_index); o.ms_leaves = std::max(o.ms_leaves, s.ms_leaves); o.ms_nodes = std::max(o.ms_nodes, s.ms_nodes);
        o.ms_call = std::max(o.ms_call, s.ms_call); o.ms_format = std::max(o.ms_format, s.ms_format);
    }
    o.ms_wall += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w_start).count();
    destroy_all();
    return rc;
}
