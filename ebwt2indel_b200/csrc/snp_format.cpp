// snp_format.cpp -- a21-a23: event classification and KisSNP2-style .snp text (host side).
//
// Replaces has_run / dH / distance (/root/reference/ebwt2InDel.cpp:143-240), event_type
// (:1102-1144) and the two to_file overloads (:1149-1252 for modes -2/-d, :1254-1330 for mode -1).
// Input: the per-cluster records produced on the device by e2i_call, in suffix-array order.
// Quirks kept on purpose (SURVEY.md §8a "parity hazards"): strict comparisons in distance(),
// the good[1] operand of event_type in mode -1, cluster numbers that advance for clusters that
// print nothing, the `right:` field printing the actual right-context length.
#include <algorithm>
#include <chrono>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include <cub/cub.cuh>

#include "common.cuh"

#define E2I_HD __host__ __device__ __forceinline__

namespace {

// ---------------------------------------------------------------------------------------------
// The classification and the text of ONE record, written once for the host formatter
// (e2i_snp_format, records in host memory) and for the device formatter (format_device, records
// still in HBM after phase 4): the two differ only in the sink the characters go to.
// ---------------------------------------------------------------------------------------------
struct Dist { int mism; int gap; };   // gap > 0: insertion in a; gap < 0: insertion in b

E2I_HD int popc64(uint64_t x) {
#ifdef __CUDA_ARCH__
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}

// right-aligned Hamming distance over the shorter length (:157-171); the host compares 8 bytes per step
E2I_HD int hamming_right(const char *a, int la, const char *b, int lb) {
    const int len = la < lb ? la : lb;
    const char *pa = a + (la - len), *pb = b + (lb - len);
    int d = 0, i = 0;
#ifndef __CUDA_ARCH__
    for (; i + 8 <= len; i += 8) {
        uint64_t x, y;
        std::memcpy(&x, pa + i, 8);
        std::memcpy(&y, pb + i, 8);
        uint64_t z = x ^ y;                               // a byte is non-zero iff the characters differ
        z |= z >> 4;
        z |= z >> 2;
        z |= z >> 1;
        d += popc64(z & 0x0101010101010101ull);
    }
#endif
    for (; i < len; ++i) d += pa[i] != pb[i];
    return d;
}

// Contexts of <= 32 bases as 2-bit codes, character k at bits [2k, 2k+2): false if a character is not A,C,G,T.
E2I_HD bool pack2(const char *s, int len, uint64_t &out) {
    uint64_t v = 0;
    int k = 0;
#ifndef __CUDA_ARCH__
    for (; k + 4 <= len; k += 4) {                      // four characters per step
        uint32_t w;
        std::memcpy(&w, s + k, 4);
        const uint32_t y = (w >> 1) & 0x03030303u, code = y ^ ((y >> 1) & 0x01010101u);
        // valid iff every byte equals "ACGT"[code]: A 0x41, C 0x43, G 0x47, T 0x54
        const uint32_t lo = code & 0x01010101u, hi = (code >> 1) & 0x01010101u;
        const uint32_t expect = 0x41414141u + lo * 2u + hi * 6u + (lo & hi) * 11u;   // +2 C, +6 G, +19 T
        if (w != expect) return false;
        v |= (uint64_t)((code * 0x01041040u) >> 24) << (2 * k);
    }
#endif
    for (; k < len; ++k) {
        const unsigned c = (unsigned char)s[k];
        const unsigned y = (c >> 1) & 3u, code = y ^ (y >> 1);              // A 0, C 1, G 2, T 3
        const unsigned lo = code & 1u, hi = code >> 1;
        if (c != 0x41u + lo * 2u + hi * 6u + (lo & hi) * 11u) return false;
        v |= (uint64_t)code << (2 * k);
    }
    out = v;
    return true;
}

E2I_HD int mismatches2(uint64_t x, uint64_t y, int chars) {             // over the lowest `chars` characters
    const uint64_t z = x ^ y, m = chars >= 32 ? ~0ull : ((1ull << (2 * chars)) - 1);
    return popc64((z | (z >> 1)) & 0x5555555555555555ull & m);
}

// :192-240.  Candidates: no indel, or drop 1..max_gap characters from the right end of a / of b.
// "No indel" wins only if strictly better than both; "insert in a" only if strictly better than b.
// distance() on packed contexts: every shifted comparison is a shift, an XOR and a popcount.
E2I_HD Dist distance_packed(uint64_t a, uint64_t b, int len, int max_gap) {
    const int plain = mismatches2(a, b, len);
    int best_a = 0, gap_a = 0, best_b = 0, gap_b = 0;
    for (int g = 1; g <= max_gap; ++g) {
        int da, db;
        if (g <= len) {
            const int keep = len - g;
            da = (keep ? mismatches2(a, g >= 32 ? 0 : (b >> (2 * g)), keep) : 0) + g;   // a[0, keep) against b[g, len)
            db = (keep ? mismatches2(g >= 32 ? 0 : (a >> (2 * g)), b, keep) : 0) + g;   // a[g, len) against b[0, keep)
        } else {
            da = db = plain + g;                     // substr(0, len-g) wraps to the whole string when g > len
        }
        if (g == 1 || da < best_a) { best_a = da; gap_a = g; }
        if (g == 1 || db < best_b) { best_b = db; gap_b = g; }
    }
    if (plain < best_a && plain < best_b) return {plain, 0};
    if (best_a < best_b) return {best_a - gap_a, gap_a};
    return {best_b - gap_b, -gap_b};
}

E2I_HD Dist distance(const char *a, const char *b, int len, int max_gap) {
    uint64_t pa, pb;
    if (max_gap > 0 && len >= 1 && len <= 32 && pack2(a, len, pa) && pack2(b, len, pb)) return distance_packed(pa, pb, len, max_gap);
    const int plain = hamming_right(a, len, b, len);
    if (max_gap <= 0) return {plain, 0};
    int best_a = 0, gap_a = 0, best_b = 0, gap_b = 0;
    for (int g = 1; g <= max_gap; ++g) {
        const int keep = g <= len ? len - g : len;   // substr(0, len-g) wraps to the whole string when g > len
        const int da = hamming_right(a, keep, b, len) + g;
        const int db = hamming_right(a, len, b, keep) + g;
        if (g == 1 || da < best_a) { best_a = da; gap_a = g; }
        if (g == 1 || db < best_b) { best_b = db; gap_b = g; }
    }
    if (plain < best_a && plain < best_b) return {plain, 0};
    if (best_a < best_b) return {best_a - gap_a, gap_a};
    return {best_b - gap_b, -gap_b};
}

// true iff s starts with a run of >= k equal characters (:144-152)
E2I_HD bool starts_with_run(const char *s, int len, int k) {
    if (k < 0 || k > len) return false;
    for (int i = 1; i < k; ++i)
        if (s[i] != s[i - 1]) return false;
    return true;
}

E2I_HD int digits10(uint64_t v) { int d = 1; while (v >= 10) { v /= 10; ++d; } return d; }

// A sink takes characters: lit("..."), put(c), put(s, len), uint(v), sint(v), cluster_nr() -- the
// number of the record's cluster, which only the sink knows (numbering is sequential over the
// whole run, cluster_nr, ebwt2InDel.cpp:1250/1328) -- and room(bytes) before every output line.
#pragma nv_exec_check_disable
template <class Sink>
E2I_HD void append_event(Sink &o, const char *l0, const char *l1, int len, Dist d) {   // :1102-1144
    o.lit("type:");
    if (d.gap != 0) o.lit("_INDEL_event:"); else o.lit("_SNP_event:");
    if (d.gap == 0) { o.put(l0[len - 1]); o.put('/'); o.put(l1[len - 1]); }
    else if (d.gap > 0) { o.put(l0 + len - d.gap, (size_t)d.gap); o.put('/'); }
    else { o.put('/'); o.put(l1 + len + d.gap, (size_t)(-d.gap)); }
}

#pragma nv_exec_check_disable
template <class Sink>
E2I_HD void append_header(Sink &o, uint64_t id, int right_len, int cov) {
    o.lit(">cluster:");
    o.cluster_nr();
    o.lit("_id:");      o.uint(id);
    o.lit("_right:");   o.sint(right_len);
    o.lit("_cov:");     o.sint(cov);
    o.put('_');
}

// One record: its text goes to the sink; `clusters` is how many cluster numbers it consumes (0 or 1),
// `events` how many events it stores (mode -1 only, :1320).  L: the record's 8 left-context slots,
// R: its right context.
#pragma nv_exec_check_disable
template <class Sink>
E2I_HD void format_record(const e2i_call_rec &rec, const char *L, const char *R, const e2i_params &p, int two_samples,
                          size_t line_max, Sink &o, uint32_t &clusters, uint32_t &events) {
    clusters = 0;
    events = 0;
    if (!rec.has_right) return;                         // empty variant vector: nothing printed, nothing counted
    const int kl = p.k_left;
    const int rlen = rec.right_len;
    if (!two_samples) {
        // to_file(vector<variant_single_t>) :1254-1330
        const int nv = rec.n0;
        if (nv < 2) return;
        int max_dist = 0, good[4], ng = 0;
        for (int i = 0; i + 1 < nv; ++i) {
            const Dist d = distance(L + i * kl, L + (i + 1) * kl, kl, p.max_gap);
            if (d.mism > max_dist) max_dist = d.mism;
            if (rec.support[i] >= p.mcov_out) good[ng++] = i;
        }
        if (rec.support[nv - 1] >= p.mcov_out) good[ng++] = nv - 1;
        if (max_dist <= p.max_snvs && ng >= 2 && !starts_with_run(R, rlen, p.complexity)) {
            uint64_t id = 1;
            for (int g = 0; g < ng; ++g) {
                const char *me = L + good[g] * kl;
                o.room(line_max);
                append_header(o, id++, rlen, rec.support[good[g]]);
                const char *x = g == 0 ? me : L + good[g - 1] * kl;   // :1299-1307
                const char *y = L + good[1] * kl;
                append_event(o, x, y, kl, distance(x, y, kl, p.max_gap));
                o.put('\n');
                o.put(me, (size_t)kl);
                o.put(R, (size_t)rlen);
                o.put('\n');
                events++;
            }
        }
        clusters = 1;                                                // :1328
    } else {
        // find_variants' cross product (:915-928, 1077-1090) + to_file(vector<variant_t>) :1149-1252
        bool found = false;
        uint64_t id = 1;
        if (starts_with_run(R, rlen, p.complexity)) return;          // :1159 rejects every pair of the cluster
        for (int i0 = 0; i0 < rec.n0; ++i0) for (int i1 = 0; i1 < rec.n1; ++i1) {
            const char *l0 = L + i0 * kl, *l1 = L + (4 + i1) * kl;
            if (l0[kl - 1] == l1[kl - 1]) continue;
            const int s0 = rec.support[i0], s1 = rec.support[4 + i1];
            if (s0 < p.mcov_out || s1 < p.mcov_out) continue;
            const Dist d = distance(l0, l1, kl, p.max_gap);
            if (d.mism > p.max_snvs) continue;
            found = true;
            for (int side = 0; side < 2; ++side) {
                o.room(line_max);
                append_header(o, id, rlen, side ? s1 : s0);
                append_event(o, l0, l1, kl, d);
                o.put('\n');
                int skip = 0;
                if (side == 0 && d.gap < 0) skip = -d.gap;           // :1199
                if (side == 1 && d.gap > 0) skip = d.gap;            // :1233
                o.put((side ? l1 : l0) + skip, (size_t)(kl - skip));
                o.put(R, (size_t)rlen);
                o.put('\n');
            }
            id++;
        }
        clusters = found ? 1 : 0;                                    // :1250
    }
}

// ---- host sink ----------------------------------------------------------------------------
// Growable text buffer with unchecked appends after ensure(): the formatter writes a few hundred
// bytes per record, and std::string's per-character capacity checks were half of its time.
// Buffers are kept (with their touched pages) in a process-wide pool between calls.
struct TextBuf {
    char *p = nullptr;
    size_t n = 0, cap = 0;
    TextBuf() = default;
    TextBuf(const TextBuf &) = delete;
    TextBuf &operator=(const TextBuf &) = delete;
    TextBuf(TextBuf &&o) noexcept : p(o.p), n(o.n), cap(o.cap) { o.p = nullptr; o.n = o.cap = 0; }
    ~TextBuf() { std::free(p); }
    void ensure(size_t extra) {
        if (n + extra <= cap) return;
        size_t want = std::max(cap * 2, n + extra + (1u << 16));
        p = static_cast<char *>(std::realloc(p, want));
        cap = want;
    }
};

// Record ranges are formatted in parallel with local cluster indices behind a placeholder; the
// numbers are filled in once the per-range totals are known.
struct Piece {
    TextBuf text;                                       // '\x01' marks where a cluster number goes
    std::vector<std::pair<size_t, uint64_t>> marks;     // (offset of the placeholder, local cluster index)
    uint64_t clusters = 0, events = 0;
};

struct HostSink {
    Piece &pc;
    uint64_t local = 0;                                 // local cluster index of the current record
    explicit HostSink(Piece &x) : pc(x) {}
    void room(size_t bytes) { pc.text.ensure(bytes); }
    void put(char c) { pc.text.p[pc.text.n++] = c; }
    void put(const char *s, size_t len) { std::memcpy(pc.text.p + pc.text.n, s, len); pc.text.n += len; }
    template <size_t N> void lit(const char (&s)[N]) { std::memcpy(pc.text.p + pc.text.n, s, N - 1); pc.text.n += N - 1; }
    void uint(uint64_t v) {
        char buf[24];
        int k = 0;
        do { buf[k++] = (char)('0' + v % 10); v /= 10; } while (v);
        while (k) pc.text.p[pc.text.n++] = buf[--k];
    }
    void sint(int v) {
        if (v < 0) { put('-'); uint((uint64_t)(-(int64_t)v)); } else uint((uint64_t)v);
    }
    void cluster_nr() { pc.marks.emplace_back(pc.text.n, local); put('\x01'); }
};

inline size_t line_bytes(const e2i_params *p) {        // one header + sequence line, at most
    return 160 + (size_t)std::max(0, p->max_gap) + (size_t)p->k_left + (size_t)p->k_right;
}

void format_range(const e2i_call_rec *recs, const char *left, const char *right, uint64_t r0, uint64_t r1,
                  const e2i_params *p, int two_samples, Piece &out) {
    const int kl = p->k_left, kr = p->k_right;
    out.text.n = 0;
    out.marks.clear();
    out.clusters = out.events = 0;
    const size_t line_max = line_bytes(p);
    out.text.ensure((size_t)(r1 - r0) * 2 * (size_t)(kl + kr + 72) + line_max);
    out.marks.reserve((size_t)(r1 - r0) * 2);
    HostSink o(out);
    for (uint64_t r = r0; r < r1; ++r) {
        uint32_t c = 0, e = 0;
        o.local = out.clusters;                         // global number = first of the range + local index
        format_record(recs[r], left + r * 8 * (size_t)kl, right + r * (size_t)kr, *p, two_samples, line_max, o, c, e);
        out.clusters += c;
        out.events += e;
    }
}

// ---- device sinks -------------------------------------------------------------------------
// Pass 1 measures a record (bytes outside the cluster numbers + how many cluster numbers it prints),
// a scan of the cluster flags numbers the clusters, a scan of the lengths places the records, pass 2
// writes the characters: one thread per record, each into its own stretch of the text.
struct CountSink {
    uint32_t fixed = 0, headers = 0;
    __device__ void room(size_t) {}
    __device__ void put(char) { fixed++; }
    __device__ void put(const char *, size_t len) { fixed += (uint32_t)len; }
    template <size_t N> __device__ void lit(const char (&)[N]) { fixed += (uint32_t)(N - 1); }
    __device__ void uint(uint64_t v) { fixed += (uint32_t)digits10(v); }
    __device__ void sint(int v) { if (v < 0) { fixed++; uint((uint64_t)(-(int64_t)v)); } else uint((uint64_t)v); }
    __device__ void cluster_nr() { headers++; }
};

struct WriteSink {
    char *p;
    uint64_t number;                                    // the record's cluster number
    __device__ void room(size_t) {}
    __device__ void put(char c) { *p++ = c; }
    __device__ void put(const char *s, size_t len) { for (size_t i = 0; i < len; ++i) p[i] = s[i]; p += len; }
    template <size_t N> __device__ void lit(const char (&s)[N]) {
#pragma unroll
        for (size_t i = 0; i + 1 < N; ++i) p[i] = s[i];
        p += N - 1;
    }
    __device__ void uint(uint64_t v) {
        const int d = digits10(v);
        for (int i = d - 1; i >= 0; --i) { p[i] = (char)('0' + v % 10); v /= 10; }
        p += d;
    }
    __device__ void sint(int v) { if (v < 0) { put('-'); uint((uint64_t)(-(int64_t)v)); } else uint((uint64_t)v); }
    __device__ void cluster_nr() { uint(number); }
};

constexpr int kFmtThreads = 128;

__global__ void __launch_bounds__(kFmtThreads)
snp_measure_kernel(const e2i_call_rec *recs, const char *left, const char *right, uint32_t n, const e2i_params p, int two_samples,
                   uint32_t *clus, uint32_t *fixed, uint32_t *headers, unsigned long long *totals) {
    const uint32_t r = blockIdx.x * kFmtThreads + threadIdx.x;
    uint32_t c = 0, e = 0;
    if (r < n) {
        CountSink o;
        format_record(recs[r], left + (size_t)r * 8 * p.k_left, right + (size_t)r * p.k_right, p, two_samples, 0, o, c, e);
        clus[r] = c;
        fixed[r] = o.fixed;
        headers[r] = o.headers;
    }
    e = __reduce_add_sync(0xffffffffu, e);
    if ((threadIdx.x & 31) == 0 && e) atomicAdd(&totals[2], (unsigned long long)e);
}

// length of a record's text now that its cluster number is known (cpre: exclusive scan of clus)
__global__ void snp_length_kernel(const uint32_t *fixed, const uint32_t *headers, const uint32_t *cpre, uint32_t n, uint64_t first, uint64_t *len) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) len[r] = (uint64_t)fixed[r] + (headers[r] ? (uint64_t)headers[r] * (uint64_t)digits10(first + cpre[r]) : 0ull);
}

__global__ void snp_totals_kernel(const uint32_t *clus, const uint32_t *cpre, const uint64_t *len, const uint64_t *off, uint32_t n, unsigned long long *totals) {
    totals[0] = off[n - 1] + len[n - 1];
    totals[1] = (unsigned long long)cpre[n - 1] + clus[n - 1];
}

__global__ void __launch_bounds__(kFmtThreads)
snp_write_kernel(const e2i_call_rec *recs, const char *left, const char *right, uint32_t n, const e2i_params p, int two_samples,
                 const uint32_t *cpre, uint64_t first, const uint64_t *off, const uint64_t *len, char *text) {
    const uint32_t r = blockIdx.x * kFmtThreads + threadIdx.x;
    if (r >= n || len[r] == 0) return;
    WriteSink o{text + off[r], first + cpre[r]};
    uint32_t c, e;
    format_record(recs[r], left + (size_t)r * 8 * p.k_left, right + (size_t)r * p.k_right, p, two_samples, 0, o, c, e);
}

// Number of cluster numbers a range of records consumes (what format_range adds to cluster_nr),
// without building any text.  Mode -1: every record with >= 2 variants; two samples: records with
// at least one emitted pair.
uint64_t count_range(const e2i_call_rec *recs, const char *left, const char *right, uint64_t r0, uint64_t r1,
                     const e2i_params *p, int two_samples) {
    const int kl = p->k_left, kr = p->k_right;
    uint64_t cluster = 0;
    for (uint64_t r = r0; r < r1; ++r) {
        const e2i_call_rec &rec = recs[r];
        if (!rec.has_right) continue;
        if (!two_samples) { cluster += rec.n0 >= 2; continue; }
        const char *L = left + r * 8 * (size_t)kl;
        if (starts_with_run(right + r * (size_t)kr, rec.right_len, p->complexity)) continue;
        bool found = false;
        for (int i0 = 0; i0 < rec.n0 && !found; ++i0) for (int i1 = 0; i1 < rec.n1 && !found; ++i1) {
            const char *l0 = L + i0 * kl, *l1 = L + (4 + i1) * kl;
            if (l0[kl - 1] == l1[kl - 1]) continue;
            if (rec.support[i0] < p->mcov_out || rec.support[4 + i1] < p->mcov_out) continue;
            found = distance(l0, l1, kl, p->max_gap).mism <= p->max_snvs;
        }
        cluster += found;
    }
    return cluster;
}


}  // namespace

// ---- the device formatter ------------------------------------------------------------------
// Records, left and right contexts in device memory (the layout of e2i_call_rec / e2i_calls_view) -> the
// text in device memory (*d_text, to be released with dfree; nullptr when the records print nothing).
int e2i::format_device(e2i_ctx *ctx, const e2i_call_rec *d_recs, const char *d_left, const char *d_right, uint64_t n_recs,
                       const e2i_params *p, int two_samples, uint64_t first_cluster_nr,
                       char **d_text, uint64_t *text_len, uint64_t *clusters, uint64_t *events, bool want_text) {
    *d_text = nullptr;
    *text_len = *clusters = *events = 0;
    if (n_recs == 0) return E2I_OK;
    if (n_recs >= (1ull << 31)) { set_error("format_device: too many records in one batch"); return E2I_ERR_ARG; }
    const uint32_t n = (uint32_t)n_recs;
    cudaStream_t s = ctx->stream;
    size_t tmp32 = 0, tmp64 = 0;
    E2I_CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp32, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)n, s));
    E2I_CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp64, (const uint64_t *)nullptr, (uint64_t *)nullptr, (int)n, s));
    const size_t tmp_bytes = (std::max(tmp32, tmp64) + 255) & ~(size_t)255;
    const size_t n_pad = ((size_t)n + 63) & ~(size_t)63;
    // [totals 256 B | len u64 | off u64 | clus, cpre, fixed, headers u32 | scan scratch]
    char *work = nullptr;
    if (dmalloc(ctx, &work, 256 + n_pad * 32 + tmp_bytes) != cudaSuccess) { cudaGetLastError(); set_error("format_device: out of device memory"); return E2I_ERR_MEMORY; }
    unsigned long long *totals = reinterpret_cast<unsigned long long *>(work);
    uint64_t *len = reinterpret_cast<uint64_t *>(work + 256), *off = len + n_pad;
    uint32_t *clus = reinterpret_cast<uint32_t *>(off + n_pad), *cpre = clus + n_pad, *fixed = cpre + n_pad, *headers = fixed + n_pad;
    void *scratch = headers + n_pad;
    auto fail = [&](int rc) { dfree(ctx, work); return rc; };
#define TRYF(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__, cudaGetErrorString(_e)); return fail(E2I_ERR_CUDA); } } while (0)
    const unsigned grid = (n + kFmtThreads - 1) / kFmtThreads;
    size_t tb = tmp_bytes;
    TRYF(cudaMemsetAsync(totals, 0, 256, s));
    snp_measure_kernel<<<grid, kFmtThreads, 0, s>>>(d_recs, d_left, d_right, n, *p, two_samples, clus, fixed, headers, totals);
    TRYF(cub::DeviceScan::ExclusiveSum(scratch, tb, clus, cpre, (int)n, s));
    snp_length_kernel<<<(n + 255) / 256, 256, 0, s>>>(fixed, headers, cpre, n, first_cluster_nr, len);
    tb = tmp_bytes;
    TRYF(cub::DeviceScan::ExclusiveSum(scratch, tb, len, off, (int)n, s));
    snp_totals_kernel<<<1, 1, 0, s>>>(clus, cpre, len, off, n, totals);
    TRYF(cudaGetLastError());
    unsigned long long h[3];
    TRYF(cudaMemcpyAsync(h, totals, sizeof h, cudaMemcpyDeviceToHost, s));
    TRYF(cudaStreamSynchronize(s));
    ctx->n_launch += 5;
    ctx->n_d2h += sizeof h;
    if (h[0] && want_text) {
        char *text = nullptr;
        if (dmalloc(ctx, &text, h[0]) != cudaSuccess) { cudaGetLastError(); set_error("format_device: out of device memory (%llu bytes of text)", h[0]); return fail(E2I_ERR_MEMORY); }
        snp_write_kernel<<<grid, kFmtThreads, 0, s>>>(d_recs, d_left, d_right, n, *p, two_samples, cpre, first_cluster_nr, off, len, text);
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { dfree(ctx, text); set_error("CUDA error in the .snp writer: %s", cudaGetErrorString(e)); return fail(E2I_ERR_CUDA); }
        ctx->n_launch++;
        *d_text = text;
    }
#undef TRYF
    dfree(ctx, work);                                   // stream-ordered: after the writer
    *text_len = h[0];
    *clusters = h[1];
    *events = h[2];
    return E2I_OK;
}

// ---- page-locked text buffers, cached process-wide ---------------------------------------------
namespace {
struct TextBlock { char *p; size_t cap; bool used; };
std::mutex g_text_mutex;
std::vector<TextBlock> g_text_blocks;
constexpr size_t kTextCached = 3;                       // buffers kept when they come back
}  // namespace

char *e2i::text_alloc(size_t bytes) {
    std::lock_guard<std::mutex> hold(g_text_mutex);
    TextBlock *best = nullptr;
    for (auto &b : g_text_blocks) if (!b.used && b.cap >= bytes && (!best || b.cap < best->cap)) best = &b;
    if (best) { best->used = true; return best->p; }
    const size_t cap = std::max<size_t>(bytes + bytes / 8, 1u << 20);
    void *p = nullptr;
    if (cudaMallocHost(&p, cap) != cudaSuccess) {       // no page-locked memory left: an ordinary buffer (the copy is staged by the driver)
        cudaGetLastError();
        return static_cast<char *>(std::malloc(bytes ? bytes : 1));
    }
    g_text_blocks.push_back({static_cast<char *>(p), cap, true});
    return static_cast<char *>(p);
}

bool e2i::text_release(void *p) {
    std::lock_guard<std::mutex> hold(g_text_mutex);
    for (size_t i = 0; i < g_text_blocks.size(); ++i) {
        if (g_text_blocks[i].p != p) continue;
        g_text_blocks[i].used = false;
        size_t idle = 0, smallest = i;
        for (size_t k = 0; k < g_text_blocks.size(); ++k)
            if (!g_text_blocks[k].used) { ++idle; if (g_text_blocks[k].cap < g_text_blocks[smallest].cap) smallest = k; }
        if (idle > kTextCached) {                       // keep the large ones
            cudaFreeHost(g_text_blocks[smallest].p);
            g_text_blocks.erase(g_text_blocks.begin() + (long)smallest);
        }
        return true;
    }
    return false;
}

void e2i::text_cache_trim() {
    std::lock_guard<std::mutex> hold(g_text_mutex);
    for (size_t i = g_text_blocks.size(); i-- > 0;)
        if (!g_text_blocks[i].used) { cudaFreeHost(g_text_blocks[i].p); g_text_blocks.erase(g_text_blocks.begin() + (long)i); }
}

extern "C" void e2i_buffer_free(void *p);

// host text buffer filled by one copy
static int text_to_host(e2i_ctx *ctx, char *d_text, uint64_t len, char **snp, size_t *snp_len) {
    char *buf = e2i::text_alloc(len + 1);
    if (!buf) { e2i::dfree(ctx, d_text); e2i::set_error("out of host memory (%llu bytes of .snp text)", (unsigned long long)len); return E2I_ERR_MEMORY; }
    if (len) {
        cudaError_t e = cudaMemcpyAsync(buf, d_text, len, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { e2i_buffer_free(buf); e2i::dfree(ctx, d_text); e2i::set_error("CUDA error copying the .snp text: %s", cudaGetErrorString(e)); return E2I_ERR_CUDA; }
        ctx->n_d2h += len;
    }
    e2i::dfree(ctx, d_text);
    buf[len] = 0;
    *snp = buf;
    *snp_len = (size_t)len;
    return E2I_OK;
}

extern "C" int e2i_snp_format_gpu(e2i_ctx *ctx, const e2i_call_rec *recs, const char *left, const char *right, uint64_t n_recs,
                                  const e2i_params *p, int two_samples, uint64_t first_cluster_nr,
                                  char **snp, size_t *snp_len, e2i_stats *st) {
    using namespace e2i;
    if (!ctx || !p || !snp || !snp_len || (n_recs && (!recs || !left || !right))) { set_error("e2i_snp_format_gpu: null argument"); return E2I_ERR_ARG; }
    if (p->k_left < 1 || p->k_left > 255 || p->k_right < 1 || p->k_right > 255) { set_error("e2i_snp_format_gpu: k_left and k_right must be in [1,255]"); return E2I_ERR_ARG; }
    E2I_CUDA_TRY(cudaSetDevice(ctx->device));
    Accounting acct(ctx, st);
    const size_t kl = (size_t)p->k_left, kr = (size_t)p->k_right;
    const size_t b_rec = (n_recs * sizeof(e2i_call_rec) + 255) & ~(size_t)255, b_left = (n_recs * 8 * kl + 255) & ~(size_t)255, b_right = n_recs * kr;
    char *dev = nullptr, *d_text = nullptr;
    uint64_t len = 0, clusters = 0, events = 0;
    if (n_recs) {
        if (dmalloc(ctx, &dev, b_rec + b_left + b_right + 256) != cudaSuccess) { cudaGetLastError(); set_error("e2i_snp_format_gpu: out of device memory"); return E2I_ERR_MEMORY; }
        cudaError_t e = cudaMemcpyAsync(dev, recs, n_recs * sizeof(e2i_call_rec), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(dev + b_rec, left, n_recs * 8 * kl, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(dev + b_rec + b_left, right, n_recs * kr, cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) { dfree(ctx, dev); set_error("CUDA error uploading the call records: %s", cudaGetErrorString(e)); return E2I_ERR_CUDA; }
        ctx->n_h2d += n_recs * (sizeof(e2i_call_rec) + 8 * kl + kr);
    }
    const int rc = format_device(ctx, reinterpret_cast<const e2i_call_rec *>(dev), dev + b_rec, dev + b_rec + b_left, n_recs, p, two_samples,
                                 first_cluster_nr ? first_cluster_nr : 1, &d_text, &len, &clusters, &events);
    if (dev) dfree(ctx, dev);
    if (rc != E2I_OK) return rc;
    E2I_TRY(text_to_host(ctx, d_text, len, snp, snp_len));
    if (st) { st->events += events; st->clusters_out += clusters; }
    return E2I_OK;
}

extern "C" void e2i_distance(const char *a, const char *b, int32_t len, int32_t max_gap, int32_t out[2]) {
    const Dist d = distance(a, b, len, max_gap);
    out[0] = d.mism;
    out[1] = d.gap;
}

extern "C" int e2i_snp_count(const e2i_call_rec *recs, const char *left, const char *right, uint64_t n_recs,
                             const e2i_params *p, int two_samples, uint64_t *clusters) {
    if (!p || !clusters || (n_recs && (!recs || !left || !right))) { e2i::set_error("e2i_snp_count: null argument"); return E2I_ERR_ARG; }
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 1;
    const uint64_t nt = std::max<uint64_t>(1, std::min<uint64_t>({(uint64_t)hw, 64, n_recs / 4096}));
    std::vector<uint64_t> part(nt, 0);
    std::vector<std::thread> th;
    auto work = [&](uint64_t t) { part[t] = count_range(recs, left, right, n_recs * t / nt, n_recs * (t + 1) / nt, p, two_samples); };
    for (uint64_t t = 1; t < nt; ++t) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
    uint64_t tot = 0;
    for (uint64_t v : part) tot += v;
    *clusters = tot;
    return E2I_OK;
}

extern "C" int e2i_snp_format(const e2i_call_rec *recs, const char *left, const char *right, uint64_t n_recs,
                              const e2i_params *p, int two_samples, uint64_t first_cluster_nr,
                              char **snp, size_t *snp_len, e2i_stats *st) {
    if (!p || !snp || !snp_len || (n_recs && (!recs || !left || !right))) { e2i::set_error("e2i_snp_format: null argument"); return E2I_ERR_ARG; }
    // -g larger than -L is accepted like the reference does (its substr(0, len - g) wraps to the whole string,
    // ebwt2InDel.cpp:208-219; such a gap never wins, see distance())
    const uint64_t first = first_cluster_nr ? first_cluster_nr : 1;
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 1;
    const uint64_t nt = std::max<uint64_t>(1, std::min<uint64_t>({(uint64_t)hw, 64, n_recs / 2048}));
    // formatter buffers survive between calls (their pages stay touched); a concurrent caller gets fresh ones
    static std::mutex pool_mutex;
    static std::vector<Piece> pool;
    std::vector<Piece> local;
    std::unique_lock<std::mutex> hold(pool_mutex, std::try_to_lock);
    std::vector<Piece> &pieces = hold.owns_lock() ? pool : local;
    if (pieces.size() < nt) pieces.resize(nt);
    const auto t_a = std::chrono::steady_clock::now();
    {
        std::vector<std::thread> th;
        for (uint64_t t = 1; t < nt; ++t)
            th.emplace_back(format_range, recs, left, right, n_recs * t / nt, n_recs * (t + 1) / nt, p, two_samples, std::ref(pieces[t]));
        format_range(recs, left, right, 0, n_recs / nt, p, two_samples, pieces[0]);
        for (auto &x : th) x.join();
    }
    const auto t_b = std::chrono::steady_clock::now();
    // global cluster numbers and output offsets
    std::vector<uint64_t> start(nt + 1, first);
    std::vector<size_t> off(nt + 1, 0);
    uint64_t events = 0;
    for (uint64_t t = 0; t < nt; ++t) {
        start[t + 1] = start[t] + pieces[t].clusters;
        size_t len = pieces[t].text.n - pieces[t].marks.size();
        for (const auto &m : pieces[t].marks) len += (size_t)digits10(start[t] + m.second);
        off[t + 1] = off[t] + len;
        events += pieces[t].events;
    }
    char *buf = static_cast<char *>(std::malloc(off[nt] + 1));
    if (!buf) { e2i::set_error("e2i_snp_format: out of host memory"); return E2I_ERR_MEMORY; }
    auto emit = [&](uint64_t t) {
        const Piece &pc = pieces[t];
        char *w = buf + off[t];
        size_t pos = 0;
        for (const auto &m : pc.marks) {
            std::memcpy(w, pc.text.p + pos, m.first - pos);
            w += m.first - pos;
            char num[24];
            const int nd = std::snprintf(num, sizeof num, "%llu", (unsigned long long)(start[t] + m.second));
            std::memcpy(w, num, (size_t)nd);
            w += nd;
            pos = m.first + 1;
        }
        std::memcpy(w, pc.text.p + pos, pc.text.n - pos);
    };
    {
        std::vector<std::thread> th;
        for (uint64_t t = 1; t < nt; ++t) th.emplace_back(emit, t);
        emit(0);
        for (auto &x : th) x.join();
    }
    if (std::getenv("E2I_DEBUG")) {
        const auto t_c = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[e2i] format: %llu threads, pass 1 %.1f ms, numbering + copy %.1f ms, %zu bytes\n", (unsigned long long)nt,
                     std::chrono::duration<double, std::milli>(t_b - t_a).count(), std::chrono::duration<double, std::milli>(t_c - t_b).count(), off[nt]);
    }
    buf[off[nt]] = 0;
    *snp = buf;
    *snp_len = off[nt];
    if (st) {
        st->events += events;
        st->clusters_out += start[nt] - first;
    }
    return E2I_OK;
}

// Coverage filter of filter_snp.cpp:38-77.  Lines are taken in pairs (header, sequence); the
// coverage is the integer after the first ':' of the 4th '_'-separated token of the header
// (atoi semantics: 0 when absent); a pair is kept iff cov >= m and (M == 0 or cov <= M).  A
// trailing header without its sequence line prints nothing, like the reference.
extern "C" int e2i_filter_snp(const char *snp, size_t len, int32_t m, int32_t M, char **out, size_t *out_len) {
    if (!out || !out_len || (len && !snp)) { e2i::set_error("e2i_filter_snp: null argument"); return E2I_ERR_ARG; }
    char *buf = static_cast<char *>(std::malloc(len + 2));
    if (!buf) { e2i::set_error("e2i_filter_snp: out of host memory"); return E2I_ERR_MEMORY; }
    size_t w = 0, pos = 0, header_b = 0, header_e = 0;
    bool have_header = false;
    int cov = 0;
    while (pos < len) {
        size_t e = pos;
        while (e < len && snp[e] != '\n') ++e;                   // getline: [pos, e)
        if (!have_header) {
            header_b = pos; header_e = e; have_header = true;
            size_t t = pos;                                        // 4th '_' token
            for (int k = 0; k < 3 && t < e; ++k) { while (t < e && snp[t] != '_') ++t; if (t < e) ++t; }
            size_t te = t;
            while (te < e && snp[te] != '_') ++te;
            size_t c = t;
            while (c < te && snp[c] != ':') ++c;
            cov = 0;
            if (c < te) {                                          // atoi of the text after the first ':' up to the next ':'
                size_t v = c + 1, ve = v;
                while (ve < te && snp[ve] != ':') ++ve;
                cov = atoi(std::string(snp + v, ve - v).c_str());
            }
        } else {
            if (cov >= m && (M == 0 || cov <= M)) {
                std::memcpy(buf + w, snp + header_b, header_e - header_b); w += header_e - header_b; buf[w++] = '\n';
                std::memcpy(buf + w, snp + pos, e - pos); w += e - pos; buf[w++] = '\n';
            }
            have_header = false;
            cov = 0;
        }
        pos = e + 1;
    }
    buf[w] = 0;
    *out = buf;
    *out_len = w;
    return E2I_OK;
}

extern "C" void e2i_buffer_free(void *p) {
    if (p && !e2i::text_release(p)) std::free(p);
}
