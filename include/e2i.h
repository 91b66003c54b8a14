/*
 * e2i.h -- C ABI of the B200-native ebwt2InDel hot path (libe2i.so).
 *
 * The reference (nicolaprezza/ebwt2InDel) exposes no FFI: the path sits behind the process
 * boundary and the header-only class dna_bwt_t.  These entry points are what a binding of that
 * path would bind; each cites the reference interface it replaces (paths relative to the
 * reference root).  Plain pointers and sizes only; no C++ or torch types cross this boundary.
 *
 * Conventions: every function returning int returns 0 on success and a non-zero E2I_ERR_* code
 * on failure; e2i_last_error() returns a thread-local message.  Handles are opaque and owned by
 * the caller (free with the matching *_free).  One host thread per e2i_ctx.  Buffers named
 * host_* are caller-owned host memory (pageable or pinned); buffers named dev_* are device
 * pointers on the context's device.  Packed bitvectors are little-endian arrays of uint64_t
 * words: bit i = word i/64, bit i%64.  There is no CPU fallback: without a CUDA device every
 * compute entry point fails with E2I_ERR_CUDA.
 */
#ifndef E2I_H_
#define E2I_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define E2I_OK 0
#define E2I_ERR_CUDA 1       /* CUDA runtime error / no device */
#define E2I_ERR_SYMBOL 2     /* forbidden symbol in the input BWT (dna_string.hpp:90-96) */
#define E2I_ERR_ARG 3        /* invalid argument */
#define E2I_ERR_MEMORY 4     /* frontier / output capacity exceeded */
#define E2I_ERR_IO 5         /* file could not be read / written */

typedef struct e2i_ctx e2i_ctx;         /* device + streams + scratch */
typedef struct e2i_index e2i_index;     /* rank-indexed BWT in HBM: replaces dna_bwt_t (internal/dna_bwt.hpp:24-420) */
typedef struct e2i_bits e2i_bits;       /* packed bitvector in HBM: replaces vector<bool> DA (ebwt2InDel.cpp:58) */
typedef struct e2i_lcpbits e2i_lcpbits; /* LCP_threshold (2n bits) + LCP_minima (n bits) (ebwt2InDel.cpp:56-57) */
typedef struct e2i_comm e2i_comm;       /* ranks of a multi-GPU run: barrier + publish slots + peer memory mapping */
typedef struct e2i_calls e2i_calls;     /* per-cluster variant records, SA order (variant_t / variant_single_t, :115-141) */

/* Resolved parameters: the globals of ebwt2InDel.cpp:20-74 after the "0 means default" rule (:1740-1746). */
typedef struct {
    int32_t k_left;     /* -L, default 31: left-context length, SNP included */
    int32_t k_right;    /* -R, default 30: right-context length */
    int32_t K;          /* -k, default 16: minimum LCP inside clusters */
    int32_t max_gap;    /* -g, default 10: maximum indel length */
    int32_t max_snvs;   /* -v, default 2 */
    int32_t mcov_out;   /* -m, default 3: minimum coverage */
    int32_t complexity; /* -c, default 20 */
    int32_t max_variants_per_position; /* -q, default 0 = unlimited */
    int32_t term;       /* -t, default '#' */
} e2i_params;

/* Counters: the first block must equal what the reference prints on stdout. */
typedef struct {
    uint64_t leaves;             /* "Processed N suffix-tree leaves."    ebwt2InDel.cpp:620/762 */
    uint64_t nodes;              /* "Processed N suffix-tree nodes."     :673/829 */
    uint64_t lcp_values;         /* "Computed N/n LCP values."           :670/826 */
    uint64_t lcp_values_leaves;  /* "Computed N/n LCP threshold values." :617/759 */
    uint64_t n_min;              /* "Found N LCP minima."                :671/827 */
    uint64_t da_values;          /* "Computed N/n DA values."            :825 */
    uint64_t n_clusters;         /* "Analyzed N clusters."               :1448/1563/1658 */
    uint64_t clust_size;         /* cumulative cluster length (average = clust_size / n_clusters) */
    uint64_t events;             /* "Stored to file N events" (mode -1)  :1320 */
    uint64_t clusters_out;       /* cluster_nr - 1                       :1250/1328 */
    uint64_t clust_sizes[201];   /* CLUST_SIZES histogram                :1414/1532/1627 */
    /* work counters (same definitions as the reference's call counts) */
    uint64_t rank_leaves;        /* parallel_rank queries in phase 2 */
    uint64_t rank_nodes;         /* parallel_rank queries in phase 3 (distinct boundaries per node, dna_bwt.hpp:332-347) */
    uint64_t rank_call;          /* rank queries issued by phase 4 */
    uint64_t bit_updates;        /* LCP border + minima bit writes in phase 3 */
    uint64_t candidates;         /* clusters that passed the reference's frequent-allele filter (:870-880, :961-966) */
    uint64_t levels_leaves;      /* frontier sweeps of phase 2 */
    uint64_t levels_nodes;       /* frontier sweeps of phase 3 */
    uint64_t max_frontier;       /* largest frontier chunk (nodes) */
    /* device time per phase, milliseconds (CUDA events on the context's stream) */
    double ms_index;
    double ms_leaves;
    double ms_nodes;
    double ms_call;
    double ms_h2d;
    double ms_d2h;
    /* launch / transfer accounting of the calls that filled this struct */
    uint64_t kernel_launches;    /* kernels of this library launched */
    uint64_t h2d_bytes;          /* host -> device bytes copied (inputs) */
    uint64_t d2h_bytes;          /* device -> host bytes copied (call records, counters) */
    /* host wall-clock milliseconds (e2i_run / e2i_run_device only) */
    double ms_format;            /* e2i_snp_format */
    double ms_wall;              /* the whole call */
    uint64_t da_values_leaves;   /* "Computed N/n DA values." of the leaf pass (mode -2) :757 */
} e2i_stats;

/* One analysed cluster that passed the allele filter (and, with two samples, can emit a pair).
 * Left contexts live in a separate char array: 8 slots of k_left chars per record
 * (slots 0-3: individual 0 / the single sample; slots 4-7: individual 1), packed to the front
 * of each group in A,C,G,T order.  Records without a right context (has_right == 0) are kept in
 * the list but produce no output and advance no counter, like the reference's empty variant
 * vector (ebwt2InDel.cpp:909, 985, 1071). */
typedef struct {
    uint64_t begin;          /* merged SA position of the first flagged position */
    uint64_t end;            /* merged SA position one past the last */
    uint8_t n0, n1;          /* left contexts of individual 0 / 1 that reached k_left chars */
    uint8_t right_len;       /* chars in the right context (stops early at a terminator) */
    uint8_t has_right;       /* a position with LCP >= k_right exists in the cluster */
    int32_t support[8];      /* |LF(range, c)| per left context (ebwt2InDel.cpp:310) */
} e2i_call_rec;

const char *e2i_last_error(void);
const char *e2i_version(void);
void e2i_params_default(e2i_params *p);

/* Flag resolution of main(): 0 -> default for -L -R -k -g -v -m -c (ebwt2InDel.cpp:1740-1746). */
void e2i_params_resolve(e2i_params *p);

/* ---- context ---------------------------------------------------------------------------- */
int e2i_create(int device, e2i_ctx **out);
void e2i_destroy(e2i_ctx *ctx);
/* Upper bound on frontier memory in bytes (0 = use what is free on the device). */
int e2i_set_frontier_budget(e2i_ctx *ctx, uint64_t bytes);
/* Page-locked host staging buffers for the ASCII inputs (what the CLI reads the files into;
 * replaces the byte-at-a-time ifstream loops of dna_string.hpp:82-101 and ebwt2InDel.cpp:1503-1508). */
/* Return the device memory cached by the library's stream-ordered pool to the driver. */
int e2i_trim(e2i_ctx *ctx);
/* The context's CUDA stream (a cudaStream_t), so that callers can record their own events on it. */
void *e2i_stream(const e2i_ctx *ctx);
int e2i_host_alloc(uint64_t bytes, void **out);
void e2i_host_free(void *p);

/* ---- a1/a5: index build.  Replaces dna_bwt_t(path, TERM) (dna_bwt.hpp:36-62) and
 *      dna_string(path, TERM) (dna_string.hpp:55-110, 275-315).  Input: raw ASCII BWT over
 *      {A,C,G,T,term}; any other byte -> E2I_ERR_SYMBOL, *bad_pos = its position. ------------ */
int e2i_index_build(e2i_ctx *ctx, const uint8_t *host_ascii, uint64_t n, uint8_t term,
                    e2i_index **out, uint64_t *bad_pos);
int e2i_index_build_device(e2i_ctx *ctx, const uint8_t *dev_ascii, uint64_t n, uint8_t term,
                           e2i_index **out, uint64_t *bad_pos);
void e2i_index_free(e2i_index *ix);
uint64_t e2i_index_size(const e2i_index *ix);                   /* dna_bwt::size()          :231-236 */
int e2i_index_F(const e2i_index *ix, uint64_t F[4]);            /* F_A,F_C,F_G,F_T          :412-415 */
uint64_t e2i_index_bytes(const e2i_index *ix);                  /* HBM footprint */

/* ---- slice-wise index construction for multi-GPU runs: the string is cut into
 *      slices that start on multiples of e2i_index_slice_align() symbols; every rank counts and
 *      packs ONE slice into a pre-allocated index, the block ranges are exchanged between the GPUs
 *      (all-gather over NVLink, ebwt2indel_b200/distributed.py) and e2i_index_finish sets F.
 *      Same result as e2i_index_build_device on the whole string.
 *      tile_multiple: the block array is padded to a multiple of that many slices' worth of tiles.
 *      slice_count: counts[4] = #A,#C,#G,#T of the slice; n_tiles = the tiles of e2i_index_slice_align()
 *      symbols the slice owns (the last tile of the string may hold no symbol, only the block that makes
 *      rank(n) addressable; exactly one slice owns it); n_tiles = 0: an empty slice, nothing to do.  slice_super: host_super[n_super*4]
 *      receives the superblock entries this slice owns (zeros elsewhere) given the counts before
 *      the slice; the SUM over ranks is the complete table that slice_pack takes. --------------- */
uint64_t e2i_index_slice_align(void);
uint64_t e2i_index_super_count(const e2i_index *ix);   /* entries (4 x u64 each) of the superblock table */
int e2i_index_alloc(e2i_ctx *ctx, uint64_t n, uint8_t term, uint64_t tile_multiple, e2i_index **out);
int e2i_index_slice_count(e2i_ctx *ctx, e2i_index *ix, const uint8_t *dev_slice, uint64_t begin, uint64_t len,
                          uint64_t n_tiles, uint64_t counts[4], uint64_t *bad_pos);
int e2i_index_slice_super(e2i_ctx *ctx, e2i_index *ix, const uint64_t before[4], uint64_t *host_super);
int e2i_index_slice_pack(e2i_ctx *ctx, e2i_index *ix, const uint8_t *dev_slice, const uint64_t before[4],
                         const uint64_t *host_super);
int e2i_index_finish(e2i_index *ix, const uint64_t totals[4]);
int e2i_index_device(const e2i_index *ix, void **dev_blocks, uint64_t *block_bytes);

/* ---- a2/a3/a4 test hooks.  parallel_rank (dna_string.hpp:140-152), operator[] (:113-135),
 *      FL = select (dna_bwt.hpp:115-133, dna_string.hpp:254-272), batched. ------------------ */
int e2i_rank_batch(e2i_ctx *ctx, const e2i_index *ix, const uint64_t *host_pos, uint64_t m, uint64_t *host_out4);
int e2i_access_batch(e2i_ctx *ctx, const e2i_index *ix, const uint64_t *host_pos, uint64_t m, uint8_t *host_out);
int e2i_fl_batch(e2i_ctx *ctx, const e2i_index *ix, const uint64_t *host_pos, uint64_t m, uint64_t *host_out);
/* device-resident variant used by bench.py for the kernel-only rank throughput */
int e2i_rank_batch_device(e2i_ctx *ctx, const e2i_index *ix, const uint64_t *dev_pos, uint64_t m,
                          uint64_t *dev_out4, float *ms);

/* ---- document array (mode -d).  Replaces the loader at ebwt2InDel.cpp:1495-1508:
 *      one ASCII byte per BWT position, '1' -> 1, anything else -> 0. ---------------------- */
int e2i_da_load(e2i_ctx *ctx, const uint8_t *host_ascii01, uint64_t n, e2i_bits **out);
int e2i_da_load_device(e2i_ctx *ctx, const uint8_t *dev_ascii01, uint64_t n, e2i_bits **out);
int e2i_bits_fetch(e2i_ctx *ctx, const e2i_bits *b, uint64_t *host_words, uint64_t n_words);
uint64_t e2i_bits_size(const e2i_bits *b);
void e2i_bits_free(e2i_bits *b);

/* ---- a6-a15: phases 2+3.  Replaces navigate_one_bwt (ebwt2InDel.cpp:555-676) when b2 == NULL
 *      and navigate_two_bwts (:679-831) otherwise (then *da_out receives the merged DA).
 *      shard/n_shards: this call traverses only the subtrees dealt to `shard` (0 <= shard <
 *      n_shards); the bitvectors of all shards must be OR-combined before e2i_call. */
int e2i_navigate(e2i_ctx *ctx, const e2i_index *b1, const e2i_index *b2, const e2i_params *p,
                 e2i_lcpbits **out, e2i_bits **da_out, e2i_stats *st);
int e2i_navigate_shard(e2i_ctx *ctx, const e2i_index *b1, const e2i_index *b2, const e2i_params *p,
                       int shard, int n_shards, e2i_lcpbits **out, e2i_bits **da_out, e2i_stats *st);
int e2i_lcpbits_fetch(e2i_ctx *ctx, const e2i_lcpbits *l, uint64_t *host_thr_words, uint64_t *host_min_words);
/* raw device words (uint32, zero-padded to a multiple of 64 words) for the cross-GPU OR-reduce */
int e2i_lcpbits_device(const e2i_lcpbits *l, void **dev_thr, uint64_t *thr_words32,
                       void **dev_min, uint64_t *min_words32);
int e2i_bits_device(const e2i_bits *b, void **dev_words, uint64_t *words32);
void e2i_lcpbits_free(e2i_lcpbits *l);

/* ---- a16-a20: phase 4 on the device.  Replaces the cluster scans (ebwt2InDel.cpp:1609-1655,
 *      1395-1445, 1510-1560) and find_variants x3 (:840-934, 941-1005, 1013-1096).
 *      mode -1: b2 = NULL, da = NULL;  mode -2: b2, da = navigate's DA;  mode -d: b2 = NULL, da.
 *      [pos_begin, pos_end): only clusters that START in this merged SA range are analysed
 *      (0, UINT64_MAX = all); used to shard phase 4. ---------------------------------------- */
/* The records of an e2i_calls handle live in the context's page-locked staging buffer: a handle is
 * valid until the next e2i_call on the same context (e2i_calls_fetch / _view check this). */
int e2i_call(e2i_ctx *ctx, const e2i_index *b1, const e2i_index *b2, const e2i_bits *da,
             const e2i_lcpbits *l, const e2i_params *p, uint64_t pos_begin, uint64_t pos_end,
             e2i_calls **out, e2i_stats *st);
uint64_t e2i_calls_count(const e2i_calls *c);
/* host_left: cap * 8 * k_left chars; host_right: cap * k_right chars */
int e2i_calls_fetch(const e2i_calls *c, e2i_call_rec *host_recs, char *host_left, char *host_right,
                    uint64_t cap, uint64_t *n);
/* zero-copy view of the page-locked result arrays (valid until the next e2i_call on the context) */
int e2i_calls_view(const e2i_calls *c, const e2i_call_rec **recs, const char **left, const char **right, uint64_t *n);
void e2i_calls_free(e2i_calls *c);

/* ---- a21-a23: classification + .snp text.  Replaces distance/event_type/to_file x2
 *      (ebwt2InDel.cpp:143-240, 1102-1330).  two_samples = 0 for mode -1, 1 for modes -2/-d.
 *      first_cluster_nr: cluster number of the first emitted cluster (1 in a single-shard run);
 *      *snp is malloc'ed (free with e2i_buffer_free); st->events / clusters_out are updated. -- */
int e2i_snp_format(const e2i_call_rec *recs, const char *left, const char *right, uint64_t n_recs,
                   const e2i_params *p, int two_samples, uint64_t first_cluster_nr,
                   char **snp, size_t *snp_len, e2i_stats *st);
/* The same text, produced by the device formatter (one thread per record: measure, number the clusters with a
 * scan, place the records with a scan, write) from host records -- what e2i_call_snp runs on the records while they
 * are still in HBM; this entry exists so that the two formatters can be compared on any records. */
int e2i_snp_format_gpu(e2i_ctx *ctx, const e2i_call_rec *recs, const char *left, const char *right, uint64_t n_recs,
                       const e2i_params *p, int two_samples, uint64_t first_cluster_nr,
                       char **snp, size_t *snp_len, e2i_stats *st);
/* Phase 4 and the .snp text in one call: find_variants + to_file (ebwt2InDel.cpp:1344-1660, 1149-1330) for the
 * positions [pos_begin, pos_end); the records never leave the device, only the text is copied out (*snp is a
 * page-locked buffer owned by the library's cache: give it back with e2i_buffer_free).  This is what e2i_run / e2i_run_files / e2i_run_device use. */
int e2i_call_snp(e2i_ctx *ctx, const e2i_index *b1, const e2i_index *b2, const e2i_bits *da,
                 const e2i_lcpbits *l, const e2i_params *p, uint64_t pos_begin, uint64_t pos_end,
                 uint64_t first_cluster_nr, char **snp, size_t *snp_len, e2i_stats *st);
/* Phase 4 with the records left in HBM (no host copy): for callers that number the clusters of several position
 * ranges consecutively (one range per GPU).  e2i_calls_count works on the handle; the records are turned into text by
 *   e2i_calls_clusters   - how many cluster numbers the range consumes (device pass, no text)
 *   e2i_calls_snp        - the text from a given first cluster number (*snp: e2i_buffer_free)
 *   e2i_calls_snp_device - the same text left in device memory (for a gather over NVLink; e2i_device_free) */
int e2i_call_device(e2i_ctx *ctx, const e2i_index *b1, const e2i_index *b2, const e2i_bits *da,
                    const e2i_lcpbits *l, const e2i_params *p, uint64_t pos_begin, uint64_t pos_end,
                    e2i_calls **out, e2i_stats *st);
int e2i_calls_clusters(const e2i_calls *c, const e2i_params *p, uint64_t *clusters);
int e2i_calls_snp(const e2i_calls *c, const e2i_params *p, uint64_t first_cluster_nr, char **snp, size_t *snp_len, e2i_stats *st);
int e2i_calls_snp_device(const e2i_calls *c, const e2i_params *p, uint64_t first_cluster_nr, void **dev_text, uint64_t *len,
                         e2i_stats *st);
void e2i_device_free(e2i_ctx *ctx, void *p);
/* How many cluster numbers the records consume (= what e2i_snp_format adds to clusters_out),
 * without building text: lets every GPU rank learn its first cluster number (distributed.py). */
int e2i_snp_count(const e2i_call_rec *recs, const char *left, const char *right, uint64_t n_recs,
                  const e2i_params *p, int two_samples, uint64_t *clusters);
void e2i_distance(const char *a, const char *b, int32_t len, int32_t max_gap, int32_t out[2]);
/* ---- output-side coverage filter.  Replaces filter_snp (filter_snp.cpp:17-81): keeps the records
 *      whose `cov:` field is >= m and (M == 0 or <= M); *out is malloc'ed (e2i_buffer_free). ------- */
int e2i_filter_snp(const char *snp, size_t len, int32_t m, int32_t M, char **out, size_t *out_len);
void e2i_buffer_free(void *p);

/* ---- the whole path with host buffers: what bin/ebwt2InDel calls (run_one_dataset :1584,
 *      run_two_datasets :1344, run_two_datasets_da :1471).  host_bwt2 / host_da may be NULL. -- */
int e2i_run(e2i_ctx *ctx, const uint8_t *host_bwt1, uint64_t n1, const uint8_t *host_bwt2, uint64_t n2,
            const uint8_t *host_da, const e2i_params *p, char **snp, size_t *snp_len, e2i_stats *st);
/* The same from files (what bin/ebwt2InDel calls): streaming ingest -- a reader thread fills a ring of page-
 * locked buffers, the chunks go up on a copy stream and the counting pass of the index build follows them, so
 * disk reads, PCIe copies and kernels overlap (replaces the byte-at-a-time loops of dna_string.hpp:82-101 and
 * ebwt2InDel.cpp:1503-1508; a DA file shorter than the eBWT repeats its last byte like that loop does).  With
 * the environment variable E2I_INDEX_CACHE=1 the packed index is kept as <file>.e2ix and reused by later runs
 * (e2i_index_save / e2i_index_load: the reference's unused serialize / load, dna_bwt.hpp:238-289).
 * n1_out / n2_out: the eBWT lengths; bad_pos[2]: on E2I_ERR_SYMBOL the position and the input (1 or 2). */
int e2i_run_files(e2i_ctx *ctx, const char *path_bwt1, const char *path_bwt2, const char *path_da, const e2i_params *p,
                  char **snp, size_t *snp_len, e2i_stats *st, uint64_t *n1_out, uint64_t *n2_out, uint64_t *bad_pos);
int e2i_index_build_file(e2i_ctx *ctx, const char *path, uint8_t term, e2i_index **out, uint64_t *bad_pos);
int e2i_da_load_file(e2i_ctx *ctx, const char *path, uint64_t n, e2i_bits **out);
int e2i_index_save(const e2i_index *ix, const char *path);
int e2i_index_load(e2i_ctx *ctx, const char *path, e2i_index **out);
/* same, inputs already resident in HBM (kernel-side throughput in bench.py) */
int e2i_run_device(e2i_ctx *ctx, const uint8_t *dev_bwt1, uint64_t n1, const uint8_t *dev_bwt2, uint64_t n2,
                   const uint8_t *dev_da, const e2i_params *p, char **snp, size_t *snp_len, e2i_stats *st);

/* ---- eBWT (+ document array) construction on the GPU.  The reference takes its input from an
 *      external builder (BCR_LCP_GSA / egap, README.md:38, 91-92); this entry point makes the chain self-contained.
 *      host_reads: m x L ASCII matrix (row-major, A/C/G/T only, one fixed length L); reads with index >=
 *      second_from belong to the second individual (pass m for a single set).  Convention: '#'_i < '#'_j for
 *      i < j, '#' < A < C < G < T.  host_bwt: m * (L + 1) bytes, caller-allocated; host_da (nullable): the
 *      document array as ASCII '0' / '1', same size.  bin/ebwt_build is its command line. ---------------------- */
int e2i_ebwt_build(e2i_ctx *ctx, const uint8_t *host_reads, uint64_t m, uint32_t L, uint64_t second_from, uint8_t term,
                   uint8_t *host_bwt, uint8_t *host_da);

/* ---- several GPUs of one box, one process (replaces the parallel wrapper pebwt2InDel.sh:45-88 without its
 *      loss of cross-piece coverage: the output equals the single-GPU run's).  e2i_run_multi: the whole path
 *      on devices[0..n_devices); one host thread and one context per GPU inside the call; the index is built
 *      slice-wise and replicated through peer copies, the traversal is sharded, the bit vectors are OR-combined
 *      by e2i_or_allreduce over peer memory, phase 4 runs per suffix-array range.  frontier_budget: bytes per
 *      GPU (0 = what is free).
 *      e2i_enable_peers: mutual peer access (device + memory pool) between the contexts of ONE process.
 *      e2i_or_allreduce: rank's share of the OR of dev_words[0..n_ranks) (words32 32-bit words each, a
 *      multiple of 4): its 1/n_ranks slice is read from every peer, combined and stored back to every peer;
 *      the caller synchronises the ranks before (all inputs written) and after (all slices done). ------------ */
int e2i_run_multi(const int *devices, int n_devices, const uint8_t *host_bwt1, uint64_t n1, const uint8_t *host_bwt2,
                  uint64_t n2, const uint8_t *host_da, const e2i_params *p, uint64_t frontier_budget,
                  char **snp, size_t *snp_len, e2i_stats *st);
int e2i_enable_peers(e2i_ctx **ctxs, int n);
/* Position-range sharded traversal (the dense alternative to e2i_navigate_shard): rank r of `comm`
 * processes the nodes / leaves whose first suffix-array position lies in its 1/world range; after every level the
 * ranks publish where each (destination, queue) piece of their frame starts and the next sweep pulls its input
 * straight out of the peers' frames over NVLink (P2P loads); one barrier per level, no collective.  All ranks call
 * it together.  The bit vectors still have to be OR-combined afterwards.  Communicators: e2i_comm_local (world
 * handles for the threads of one process) or e2i_comm_shm (one handle per process of one box; `name` is a POSIX
 * shared-memory name chosen by the launcher, rank 0 creates it; peer device memory is mapped by CUDA IPC). */
int e2i_navigate_ranged(e2i_ctx *ctx, e2i_comm *comm, const e2i_index *b1, const e2i_index *b2, const e2i_params *p,
                        e2i_lcpbits **out, e2i_bits **da_out, e2i_stats *st);
int e2i_comm_local(int world, e2i_comm **out);
int e2i_comm_shm(const char *name, int rank, int world, e2i_comm **out);
void e2i_comm_barrier(e2i_comm *c);
void e2i_comm_free(e2i_comm *c);
int e2i_or_allreduce(e2i_ctx *ctx, void *const *dev_words, int n_ranks, int rank, uint64_t words32);

#ifdef __cplusplus
}
#endif
#endif /* E2I_H_ */
