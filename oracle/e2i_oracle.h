/*
 * e2i_oracle.h -- CPU oracle for the ebwt2InDel hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference algorithm (nicolaprezza/ebwt2InDel:
 * ebwt2InDel.cpp + internal/{dna_string,dna_bwt,include}.hpp).  It exists to check the CUDA
 * path; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
 * The product (libe2i.so, bin/ebwt2InDel) never links or calls it.
 *
 * Parity status: PINNED.  The reference has no golden vectors except the
 * distance() example at ebwt2InDel.cpp:186-189; the oracle is pinned against (1) that example,
 * (2) byte-identical .snp output and identical printed counters of the compiled reference
 * (oracle/_ref/ebwt2InDel, built by oracle/Makefile from /root/reference) on seeded inputs in
 * all three modes -- see tests/test_oracle_vs_ref.py and tests/golden/.
 *
 * All positions are uint64; bitvectors are little-endian arrays of uint64 words
 * (bit i = word i/64, bit i%64), the same layout the CUDA library exports.
 */
#ifndef E2I_ORACLE_H_
#define E2I_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_bwt orc_bwt;

/* Resolved parameters (after the "0 means default" rule of ebwt2InDel.cpp:1740-1746). */
typedef struct {
    int32_t k_left;     /* -L, default 31 */
    int32_t k_right;    /* -R, default 30 */
    int32_t K;          /* -k, default 16 */
    int32_t max_gap;    /* -g, default 10 */
    int32_t max_snvs;   /* -v, default 2  */
    int32_t mcov_out;   /* -m, default 3  */
    int32_t complexity; /* -c, default 20 */
    int32_t max_variants_per_position; /* -q, default 0 = unlimited */
    int32_t term;       /* -t, default '#' */
} orc_params;

typedef struct {
    uint64_t leaves;            /* "Processed N suffix-tree leaves."      ebwt2InDel.cpp:620/762 */
    uint64_t nodes;             /* "Processed N suffix-tree nodes."       :673/829 */
    uint64_t lcp_values;        /* "Computed N/n LCP values."             :670/826 */
    uint64_t lcp_values_leaves; /* "Computed N/n LCP threshold values."   :617/759 */
    uint64_t n_min;             /* "Found N LCP minima."                  :671/827 */
    uint64_t da_values;         /* "Computed N/n DA values."              :825 (mode 2) */
    uint64_t max_stack_leaves;  /* "Max stack depth"                      :619/761 */
    uint64_t max_stack_nodes;   /*                                        :672/828 */
    uint64_t n_clusters;        /* "Analyzed N clusters."                 :1448/1563/1658 */
    uint64_t clust_size;        /* cumulative cluster length              :1412/1530/1625 */
    uint64_t events;            /* "Stored to file N events" (mode 1 only) :1320 */
    uint64_t clusters_out;      /* cluster_nr - 1                          :1250/1328 */
    uint64_t rank_leaves;       /* parallel_rank calls in phase 2 */
    uint64_t rank_nodes;        /* parallel_rank calls in phase 3 */
    uint64_t rank_call;         /* parallel_rank calls in phase 4 */
    uint64_t lcp_border_updates;/* update_lcp_threshold writes in phase 3 */
    uint64_t clust_sizes[201];  /* CLUST_SIZES histogram                   :1414/1532/1627 */
} orc_stats;

/* a1/a5: dna_string ctor + dna_bwt ctor.  Returns NULL on a forbidden symbol; *bad_pos gets its position. */
orc_bwt *orc_bwt_build(const uint8_t *ascii, uint64_t n, uint8_t term, uint64_t *bad_pos);
void orc_bwt_free(orc_bwt *b);
uint64_t orc_bwt_size(const orc_bwt *b);
void orc_bwt_F(const orc_bwt *b, uint64_t F[4]);                 /* F_A, F_C, F_G, F_T */
void orc_rank4(const orc_bwt *b, uint64_t i, uint64_t out[4]);   /* a2: parallel_rank */
void orc_rank4_batch(const orc_bwt *b, const uint64_t *pos, uint64_t m, uint64_t *out4);
uint8_t orc_access(const orc_bwt *b, uint64_t i);                /* a3: operator[] */
uint64_t orc_select(const orc_bwt *b, uint64_t r, uint8_t c);    /* a4: select */
uint64_t orc_FL(const orc_bwt *b, uint64_t i);                   /* a20: FL */

/* a21 (known-answer: distance("ACCTACTG","TTACTTAC") = <1,2>). */
void orc_distance(const char *a, const char *b, int32_t len, int32_t max_gap, int32_t out[2]);

/* a11-a15: phases 2+3.  thr must hold 2n bits, minima n bits, da n bits (n = n1+n2 in mode 2); zeroed by callee. */
int orc_navigate_one(const orc_bwt *b, const orc_params *p, uint64_t *thr, uint64_t *minima, orc_stats *st);
int orc_navigate_two(const orc_bwt *b1, const orc_bwt *b2, const orc_params *p,
                     uint64_t *thr, uint64_t *minima, uint64_t *da, orc_stats *st);

/* a16-a23: phase 4 -> .snp text (malloc'ed, caller frees with orc_free).
 * mode 1: b2 = NULL, da = NULL.  mode 2: b2 != NULL, da = navigate_two's DA.  mode 3: b2 = NULL, da = packed DA. */
int orc_call(const orc_bwt *b1, const orc_bwt *b2, const uint64_t *da, const uint64_t *thr,
             const uint64_t *minima, const orc_params *p, char **snp, size_t *snp_len, orc_stats *st);

/* Whole path from ASCII inputs (mode by which of ascii2 / da_ascii is non-NULL). */
int orc_run(const uint8_t *ascii1, uint64_t n1, const uint8_t *ascii2, uint64_t n2,
            const uint8_t *da_ascii, const orc_params *p, char **snp, size_t *snp_len, orc_stats *st);

void orc_params_default(orc_params *p);
void orc_free(void *ptr);

#ifdef __cplusplus
}
#endif
#endif
