/*
 * e2i_oracle.c -- CPU oracle for the ebwt2InDel hot path.  TEST INFRASTRUCTURE ONLY
 * (see e2i_oracle.h for who may use it and how it is pinned against the compiled reference).
 *
 * Each function cites the reference lines it restates; paths are relative to /root/reference.
 */
#include "e2i_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * a1/a2/a3/a4: rank string.  Restates internal/dna_string.hpp.
 * 64-byte block = 128 symbols: w[1]|w[0] plane 0, w[3]|w[2] plane 1, w[5]|w[4] plane 2 (TERM),
 * symbol j of the block at bit 63-(j%64) of the odd word (j<64) or the even word (j>=64)
 * (dna_string.hpp:320-369, 408-434); w[6], w[7] = four u32 counters relative to the 2^32-symbol
 * superblock (dna_string.hpp:554-585); superblock table of absolute counts (:65, 275-315).
 * ---------------------------------------------------------------------------------------- */
#define SB_SHIFT 32
#define BLK 128u

struct orc_bwt {
    uint64_t n, n_blocks, n_super;
    uint64_t *w;        /* 8 words per block */
    uint64_t *super;    /* 4 words per superblock */
    uint64_t F[4];      /* F_A, F_C, F_G, F_T  (dna_bwt.hpp:47-60) */
    uint8_t term;
    uint64_t *rank_calls; /* points at the counter of the running phase (may be NULL) */
};

static uint64_t g_rank_sink;

static inline int popc64(uint64_t x) { return __builtin_popcountll(x); }

static inline int sym_code(uint8_t c, uint8_t term) {
    if (c == term) return 4;
    switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; }
    return -1;
}

/* whole-block A,C,G,T counts (dna_string.hpp:439-462) */
static void block_counts(const uint64_t *w, uint64_t out[4]) {
    for (int h = 0; h < 2; ++h) {
        uint64_t p0 = w[h], p1 = w[2 + h], nt = ~w[4 + h];
        out[0] += popc64(nt & ~p1 & ~p0);
        out[1] += popc64(nt & ~p1 & p0);
        out[2] += popc64(nt & p1 & ~p0);
        out[3] += popc64(nt & p1 & p0);
    }
}

orc_bwt *orc_bwt_build(const uint8_t *ascii, uint64_t n, uint8_t term, uint64_t *bad_pos) {
    orc_bwt *b = (orc_bwt *)calloc(1, sizeof(orc_bwt));
    b->n = n;
    b->term = term;
    b->n_blocks = (n + 1) / BLK + ((n + 1) % BLK != 0);                  /* :62 */
    b->n_super = ((n + 1) >> SB_SHIFT) + (((n + 1) & 0xffffffffull) != 0); /* :61 */
    b->w = (uint64_t *)calloc(b->n_blocks * 8, sizeof(uint64_t));
    b->super = (uint64_t *)calloc(b->n_super * 4, sizeof(uint64_t));
    b->rank_calls = &g_rank_sink;
    uint64_t cnt[5] = {0, 0, 0, 0, 0};
    for (uint64_t i = 0; i < n; ++i) {                                    /* :82-101, set :320-369 */
        int c = sym_code(ascii[i], term);
        if (c < 0) {                                                      /* :90-96 */
            if (bad_pos) *bad_pos = i;
            orc_bwt_free(b);
            return NULL;
        }
        cnt[c]++;
        uint64_t *w = b->w + (i / BLK) * 8;
        unsigned j = (unsigned)(i % BLK);
        unsigned word = j < 64 ? 1 : 0;
        uint64_t bit = 1ull << (63 - (j % 64));
        if (c & 1) w[word] |= bit;
        if (c & 2) w[2 + word] |= bit;
        if (c & 4) w[4 + word] |= bit;
    }
    /* build_rank_support (:275-315) */
    uint64_t sb_r[4] = {0, 0, 0, 0}, bl_r[4] = {0, 0, 0, 0};
    const uint64_t blocks_per_super = 1ull << (SB_SHIFT - 7);
    for (uint64_t bl = 0; bl < b->n_blocks; ++bl) {
        if (bl % blocks_per_super == 0) {
            memcpy(b->super + (bl / blocks_per_super) * 4, sb_r, sizeof sb_r);
            memset(bl_r, 0, sizeof bl_r);
        }
        uint64_t *w = b->w + bl * 8;
        w[6] = bl_r[0] | (bl_r[1] << 32);                                 /* :554-567 */
        w[7] = bl_r[2] | (bl_r[3] << 32);
        if (bl + 1 < b->n_blocks) {
            uint64_t loc[4] = {0, 0, 0, 0};
            block_counts(w, loc);
            for (int k = 0; k < 4; ++k) { bl_r[k] += loc[k]; sb_r[k] += loc[k]; }
        }
    }
    b->F[0] = cnt[4];                                                     /* dna_bwt.hpp:47-60 */
    b->F[1] = b->F[0] + cnt[0];
    b->F[2] = b->F[1] + cnt[1];
    b->F[3] = b->F[2] + cnt[2];
    return b;
}

void orc_bwt_free(orc_bwt *b) {
    if (!b) return;
    free(b->w);
    free(b->super);
    free(b);
}

uint64_t orc_bwt_size(const orc_bwt *b) { return b->n; }
void orc_bwt_F(const orc_bwt *b, uint64_t F[4]) { memcpy(F, b->F, sizeof b->F); }

/* parallel_rank (dna_string.hpp:140-152) = superblock + block counters + block_rank (:375-434) */
void orc_rank4(const orc_bwt *b, uint64_t i, uint64_t out[4]) {
    ++*b->rank_calls;
    uint64_t sb = i >> SB_SHIFT;
    const uint64_t *w = b->w + (i / BLK) * 8;
    unsigned off = (unsigned)(i % BLK);
    /* positions < off: bits 63..(64-off) of the odd word, then of the even word */
    uint64_t keep_hi = off >= 64 ? ~0ull : (off == 0 ? 0 : ~(~0ull >> off));
    uint64_t keep_lo = off <= 64 ? 0 : ~(~0ull >> (off - 64));
    uint64_t keep[2] = {keep_lo, keep_hi};
    uint64_t r[4] = {0, 0, 0, 0};
    for (int h = 0; h < 2; ++h) {
        uint64_t p0 = w[h], p1 = w[2 + h], nt = ~w[4 + h] & keep[h];
        r[0] += popc64(nt & ~p1 & ~p0);
        r[1] += popc64(nt & ~p1 & p0);
        r[2] += popc64(nt & p1 & ~p0);
        r[3] += popc64(nt & p1 & p0);
    }
    out[0] = b->super[sb * 4 + 0] + (w[6] & 0xffffffffull) + r[0];
    out[1] = b->super[sb * 4 + 1] + (w[6] >> 32) + r[1];
    out[2] = b->super[sb * 4 + 2] + (w[7] & 0xffffffffull) + r[2];
    out[3] = b->super[sb * 4 + 3] + (w[7] >> 32) + r[3];
}

void orc_rank4_batch(const orc_bwt *b, const uint64_t *pos, uint64_t m, uint64_t *out4) {
    for (uint64_t k = 0; k < m; ++k) orc_rank4(b, pos[k], out4 + 4 * k);
}

/* operator[] (dna_string.hpp:113-135) */
uint8_t orc_access(const orc_bwt *b, uint64_t i) {
    const uint64_t *w = b->w + (i / BLK) * 8;
    unsigned j = (unsigned)(i % BLK);
    unsigned word = j < 64 ? 1 : 0, sh = 63 - (j % 64);
    unsigned code = (unsigned)((w[word] >> sh) & 1) | (unsigned)(((w[2 + word] >> sh) & 1) << 1) |
                    (unsigned)(((w[4 + word] >> sh) & 1) << 2);
    return code == 4 ? b->term : (uint8_t)"ACGT"[code & 3];
}

/* rank(i,c) (dna_string.hpp:157-174), rank_non_dna (:194-203) */
static uint64_t rank1(const orc_bwt *b, uint64_t i, uint8_t c) {
    uint64_t r[4];
    orc_rank4(b, i, r);
    if (c == b->term) {
        uint64_t r2[4];
        orc_rank4(b, i, r2); /* the reference calls parallel_rank twice here (:161, :197) */
        return i - (r2[0] + r2[1] + r2[2] + r2[3]);
    }
    switch (c) { case 'A': return r[0]; case 'C': return r[1]; case 'G': return r[2]; case 'T': return r[3]; }
    return 0;
}

/* select by binary search on rank (dna_string.hpp:182-188, 254-272) */
uint64_t orc_select(const orc_bwt *b, uint64_t i, uint8_t c) {
    uint64_t begin = 0, end = b->n;
    while (end != begin + 1) {
        uint64_t m = (begin + end) / 2;
        if (rank1(b, m, c) > i) end = m; else begin = m;
    }
    return begin;
}

/* F (dna_bwt.hpp:100-110) and FL (:115-133) */
static uint8_t F_col(const orc_bwt *b, uint64_t i) {
    return i < b->F[0] ? b->term : i < b->F[1] ? 'A' : i < b->F[2] ? 'C' : i < b->F[3] ? 'G' : 'T';
}

uint64_t orc_FL(const orc_bwt *b, uint64_t i) {
    uint8_t c = F_col(b, i);
    uint64_t r = i < b->F[0] ? i : i < b->F[1] ? i - b->F[0] : i < b->F[2] ? i - b->F[1]
               : i < b->F[3] ? i - b->F[2] : i - b->F[3];
    return orc_select(b, r, c);
}

/* ------------------------------------------------------------------------------------------
 * a6-a10: node algebra.  Restates internal/include.hpp:394-559, 672-815 and dna_bwt.hpp:138-404.
 * node = 6 boundaries {first_TERM, first_A, first_C, first_G, first_T, last} + depth.
 * ---------------------------------------------------------------------------------------- */
typedef struct { uint64_t b[6]; uint64_t depth; } node_t;
typedef struct { uint64_t first, second, depth; } leaf_t;

/* LF(range_t) -> 4 ranges (dna_bwt.hpp:138-166); an empty range reuses the first rank (:148-151) */
static void lf_range(const orc_bwt *b, uint64_t first, uint64_t second, uint64_t lo[4], uint64_t hi[4]) {
    uint64_t s[4], e[4];
    orc_rank4(b, first, s);
    if (second > first) orc_rank4(b, second, e); else memcpy(e, s, sizeof s);
    for (int c = 0; c < 4; ++c) { lo[c] = b->F[c] + s[c]; hi[c] = b->F[c] + e[c]; }
}

/* LF(range_t, c) (dna_bwt.hpp:168-192): two single-symbol ranks */
static void lf_range_c(const orc_bwt *b, uint64_t first, uint64_t second, int c, uint64_t *lo, uint64_t *hi) {
    uint64_t s = rank1(b, first, (uint8_t)"ACGT"[c]);
    uint64_t e = second > first ? rank1(b, second, (uint8_t)"ACGT"[c]) : s;
    *lo = b->F[c] + s;
    *hi = b->F[c] + e;
}

/* LF(sa_node) (dna_bwt.hpp:323-356): equal neighbouring boundaries reuse the previous rank */
static void lf_node(const orc_bwt *b, const node_t *N, node_t out[4]) {
    uint64_t r[6][4];
    orc_rank4(b, N->b[0], r[0]);
    for (int j = 1; j < 6; ++j) {
        if (N->b[j] == N->b[j - 1]) memcpy(r[j], r[j - 1], sizeof r[j]);
        else orc_rank4(b, N->b[j], r[j]);
    }
    for (int c = 0; c < 4; ++c) {
        for (int j = 0; j < 6; ++j) out[c].b[j] = b->F[c] + r[j][c];
        out[c].depth = N->depth + 1;
    }
}

static int n_children(const node_t *N) {                       /* include.hpp:760-768 */
    int k = 0;
    for (int j = 0; j < 5; ++j) k += N->b[j + 1] > N->b[j];
    return k;
}
static int n_children2(const node_t *A, const node_t *B) {     /* include.hpp:784-792 */
    int k = 0;
    for (int j = 0; j < 5; ++j) k += (A->b[j + 1] > A->b[j]) || (B->b[j + 1] > B->b[j]);
    return k;
}
static uint64_t node_size(const node_t *N) { return N->b[5] - N->b[0]; }   /* include.hpp:437-439 */

/* std::sort on <=4 elements is an insertion sort => stable; restated as a stable insertion sort
 * on precomputed keys, ascending (dna_bwt.hpp:374-377, 399-402; ebwt2InDel.cpp:467-470, 548-551). */
static void stable_order4(const uint64_t *key, int t, int *order) {
    for (int i = 0; i < t; ++i) order[i] = i;
    for (int i = 1; i < t; ++i) {
        int v = order[i], j = i - 1;
        while (j >= 0 && key[order[j]] > key[v]) { order[j + 1] = order[j]; --j; }
        order[j + 1] = v;
    }
}

static inline void bit_set(uint64_t *bv, uint64_t i, int v) {
    if (v) bv[i >> 6] |= 1ull << (i & 63); else bv[i >> 6] &= ~(1ull << (i & 63));
}
static inline int bit_get(const uint64_t *bv, uint64_t i) { return (int)((bv[i >> 6] >> (i & 63)) & 1); }

/* growable LIFO */
typedef struct { char *data; size_t size, cap, esz; } stack_t_;
static void st_init(stack_t_ *s, size_t esz) { s->data = NULL; s->size = s->cap = 0; s->esz = esz; }
static void st_push(stack_t_ *s, const void *e) {
    if (s->size == s->cap) { s->cap = s->cap ? 2 * s->cap : 64; s->data = (char *)realloc(s->data, s->cap * s->esz); }
    memcpy(s->data + s->size * s->esz, e, s->esz);
    s->size++;
}
static void st_pop(stack_t_ *s, void *e) { s->size--; memcpy(e, s->data + s->size * s->esz, s->esz); }

/* update_lcp_threshold (include.hpp:826-860) */
static void update_lcp_threshold(const node_t *x, uint64_t *thr, uint64_t *lcp_values, const orc_params *p, orc_stats *st) {
    for (int j = 1; j <= 4; ++j) {
        if (x->b[j] > x->b[j - 1] && x->b[j] != x->b[5]) {
            bit_set(thr, 2 * x->b[j], x->depth >= (uint64_t)p->K);
            bit_set(thr, 2 * x->b[j] + 1, x->depth >= (uint64_t)p->k_right);
            ++*lcp_values;
            st->lcp_border_updates++;
        }
    }
}

/* update_lcp_minima (ebwt2InDel.cpp:357-391): children A, C, G only */
static void update_lcp_minima(const node_t *x, uint64_t *minima, uint64_t *n_min) {
    for (int j = 2; j <= 4; ++j) {
        if (x->b[j] - x->b[j - 1] >= 2 && x->b[j] < x->b[5] - 1) {
            bit_set(minima, x->b[j], 1);
            ++*n_min;
        }
    }
}

/* navigate_one_bwt (ebwt2InDel.cpp:555-676) */
int orc_navigate_one(const orc_bwt *b, const orc_params *p, uint64_t *thr, uint64_t *minima, orc_stats *st) {
    orc_bwt *mb = (orc_bwt *)b;
    uint64_t n = b->n;
    memset(thr, 0, ((2 * n + 63) / 64) * 8);
    memset(minima, 0, ((n + 63) / 64) * 8);
    uint64_t lcp_values = 1;

    /* Phase 2: leaves (:577-615); update_LCP_leaf (:344-355); next_leaves (dna_bwt.hpp:358-379) */
    mb->rank_calls = &st->rank_leaves;
    stack_t_ S;
    st_init(&S, sizeof(leaf_t));
    leaf_t L0 = {0, b->F[0], 0};                                    /* first_leaf, dna_bwt.hpp:313-317 */
    st_push(&S, &L0);
    while (S.size) {
        leaf_t L;
        st_pop(&S, &L);
        st->leaves++;
        if (S.size > st->max_stack_leaves) st->max_stack_leaves = S.size;
        for (uint64_t i = L.first + 1; i < L.second; ++i) {
            bit_set(thr, 2 * i, L.depth >= (uint64_t)p->K);
            bit_set(thr, 2 * i + 1, L.depth >= (uint64_t)p->k_right);
            lcp_values++;
        }
        uint64_t lo[4], hi[4], key[4];
        leaf_t tmp[4];
        int t = 0, order[4];
        lf_range(b, L.first, L.second, lo, hi);
        for (int c = 0; c < 4; ++c)
            if (hi[c] - lo[c] >= 2) { tmp[t].first = lo[c]; tmp[t].second = hi[c]; tmp[t].depth = L.depth + 1; key[t] = hi[c] - lo[c]; t++; }
        stable_order4(key, t, order);
        for (int i = t - 1; i >= 0; --i) st_push(&S, &tmp[order[i]]);
    }
    free(S.data);
    st->lcp_values_leaves = lcp_values;

    /* Phase 3: internal nodes (:631-668); next_nodes (dna_bwt.hpp:381-404) */
    mb->rank_calls = &st->rank_nodes;
    st_init(&S, sizeof(node_t));
    node_t root = {{0, b->F[0], b->F[1], b->F[2], b->F[3], n}, 0};  /* dna_bwt.hpp:296-308 */
    st_push(&S, &root);
    while (S.size) {
        if (S.size > st->max_stack_nodes) st->max_stack_nodes = S.size;
        node_t N;
        st_pop(&S, &N);
        st->nodes++;
        update_lcp_threshold(&N, thr, &lcp_values, p, st);
        update_lcp_minima(&N, minima, &st->n_min);
        node_t ch[4], tmp[4];
        uint64_t key[4];
        int t = 0, order[4];
        lf_node(b, &N, ch);
        for (int c = 0; c < 4; ++c)
            if (n_children(&ch[c]) >= 2) { tmp[t] = ch[c]; key[t] = node_size(&ch[c]); t++; }
        stable_order4(key, t, order);
        for (int i = t - 1; i >= 0; --i) st_push(&S, &tmp[order[i]]);
    }
    free(S.data);
    st->lcp_values = lcp_values;
    mb->rank_calls = &g_rank_sink;
    return 0;
}

/* update_DA (ebwt2InDel.cpp:394-449) */
static void update_DA(const leaf_t *L1, const leaf_t *L2, uint64_t *da, uint64_t *thr, uint64_t *lcp_values,
                      uint64_t *m, const orc_params *p, int with_lcp) {
    uint64_t start1 = L1->first + L2->first, start2 = L2->first + L1->second, end = L1->second + L2->second;
    for (uint64_t i = start1; i < start2; ++i) { bit_set(da, i, 0); ++*m; }
    for (uint64_t i = start2; i < end; ++i) { bit_set(da, i, 1); ++*m; }
    if (!with_lcp) return;
    for (uint64_t i = start1 + 1; i < end; ++i) {
        bit_set(thr, 2 * i, L1->depth >= (uint64_t)p->K);
        bit_set(thr, 2 * i + 1, L1->depth >= (uint64_t)p->k_right);
        ++*lcp_values;
    }
}

/* navigate_two_bwts (ebwt2InDel.cpp:679-831) */
int orc_navigate_two(const orc_bwt *b1, const orc_bwt *b2, const orc_params *p,
                     uint64_t *thr, uint64_t *minima, uint64_t *da, orc_stats *st) {
    orc_bwt *m1 = (orc_bwt *)b1, *m2 = (orc_bwt *)b2;
    uint64_t n = b1->n + b2->n;
    memset(thr, 0, ((2 * n + 63) / 64) * 8);
    memset(minima, 0, ((n + 63) / 64) * 8);
    memset(da, 0, ((n + 63) / 64) * 8);
    uint64_t lcp_values = 1, da_values = 0;

    m1->rank_calls = m2->rank_calls = &st->rank_leaves;
    typedef struct { leaf_t a, b; } lpair;
    stack_t_ S;
    st_init(&S, sizeof(lpair));
    lpair P0 = {{0, b1->F[0], 0}, {0, b2->F[0], 0}};
    st_push(&S, &P0);
    while (S.size) {
        lpair L;
        st_pop(&S, &L);
        st->leaves++;
        if (S.size > st->max_stack_leaves) st->max_stack_leaves = S.size;
        update_DA(&L.a, &L.b, da, thr, &lcp_values, &da_values, p, 1);
        /* next_leaves, two BWTs (:452-472): keep children whose summed size >= 2 */
        uint64_t lo1[4], hi1[4], lo2[4], hi2[4], key[4];
        lpair tmp[4];
        int t = 0, order[4];
        lf_range(b1, L.a.first, L.a.second, lo1, hi1);
        lf_range(b2, L.b.first, L.b.second, lo2, hi2);
        for (int c = 0; c < 4; ++c) {
            uint64_t sz = (hi1[c] - lo1[c]) + (hi2[c] - lo2[c]);
            if (sz >= 2) {
                tmp[t].a.first = lo1[c]; tmp[t].a.second = hi1[c]; tmp[t].a.depth = L.a.depth + 1;
                tmp[t].b.first = lo2[c]; tmp[t].b.second = hi2[c]; tmp[t].b.depth = L.b.depth + 1;
                key[t] = sz; t++;
            }
        }
        stable_order4(key, t, order);
        for (int i = t - 1; i >= 0; --i) st_push(&S, &tmp[order[i]]);
    }
    free(S.data);
    st->lcp_values_leaves = lcp_values;

    m1->rank_calls = m2->rank_calls = &st->rank_nodes;
    typedef struct { node_t a, b; } npair;
    st_init(&S, sizeof(npair));
    npair R0 = {{{0, b1->F[0], b1->F[1], b1->F[2], b1->F[3], b1->n}, 0},
                {{0, b2->F[0], b2->F[1], b2->F[2], b2->F[3], b2->n}, 0}};
    st_push(&S, &R0);
    while (S.size) {
        if (S.size > st->max_stack_nodes) st->max_stack_nodes = S.size;
        npair N;
        st_pop(&S, &N);
        st->nodes++;
        node_t merged;                                               /* merge_nodes, include.hpp:476-490 */
        for (int j = 0; j < 6; ++j) merged.b[j] = N.a.b[j] + N.b.b[j];
        merged.depth = N.a.depth;
        /* find_leaves (:474-527): children of summed size exactly 1 were skipped by the leaf pass */
        for (int j = 0; j < 5; ++j) {
            uint64_t s1 = N.a.b[j + 1] - N.a.b[j], s2 = N.b.b[j + 1] - N.b.b[j];
            if (s1 + s2 == 1) {
                leaf_t l1 = {N.a.b[j], N.a.b[j + 1], 0}, l2 = {N.b.b[j], N.b.b[j + 1], 0};
                update_DA(&l1, &l2, da, thr, &lcp_values, &da_values, p, 0);
            }
        }
        update_lcp_threshold(&merged, thr, &lcp_values, p, st);
        update_lcp_minima(&merged, minima, &st->n_min);
        /* next_nodes, two BWTs (:529-553) */
        node_t c1[4], c2[4];
        npair tmp[4];
        uint64_t key[4];
        int t = 0, order[4];
        lf_node(b1, &N.a, c1);
        lf_node(b2, &N.b, c2);
        for (int c = 0; c < 4; ++c)
            if (n_children2(&c1[c], &c2[c]) >= 2) { tmp[t].a = c1[c]; tmp[t].b = c2[c]; key[t] = node_size(&c1[c]) + node_size(&c2[c]); t++; }
        stable_order4(key, t, order);
        for (int i = t - 1; i >= 0; --i) st_push(&S, &tmp[order[i]]);
    }
    free(S.data);
    st->lcp_values = lcp_values;
    st->da_values = da_values;
    m1->rank_calls = m2->rank_calls = &g_rank_sink;
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * a21: has_run / dH / distance / event_type.  Restates ebwt2InDel.cpp:143-240, 1102-1144.
 * ---------------------------------------------------------------------------------------- */
static int has_run(const char *s, int len, int k) {                  /* :144-152 */
    if (k < 0 || k > len) return 0;
    for (int i = 1; i < k; ++i) if (s[i] != s[i - 1]) return 0;
    return 1;
}

static int dH(const char *a, int la, const char *b, int lb) {        /* :157-171 */
    int len = la < lb ? la : lb, d = 0;
    for (int i = 0; i < len; ++i) d += a[la - i - 1] != b[lb - i - 1];
    return d;
}

void orc_distance(const char *a, const char *b, int32_t len, int32_t max_gap, int32_t out[2]) { /* :192-240 */
    int no_indel = dH(a, len, b, len);
    if (max_gap == 0) { out[0] = no_indel; out[1] = 0; return; }
    int best_ab = 0, best_ba = 0, iab = 0, iba = 0;
    for (int i = 1; i <= max_gap; ++i) {
        int la = len - i < 0 ? len : len - i; /* substr(0, len-i) with size_t wrap keeps the whole string */
        int dab = dH(a, la, b, len) + i;
        int dba = dH(a, len, b, la) + i;
        if (i == 1 || dab < best_ab) { best_ab = dab; iab = i; }    /* first minimum (:220-221) */
        if (i == 1 || dba < best_ba) { best_ba = dba; iba = i; }
    }
    if (no_indel < best_ab && no_indel < best_ba) { out[0] = no_indel; out[1] = 0; }
    else if (best_ab < best_ba) { out[0] = best_ab - iab; out[1] = iab; }
    else { out[0] = best_ba - iba; out[1] = -iba; }
}

/* growable text buffer */
typedef struct { char *s; size_t len, cap; } buf_t;
static void buf_add(buf_t *o, const char *s, size_t n) {
    if (o->len + n + 1 > o->cap) { o->cap = (o->len + n + 1) * 2 + 256; o->s = (char *)realloc(o->s, o->cap); }
    memcpy(o->s + o->len, s, n);
    o->len += n;
    o->s[o->len] = 0;
}
static void buf_str(buf_t *o, const char *s) { buf_add(o, s, strlen(s)); }
static void buf_u64(buf_t *o, uint64_t v) { char t[32]; snprintf(t, sizeof t, "%llu", (unsigned long long)v); buf_str(o, t); }
static void buf_int(buf_t *o, int v) { char t[32]; snprintf(t, sizeof t, "%d", v); buf_str(o, t); }

/* event_type (:1102-1144) */
static void event_type(buf_t *o, const char *l0, const char *l1, int len, const int32_t d[2]) {
    buf_str(o, "type:");
    buf_str(o, d[1] != 0 ? "_INDEL_event:" : "_SNP_event:");
    if (d[1] == 0) { buf_add(o, l0 + len - 1, 1); buf_str(o, "/"); buf_add(o, l1 + len - 1, 1); }
    else if (d[1] > 0) { buf_add(o, l0 + len - d[1], (size_t)d[1]); buf_str(o, "/"); }
    else { buf_str(o, "/"); buf_add(o, l1 + len + d[1], (size_t)(-d[1])); }
}

/* ------------------------------------------------------------------------------------------
 * a17-a20: find_variants and its helpers.  Restates ebwt2InDel.cpp:243-342, 840-1096.
 * ---------------------------------------------------------------------------------------- */
#define MAX_CTX 4096
typedef struct { char ctx[MAX_CTX]; int len; int support; } lctx_t;

/* extract_consensus (:265-282, 303-319) + consensus_letter (:243-261: largest child, ties -> A,C,G,T order) */
static int extract_consensus(const orc_bwt *b, uint64_t first, uint64_t second, int c, int len, lctx_t *out) {
    char rev[MAX_CTX];
    int k = 0;
    rev[k++] = "ACGT"[c];
    uint64_t lo, hi;
    lf_range_c(b, first, second, c, &lo, &hi);
    int freq = (int)(hi - lo);
    for (int rem = len - 1; rem > 0; --rem) {
        uint64_t l4[4], h4[4];
        lf_range(b, lo, hi, l4, h4);
        int best = 0;
        for (int x = 1; x < 4; ++x) if (h4[x] - l4[x] > h4[best] - l4[best]) best = x;
        if (h4[best] - l4[best] == 0) break;
        rev[k++] = "ACGT"[best];
        lo = l4[best];
        hi = h4[best];
    }
    if (k != len) return 0;                                           /* :317 */
    for (int i = 0; i < k; ++i) out->ctx[i] = rev[k - 1 - i];
    out->ctx[k] = 0;
    out->len = k;
    out->support = freq;
    return 1;
}

/* extract_dna (:325-342) */
static int extract_dna(const orc_bwt *b, uint64_t i, int len, char *out) {
    int k = 0;
    uint8_t c = F_col(b, i);
    while (c != b->term && len > 0) {
        out[k++] = (char)c;
        i = orc_FL(b, i);
        c = F_col(b, i);
        len--;
    }
    out[k] = 0;
    return k;
}

static int base_to_int(uint8_t c) {                                   /* include.hpp:275-289 (TERM -> 0) */
    switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; }
    return 0;
}

typedef struct {
    lctx_t l0[4], l1[4];
    int n0, n1;
    char right[MAX_CTX];
    int right_len;
    int has_right;
} cluster_t;

/* mode 1: find_variants(bwt, range) (:941-1005) + to_file(vector<variant_single_t>) (:1254-1330) */
static void call_single(const orc_bwt *b, const uint64_t *thr, uint64_t first, uint64_t second,
                        const orc_params *p, buf_t *o, orc_stats *st, uint64_t *cluster_nr, cluster_t *cl) {
    unsigned counts[4] = {0, 0, 0, 0};
    for (uint64_t i = first; i < second; ++i) counts[base_to_int(orc_access(b, i))]++;
    int freq[4], nf = 0;
    for (int c = 0; c < 4; ++c) if (counts[c] >= (unsigned)p->mcov_out) freq[nf++] = c;
    if (nf < 2 || (p->max_variants_per_position > 0 && nf > p->max_variants_per_position)) return;
    cl->n0 = 0;
    for (int k = 0; k < nf; ++k) cl->n0 += extract_consensus(b, first, second, freq[k], p->k_left, &cl->l0[cl->n0]);
    uint64_t i = first;
    while (i < second && !bit_get(thr, 2 * i + 1)) ++i;
    if (!(i < second)) return;
    cl->right_len = extract_dna(b, i, p->k_right, cl->right);
    int nv = cl->n0;
    /* to_file */
    if (nv < 2) return;
    int max_dist = 0, good[4], ng = 0;
    for (int k = 0; k + 1 < nv; ++k) {
        int32_t d[2];
        orc_distance(cl->l0[k].ctx, cl->l0[k + 1].ctx, p->k_left, p->max_gap, d);
        if (d[0] > max_dist) max_dist = d[0];
        if (cl->l0[k].support >= p->mcov_out) good[ng++] = k;
    }
    if (cl->l0[nv - 1].support >= p->mcov_out) good[ng++] = nv - 1;
    if (max_dist <= p->max_snvs && ng >= 2) {
        uint64_t id_nr = 1;
        for (int g = 0; g < ng; ++g) {
            lctx_t *v = &cl->l0[good[g]];
            if (has_run(cl->right, cl->right_len, p->complexity)) continue;
            buf_str(o, ">cluster:"); buf_u64(o, *cluster_nr);
            buf_str(o, "_id:"); buf_u64(o, id_nr++);
            buf_str(o, "_right:"); buf_int(o, cl->right_len);
            buf_str(o, "_cov:"); buf_int(o, v->support);
            buf_str(o, "_");
            const char *a = g == 0 ? v->ctx : cl->l0[good[g - 1]].ctx;   /* :1299-1307 quirk */
            const char *bb = cl->l0[good[1]].ctx;
            int32_t d[2];
            orc_distance(a, bb, p->k_left, p->max_gap, d);
            event_type(o, a, bb, p->k_left, d);
            buf_str(o, "\n");
            buf_add(o, v->ctx, (size_t)v->len);
            buf_add(o, cl->right, (size_t)cl->right_len);
            buf_str(o, "\n");
            st->events++;
        }
    }
    ++*cluster_nr;
}

/* to_file(vector<variant_t>) (:1149-1252) on the cross product built at :915-928 / :1077-1090 */
static void emit_pairs(const cluster_t *cl, const orc_params *p, buf_t *o, uint64_t *cluster_nr) {
    int found = 0;
    uint64_t id_nr = 1;
    for (int a = 0; a < cl->n0; ++a) for (int b = 0; b < cl->n1; ++b) {
        const lctx_t *L0 = &cl->l0[a], *L1 = &cl->l1[b];
        if (L0->len == 0 || L1->len == 0) continue;
        if (L0->ctx[L0->len - 1] == L1->ctx[L1->len - 1]) continue;
        int32_t d[2];
        orc_distance(L0->ctx, L1->ctx, p->k_left, p->max_gap, d);
        if (has_run(cl->right, cl->right_len, p->complexity) || d[0] > p->max_snvs ||
            L0->support < p->mcov_out || L1->support < p->mcov_out) continue;
        found = 1;
        for (int side = 0; side < 2; ++side) {
            const lctx_t *L = side ? L1 : L0;
            buf_str(o, ">cluster:"); buf_u64(o, *cluster_nr);
            buf_str(o, "_id:"); buf_u64(o, id_nr);
            buf_str(o, "_right:"); buf_int(o, cl->right_len);
            buf_str(o, "_cov:"); buf_int(o, L->support);
            buf_str(o, "_");
            event_type(o, L0->ctx, L1->ctx, p->k_left, d);
            buf_str(o, "\n");
            int skip = 0;
            if (side == 0 && d[1] < 0) skip = -d[1];                    /* :1199 */
            if (side == 1 && d[1] > 0) skip = d[1];                     /* :1233 */
            buf_add(o, L->ctx + skip, (size_t)(L->len - skip));
            buf_add(o, cl->right, (size_t)cl->right_len);
            buf_str(o, "\n");
        }
        id_nr++;
    }
    *cluster_nr += (uint64_t)found;
}

/* mode 3: find_variants(bwt, DA, range) (:1013-1096) */
static void call_da(const orc_bwt *b, const uint64_t *da, const uint64_t *thr, uint64_t first, uint64_t second,
                    const orc_params *p, buf_t *o, uint64_t *cluster_nr, cluster_t *cl) {
    unsigned counts[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
    for (uint64_t i = first; i < second; ++i) counts[bit_get(da, i)][base_to_int(orc_access(b, i))]++;
    int f0[4], f1[4], n0 = 0, n1 = 0;
    for (int c = 0; c < 4; ++c) {
        if (counts[0][c] >= (unsigned)p->mcov_out) f0[n0++] = c;
        if (counts[1][c] >= (unsigned)p->mcov_out) f1[n1++] = c;
    }
    int q = p->max_variants_per_position;
    if (n0 == 0 || n1 == 0 || (q > 0 && n0 > q) || (q > 0 && n1 > q)) return;
    cl->n0 = cl->n1 = 0;
    for (int k = 0; k < n0; ++k) cl->n0 += extract_consensus(b, first, second, f0[k], p->k_left, &cl->l0[cl->n0]);
    for (int k = 0; k < n1; ++k) cl->n1 += extract_consensus(b, first, second, f1[k], p->k_left, &cl->l1[cl->n1]);
    uint64_t i = first;
    while (i < second && !bit_get(thr, 2 * i + 1)) ++i;
    if (!(i < second)) return;
    cl->right_len = extract_dna(b, i, p->k_right, cl->right);
    emit_pairs(cl, p, o, cluster_nr);
}

/* mode 2: find_variants(bwt1, bwt2, range1, range2) (:840-934) */
static void call_two(const orc_bwt *b1, const orc_bwt *b2, const uint64_t *da, const uint64_t *thr,
                     uint64_t f1, uint64_t s1, uint64_t f2, uint64_t s2,
                     const orc_params *p, buf_t *o, uint64_t *cluster_nr, cluster_t *cl) {
    unsigned counts[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
    for (uint64_t i = f1; i < s1; ++i) counts[0][base_to_int(orc_access(b1, i))]++;
    for (uint64_t i = f2; i < s2; ++i) counts[1][base_to_int(orc_access(b2, i))]++;
    int fr0[4], fr1[4], n0 = 0, n1 = 0;
    for (int c = 0; c < 4; ++c) {
        if (counts[0][c] >= (unsigned)p->mcov_out) fr0[n0++] = c;
        if (counts[1][c] >= (unsigned)p->mcov_out) fr1[n1++] = c;
    }
    int q = p->max_variants_per_position;
    if (n0 == 0 || n1 == 0 || (q > 0 && n0 > q) || (q > 0 && n1 > q)) return;
    cl->n0 = cl->n1 = 0;
    for (int k = 0; k < n0; ++k) cl->n0 += extract_consensus(b1, f1, s1, fr0[k], p->k_left, &cl->l0[cl->n0]);
    for (int k = 0; k < n1; ++k) cl->n1 += extract_consensus(b2, f2, s2, fr1[k], p->k_left, &cl->l1[cl->n1]);
    uint64_t i0 = f1, i1 = f2, i = i0 + i1, end = s1 + s2;
    while (i < end && !bit_get(thr, 2 * i + 1)) {                      /* :901-906 */
        if (bit_get(da, i)) i1++; else i0++;
        ++i;
    }
    if (!(i < end)) return;
    if (bit_get(da, i)) cl->right_len = extract_dna(b2, i1, p->k_right, cl->right);
    else cl->right_len = extract_dna(b1, i0, p->k_right, cl->right);
    emit_pairs(cl, p, o, cluster_nr);
}

/* cluster scan: run_one_dataset (:1609-1655), run_two_datasets (:1395-1445), run_two_datasets_da (:1510-1560) */
int orc_call(const orc_bwt *b1, const orc_bwt *b2, const uint64_t *da, const uint64_t *thr,
             const uint64_t *minima, const orc_params *p, char **snp, size_t *snp_len, orc_stats *st) {
    if (p->k_left >= MAX_CTX || p->k_right >= MAX_CTX || p->k_left < 1) return 2;   /* -g > -L is accepted like the reference (orc_distance) */
    orc_bwt *m1 = (orc_bwt *)b1, *m2 = (orc_bwt *)b2;
    m1->rank_calls = &st->rank_call;
    if (m2) m2->rank_calls = &st->rank_call;
    uint64_t n = b1->n + (b2 ? b2->n : 0);
    buf_t o = {NULL, 0, 0};
    buf_add(&o, "", 0);
    cluster_t *cl = (cluster_t *)malloc(sizeof(cluster_t));
    uint64_t cluster_nr = 1, begin = 0, begin0 = 0, begin1 = 0, i0 = 0, i1 = 0, clust_len = 0;
    int open = 0;
    for (uint64_t i = 0; i < n; ++i) {
        if (bit_get(thr, 2 * i) && !bit_get(minima, i)) {
            if (open) clust_len++;
            else { open = 1; clust_len = 1; begin = i; begin0 = i0; begin1 = i1; }
        } else {
            if (open) {
                st->clust_size += clust_len;
                if (clust_len <= 200) st->clust_sizes[clust_len] += clust_len;
                if (clust_len >= 2 * (uint64_t)p->mcov_out) {
                    st->n_clusters++;
                    if (b2) call_two(b1, b2, da, thr, begin0, i0, begin1, i1, p, &o, &cluster_nr, cl);
                    else if (da) call_da(b1, da, thr, begin, i, p, &o, &cluster_nr, cl);
                    else call_single(b1, thr, begin, i, p, &o, st, &cluster_nr, cl);
                }
            }
            open = 0;
            clust_len = 0;
        }
        if (b2) { if (bit_get(da, i)) i1++; else i0++; }
    }
    free(cl);
    st->clusters_out = cluster_nr - 1;
    *snp = o.s;
    *snp_len = o.len;
    m1->rank_calls = &g_rank_sink;
    if (m2) m2->rank_calls = &g_rank_sink;
    return 0;
}

void orc_params_default(orc_params *p) {                               /* ebwt2InDel.cpp:20-74 */
    p->k_left = 31; p->k_right = 30; p->K = 16; p->max_gap = 10; p->max_snvs = 2; p->mcov_out = 3;
    p->complexity = 20; p->max_variants_per_position = 0; p->term = '#';
}

int orc_run(const uint8_t *ascii1, uint64_t n1, const uint8_t *ascii2, uint64_t n2,
            const uint8_t *da_ascii, const orc_params *p, char **snp, size_t *snp_len, orc_stats *st) {
    memset(st, 0, sizeof *st);
    orc_bwt *b1 = orc_bwt_build(ascii1, n1, (uint8_t)p->term, NULL), *b2 = NULL;
    if (!b1) return 1;
    if (ascii2) {
        b2 = orc_bwt_build(ascii2, n2, (uint8_t)p->term, NULL);
        if (!b2) { orc_bwt_free(b1); return 1; }
    }
    uint64_t n = n1 + (b2 ? n2 : 0);
    uint64_t *thr = (uint64_t *)malloc(((2 * n + 63) / 64 + 1) * 8);
    uint64_t *mn = (uint64_t *)malloc(((n + 63) / 64 + 1) * 8);
    uint64_t *da = (uint64_t *)calloc((n + 63) / 64 + 1, 8);
    int rc;
    if (b2) rc = orc_navigate_two(b1, b2, p, thr, mn, da, st);
    else {
        rc = orc_navigate_one(b1, p, thr, mn, st);
        if (da_ascii) for (uint64_t i = 0; i < n; ++i) bit_set(da, i, da_ascii[i] == '1');  /* :1503-1508 */
    }
    if (!rc) rc = orc_call(b1, b2, (b2 || da_ascii) ? da : NULL, thr, mn, p, snp, snp_len, st);
    free(thr); free(mn); free(da);
    orc_bwt_free(b1);
    orc_bwt_free(b2);
    return rc;
}

void orc_free(void *ptr) { free(ptr); }
