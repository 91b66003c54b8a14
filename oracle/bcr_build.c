/*
 * bcr_build.c -- CPU eBWT builder for LARGE parity inputs.  TEST INFRASTRUCTURE ONLY (oracle/).
 *
 * The reference reads an eBWT produced by an external tool (BCR_LCP_GSA / egap, README.md:38,
 * 91-92 of the reference) that is not vendored.  The parity corpus therefore needs its own
 * builder; tests/golden/make_big_golden.py uses this one to make the multi-gigasymbol inputs the
 * compiled reference (oracle/_ref/ebwt2InDel) is run on.  It is independent of the GPU builder
 * of the bench tooling (ebwt2indel_b200/synth.py, csrc/tools.cu): the GPU tests compare a checksum
 * of the GPU-built eBWT with the one recorded here before comparing outputs.
 *
 * Algorithm: BCR-style column insertion (Bauer, Cox, Rosone 2013).  After iteration k the array
 * holds the symbols preceding all read suffixes of length <= k in suffix order ('#'_i < '#'_j for
 * i < j, '#' < A < C < G < T).  Reads are kept sorted by the position P of their newest suffix; the
 * LF step of all reads is ONE sequential counting pass over the array, the insertion ONE merge.
 * Both passes are split over threads (pthreads fork-join) at read-index boundaries.
 *
 * Reads are given implicitly (nothing is materialised): read r = hap[start[r] .. start[r]+L) or, when
 * rc[r] != 0, the reverse complement of that window.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

typedef struct {
    /* inputs */
    const uint8_t *hap;
    const int64_t *start;
    const uint8_t *rc;
    uint64_t m, second_from;
    int L, T, k;
    uint8_t term;
    uint8_t comp[256], code[256];
    /* state */
    uint8_t *cur, *nxt, *ocur, *onxt;       /* symbol arrays (double buffer), owner arrays (optional) */
    uint32_t *order, *order2;               /* read ids sorted by P */
    uint64_t *P, *rank_local, *newP;
    uint8_t *csym;
    uint64_t S;
    uint64_t (*tcnt)[4], (*toff)[4], (*dst)[4];
    uint64_t base[4];
} Bcr;

typedef struct { Bcr *b; int phase, t; } Job;

static inline uint8_t read_sym(const Bcr *b, uint32_t r, int j) {
    return b->rc[r] ? b->comp[b->hap[b->start[r] + (b->L - 1 - j)]] : b->hap[b->start[r] + j];
}

static void phase_init(Bcr *b, int t) {
    const uint64_t i0 = b->m * (uint64_t)t / b->T, i1 = b->m * (uint64_t)(t + 1) / b->T;
    for (uint64_t i = i0; i < i1; ++i) {
        b->cur[i] = read_sym(b, (uint32_t)i, b->L - 1);
        if (b->ocur) b->ocur[i] = i >= b->second_from;
        b->order[i] = (uint32_t)i;
        b->P[i] = i;
    }
}

/* LF of every read's newest suffix: rank of its symbol at P.  Thread t counts the array range
 * [P[i0], P[i1]) (thread 0 from 0, the last one up to S); the ranks are local to that range. */
static void phase_count(Bcr *b, int t) {
    const uint64_t i0 = b->m * (uint64_t)t / b->T, i1 = b->m * (uint64_t)(t + 1) / b->T;
    const uint8_t *cur = b->cur, *code = b->code;
    uint64_t pos = t == 0 ? 0 : b->P[i0];
    const uint64_t pend = t == b->T - 1 ? b->S : b->P[i1];
    uint64_t cnt[4] = {0, 0, 0, 0}, h[4] = {0, 0, 0, 0};
    for (uint64_t i = i0; i < i1; ++i) {
        const uint64_t p = b->P[i];
        for (; pos < p; ++pos) cnt[code[cur[pos]]]++;
        const uint8_t ch = cur[p];
        b->csym[i] = ch;
        b->rank_local[i] = cnt[code[ch]];
        h[code[ch]]++;
    }
    for (; pos < pend; ++pos) cnt[code[cur[pos]]]++;
    for (int c = 0; c < 4; ++c) { b->tcnt[t][c] = cnt[c]; b->dst[t][c] = h[c]; }
}

/* stable 4-way partition of the reads by symbol: the new positions come out sorted */
static void phase_partition(Bcr *b, int t) {
    const uint64_t i0 = b->m * (uint64_t)t / b->T, i1 = b->m * (uint64_t)(t + 1) / b->T;
    uint64_t d[4] = {b->dst[t][0], b->dst[t][1], b->dst[t][2], b->dst[t][3]};
    for (uint64_t i = i0; i < i1; ++i) {
        const int c = b->code[b->csym[i]];
        const uint64_t j = d[c]++;
        b->order2[j] = b->order[i];
        b->newP[j] = b->base[c] + b->toff[t][c] + b->rank_local[i];
    }
}

/* merge: old symbols keep their order, the m new ones go to newP (ascending) */
static void phase_merge(Bcr *b, int t) {
    const uint64_t j0 = b->m * (uint64_t)t / b->T, j1 = b->m * (uint64_t)(t + 1) / b->T;
    const uint64_t *newP = b->newP;
    uint64_t opos = t == 0 ? 0 : newP[j0];
    uint64_t ipos = opos - (t == 0 ? 0 : j0);
    const int last_col = b->k + 1 >= b->L;
    for (uint64_t j = j0; j < j1; ++j) {
        const uint64_t np = newP[j], seg = np - opos;
        if (seg) {
            memcpy(b->nxt + opos, b->cur + ipos, seg);
            if (b->onxt) memcpy(b->onxt + opos, b->ocur + ipos, seg);
            ipos += seg;
        }
        const uint32_t r = b->order2[j];
        b->nxt[np] = last_col ? b->term : read_sym(b, r, b->L - 2 - b->k);
        if (b->onxt) b->onxt[np] = r >= b->second_from;
        opos = np + 1;
    }
    const uint64_t oend = t == b->T - 1 ? b->S + b->m : newP[j1];
    if (oend > opos) {
        memcpy(b->nxt + opos, b->cur + ipos, oend - opos);
        if (b->onxt) memcpy(b->onxt + opos, b->ocur + ipos, oend - opos);
    }
}

static void *worker(void *arg) {
    Job *j = (Job *)arg;
    switch (j->phase) {
    case 0: phase_init(j->b, j->t); break;
    case 1: phase_count(j->b, j->t); break;
    case 2: phase_partition(j->b, j->t); break;
    default: phase_merge(j->b, j->t); break;
    }
    return NULL;
}

static void run_phase(Bcr *b, int phase) {
    pthread_t th[256];
    Job jobs[256];
    for (int t = 0; t < b->T; ++t) {
        jobs[t].b = b; jobs[t].phase = phase; jobs[t].t = t;
        if (t + 1 < b->T) pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    worker(&jobs[b->T - 1]);
    for (int t = 0; t + 1 < b->T; ++t) pthread_join(th[t], NULL);
}

/* Builds the eBWT of m reads of length L into out_bwt[m*(L+1)] (caller-allocated).  When out_owner
 * is not NULL it receives, per position, 1 if the suffix belongs to a read with index >= second_from.
 * Allocates one more m*(L+1)-byte working array (two with owners).  threads <= 0: all online CPUs. */
int orc_bcr_build(const uint8_t *hap, const int64_t *start, const uint8_t *rc, uint64_t m, int L, uint8_t term,
                  uint64_t second_from, uint8_t *out_bwt, uint8_t *out_owner, int threads, int verbose) {
    if (m == 0 || L < 1 || m >= 0xffffffffull) return 1;
    Bcr b;
    memset(&b, 0, sizeof b);
    b.hap = hap; b.start = start; b.rc = rc; b.m = m; b.second_from = second_from; b.L = L; b.term = term;
    memset(b.comp, 'N', sizeof b.comp);
    b.comp['A'] = 'T'; b.comp['C'] = 'G'; b.comp['G'] = 'C'; b.comp['T'] = 'A';
    memset(b.code, 3, sizeof b.code);
    b.code['A'] = 0; b.code['C'] = 1; b.code['G'] = 2; b.code['T'] = 3;
    long ncpu = threads > 0 ? threads : sysconf(_SC_NPROCESSORS_ONLN);
    if (ncpu < 1) ncpu = 1;
    if (ncpu > 256) ncpu = 256;
    b.T = (uint64_t)ncpu > m ? (int)m : (int)ncpu;
    const uint64_t total = m * (uint64_t)(L + 1);
    uint8_t *bufB = (uint8_t *)malloc(total), *ownB = out_owner ? (uint8_t *)malloc(total) : NULL;
    b.cur = out_bwt; b.nxt = bufB; b.ocur = out_owner; b.onxt = ownB;
    b.order = (uint32_t *)malloc(m * 4); b.order2 = (uint32_t *)malloc(m * 4);
    b.P = (uint64_t *)malloc(m * 8); b.rank_local = (uint64_t *)malloc(m * 8); b.newP = (uint64_t *)malloc(m * 8);
    b.csym = (uint8_t *)malloc(m);
    b.tcnt = (uint64_t(*)[4])calloc((size_t)b.T, sizeof(uint64_t[4]));
    b.toff = (uint64_t(*)[4])calloc((size_t)b.T, sizeof(uint64_t[4]));
    b.dst = (uint64_t(*)[4])calloc((size_t)b.T, sizeof(uint64_t[4]));
    if (!bufB || (out_owner && !ownB) || !b.order || !b.order2 || !b.P || !b.rank_local || !b.newP || !b.csym) return 2;

    run_phase(&b, 0);                     /* the m suffixes '#'_i, preceded by the last symbol of read i */
    b.S = m;
    for (int k = 0; k < L; ++k) {
        b.k = k;
        run_phase(&b, 1);
        uint64_t tot[4] = {0, 0, 0, 0}, run[4] = {0, 0, 0, 0};
        for (int t = 0; t < b.T; ++t)
            for (int c = 0; c < 4; ++c) { b.toff[t][c] = tot[c]; tot[c] += b.tcnt[t][c]; }
        b.base[0] = m;                    /* the '#' suffixes come first, then A, C, G, T */
        for (int c = 1; c < 4; ++c) b.base[c] = b.base[c - 1] + tot[c - 1];
        uint64_t sym_tot[4] = {0, 0, 0, 0};
        for (int t = 0; t < b.T; ++t) for (int c = 0; c < 4; ++c) sym_tot[c] += b.dst[t][c];
        for (int c = 1; c < 4; ++c) run[c] = run[c - 1] + sym_tot[c - 1];
        for (int t = 0; t < b.T; ++t)
            for (int c = 0; c < 4; ++c) { const uint64_t h = b.dst[t][c]; b.dst[t][c] = run[c]; run[c] += h; }
        run_phase(&b, 2);
        run_phase(&b, 3);
        b.S += m;
        { uint8_t *x = b.cur; b.cur = b.nxt; b.nxt = x; x = b.ocur; b.ocur = b.onxt; b.onxt = x; }
        { uint32_t *x = b.order; b.order = b.order2; b.order2 = x; }
        { uint64_t *x = b.P; b.P = b.newP; b.newP = x; }
        if (verbose && (k % 16 == 15 || k + 1 == L))
            fprintf(stderr, "[bcr_build] column %d/%d, %llu symbols\n", k + 1, L, (unsigned long long)b.S);
    }
    if (b.cur != out_bwt) {
        memcpy(out_bwt, b.cur, total);
        if (out_owner) memcpy(out_owner, b.ocur, total);
    }
    free(bufB); free(ownB);
    free(b.order); free(b.order2); free(b.P); free(b.rank_local); free(b.newP); free(b.csym);
    free(b.tcnt); free(b.toff); free(b.dst);
    return 0;
}

/* position-weighted checksum of a byte string, the same formula tests/bigcase.py evaluates on the
 * GPU with torch: sum_i b[i] * ((i mod 2^20) + 1)  +  (sum_i b[i]) * 2^40   (mod 2^64) */
uint64_t orc_checksum(const uint8_t *b, uint64_t n) {
    uint64_t a = 0, s = 0;
    for (uint64_t i = 0; i < n; ++i) {
        a += (uint64_t)b[i] * ((i & 0xfffffull) + 1);
        s += b[i];
    }
    return a + (s << 40);
}
