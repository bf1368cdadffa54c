/*
 * oracle/quicked_oracle.c — CPU restatement of QuickEd's bound-and-align hot path (plain C, scalar).
 *
 * TEST INFRASTRUCTURE ONLY — see quicked_oracle.h.  Parity status: PINNED against the reference's known
 * answers and against oracle/_ref/libquicked_ref.so (tests/test_oracle_vs_reference.py, tests/golden/).
 *
 * This is a from-scratch restatement: one column-major loop per kernel, no SIMD skewing (the reference's
 * 2/4/8-column skews are pure re-orderings of the same block updates), explicit structs
 * for band / window state.  Every function cites the reference lines whose behaviour it reproduces.
 * Citations are file:line under /root/reference.
 */
#include "quicked_oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define W64 64
#define ALPHA 5
#define CEILDIV(a, b) (((a) + ((b) - 1)) / (b))
#define MAXI(a, b) ((a) >= (b) ? (a) : (b))
#define MINI(a, b) ((a) <= (b) ? (a) : (b))
#define ABSI(a) ((a) >= 0 ? (a) : -(a))

static __thread uint64_t g_ws_windowed, g_ws_banded;     /* per thread: qo_batch_align runs qo_align concurrently */
static __thread int g_splits;
static __thread int g_ref_undefined;   /* set when the reference would read uninitialised memory (see banded_score_run) */

uint64_t qo_word_steps_total(void) { return g_ws_windowed + g_ws_banded; }

/* ------------------------------------------------------------------------------------------------
 * Encoding: quicked_utils/src/dna_text.c:41-46 — A/a 0, C/c 1, G/g 2, T/t 3, everything else 4.
 * Bytes >= 0x80 index the reference table negatively (UB); we define them as 4.
 * ---------------------------------------------------------------------------------------------- */
static inline int enc(char ch)
{
    switch (ch) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
    }
}

/* ------------------------------------------------------------------------------------------------
 * One Myers block update: quicked/include/bpm_commons.h:49-68 (masked carry-out) and :82-101 (bit 63).
 * ---------------------------------------------------------------------------------------------- */
static inline void myers_block(uint64_t eq, uint64_t *pv_io, uint64_t *mv_io, unsigned hp_in, unsigned hm_in,
                               uint64_t out_mask, unsigned *hp_out, unsigned *hm_out)
{
    const uint64_t pv = *pv_io, mv = *mv_io;
    const uint64_t xv = eq | mv;
    const uint64_t eqh = eq | (uint64_t)hm_in;
    const uint64_t xh = (((eqh & pv) + pv) ^ pv) | eqh;
    uint64_t ph = mv | ~(xh | pv);
    uint64_t mh = pv & xh;
    *hp_out = (ph & out_mask) != 0;
    *hm_out = (mh & out_mask) != 0;
    ph = (ph << 1) | (uint64_t)hp_in;
    mh = (mh << 1) | (uint64_t)hm_in;
    *pv_io = mh | ~(xv | ph);
    *mv_io = ph & xv;
}

/* ------------------------------------------------------------------------------------------------
 * Pattern match masks: bpm_banded.c:40-103 == bpm_windowed.c:41-122.
 * peq[blk*5+code]; rows >= m of the last block match all 5 codes; lvl[blk] = bit 63, last block bit m%64-1.
 * Two extra all-zero blocks are appended: the reference reads (in-allocation) garbage there in score-only
 * mode and those values provably cannot reach any output.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int64_t m, nblk;
    uint64_t *peq, *lvl;
    const char *raw;
} pat_t;

static void pat_build(pat_t *p, const char *pattern, int64_t m)
{
    p->m = m;
    p->nblk = CEILDIV(m, W64);
    p->raw = pattern;
    p->peq = (uint64_t *)calloc((size_t)(p->nblk + 2) * ALPHA, sizeof(uint64_t));
    p->lvl = (uint64_t *)calloc((size_t)(p->nblk + 2), sizeof(uint64_t));
    for (int64_t i = 0; i < m; ++i)
        p->peq[(i / W64) * ALPHA + enc(pattern[i])] |= 1ull << (i % W64);
    for (int64_t i = m; i < p->nblk * W64; ++i)
        for (int c = 0; c < ALPHA; ++c) p->peq[(i / W64) * ALPHA + c] |= 1ull << (i % W64);
    for (int64_t b = 0; b < p->nblk + 2; ++b) p->lvl[b] = 1ull << 63;
    if (p->nblk > 0 && (m % W64)) p->lvl[p->nblk - 1] = 1ull << ((m % W64) - 1);
}
static void pat_free(pat_t *p) { free(p->peq); free(p->lvl); }

/* ------------------------------------------------------------------------------------------------
 * BandEd geometry: bpm_banded.c:121-135 (allocate), :359-361 / :801-803 (score-only height).
 * ---------------------------------------------------------------------------------------------- */
qo_band_geom_t qo_band_geometry(int64_t m, int64_t n, int64_t cutoff)
{
    qo_band_geom_t g;
    g.k = MAXI(MAXI(ABSI(n - m) + 1, cutoff), 65);
    g.d = m - n;
    g.rel = CEILDIV(g.k - ABSI(g.d), 2);
    if (g.d >= 0) {
        g.prolog = CEILDIV(g.rel, W64);
        g.B_cigar = CEILDIV(g.rel + g.d, W64) + 1 + g.prolog;
    } else {
        g.prolog = CEILDIV(g.rel - g.d, W64);
        g.B_cigar = CEILDIV(g.rel, W64) + 1 + g.prolog;
    }
    g.B_score = CEILDIV(g.k, W64) + 1;
    g.fin = g.prolog * W64 + g.d;
    return g;
}

/* Sliding band state.  Word i of the band covers pattern block (i + pos_v). */
typedef struct {
    qo_band_geom_t g;
    int64_t first, last, pos_v, pos_h;
    int64_t *scores;   /* per absolute pattern block */
    int64_t scores_init;   /* highest scores[] index written so far */
    int consumed;          /* the caller uses the returned score (stage 3 / BANDED only_score), not just the band state */
} band_t;

static void band_reset(band_t *b, int64_t nwords, uint64_t *pv, uint64_t *mv)
{   /* bpm_banded.c:180-197, :222-225 */
    b->pos_v = -b->g.prolog;
    b->pos_h = 0;
    b->first = b->g.prolog;
    b->last = nwords - 1;
    b->scores_init = nwords - 1;
    for (int64_t i = 0; i < nwords; ++i) { pv[i] = ~0ull; mv[i] = 0; b->scores[i] = W64 * (i + 1); }
}

/* One text column over the live words: bpm_banded.c:232-261 / :928-950. src/dst may alias. */
static void band_column(band_t *b, const pat_t *p, int code, const uint64_t *pv_src, const uint64_t *mv_src,
                        uint64_t *pv_dst, uint64_t *mv_dst)
{
    unsigned hp = 1, hm = 0;
    for (int64_t i = b->first; i <= b->last; ++i) {
        const int64_t blk = i + b->pos_v;
        uint64_t pv = pv_src[i], mv = mv_src[i];
        unsigned hpo, hmo;
        /* blocks past the pattern: the reference reads whatever follows its PEQ table there (in-allocation garbage,
         * SURVEY App. B.4).  One block past is harmless (never read back); further out the garbage reaches the cut
         * tests, so the reference is undefined.  We define every such block as "no match, carry at bit 63". */
        const int past = blk >= p->nblk + 2;
        if (blk > p->nblk) g_ref_undefined = 1;
        myers_block(past ? 0 : p->peq[blk * ALPHA + code], &pv, &mv, hp, hm, past ? (1ull << 63) : p->lvl[blk], &hpo, &hmo);
        pv_dst[i] = pv; mv_dst[i] = mv;
        hp = hpo; hm = hmo;
        b->scores[blk] += (int64_t)hpo - (int64_t)hmo;
        ++g_ws_banded;
    }
}

/* End of a 64-column block: bpm_banded.c:264-301 (CIGAR mode, clamp = nblk-1) and :889-922 (score-only,
 * clamp = nblk).  Order matters: lower cut / prolog widen, shift, new bottom word, upper cut, advance. */
static void band_shift(band_t *b, uint64_t *pv, uint64_t *mv, int64_t bottom_clamp)
{
    const int64_t k = b->g.k, fin = b->g.fin;
    const int cut_lo = (b->first + 2 < b->last) && (fin > W64 * (b->first + 1)) &&
                       (b->scores[b->first + b->pos_v + 1] + (fin - W64 * (b->first + 1)) > k);
    if (cut_lo && b->pos_h >= b->g.prolog) b->first++;
    else if (!cut_lo && b->pos_h < b->g.prolog) b->first--;
    for (int64_t j = b->first; j < b->last; ++j) { pv[j] = pv[j + 1]; mv[j] = mv[j + 1]; }
    pv[b->last] = ~0ull; mv[b->last] = 0;
    b->scores[b->last + b->pos_v + 1] = b->scores[b->last + b->pos_v] + W64;
    if (b->last + b->pos_v + 1 > b->scores_init) b->scores_init = b->last + b->pos_v + 1;
    const int cut_hi = (b->first + 2 < b->last) && (W64 * (b->last - 1) > fin) &&
                       (b->scores[b->last + b->pos_v - 1] + (W64 * (b->last - 1) - fin) > k);
    if (cut_hi || (b->pos_v + b->last >= bottom_clamp)) b->last--;
    b->pos_v++; b->pos_h++;
}

static int64_t band_final_score(const band_t *b, const pat_t *p)
{   /* bpm_banded.c:303-312 / :952-961.
     * If the band was cut away before it reached the last pattern block, the reference reads a scores[]
     * entry it never wrote (arena memory, not cleared): undefined there.  We define the entry as 0 (what a
     * fresh arena holds) and flag the case so tests do not compare against the reference on it. */
    if (b->consumed && (p->m - 1) / W64 > b->scores_init) g_ref_undefined = 1;
    if (p->m % W64) return b->scores[p->m / W64] - (W64 - (p->m % W64));
    return b->scores[(p->m - 1) / W64];
}

/* BandEd score-only: bpm_banded.c:791-964 (scalar) == :349-788 (AVX2), re-ordered column-major. */
static int64_t banded_score_run(const pat_t *p, const char *text, int64_t n, int64_t cutoff, int64_t finish,
                                band_t *b, uint64_t *pv, uint64_t *mv)
{
    b->g = qo_band_geometry(p->m, n, cutoff);
    b->consumed = (finish == n);
    band_reset(b, b->g.B_score, pv, mv);
    const int64_t full_blocks = finish / W64;
    int64_t col = 0;
    for (int64_t kb = 0; kb < full_blocks; ++kb) {
        for (; col < (kb + 1) * W64; ++col) band_column(b, p, enc(text[col]), pv, mv, pv, mv);
        band_shift(b, pv, mv, p->nblk);
    }
    for (; col < finish; ++col) band_column(b, p, enc(text[col]), pv, mv, pv, mv);
    return band_final_score(b, p);
}

int64_t qo_banded_score(const char *pattern, int m, const char *text, int n, int64_t cutoff, int64_t finish,
                        uint64_t *pv_out, uint64_t *mv_out, int64_t *scores_out, int64_t *lower_block,
                        int64_t *higher_block)
{
    pat_t p; pat_build(&p, pattern, m);
    const qo_band_geom_t g = qo_band_geometry(m, n, cutoff);
    const int64_t nsc = p.nblk + g.B_score + 2;
    uint64_t *pv = (uint64_t *)calloc((size_t)g.B_score + 1, 8), *mv = (uint64_t *)calloc((size_t)g.B_score + 1, 8);
    band_t b; b.scores = (int64_t *)calloc((size_t)nsc, 8);
    const int64_t s = banded_score_run(&p, text, n, cutoff, finish, &b, pv, mv);
    if (pv_out) memcpy(pv_out, pv, (size_t)g.B_score * 8);
    if (mv_out) memcpy(mv_out, mv, (size_t)g.B_score * 8);
    if (scores_out) memcpy(scores_out, b.scores, (size_t)nsc * 8);
    if (lower_block) *lower_block = b.first;
    if (higher_block) *higher_block = b.last;
    free(pv); free(mv); free(b.scores); pat_free(&p);
    return s;
}

/* ------------------------------------------------------------------------------------------------
 * Op buffer filled right-to-left (cigar_t begin/end offsets, quicked_utils/include/cigar.h:33-47).
 * ---------------------------------------------------------------------------------------------- */
typedef struct { char *ops; int64_t begin, end; } ops_t;
static inline void ops_push_front(ops_t *o, char c) { o->ops[--o->begin] = c; }

static __thread uint64_t *g_dump_pv, *g_dump_mv;   /* test hooks, see qo_banded_full_dump */
static __thread int64_t *g_dump_ranges;

/* BandEd full matrix + traceback (a leaf): bpm_banded.c:199-316 fill, :967-1036 walk.
 * The walk's ops are pushed in front of whatever `out` already holds (cigar.c:179-188). */
static int64_t banded_full(const pat_t *p, const char *text, int64_t n, int64_t cutoff, ops_t *out)
{
    band_t b;
    b.consumed = 0;
    b.g = qo_band_geometry(p->m, n, cutoff);
    const int64_t B = b.g.B_cigar, prolog = b.g.prolog;
    uint64_t *PV = (uint64_t *)malloc((size_t)B * (size_t)(n + 1) * 8);
    uint64_t *MV = (uint64_t *)malloc((size_t)B * (size_t)(n + 1) * 8);
    /* the reference leaves words outside [first,last] uninitialised; zero them so runs are deterministic */
    memset(PV, 0, (size_t)B * (size_t)(n + 1) * 8);
    memset(MV, 0, (size_t)B * (size_t)(n + 1) * 8);
    b.scores = (int64_t *)calloc((size_t)(p->nblk + B + 2), 8);
    band_reset(&b, B, PV, MV);
    if (g_dump_ranges) { g_dump_ranges[0] = b.first; g_dump_ranges[1] = b.last; }
    for (int64_t col = 0; col < n; ++col) {
        band_column(&b, p, enc(text[col]), PV + col * B, MV + col * B, PV + (col + 1) * B, MV + (col + 1) * B);
        if ((col + 1) % W64 == 0) {
            band_shift(&b, PV + (col + 1) * B, MV + (col + 1) * B, p->nblk - 1);
            if (g_dump_ranges) { g_dump_ranges[2 * b.pos_h] = b.first; g_dump_ranges[2 * b.pos_h + 1] = b.last; }
        }
    }
    const int64_t band_score = band_final_score(&b, p);
    if (g_dump_pv) {   /* test hook (qo_banded_full_dump): the stored matrix exactly as the walk below reads it */
        memcpy(g_dump_pv, PV, (size_t)B * (size_t)(n + 1) * 8);
        memcpy(g_dump_mv, MV, (size_t)B * (size_t)(n + 1) * 8);
    }
    /* walk: D if Pv[col h+1] bit v; else I if Mv[col h] bit v; else M/X on RAW bytes (:1002-1023) */
    int64_t h = n - 1, v = p->m - 1;
    while (v >= 0 && h >= 0) {
        const int64_t ev = v - W64 * (h / W64 - prolog);
        const int64_t ev_r = v - W64 * ((h + 1) / W64 - prolog);
        /* A too-narrow band can put v outside the band's coordinates: the reference then indexes the flat
         * [column][word] array with a C-truncated (possibly negative) word number and shifts by a count the CPU
         * masks to 6 bits (bpm_banded.c:995-1000).  Reproduce exactly that; outside the allocation read 0. */
        const int64_t fr = (h + 1) * B + ev_r / W64, fl = h * B + ev / W64;
        const int64_t cells = B * (n + 1);
        const uint64_t pvw = (fr >= 0 && fr < cells) ? PV[fr] : 0, mvw = (fl >= 0 && fl < cells) ? MV[fl] : 0;
        if (pvw & (1ull << (ev_r & 63))) { ops_push_front(out, 'D'); --v; }
        else if (mvw & (1ull << (ev & 63))) { ops_push_front(out, 'I'); --h; }
        else { ops_push_front(out, text[h] == p->raw[v] ? 'M' : 'X'); --h; --v; }
    }
    while (h >= 0) { ops_push_front(out, 'I'); --h; }
    while (v >= 0) { ops_push_front(out, 'D'); --v; }
    free(PV); free(MV); free(b.scores);
    return band_score;
}

/* Test hook: one BandEd leaf (full matrix + walk) with its stored matrix and live ranges exported.
 * pv/mv: B_cigar*(n+1) words each ([column][band word], never-written words 0); ranges: (first,last) per 64-column
 * block, n/64+1 pairs; ops: the walk's op string (m+n+1 bytes).  Returns the band score. */
int64_t qo_banded_full_dump(const char *pattern, int m, const char *text, int n, int64_t cutoff, uint64_t *pv,
                            uint64_t *mv, int64_t *ranges, char *ops)
{
    pat_t p; pat_build(&p, pattern, m);
    ops_t o; o.ops = (char *)malloc((size_t)m + (size_t)n + 1); o.begin = o.end = (int64_t)m + n;
    g_dump_pv = pv; g_dump_mv = mv; g_dump_ranges = ranges;
    const int64_t s = banded_full(&p, text, n, cutoff, &o);
    g_dump_pv = g_dump_mv = NULL; g_dump_ranges = NULL;
    if (ops) { memcpy(ops, o.ops + o.begin, (size_t)(o.end - o.begin)); ops[o.end - o.begin] = 0; }
    free(o.ops); pat_free(&p);
    return s;
}

/* ------------------------------------------------------------------------------------------------
 * Hirschberg: bpm_hirschberg.c:33-270.  text_r / pattern_r are the reversed sequences, sliced exactly
 * like the reference slices them (:183-191).
 * ---------------------------------------------------------------------------------------------- */
static int hirschberg(const char *text, const char *text_r, int64_t n, const char *pattern, const char *pattern_r,
                      int64_t m, int64_t cutoff, ops_t *out)
{
    const qo_band_geom_t g = qo_band_geometry(m, n, cutoff);
    const uint64_t footprint = (uint64_t)g.B_cigar * (uint64_t)n * 8u * 2u;     /* :63 */
    if (footprint <= (1ull << 24)) {                                              /* :65, leaf :242-268 */
        pat_t p; pat_build(&p, pattern, m);
        banded_full(&p, text, n, cutoff, out);
        pat_free(&p);
        return QO_OK;
    }
    ++g_splits;
    const int64_t n_l = (n + 1) / 2, n_r = n - n_l;                               /* :68-69 */
    pat_t pf, pr; pat_build(&pf, pattern, m); pat_build(&pr, pattern_r, m);
    const int64_t nsc = pf.nblk + g.B_score + 2;
    uint64_t *pv = (uint64_t *)calloc((size_t)g.B_score + 1, 8), *mv = (uint64_t *)calloc((size_t)g.B_score + 1, 8);
    uint64_t *pvr = (uint64_t *)calloc((size_t)g.B_score + 1, 8), *mvr = (uint64_t *)calloc((size_t)g.B_score + 1, 8);
    band_t bf, br;
    bf.scores = (int64_t *)calloc((size_t)nsc, 8); br.scores = (int64_t *)calloc((size_t)nsc, 8);
    banded_score_run(&pf, text, n, cutoff, n_l, &bf, pv, mv);                     /* :85-91 */
    banded_score_run(&pr, text_r, n, cutoff, n_r, &br, pvr, mvr);                 /* :94-100 */

    /* band origins on the middle column, clamped like :103-104 */
    const int64_t org = n_l < g.prolog * W64 ? 0 : n_l / W64 - g.prolog;
    const int64_t org_r = n_r < g.prolog * W64 ? 0 : n_r / W64 - g.prolog;
    const int64_t lo_f = bf.first * 64 + 63 + org * 64;                           /* :110 */
    const int64_t lo_r = (m - 1) - (br.last * 64 + 63 + org_r * 64);              /* :111 */
    const int64_t hi_f = bf.last * 64 + 63 + org * 64;                            /* :112 */
    const int64_t hi_r = (m - 1) - (br.first * 64 + 63 + org_r * 64);             /* :113 */
    int status = QO_OK;
    if (lo_f > hi_r || lo_r > hi_f) {                                             /* :116-122 */
        status = QO_FAIL_NON_CONVERGENCE;
    } else {
        int64_t cell0, start, top, top_r;
        if (lo_f > lo_r) { cell0 = bf.first * 64 + 63; start = lo_f; }            /* :125-134 */
        else { cell0 = lo_r - org * 64; start = lo_r; }
        if (hi_f < hi_r) { top = bf.last * 64 + 63; top_r = (m - 1) - hi_f - org_r * 64; }   /* :137-146 */
        else { top = hi_r - org * 64; top_r = br.first * 64 + 63; }
        const int64_t ncell = top - cell0 + 2;                                    /* :147 */
        int32_t *cs = (int32_t *)malloc((size_t)(ncell + 1) * 4), *csr = (int32_t *)malloc((size_t)(ncell + 1) * 4);
        cs[0] = 0; csr[0] = 0;
        for (int64_t i = 0; i < ncell; ++i) {                                     /* :152-167 */
            const int64_t c = cell0 + i, cr = top_r + i;
            cs[i + 1] = cs[i] + (int)((pv[c / 64] >> (c % 64)) & 1) - (int)((mv[c / 64] >> (c % 64)) & 1);
            csr[i + 1] = csr[i] + (int)((pvr[cr / 64] >> (cr % 64)) & 1) - (int)((mvr[cr / 64] >> (cr % 64)) & 1);
        }
        int64_t best = 0, best_score = (int64_t)csr[ncell - 1] + cs[0];           /* :170-180, strict '<' */
        for (int64_t i = 1; i < ncell; ++i) {
            const int64_t s = (int64_t)csr[ncell - 1 - i] + cs[i];
            if (s < best_score) { best = i; best_score = s; }
        }
        const int64_t m_l = start + best, m_r = m - m_l;                          /* :183-184 */
        const int64_t ref_l = CEILDIV(m_l, 64) - (ncell < best + 64);             /* :194-196 */
        const int64_t sp_l = ref_l * 64 - (cell0 + org * 64);
        const int64_t score_l = cs[best] - cs[sp_l] + bf.scores[ref_l - 1];
        const int64_t ref_r = CEILDIV(m_r, 64) - (best < 64);                     /* :198-200 */
        const int64_t sp_r = ref_r * 64 - (top_r + org_r * 64);
        const int64_t score_r = csr[ncell - 1 - best] - csr[sp_r] + br.scores[ref_r - 1];
        free(cs); free(csr);
        /* right half first, then left (:212-239): the op buffer fills back to front */
        status = hirschberg(text + n_l, text_r, n - n_l, pattern + m_l, pattern_r, m_r, score_r, out);
        if (status >= 0)
            status = hirschberg(text, text_r + (n - n_l), n_l, pattern, pattern_r + m_r, m_l, score_l, out);
    }
    free(pv); free(mv); free(pvr); free(mvr); free(bf.scores); free(br.scores);
    pat_free(&pf); pat_free(&pr);
    return status;
}

/* ------------------------------------------------------------------------------------------------
 * WindowEd: bpm_windowed.c.  A window is anchored at its bottom-right corner (pos_v, pos_h).
 * Stored matrix: column index c holds the state after c window columns, [c*W + word].
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int W;
    int64_t corner_v, corner_h;     /* pos_v, pos_h: current corner */
    int64_t score, hew;
    uint64_t *pv, *mv;              /* (64W+2) columns x W words */
    uint64_t *weq;                  /* window-aligned match masks [W*5] */
} win_t;

typedef struct { int64_t v0, h0, words, cols; } win_geom_t;

static win_geom_t window_geometry(const win_t *w)
{   /* :219-232 */
    win_geom_t g;
    g.v0 = MAXI(w->corner_v - (int64_t)W64 * w->W + 1, 0);
    g.h0 = MAXI(w->corner_h - (int64_t)W64 * w->W + 1, 0);
    g.words = (w->corner_v - g.v0) / W64 + 1;
    g.cols = w->corner_h - g.h0 + 1;
    return g;
}

static void window_prepare(win_t *w, const pat_t *p, const win_geom_t *g)
{   /* column 0 (:225-229) and the funnel-shifted match masks (:237-244) */
    for (int i = 0; i < w->W; ++i) { w->pv[i] = (g->h0 == 0) ? ~0ull : 0; w->mv[i] = 0; }
    const unsigned sh = (unsigned)(g->v0 % W64);
    const int64_t blk0 = g->v0 / W64;
    for (int64_t i = 0; i < g->words; ++i)
        for (int c = 0; c < ALPHA; ++c) {
            uint64_t e = p->peq[(blk0 + i) * ALPHA + c] >> sh;
            if (sh) e |= p->peq[(blk0 + i + 1) * ALPHA + c] << (W64 - sh);
            w->weq[i * ALPHA + c] = e;
        }
}

/* scalar fill: windowed_compute_window, bpm_windowed.c:202-280 */
static void window_fill_scalar(win_t *w, const pat_t *p, const char *text)
{
    const win_geom_t g = window_geometry(w);
    window_prepare(w, p, &g);
    const unsigned top_in = (g.v0 == 0);                                          /* :247-252 */
    for (int64_t c = 0; c < g.cols; ++c) {
        const int code = enc(text[g.h0 + c]);
        unsigned hp = top_in, hm = 0;
        for (int64_t i = 0; i < g.words; ++i) {
            uint64_t pv = w->pv[c * w->W + i], mv = w->mv[c * w->W + i];
            unsigned hpo, hmo;
            myers_block(w->weq[i * ALPHA + code], &pv, &mv, hp, hm, 1ull << 63, &hpo, &hmo);
            w->pv[(c + 1) * w->W + i] = pv; w->mv[(c + 1) * w->W + i] = mv;
            hp = hpo; hm = hmo;
            ++g_ws_windowed;
        }
    }
}

/* SSE4.1 fill for W == 2 (windowed_compute_window_sse, bpm_windowed.c:283-445), restated in scalar form.
 * Observable differences from the scalar fill:
 *   - word 0's top carry-in is (v0==0) on column 0 only, then 1 on column 1 and on even columns, 0 on odd
 *     columns >= 3 (:348, :393, :424);
 *   - one look-ahead column of word 0 is computed from text[corner_h+1] when cols is even (:361) — one byte
 *     past the text for the first window; we define that byte as code 4 ('\0' in every reference caller);
 *   - when cols is even the last column of word 1 is recomputed with the look-ahead column's carry (:428-444).
 * For a single-column window the reference uses an uninitialised carry (:429-430); we define it as (0,0). */
static void window_fill_sse2(win_t *w, const pat_t *p, const char *text, int64_t n)
{
    const win_geom_t g = window_geometry(w);
    window_prepare(w, p, &g);
    const int64_t steps_h = g.cols - 1;
    const int64_t ncol0 = (steps_h == 0) ? 1 : ((steps_h & 1) ? steps_h + 2 : steps_h + 1);
    unsigned *c_hp = (unsigned *)calloc((size_t)ncol0 + 1, sizeof(unsigned));
    unsigned *c_hm = (unsigned *)calloc((size_t)ncol0 + 1, sizeof(unsigned));
    for (int64_t c = 0; c < ncol0; ++c) {               /* word 0, including the look-ahead column */
        const int64_t ti = g.h0 + c;
        const int code = (ti < n) ? enc(text[ti]) : 4;
        unsigned hp;
        if (c == 0) hp = (g.v0 == 0);
        else if (c == 1) hp = 1;
        else hp = (c & 1) ? 0 : 1;
        uint64_t pv = w->pv[c * 2], mv = w->mv[c * 2];
        myers_block(w->weq[code], &pv, &mv, hp, 0, 1ull << 63, &c_hp[c], &c_hm[c]);
        w->pv[(c + 1) * 2] = pv; w->mv[(c + 1) * 2] = mv;
        if (c <= steps_h) ++g_ws_windowed;
    }
    if (g.words == 2) {
        for (int64_t c = 0; c <= steps_h; ++c) {        /* word 1 with the true carries */
            const int code = enc(text[g.h0 + c]);
            unsigned hp = c_hp[c], hm = c_hm[c], o1, o2;
            if (steps_h == 0) { hp = 0; hm = 0; }
            uint64_t pv = w->pv[c * 2 + 1], mv = w->mv[c * 2 + 1];
            myers_block(w->weq[ALPHA + code], &pv, &mv, hp, hm, 1ull << 63, &o1, &o2);
            w->pv[(c + 1) * 2 + 1] = pv; w->mv[(c + 1) * 2 + 1] = mv;
            ++g_ws_windowed;
        }
        if (steps_h & 1) {                              /* last column of word 1 redone with look-ahead carry */
            const int64_t c = steps_h;
            const int code = enc(text[g.h0 + c]);
            unsigned o1, o2;
            uint64_t pv = w->pv[c * 2 + 1], mv = w->mv[c * 2 + 1];
            myers_block(w->weq[ALPHA + code], &pv, &mv, c_hp[c + 1], c_hm[c + 1], 1ull << 63, &o1, &o2);
            w->pv[(c + 1) * 2 + 1] = pv; w->mv[(c + 1) * 2 + 1] = mv;
        }
    }
    free(c_hp); free(c_hm);
}

/* Walk back from the corner through the non-overlapping part of the window.
 * score_only: priority D, I, M, X (bpm_windowed.c:504-561); else M(raw) first, then D, I, X (:448-502). */
static void window_walk(win_t *w, const pat_t *p, const char *text, int O, int hew_threshold, int score_only,
                        ops_t *out)
{
    const int64_t v0 = MAXI(w->corner_v - (int64_t)W64 * w->W + 1, 0);
    const int64_t h0 = MAXI(w->corner_h - (int64_t)W64 * w->W + 1, 0);
    const int64_t v_stop = MAXI(w->corner_v - (int64_t)W64 * (w->W - O) + 1, 0);
    const int64_t h_stop = MAXI(w->corner_h - (int64_t)W64 * (w->W - O) + 1, 0);
    int64_t v = w->corner_v, h = w->corner_h, cost = 0;
    while (v >= v_stop && h >= h_stop) {
        const int64_t word = (v - v0) / W64;
        const uint64_t bit = 1ull << ((v - v0) & 63);           /* x86 masks the shift count (:474,:530) */
        const int del = (w->pv[(h - h0 + 1) * w->W + word] & bit) != 0;
        const int ins = (w->mv[(h - h0) * w->W + word] & bit) != 0;
        const int same = text[h] == p->raw[v];
        if (score_only) {
            if (del) { ++cost; --v; }
            else if (ins) { ++cost; --h; }
            else { cost += !same; --h; --v; }
        } else {
            if (same) { ops_push_front(out, 'M'); --h; --v; }
            else if (del) { ops_push_front(out, 'D'); --v; }
            else if (ins) { ops_push_front(out, 'I'); --h; }
            else { ops_push_front(out, 'X'); --h; --v; }
        }
    }
    if (score_only) {
        if (cost > (int64_t)((w->W - O) * W64 * hew_threshold / 100)) w->hew++;   /* :555-556, int arithmetic */
        w->score += cost;
    }
    w->corner_v = v; w->corner_h = h;
}

/* windowed_compute: bpm_windowed.c:563-628 */
static void windowed_run(const pat_t *p, const char *text, int64_t n, int W, int O, int hew_threshold, int sse,
                         int score_only, win_t *w, ops_t *out)
{
    w->W = W;
    w->corner_v = p->m - 1; w->corner_h = n - 1;               /* :148-149 */
    w->score = 0; w->hew = 0;
    w->pv = (uint64_t *)calloc((size_t)(W64 * W + 3) * (size_t)W, 8);
    w->mv = (uint64_t *)calloc((size_t)(W64 * W + 3) * (size_t)W, 8);
    w->weq = (uint64_t *)calloc((size_t)W * ALPHA, 8);
    while (w->corner_v >= 0 && w->corner_h >= 0) {
        if (sse && W == 2) window_fill_sse2(w, p, text, n);      /* dispatch :577 */
        else window_fill_scalar(w, p, text);
        window_walk(w, p, text, O, hew_threshold, score_only, out);
    }
    if (score_only) {                                           /* :599-607 */
        if (w->corner_h >= 0) w->score += w->corner_h + 1;
        if (w->corner_v >= 0) w->score += w->corner_v + 1;
    } else {                                                    /* :608-627 */
        for (int64_t h = w->corner_h; h >= 0; --h) ops_push_front(out, 'I');
        for (int64_t v = w->corner_v; v >= 0; --v) ops_push_front(out, 'D');
    }
    free(w->pv); free(w->mv); free(w->weq);
}

int64_t qo_windowed_score(const char *pattern, int m, const char *text, int n, int W, int O, int hew_threshold,
                          int sse, int64_t *hew)
{
    pat_t p; pat_build(&p, pattern, m);
    win_t w;
    windowed_run(&p, text, n, W, O, hew_threshold, sse, 1, &w, NULL);
    if (hew) *hew = w.hew;
    pat_free(&p);
    return w.score;
}

/* ------------------------------------------------------------------------------------------------
 * Output helpers: cigar_score_edit (cigar.c:274-289), cigar_sprint with matches (cigar.c:453-488).
 * ---------------------------------------------------------------------------------------------- */
static int64_t ops_cost(const ops_t *o)
{
    int64_t s = 0;
    for (int64_t i = o->begin; i < o->end; ++i) s += (o->ops[i] != 'M');
    return s;
}
static char *ops_to_rle(const ops_t *o)
{
    const int64_t len = o->end - o->begin;
    if (len <= 0) return NULL;                                  /* quicked.c:45 leaves aligner->cigar untouched */
    char *buf = (char *)malloc((size_t)(2 * len + 16)), *cur = buf;
    int64_t i = o->begin;
    while (i < o->end) {
        int64_t j = i;
        while (j < o->end && o->ops[j] == o->ops[i]) ++j;
        cur += sprintf(cur, "%lld%c", (long long)(j - i), o->ops[i]);
        i = j;
    }
    *cur = 0;
    return buf;
}

static char *reversed(const char *s, int64_t len)
{   /* commons.c:82 */
    char *r = (char *)malloc((size_t)len + 1);
    for (int64_t i = 0; i < len; ++i) r[len - 1 - i] = s[i];
    r[len] = 0;
    return r;
}

/* ------------------------------------------------------------------------------------------------
 * The schedule: quicked.c:58-306.
 * ---------------------------------------------------------------------------------------------- */
qo_params_t qo_default_params(void)
{   /* quicked.c:308-321 */
    qo_params_t p;
    memset(&p, 0, sizeof p);
    p.algo = QO_QUICKED; p.bandwidth = 15; p.window_size = 9; p.overlap_size = 1;
    p.hew_threshold[0] = p.hew_threshold[1] = 40;
    p.hew_percentage[0] = p.hew_percentage[1] = 15;
    return p;
}

const char *qo_status_msg(int status)
{   /* quicked.c:382-403 */
    switch (status) {
    case QO_ERROR: return "ERROR: QuickEd has finished with unspecific error\n";
    case QO_FAIL_NON_CONVERGENCE: return "ERROR: Hirschberg algorithm can not find a middle point of subsequence division!\n";
    case QO_UNIMPLEMENTED: return "ERROR: The algorithm or parameter combination selected is not implemented\n";
    case QO_UNKNOWN_ALGO: return "ERROR: Unknown algorithm selection\n";
    case QO_EMPTY_SEQUENCE: return "ERROR: Tried to align an empty sequence\n";
    default: return "QuickEd finished without errors.\n";
    }
}

static int align_core(const qo_params_t *prm, const char *pattern, int m, const char *text, int n, ops_t *ops,
                      int64_t *band_or_window_score, qo_result_t *diag)
{
    const int64_t maxlen = MAXI(m, n);
    const int sse = !prm->force_scalar;
    *band_or_window_score = -1;
    switch (prm->algo) {
    case QO_BANDED: {                                           /* run_banded, quicked.c:58-89 */
        const int64_t cutoff = (int64_t)((unsigned)maxlen * prm->bandwidth / 100);
        pat_t p; pat_build(&p, pattern, m);
        if (prm->only_score) {
            band_t b;
            const qo_band_geom_t g = qo_band_geometry(m, n, cutoff);
            uint64_t *pv = (uint64_t *)calloc((size_t)g.B_score + 1, 8), *mv = (uint64_t *)calloc((size_t)g.B_score + 1, 8);
            b.scores = (int64_t *)calloc((size_t)(p.nblk + g.B_score + 2), 8);
            *band_or_window_score = banded_score_run(&p, text, n, cutoff, n, &b, pv, mv);
            free(pv); free(mv); free(b.scores);
        } else {
            banded_full(&p, text, n, cutoff, ops);
        }
        pat_free(&p);
        return QO_WIP;
    }
    case QO_WINDOWED: {                                         /* run_windowed, quicked.c:91-123 */
        pat_t p; pat_build(&p, pattern, m);
        win_t w;
        windowed_run(&p, text, n, (int)prm->window_size, (int)prm->overlap_size, 0, sse, prm->only_score, &w, ops);
        *band_or_window_score = w.score;
        pat_free(&p);
        return QO_WIP;
    }
    case QO_HIRSCHBERG: {                                       /* run_hirschberg, quicked.c:125-161 */
        const int64_t cutoff = (int64_t)((unsigned)maxlen * prm->bandwidth / 100);
        char *tr = reversed(text, n), *pr = reversed(pattern, m);
        const int st = hirschberg(text, tr, n, pattern, pr, m, cutoff, ops);
        free(tr); free(pr);
        return st;
    }
    case QO_QUICKED: {                                          /* run_quicked, quicked.c:163-306 */
        char *tr = reversed(text, n), *pr = reversed(pattern, m);
        pat_t p; pat_build(&p, pattern, m);
        win_t w;
        windowed_run(&p, text, n, 2, 1, (int)prm->hew_threshold[0], sse, 1, &w, NULL);       /* stage 1 */
        int64_t score = w.score;
        if (diag) { diag->bound_ws = score; diag->stage = 1; }
        if ((int64_t)(w.hew * 64) > (int64_t)((unsigned)maxlen * prm->hew_percentage[0] / 100)) {   /* :201 */
            if (diag) diag->stage = 2;
            const int W = (int)prm->window_size, O = (int)prm->overlap_size;
            windowed_run(&p, text, n, W, O, (int)prm->hew_threshold[1], sse, 1, &w, NULL);
            score = w.score;
            uint64_t hew = (uint64_t)w.hew;
            pat_t prp; pat_build(&prp, pr, m);
            windowed_run(&prp, tr, n, W, O, (int)prm->hew_threshold[1], sse, 1, &w, NULL);
            pat_free(&prp);
            score = MINI(score, w.score);                                                       /* :229 */
            if (score >= w.score) hew = (uint64_t)w.hew;                                        /* :230 */
            if (hew * 64 * (uint64_t)(prm->window_size - prm->overlap_size) >
                (uint64_t)((unsigned)maxlen * prm->hew_percentage[1] / 100)) {                  /* :237 */
                if (diag) diag->stage = 3;
                score = MINI((int64_t)((unsigned)maxlen * prm->bandwidth / 100), score);        /* :246 */
                band_t b;
                int64_t nw;
                for (;;) {                                                                       /* :248-276 */
                    const qo_band_geom_t g = qo_band_geometry(m, n, score);
                    uint64_t *pv = (uint64_t *)calloc((size_t)g.B_score + 1, 8), *mv = (uint64_t *)calloc((size_t)g.B_score + 1, 8);
                    b.scores = (int64_t *)calloc((size_t)(p.nblk + g.B_score + 2), 8);
                    nw = banded_score_run(&p, text, n, score, n, &b, pv, mv);
                    free(pv); free(mv); free(b.scores);
                    if (diag) diag->banded_tries++;
                    if (!((nw > maxlen / 4 && score * 3 / 2 < nw) || nw < 0)) break;
                    score *= 2;
                }
                score = nw;                                                                      /* :278 */
            }
        }
        if (diag) diag->bound_final = score;
        hirschberg(text, tr, n, pattern, pr, m, score, ops);    /* status ignored, :290 */
        pat_free(&p); free(tr); free(pr);
        return QO_WIP;
    }
    default:
        return QO_UNKNOWN_ALGO;
    }
}

int qo_align(const qo_params_t *prm, const char *pattern, int m, const char *text, int n, qo_result_t *out)
{
    memset(out, 0, sizeof *out);
    out->score = -1;
    g_ws_windowed = g_ws_banded = 0; g_splits = 0; g_ref_undefined = 0;
    if (m == 0 || n == 0) { out->status = QO_EMPTY_SEQUENCE; return out->status; }   /* quicked.c:411-414 */
    ops_t ops;
    ops.ops = (char *)malloc((size_t)m + (size_t)n + 1);
    ops.begin = ops.end = (int64_t)m + n;
    int64_t aux = -1;
    out->status = align_core(prm, pattern, m, text, n, &ops, &aux, out);
    if (out->status != QO_UNKNOWN_ALGO) {
        /* extract_results, quicked.c:34-56.  only_score for QUICKED/HIRSCHBERG is uninitialised in the
         * reference; we return the distance of the traced alignment instead. */
        if (prm->only_score && (prm->algo == QO_BANDED || prm->algo == QO_WINDOWED)) out->score = aux;
        else {
            out->score = ops_cost(&ops);
            if (!prm->only_score) out->cigar = ops_to_rle(&ops);
        }
    }
    free(ops.ops);
    out->word_steps_windowed = g_ws_windowed; out->word_steps_banded = g_ws_banded;
    out->word_steps = g_ws_windowed + g_ws_banded;
    out->splits = g_splits;
    out->ref_undefined = g_ref_undefined;
    return out->status;
}

char *qo_align_ops(const qo_params_t *prm, const char *pattern, int m, const char *text, int n, int *status,
                   int64_t *score)
{
    if (m == 0 || n == 0) { if (status) *status = QO_EMPTY_SEQUENCE; return NULL; }
    ops_t ops;
    ops.ops = (char *)malloc((size_t)m + (size_t)n + 1);
    ops.begin = ops.end = (int64_t)m + n;
    int64_t aux;
    qo_params_t q = *prm; q.only_score = 0;
    const int st = align_core(&q, pattern, m, text, n, &ops, &aux, NULL);
    if (status) *status = st;
    if (score) *score = ops_cost(&ops);
    const int64_t len = ops.end - ops.begin;
    char *r = (char *)malloc((size_t)len + 1);
    memcpy(r, ops.ops + ops.begin, (size_t)len);
    r[len] = 0;
    free(ops.ops);
    return r;
}

void qo_free_result(qo_result_t *r) { free(r->cigar); r->cigar = NULL; }

/* ------------------------------------------------------------------------------------------------
 * Native batch driver (bench.py cpu_baseline kind "port"): contiguous ranges over `threads` pthreads,
 * like the reference tool's OpenMP loop (tools/align_benchmark/align_benchmark.c:269-284).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    const char *seqs; const int64_t *po, *to; const int32_t *pl, *tl; int64_t lo, hi;
    qo_params_t prm; int32_t *score; int64_t bytes;
} qo_job_t;

static void *qo_batch_worker(void *arg)
{
    qo_job_t *j = (qo_job_t *)arg;
    for (int64_t i = j->lo; i < j->hi; ++i) {
        qo_result_t r;
        qo_align(&j->prm, j->seqs + j->po[i], j->pl[i], j->seqs + j->to[i], j->tl[i], &r);
        if (j->score) j->score[i] = (int32_t)r.score;
        if (r.cigar) j->bytes += (int64_t)strlen(r.cigar);
        qo_free_result(&r);
    }
    return NULL;
}

int64_t qo_batch_align(const char *seqs, const int64_t *po, const int32_t *pl, const int64_t *to, const int32_t *tl,
                       int64_t n, int threads, const qo_params_t *prm, int32_t *score_out)
{
    if (threads < 1) threads = 1;
    if (threads > n) threads = (int)(n > 0 ? n : 1);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    qo_job_t *jobs = (qo_job_t *)malloc(sizeof(qo_job_t) * (size_t)threads);
    for (int t = 0; t < threads; ++t) {
        qo_job_t jb = {seqs, po, to, pl, tl, n * t / threads, n * (t + 1) / threads, *prm, score_out, 0};
        jobs[t] = jb;
        pthread_create(&th[t], NULL, qo_batch_worker, &jobs[t]);
    }
    int64_t total = 0;
    for (int t = 0; t < threads; ++t) { pthread_join(th[t], NULL); total += jobs[t].bytes; }
    free(th); free(jobs);
    return total;
}
