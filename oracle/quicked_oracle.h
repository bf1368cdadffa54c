/*
 * oracle/quicked_oracle.h — CPU restatement of QuickEd's bound-and-align hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under quicked_b200/ (the product) may include, link or call this.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, and
 * only as the checker (or the reported CPU baseline), never as the thing shipped or measured as "ours".
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks every entry point below against
 *   (a) the reference's own known answers (GATC/GATO -> 1, ACGT/ACTT -> 1 "2M1X1M",
 *       empty -> EMPTY_SEQUENCE, ONT pair -> 39740), and
 *   (b) the unmodified reference compiled from /root/reference into oracle/_ref/libquicked_ref.so
 *       (oracle/Makefile), on seeded generate_dataset-model inputs, and
 *   (c) committed golden vectors under tests/golden/ produced by oracle/gen_golden.py from (b).
 *
 * Citations are file:line under /root/reference.
 */
#ifndef QUICKED_ORACLE_H
#define QUICKED_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* quicked/quicked.h:36-41 */
enum { QO_QUICKED = 0, QO_WINDOWED = 1, QO_BANDED = 2, QO_HIRSCHBERG = 3 };
/* quicked/quicked.h:69-79 */
enum { QO_OK = 0, QO_ERROR = -1, QO_FAIL_NON_CONVERGENCE = -2, QO_UNKNOWN_ALGO = -3,
       QO_EMPTY_SEQUENCE = -4, QO_UNIMPLEMENTED = -10, QO_WIP = 1 };

/* mirrors quicked_params_t (quicked/quicked.h:43-54) minus allocator/timer plumbing */
typedef struct {
    int      algo;
    unsigned bandwidth;
    unsigned window_size;
    unsigned overlap_size;
    unsigned hew_threshold[2];
    unsigned hew_percentage[2];
    int      only_score;
    int      force_scalar;   /* 1: scalar WindowEd(2,1); 0: emulate the SSE4.1 variant (x86 default) */
} qo_params_t;

typedef struct {
    int      status;
    int64_t  score;          /* aligner->score */
    char    *cigar;          /* malloc'd RLE text ("2M1X1M"), NULL when only_score; caller frees */
    /* diagnostics (not part of the reference API) */
    int64_t  bound_ws;       /* WindowEd(S) estimate */
    int64_t  bound_final;    /* cutoff handed to the Hirschberg stage */
    int      stage;          /* 1,2,3: last bound stage entered by QUICKED */
    int      banded_tries;   /* number of BandEd score-only runs in stage 3 */
    int      splits;         /* number of Hirschberg splits */
    int      ref_undefined;  /* 1: the reference reads uninitialised memory on this input (its result is UB) */
    uint64_t word_steps;     /* exact count of 64-row x 1-column Myers block updates (scalar schedule) */
    uint64_t word_steps_windowed;
    uint64_t word_steps_banded;
} qo_result_t;

qo_params_t qo_default_params(void);                              /* quicked.c:308-321 */
int  qo_align(const qo_params_t *params, const char *pattern, int m, const char *text, int n,
              qo_result_t *out);                                   /* quicked.c:405-437 */
const char *qo_status_msg(int status);                             /* quicked.c:382-403 */
void qo_free_result(qo_result_t *r);

/* Uncompressed op string (M/X/I/D), for tests that replay the alignment; malloc'd, NUL terminated. */
char *qo_align_ops(const qo_params_t *params, const char *pattern, int m, const char *text, int n,
                   int *status, int64_t *score);

/* ---- internals exposed so tests can pin each kernel against the reference's internal entry points ---- */

/* BandEd geometry, SURVEY App. A.2 (bpm_banded.c:121-135, 359-361) */
typedef struct { int64_t k, d, rel, prolog, B_cigar, B_score, fin; } qo_band_geom_t;
qo_band_geom_t qo_band_geometry(int64_t m, int64_t n, int64_t cutoff);

/* BandEd score-only up to column `finish` (bpm_banded.c:791-964). Returns the score; optional outputs:
 * final band state for Hirschberg. pv/mv must hold B_score words, scores ceil(m/64)+B_score+2. */
int64_t qo_banded_score(const char *pattern, int m, const char *text, int n, int64_t cutoff, int64_t finish,
                        uint64_t *pv, uint64_t *mv, int64_t *scores, int64_t *lower_block, int64_t *higher_block);

/* Test hook: one BandEd leaf (bpm_banded.c:199-316 fill + :967-1036 walk) with the stored matrix exported.
 * pv/mv: B_cigar*(n+1) words each, [column][band word]; ranges: (first,last) per 64-column block (n/64+1 pairs);
 * ops: the walk's op string, m+n+1 bytes.  Any pointer may be NULL except pattern/text. */
int64_t qo_banded_full_dump(const char *pattern, int m, const char *text, int n, int64_t cutoff, uint64_t *pv,
                            uint64_t *mv, int64_t *ranges, char *ops);

/* WindowEd score-only (bpm_windowed.c:563-628 with SCORE_ONLY). sse!=0 emulates windowed_compute_window_sse
 * (only meaningful for W==2). Returns the score estimate, *hew = high-error-window count. */
int64_t qo_windowed_score(const char *pattern, int m, const char *text, int n, int W, int O,
                          int hew_threshold, int sse, int64_t *hew);

uint64_t qo_word_steps_total(void);   /* running counters (reset by qo_align) */

/* Batch of pairs in the packed layout of include/quicked_b200.h over `threads` pthreads; returns the CIGAR bytes produced. */
int64_t qo_batch_align(const char *seqs, const int64_t *po, const int32_t *pl, const int64_t *to, const int32_t *tl,
                       int64_t n, int threads, const qo_params_t *prm, int32_t *score_out);

#ifdef __cplusplus
}
#endif
#endif
