/*
 * include/quicked.h — drop-in C API of the B200-native QuickEd bound-and-align path.
 *
 * This header is layout- and symbol-compatible with the reference's public header
 * (reference: quicked/quicked.h:33-96).  A program written against the reference header can be compiled
 * against this one and linked with libquicked_b200.so instead of libquicked.a without source changes:
 *
 *   quicked_params_t   48 bytes: algo@0 bandwidth@4 window_size@8 overlap_size@12 hew_threshold@16
 *                      hew_percentage@24 only_score@32 force_scalar@33 external_timer@34 external_allocator@40
 *   quicked_aligner_t  72 bytes: params@0 mm_allocator@8 cigar@16 score@24 timer@32 timer_windowed_s@40
 *                      timer_windowed_l@48 timer_banded@56 timer_align@64
 *   (checked by static_asserts in quicked_b200/csrc/qb_capi.cu and by tests/test_cabi.py)
 *
 * The reference header pulls in quicked_utils/include/mm_allocator.h and profiler_timer.h because those
 * types appear as pointer fields.  The GPU library never allocates from the caller's arena (device
 * workspaces are its own; the CIGAR string is malloc'd and owned by the aligner), so mm_allocator_t is
 * declared opaque here; profiler_timer_t is declared with the reference layout
 * (quicked_utils/include/profiler_timer.h:51-57, profiler_counter.h:41-50; 88 bytes) because callers
 * that set params.external_timer overwrite the five timer pointers after quicked_new()
 * (tools/align_benchmark/benchmark/benchmark_edit.c:61-65) and read the accumulated times afterwards.
 * If the reference's own utility headers were included first, their definitions are used instead.
 */
#ifndef QUICKED_H
#define QUICKED_H

#include <stdbool.h>
#include <stdint.h>
#include <time.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef MM_ALLOCATOR_H_
typedef struct mm_allocator_t mm_allocator_t;    /* opaque: quicked_utils/include/mm_allocator.h:41-52 */
#endif

#ifndef PROFILER_TIMER_H
typedef struct {                                   /* quicked_utils/include/profiler_counter.h:41-50 */
    uint64_t total, samples, min, max;
    double m_oldM, m_newM, m_oldS, m_newS;
} profiler_counter_t;
typedef struct {                                   /* quicked_utils/include/profiler_timer.h:51-57 */
    struct timespec begin_timer;
    profiler_counter_t time_ns;
    uint64_t accumulated;
} profiler_timer_t;
#endif

#define QUICKED_WINDOW_STAGES 2          /* reference quicked.h:32 */
#define QUICKED_FAST_WINDOW_SIZE 2       /* WindowEd(S): 2 words ...          reference quicked.h:33 */
#define QUICKED_FAST_WINDOW_OVERLAP 1    /* ... overlapping by 1              reference quicked.h:34 */

typedef enum {                            /* reference quicked.h:36-41 */
    QUICKED,      /* WindowEd bound -> (WindowEd(L) -> BandEd doubling) -> Hirschberg/BandEd alignment */
    WINDOWED,     /* WindowEd heuristic alignment only */
    BANDED,       /* BandEd with cutoff = max(m,n) * bandwidth / 100 */
    HIRSCHBERG,   /* Hirschberg over BandEd with that cutoff */
} quicked_algo_t;

typedef struct quicked_params_t {         /* reference quicked.h:43-54 */
    quicked_algo_t algo;
    unsigned int bandwidth;               /* percent of max(m,n) */
    unsigned int window_size;             /* WindowEd(L) window, in 64-bit words */
    unsigned int overlap_size;            /* WindowEd(L) overlap, in words */
    unsigned int hew_threshold[QUICKED_WINDOW_STAGES];    /* percent error that makes a window "high error" */
    unsigned int hew_percentage[QUICKED_WINDOW_STAGES];   /* percent of HEWs that escalates to the next stage */
    bool only_score;
    bool force_scalar;                    /* true: scalar WindowEd(2,1) semantics; false: the SSE4.1 variant's */
    bool external_timer;
    mm_allocator_t *external_allocator;   /* accepted for ABI compatibility; never used for device memory */
} quicked_params_t;

typedef struct quicked_aligner_t {        /* reference quicked.h:56-67 */
    quicked_params_t *params;             /* borrowed: re-read on every quicked_align() (reference quicked.c:327) */
    mm_allocator_t *mm_allocator;
    char *cigar;                          /* NUL-terminated "<n><op>" runs over M X I D; NULL when only_score */
    int score;
    profiler_timer_t *timer;
    profiler_timer_t *timer_windowed_s;
    profiler_timer_t *timer_windowed_l;
    profiler_timer_t *timer_banded;
    profiler_timer_t *timer_align;
} quicked_aligner_t;

typedef enum quicked_status_t {           /* reference quicked.h:69-79 */
    QUICKED_OK                   = 0,
    QUICKED_ERROR                = -1,
    QUICKED_FAIL_NON_CONVERGENCE = -2,
    QUICKED_UNKNOWN_ALGO         = -3,
    QUICKED_EMPTY_SEQUENCE       = -4,
    QUICKED_UNIMPLEMENTED        = -10,
    QUICKED_WIP                  = 1,     /* "not an error": what new/free/align return on success */
} quicked_status_t;

/* reference quicked.h:81-96 — same names, argument meaning and return conventions */
bool quicked_check_error(quicked_status_t status);
const char *quicked_status_msg(quicked_status_t status);
quicked_params_t quicked_default_params(void);
quicked_status_t quicked_new(quicked_aligner_t *aligner, quicked_params_t *params);
quicked_status_t quicked_free(quicked_aligner_t *aligner);
quicked_status_t quicked_align(quicked_aligner_t *aligner,
                               const char *pattern, const int pattern_len,
                               const char *text, const int text_len);

#ifdef __cplusplus
}
#endif
#endif /* QUICKED_H */
