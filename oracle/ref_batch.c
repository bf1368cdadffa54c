/*
 * oracle/ref_batch.c — native batch driver around the UNMODIFIED reference library (oracle/_ref/libquicked_ref.so).
 * TEST / BASELINE INFRASTRUCTURE ONLY (used by bench.py's cpu_baseline and --impl reference legs).
 *
 * Mirrors what the reference's own benchmark tool does per thread (reference tools/align_benchmark/align_benchmark.c:246-284
 * and benchmark/benchmark_edit.c:36-89): one mm_allocator per thread, and per pair quicked_new -> quicked_align ->
 * quicked_free with params.external_allocator set.  Threads split the batch into contiguous ranges (pthreads instead
 * of OpenMP).  Timing is taken by the caller around this one call, so no Python / ctypes overhead lands inside it.
 */
#include <pthread.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* the reference's public API (quicked/quicked.h:36-96); layout checked by tests/test_cabi.py */
typedef struct mm_allocator_t mm_allocator_t;
typedef struct {
    int algo; unsigned bandwidth, window_size, overlap_size, hew_threshold[2], hew_percentage[2];
    bool only_score, force_scalar, external_timer; mm_allocator_t *external_allocator;
} ref_params_t;
typedef struct {
    ref_params_t *params; mm_allocator_t *mm_allocator; char *cigar; int score; void *timers[5];
} ref_aligner_t;
extern ref_params_t quicked_default_params(void);
extern int quicked_new(ref_aligner_t *, ref_params_t *);
extern int quicked_align(ref_aligner_t *, const char *, int, const char *, int);
extern int quicked_free(ref_aligner_t *);
extern mm_allocator_t *mm_allocator_new(uint64_t segment_size);          /* quicked_utils/include/mm_allocator.h:57-58 */
extern void mm_allocator_delete(mm_allocator_t *);

typedef struct {
    const char *seqs; const int64_t *po, *to; const int32_t *pl, *tl;
    int64_t lo, hi; ref_params_t prm; int32_t *score; int64_t *cigar_bytes;
} job_t;

static void *worker(void *arg)
{
    job_t *j = (job_t *)arg;
    mm_allocator_t *mm = mm_allocator_new(1ull << 27);                   /* BUFFER_SIZE_128M, align_benchmark.c:51 */
    j->prm.external_allocator = mm;
    j->prm.external_timer = true;                                        /* benchmark_edit.c:47: the tool owns the timers */
    /* five per-thread timers the aligner points at (benchmark_edit.c:61-65); profiler_timer_t is 88 bytes */
    static __thread unsigned char timers[5][128] __attribute__((aligned(16)));
    memset(timers, 0, sizeof timers);
    int64_t bytes = 0;
    for (int64_t i = j->lo; i < j->hi; ++i) {
        ref_aligner_t a;
        quicked_new(&a, &j->prm);
        for (int t = 0; t < 5; ++t) a.timers[t] = timers[t];
        /* NUL-terminated copies are not needed: the packed buffer keeps a byte after every text (see bench.py) */
        quicked_align(&a, j->seqs + j->po[i], j->pl[i], j->seqs + j->to[i], j->tl[i]);
        if (j->score) j->score[i] = a.score;
        if (a.cigar) bytes += (int64_t)strlen(a.cigar);
        quicked_free(&a);
    }
    *j->cigar_bytes = bytes;
    mm_allocator_delete(mm);
    return NULL;
}

/* Align pairs [0, n) with `threads` host threads.  Returns the total CIGAR text bytes produced (so the work cannot be
 * optimised away) or -1 on error.  `algo`..`only_score` follow quicked_params_t. */
int64_t ref_batch_align(const char *seqs, const int64_t *po, const int32_t *pl, const int64_t *to, const int32_t *tl,
                        int64_t n, int threads, int algo, unsigned bandwidth, unsigned window_size, unsigned overlap_size,
                        int only_score, int force_scalar, int32_t *score_out)
{
    if (threads < 1) threads = 1;
    if (threads > n) threads = (int)(n > 0 ? n : 1);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    job_t *jobs = (job_t *)malloc(sizeof(job_t) * (size_t)threads);
    int64_t *bytes = (int64_t *)calloc((size_t)threads, sizeof(int64_t));
    ref_params_t prm = quicked_default_params();
    prm.algo = algo; prm.bandwidth = bandwidth; prm.window_size = window_size; prm.overlap_size = overlap_size;
    prm.only_score = only_score != 0; prm.force_scalar = force_scalar != 0;
    for (int t = 0; t < threads; ++t) {
        jobs[t] = (job_t){seqs, po, to, pl, tl, n * t / threads, n * (t + 1) / threads, prm, score_out, &bytes[t]};
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    int64_t total = 0;
    for (int t = 0; t < threads; ++t) { pthread_join(th[t], NULL); total += bytes[t]; }
    free(th); free(jobs); free(bytes);
    return total;
}
