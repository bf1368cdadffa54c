/*
 * oracle/datagen.c — seeded twin of the reference's generate_dataset edit model, for the CHECKER side.
 * TEST / BASELINE INFRASTRUCTURE ONLY: bench.py --impl reference generates its pairs here so that the reference arm
 * never loads the product library.  Same model and the same random streams as the product's qb200_generate_pairs_ex
 * (tests/test_cabi.py checks the two byte for byte): text = `length` uniform ACGT (generate_dataset.c:52-63); pattern =
 * a copy with ceil(length * error) edits, each uniformly mismatch / deletion / insertion at a uniform position
 * (:108-199; error >= 1 is an absolute count, :370); optional --indels N,LEN: a uniform count in [0, N] of LEN-long
 * deletions (:204-245).  Every pair has its own stream = hash(seed, pair index in the job).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t splitmix64(uint64_t *x)
{
    uint64_t z = (*x += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
static inline uint32_t rand_below(uint64_t *s, uint32_t n) { return (uint32_t)(((splitmix64(s) >> 32) * (uint64_t)n) >> 32); }

typedef struct {
    uint64_t seed; int64_t first, n, w, nt; int32_t length, num_errors, indels_num, indels_len;
    char *seqs; int64_t *po, *to; int32_t *pl, *tl;
} gen_job_t;

static void *gen_worker(void *arg)
{
    const gen_job_t *j = (const gen_job_t *)arg;
    static const char alphabet[4] = {'A', 'C', 'G', 'T'};
    const int32_t length = j->length, num_errors = j->num_errors;
    const int64_t stride = 2 * (int64_t)length + num_errors + 2;
    for (int64_t i = j->w; i < j->n; i += j->nt) {
        uint64_t s = j->seed ^ 0x5851f42d4c957f2dull;
        s = splitmix64(&s) ^ ((uint64_t)(j->first + i) * 0xd6e8feb86659fd93ull);
        s = splitmix64(&s);
        char *pat = j->seqs + i * stride, *txt = pat + length + num_errors + 1;
        for (int k = 0; k < length; ++k) txt[k] = alphabet[rand_below(&s, 4)];
        memcpy(pat, txt, (size_t)length);
        int len = length;
        for (int e = 0; e < num_errors; ++e) {
            const uint32_t kind = rand_below(&s, 3);
            if (kind == 0 && len > 0) {
                const uint32_t pos = rand_below(&s, (uint32_t)len);
                char c;
                do { c = alphabet[rand_below(&s, 4)]; } while (c == pat[pos]);
                pat[pos] = c;
            } else if (kind == 1 && len > 1) {
                const uint32_t pos = rand_below(&s, (uint32_t)len);
                memmove(pat + pos, pat + pos + 1, (size_t)(len - 1 - (int)pos));
                --len;
            } else {
                const uint32_t pos = rand_below(&s, (uint32_t)(len > 1 ? len : 1));
                memmove(pat + pos + 1, pat + pos, (size_t)(len - (int)pos));
                pat[pos] = alphabet[rand_below(&s, 4)];
                ++len;
            }
        }
        if (j->indels_num > 0 && j->indels_len > 0) {
            const uint32_t cnt = rand_below(&s, (uint32_t)j->indels_num + 1);
            for (uint32_t d = 0; d < cnt; ++d) {
                const uint32_t pos = rand_below(&s, (uint32_t)(len > 1 ? len : 1));
                if (j->indels_len >= len) continue;
                const int nl = len - j->indels_len;
                if ((int)pos < nl) memmove(pat + pos, pat + pos + j->indels_len, (size_t)(nl - (int)pos));
                len = nl;
            }
        }
        pat[len] = 0;
        txt[length] = 0;
        j->po[i] = i * stride; j->pl[i] = len;
        j->to[i] = i * stride + length + num_errors + 1; j->tl[i] = length;
    }
    return NULL;
}

/* Returns the bytes written to seqs (n_pairs * (2*length + num_errors + 2)) or -1. */
int64_t qo_generate_pairs(uint64_t seed, int64_t first_pair, int64_t n_pairs, int32_t length, double error, int32_t indels_num,
                          int32_t indels_len, int threads, char *seqs, int64_t *po, int32_t *pl, int64_t *to, int32_t *tl)
{
    if (n_pairs < 0 || first_pair < 0 || length <= 0 || !seqs || indels_num < 0 || indels_len < 0) return -1;
    const int32_t num_errors = error >= 1.0 ? (int32_t)error : (int32_t)ceil((double)((float)length * (float)error));
    if (threads < 1) threads = 1;
    if (threads > n_pairs) threads = (int)(n_pairs > 0 ? n_pairs : 1);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    gen_job_t *jobs = (gen_job_t *)malloc(sizeof(gen_job_t) * (size_t)threads);
    for (int t = 0; t < threads; ++t) {
        jobs[t] = (gen_job_t){seed, first_pair, n_pairs, t, threads, length, num_errors, indels_num, indels_len, seqs, po, to, pl, tl};
        pthread_create(&th[t], NULL, gen_worker, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
    free(th); free(jobs);
    return n_pairs * (2 * (int64_t)length + num_errors + 2);
}
