/*
 * include/quicked_b200.h — additive batched C-ABI of the B200-native QuickEd path.
 *
 * The reference library aligns one pair per quicked_align() call; its only batching is the OpenMP loop in
 * the benchmark tool (reference: tools/align_benchmark/align_benchmark.c:232-306) over a packed
 * sequence buffer + offsets (reference: quicked_utils/include/sequence_buffer.h:30-50).  A GPU
 * needs the whole batch at once, so this header adds a batched entry point with that same packed layout.
 * quicked_align() (include/quicked.h) is a batch of one through the same code.
 *
 * Plain C: pointers and sizes only, no torch / C++ types.  All functions return 0 on success or a
 * negative qb200 error code; per-pair outcomes are quicked_status_t values in `status[]`.
 * There is no CPU fallback: without a CUDA device every compute entry point fails with QB200_ERR_NO_DEVICE.
 */
#ifndef QUICKED_B200_H
#define QUICKED_B200_H

#include <stddef.h>
#include <stdint.h>
#include "quicked.h"

#ifdef __cplusplus
extern "C" {
#endif

enum {
    QB200_OK = 0,
    QB200_ERR_NO_DEVICE = -100,   /* no CUDA device / driver: the product path refuses to run */
    QB200_ERR_CUDA = -101,        /* a CUDA call failed; see qb200_last_error() */
    QB200_ERR_ARG = -102,
    QB200_ERR_OOM = -103,
    QB200_ERR_CAPACITY = -104,    /* caller's cigar buffer too small; *cigar_bytes holds the size needed */
};

typedef struct qb200_ctx qb200_ctx_t;

/* A batch of pairs: sequence i is seqs[off[i] .. off[i]+len[i]).  Mirrors sequence_buffer_t
 * (reference sequence_buffer.h:30-50): one packed character buffer plus per-pair offsets and lengths.
 * Bytes are raw ASCII exactly as the reference takes them (dna_encode semantics, reference dna_text.c:41-46).
 * At most 2^31 - 2^20 pairs per batch.  Big jobs are pipelined in sub-batches that upload the byte range spanned by
 * their pairs: keep pattern i and text i close to each other (as sequence_buffer_t does); layouts such as "all
 * patterns, then all texts" still work but every sub-batch then uploads nearly the whole buffer. */
typedef struct {
    const char    *seqs;          /* packed characters                                   */
    int64_t        seqs_bytes;
    int64_t        n_pairs;
    const int64_t *pattern_off;   /* [n_pairs] */
    const int32_t *pattern_len;   /* [n_pairs] */
    const int64_t *text_off;      /* [n_pairs] */
    const int32_t *text_len;      /* [n_pairs] */
} qb200_batch_t;

/* Caller-provided result buffers (host memory for *_host calls). */
typedef struct {
    int32_t *score;               /* [n_pairs] aligner->score of each pair                               */
    int32_t *status;              /* [n_pairs] quicked_status_t of each pair                             */
    char    *cigar;               /* packed NUL-terminated CIGAR strings; may be NULL when only_score     */
    int64_t  cigar_capacity;      /* bytes available at `cigar`                                           */
    int64_t *cigar_off;           /* [n_pairs+1] string i is cigar[cigar_off[i] .. cigar_off[i+1]-1) + NUL */
    int64_t  cigar_bytes;         /* out: bytes written (or needed, with QB200_ERR_CAPACITY)              */
} qb200_results_t;

typedef struct {                  /* counters of the last qb200_run(); all times are CUDA-event ms */
    int64_t n_pairs;
    int64_t kernel_launches;      /* our own kernels launched by the last run                              */
    int64_t word_steps;           /* 64-row x 1-column Myers block updates executed (device-counted)       */
    int64_t word_steps_windowed, word_steps_banded;
    int64_t cells;                /* sum of m*n over the batch (GCUPS_equiv numerator)                     */
    int64_t h2d_bytes, d2h_bytes;
    int64_t pairs_stage2, pairs_stage3, banded_tries, hirschberg_splits, leaves;
    float   ms_total, ms_prepare, ms_windowed_s, ms_windowed_l, ms_banded, ms_align_fill, ms_align_trace,
            ms_cigar;
    int64_t matrix_bytes;         /* traceback state written to HBM (Pv/Mv columns)                        */
    float   ms_fused;             /* the fused WindowEd+BandEd+traceback kernel of the QUICKED fast path   */
    int32_t leaves_punted;        /* leaves the tile path handed to the exact full-matrix kernels (too-narrow bands)   */
    int64_t pairs_fused;          /* pairs completed by that kernel                                        */
} qb200_stats_t;

/* --- context --- */
int  qb200_device_count(void);                                   /* 0 when no usable CUDA device          */
int  qb200_create(qb200_ctx_t **ctx, int device);                /* one context per GPU / host thread     */
void qb200_destroy(qb200_ctx_t *ctx);
int  qb200_set_stream(qb200_ctx_t *ctx, void *cuda_stream);     /* launch on the caller's stream (e.g. torch's) */
int  qb200_set_workspace_limit(qb200_ctx_t *ctx, size_t bytes); /* cap for the traceback-state pool (default 48 GiB) */
const char *qb200_last_error(qb200_ctx_t *ctx);

/* --- three-phase batch API: upload (H2D) -> run (kernels only) -> download (D2H) --- */
int qb200_upload(qb200_ctx_t *ctx, const qb200_batch_t *host_batch);             /* pageable or pinned host memory */
int qb200_upload_device(qb200_ctx_t *ctx, const qb200_batch_t *device_batch);    /* arrays already in HBM: no copy */
int qb200_run(qb200_ctx_t *ctx, const quicked_params_t *params);                  /* results stay in HBM           */
int qb200_download(qb200_ctx_t *ctx, qb200_results_t *host_results);
int qb200_get_stats(qb200_ctx_t *ctx, qb200_stats_t *stats);
/* After a QUICKED run: the stage-1 WindowEd(S) result of every pair of the batch — its score (the first alignment
 * bound, reference quicked.c:178-199 `aligner->score` after run_windowed_score) and its count of high-error windows
 * (reference bpm_windowed.c:555-557).  Either pointer may be NULL.  n must equal the batch's n_pairs. */
int qb200_get_bounds(qb200_ctx_t *ctx, int32_t *bound, int32_t *high_error_windows, int64_t n);

/* --- one call, host in / host out: what a reference caller's batch loop is replaced by ---
 * Big jobs (>= 200 000 pairs or >= 512 MB of characters) are cut into sub-batches and pipelined (H2D / kernels / D2H
 * overlap); CIGAR strings still come back packed in input order.  Page-locked caller buffers (qb200_host_alloc or
 * cudaHostRegister) avoid one staging copy.  On QB200_ERR_CAPACITY scores, statuses and cigar_bytes (the size needed)
 * are valid: grow the cigar buffer and call again (a pipelined job keeps nothing on the device; after a small,
 * unpipelined job qb200_download works too). */
int qb200_align_batch(qb200_ctx_t *ctx, const quicked_params_t *params,
                      const qb200_batch_t *host_batch, qb200_results_t *host_results);

/* --- 2-bit packed input (BASELINE north_star design 3): a quarter of the bytes over PCIe ---
 * The character stream of a qb200_batch_t with 4 characters per byte: character i sits in bits 2*(i&3) .. 2*(i&3)+1 of
 * packed[i>>2], A = 0, C = 1, G = 2, T = 3 (upper case).  Every other byte INSIDE a sequence (N, lower case, IUPAC, ...)
 * is stored as 0 and listed as an exception (exc_pos ascending = character index, exc_chr = the raw byte), so the
 * alignment sees exactly the caller's characters (reference dna_text.c:41-46 and the raw-byte compares of the
 * tracebacks).  Offsets are character indices into the stream; bytes between sequences are not kept.
 * The device expands the stream back to the ASCII buffer the kernels read (k_unpack2 / k_patch_exceptions). */
typedef struct {
    const uint8_t *packed;        /* [(n_chars + 3) / 4]                                 */
    int64_t        n_chars;
    int64_t        n_pairs;
    const int64_t *pattern_off;   /* [n_pairs] character index                           */
    const int32_t *pattern_len;
    const int64_t *text_off;
    const int32_t *text_len;
    const int64_t *exc_pos;       /* [n_exc] ascending character indices                 */
    const uint8_t *exc_chr;       /* [n_exc] the raw bytes there                         */
    int64_t        n_exc;
} qb200_packed_batch_t;
/* Host packer: characters of `in` -> packed (capacity >= (in->seqs_bytes + 3) / 4 + 8 bytes) + exception list; character
 * index = byte index of in->seqs, so in's offset / length arrays serve the packed batch unchanged.  `threads` host threads
 * (<= 0: all).  Returns the number of exceptions, or -(number needed) when exc_cap is too small, or QB200_ERR_ARG. */
int64_t qb200_pack_batch(const qb200_batch_t *in, uint8_t *packed, int64_t *exc_pos, uint8_t *exc_chr, int64_t exc_cap, int threads);
int qb200_upload_packed(qb200_ctx_t *ctx, const qb200_packed_batch_t *host_batch);           /* then qb200_run / qb200_download */
int qb200_align_batch_packed(qb200_ctx_t *ctx, const quicked_params_t *params,
                             const qb200_packed_batch_t *host_batch, qb200_results_t *host_results);

/* --- measured integer-ALU peak (LOP3+IADD3 mix, no memory traffic), in 10^12 int32 ops/s: the denominator of the
 * bit-op roofline (MEASURED_PEAKS.json carries only HBM and bf16 peaks) --- */
int qb200_measure_int_peak(qb200_ctx_t *ctx, double *tera_ops_per_s);

/* --- pinned host staging helpers (cudaHostAlloc / cudaFreeHost) --- */
void *qb200_host_alloc(size_t bytes);
void  qb200_host_free(void *p);

/* --- seeded twin of the reference's generate_dataset edit model (reference generate_dataset.c:108-199,366-410):
 * fills a packed batch (pattern i then text i, back to back) in host memory.  Returns bytes written to seqs,
 * or a negative error.  seqs must hold n_pairs * (2*length + ceil(length*error) + 2) bytes. */
int64_t qb200_generate_pairs(uint64_t seed, int64_t n_pairs, int32_t length, double error,
                             char *seqs, int64_t *pattern_off, int32_t *pattern_len,
                             int64_t *text_off, int32_t *text_len);

/* The same generator for pairs [first_pair, first_pair + n_pairs) of job `seed` (every pair has its own random stream,
 * so ranks can generate disjoint slices of one job), plus the reference's `--indels N,LEN` switch
 * (generate_dataset.c:204-245: a uniform count in [0, N] of LEN-long deletions).  Offsets are relative to `seqs`. */
int64_t qb200_generate_pairs_ex(uint64_t seed, int64_t first_pair, int64_t n_pairs, int32_t length, double error,
                                int32_t indels_num, int32_t indels_len, char *seqs, int64_t *pattern_off,
                                int32_t *pattern_len, int64_t *text_off, int32_t *text_len);

/* The same generator as a KERNEL: the batch is born in the context's device buffers (one CTA per pair; the text is drawn
 * counter-based by all threads, the edits are replayed in order with the tail moves done in shared memory) and is
 * byte-identical to what qb200_generate_pairs_ex writes for the same arguments.  Takes the place of qb200_upload: follow
 * with qb200_run / qb200_download.  Reads up to ~160 kbp (the pattern has to fit a CTA's shared memory). */
int qb200_generate_device(qb200_ctx_t *ctx, uint64_t seed, int64_t first_pair, int64_t n_pairs, int32_t length, double error,
                          int32_t indels_num, int32_t indels_len);
/* The batch a context holds (uploaded, unpacked from 2 bits, or generated), copied back to the host: seqs must hold
 * qb200_batch_bytes() bytes; any pointer may be NULL. */
int64_t qb200_batch_bytes(qb200_ctx_t *ctx);
int qb200_download_batch(qb200_ctx_t *ctx, char *seqs, int64_t *pattern_off, int32_t *pattern_len, int64_t *text_off, int32_t *text_len);

/* --- SAM-style CIGAR of one alignment (host-side string transform; reference cigar_compute_CIGAR /
 * cigar_sprint_SAM_CIGAR, quicked_utils/src/cigar.c:193-240, :504-529).  `cigar` is the run-length text this library
 * and the reference's quicked_align produce ("12M1X3I...": M match, X mismatch, I consumes a text character,
 * D a pattern character).  show_mismatches != 0: matches print as '=', mismatches as 'X'; == 0: both merge into 'M'
 * — except that, like the reference, the very first operation is taken as it is (an alignment that starts with a
 * mismatch begins "1X").  Returns the length written (NUL-terminated), or -(bytes needed) if `capacity` is too small. */
int64_t qb200_cigar_to_sam(const char *cigar, int show_mismatches, char *out, int64_t capacity);

#ifdef __cplusplus
}
#endif
#endif /* QUICKED_B200_H */
