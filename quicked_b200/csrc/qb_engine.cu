// qb_engine.cu — host orchestration of the batched bound-and-align schedule + the C-ABI entry points.
//
// Mirrors, per batch instead of per pair, the reference driver quicked/src/quicked.c:
//   run_quicked (:163-306), run_banded (:58-89), run_windowed (:91-123), run_hirschberg (:125-161),
//   extract_results (:34-56), and the recursion of bpm_compute_matrix_hirschberg (bpm_hirschberg.c:33-270)
//   turned into a level-synchronous work queue.
// No CPU fallback exists: every compute entry point fails with QB200_ERR_NO_DEVICE when there is no GPU.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/quicked_b200.h"
#include <cub/device/device_scan.cuh>

#include "qb_banded.cuh"
#include "qb_common.cuh"
#include "qb_plan.cuh"
#include "qb_prep.cuh"
#include "qb_traceback.cuh"
#include "qb_windowed.cuh"

using namespace qb;

// ---- ABI checks against the reference layout ----
static_assert(sizeof(quicked_params_t) == 48, "quicked_params_t must stay 48 bytes");
static_assert(offsetof(quicked_params_t, hew_threshold) == 16 && offsetof(quicked_params_t, only_score) == 32 &&
              offsetof(quicked_params_t, external_allocator) == 40, "quicked_params_t layout");
static_assert(sizeof(quicked_aligner_t) == 72 && offsetof(quicked_aligner_t, score) == 24 &&
              offsetof(quicked_aligner_t, timer) == 32, "quicked_aligner_t layout");
static_assert(sizeof(profiler_timer_t) == 88, "profiler_timer_t layout");

namespace {

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        const size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { (void)cudaGetLastError(); e = cudaMalloc(&p, bytes); if (e == cudaSuccess) cap = bytes; return e; }
        cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

enum Stage { ST_PREP = 0, ST_WS, ST_WL, ST_BANDED, ST_FILL, ST_TRACE, ST_CIGAR, ST_PLAN, ST_COUNT };

}  // namespace

struct qb200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    size_t matrix_limit = (size_t)48 << 30;

    // uploaded batch
    i64 n_pairs = 0, raw_bytes = 0;
    const unsigned char *d_raw_ext = nullptr;   // upload_device: caller-owned characters
    std::vector<PairRec> h_pairs;
    std::vector<PeqJob> h_peqjobs;
    i64 peq_words = 0, cells = 0;
    DevBuf d_raw, d_codes, d_pairs, d_peq, d_peqjobs;
    // per-run
    DevBuf d_bound, d_hew, d_score, d_status, d_textlen, d_cigoff, d_cigar, d_counters, d_scan_tmp;
    DevBuf d_leaves, d_leafout, d_pairleaves, d_work, d_bandout, d_matrix, d_scores, d_state, d_ops, d_ranges;
    DevBuf d_cls, d_cutoff, d_plan_items, d_plan_offs, d_textbytes, d_list_t, d_list_w, d_list_slow, d_gsize, d_goff, d_gB;
    unsigned char *h_pinned = nullptr;     // small pinned mailbox for totals
    bool unknown_algo = false, multi_leaf_pairs = false;
    std::vector<int> h_score, h_status;
    std::vector<i64> h_cigoff_;
    i64 cigar_total = 0;
    bool have_cigar = false, ran = false;
    qb200_stats_t stats{};
    std::vector<cudaEvent_t> ev_pool;
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> ev_spans;
    size_t ev_used = 0;

    const unsigned char *raw() const { return d_raw_ext ? d_raw_ext : d_raw.as<unsigned char>(); }
};

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            char buf_[512];                                                                           \
            snprintf(buf_, sizeof buf_, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            ctx->err = buf_;                                                                          \
            return e_ == cudaErrorMemoryAllocation ? QB200_ERR_OOM : QB200_ERR_CUDA;                  \
        }                                                                                             \
    } while (0)


// ---- integer-ALU peak microbenchmark (roofline denominator for the bit-op work; MEASURED_PEAKS.json has none) ----
// 8 independent chains per thread of LOP3 + IADD (the instruction mix of a Myers block update), no memory traffic.
__global__ void __launch_bounds__(256) k_int_peak(u32 *out, int iters, u32 a, u32 b, u32 c)
{
    u32 x[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 2654435761u + i * a; y[i] = blockIdx.x + i * b; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                x[i] = ((x[i] & a) | y[i]) ^ c;      // one LOP3
                y[i] = y[i] + x[i];                   // one IADD3
            }
        }
    }
    u32 r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r ^= x[i] + y[i];
    if (r == 0x12345678u) out[0] = r;
}

namespace {

cudaEvent_t new_event(qb200_ctx *ctx)
{
    if (ctx->ev_used == ctx->ev_pool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        ctx->ev_pool.push_back(e);
    }
    return ctx->ev_pool[ctx->ev_used++];
}

struct Span {
    qb200_ctx *ctx; int stage; cudaEvent_t a, b;
    Span(qb200_ctx *c, int st) : ctx(c), stage(st) { a = new_event(c); b = new_event(c); cudaEventRecord(a, c->stream); }
    ~Span() { cudaEventRecord(b, ctx->stream); ctx->ev_spans.push_back({stage, {a, b}}); }
};

int build_pair_records(qb200_ctx *ctx, i64 n, const int64_t *poff, const int32_t *plen, const int64_t *toff,
                       const int32_t *tlen, i64 seqs_bytes)
{
    ctx->h_pairs.resize((size_t)n);
    ctx->h_peqjobs.clear();
    ctx->h_peqjobs.reserve((size_t)n);
    i64 words = 0, cells = 0;
    for (i64 i = 0; i < n; ++i) {
        PairRec &r = ctx->h_pairs[(size_t)i];
        r.p_off = poff[i]; r.t_off = toff[i]; r.m = plen[i]; r.n = tlen[i];
        if (r.m < 0 || r.n < 0 || r.p_off < 0 || r.t_off < 0 || r.p_off + r.m > seqs_bytes || r.t_off + r.n > seqs_bytes) {
            ctx->err = "pair " + std::to_string(i) + ": offsets/lengths outside the packed buffer";
            return QB200_ERR_ARG;
        }
        r.nbp = (r.m + 63) / 64 + 2;
        r.peq_off = words;
        r.pad_ = 0;
        if (r.m > 0 && r.n > 0) {
            PeqJob j; j.src_off = r.p_off; j.m = r.m; j.rev = 0; j.peq_off = words;
            ctx->h_peqjobs.push_back(j);
            words += (i64)kAlpha * r.nbp;
            cells += (i64)r.m * r.n;
        }
    }
    ctx->peq_words = words;
    ctx->cells = cells;
    return 0;
}

int finish_upload(qb200_ctx *ctx)
{
    const i64 n = ctx->n_pairs;
    CK(ctx->d_pairs.reserve(sizeof(PairRec) * (size_t)std::max<i64>(n, 1)));
    CK(cudaMemcpyAsync(ctx->d_pairs.p, ctx->h_pairs.data(), sizeof(PairRec) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    ctx->stats.h2d_bytes += (i64)sizeof(PairRec) * n;
    CK(ctx->d_peqjobs.reserve(sizeof(PeqJob) * std::max<size_t>(ctx->h_peqjobs.size(), 1)));
    CK(cudaMemcpyAsync(ctx->d_peqjobs.p, ctx->h_peqjobs.data(), sizeof(PeqJob) * ctx->h_peqjobs.size(), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->ran = false;
    return 0;
}

template <int R, bool FULL>
int launch_banded_r(qb200_ctx *ctx, const BandTask *d_tasks, const int *d_list, int begin, int n_tasks, i64 mat_sub)
{
    if (n_tasks <= 0) return 0;
    const int bpw = BandedSmem<R>::kBytesPerWarp;
    int wpb = std::max(1, std::min(4, (200 * 1024) / bpw));
    const size_t smem = (size_t)wpb * bpw;
    auto kern = k_banded_warp<R, FULL>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int blocks = (n_tasks + wpb - 1) / wpb;
    kern<<<blocks, wpb * 32, smem, ctx->stream>>>(d_tasks, d_list, begin, n_tasks, mat_sub, ctx->d_codes.as<unsigned char>(),
                                                   ctx->d_peq.as<u64>(), ctx->d_matrix.as<ulonglong2>(), ctx->d_scores.as<int>(),
                                                   ctx->d_state.as<u64>(), ctx->d_ranges.as<int2>(), ctx->d_bandout.as<BandOut>(),
                                                   ctx->d_counters.as<u64>());
    CK(cudaGetLastError());
    ctx->stats.kernel_launches++;
    return 0;
}

int rounds_for(i64 B)
{
    for (int r : {1, 2, 4, 8, 16, 32}) if (B <= 32 * r) return r;
    return 0;
}

// One launch per band-height class present in the list (each warp exits at once if its task belongs to another class).
template <bool FULL>
int launch_banded(qb200_ctx *ctx, unsigned r_mask, const BandTask *d_tasks, const int *d_list, int begin, int n_tasks, i64 mat_sub)
{
    int rc = 0;
    if (!rc && (r_mask & 1)) rc = launch_banded_r<1, FULL>(ctx, d_tasks, d_list, begin, n_tasks, mat_sub);
    if (!rc && (r_mask & 2)) rc = launch_banded_r<2, FULL>(ctx, d_tasks, d_list, begin, n_tasks, mat_sub);
    if (!rc && (r_mask & 4)) rc = launch_banded_r<4, FULL>(ctx, d_tasks, d_list, begin, n_tasks, mat_sub);
    if (!rc && (r_mask & 8)) rc = launch_banded_r<8, FULL>(ctx, d_tasks, d_list, begin, n_tasks, mat_sub);
    if (!rc && (r_mask & 16)) rc = launch_banded_r<16, FULL>(ctx, d_tasks, d_list, begin, n_tasks, mat_sub);
    if (!rc && (r_mask & 32)) rc = launch_banded_r<32, FULL>(ctx, d_tasks, d_list, begin, n_tasks, mat_sub);
    return rc;
}

constexpr int kThreadBandMax = 4;

int launch_thread_fill(qb200_ctx *ctx, const int *d_list, int begin, int n_tasks, i64 mat_sub)
{
    if (n_tasks <= 0) return 0;
    const int T = 128;
    const size_t smem = (size_t)kThreadBandMax * kAlpha * T * 8;
    k_banded_thread<kThreadBandMax><<<(n_tasks + T - 1) / T, T, smem, ctx->stream>>>(
        ctx->d_leaves.as<BandTask>(), d_list, begin, n_tasks, mat_sub, ctx->d_codes.as<unsigned char>(), ctx->d_peq.as<u64>(),
        ctx->d_matrix.as<ulonglong2>(), ctx->d_ranges.as<int2>(), ctx->d_counters.as<u64>());
    CK(cudaGetLastError());
    ctx->stats.kernel_launches++;
    return 0;
}

int launch_traceback(qb200_ctx *ctx, const int *d_list, int begin, int n_tasks, i64 mat_sub)
{
    if (n_tasks <= 0) return 0;
    k_traceback_thread<<<(n_tasks + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_leaves.as<BandTask>(), d_list, begin, n_tasks, mat_sub,
                                                                        ctx->raw(), ctx->d_matrix.as<ulonglong2>(), ctx->d_ranges.as<int2>(),
                                                                        ctx->d_ops.as<u32>(), ctx->d_leafout.as<LeafOut>());
    CK(cudaGetLastError());
    ctx->stats.kernel_launches++;
    return 0;
}

// Largest e in (s, n] such that the entries of items s..e-1 fit `limit` (at least one item).  One thread, binary search.
__global__ void k_chunk_end_groups(const i64 *goff, const i64 *gsize, int n, int s, i64 limit, int *out)
{
    int lo = s + 1, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (goff[mid - 1] + gsize[mid - 1] - goff[s] <= limit) lo = mid; else hi = mid - 1;
    }
    out[0] = lo;
}
__global__ void k_chunk_end_leaves(const BandTask *leaves, const int *list, int n, int s, i64 limit, int *out, i64 *start_off)
{
    const i64 base = leaves[list[s]].mat_off;
    int lo = s + 1, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        const BandTask &t = leaves[list[mid - 1]];
        if (t.mat_off + (i64)(t.n + 1) * t.mat_cs - base <= limit) lo = mid; else hi = mid - 1;
    }
    out[0] = lo; start_off[0] = base;
}

struct RunPlan {
    PlanSum tot;          // totals of the fast path
    i64 mat_t = 0;        // entries of the thread-kernel groups
    int n_groups = 0;
};

}  // namespace

// =====================================================================================================
//                                              C ABI
// =====================================================================================================
extern "C" {

int qb200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    return n;
}

int qb200_create(qb200_ctx_t **out, int device)
{
    if (!out) return QB200_ERR_ARG;
    *out = nullptr;
    if (qb200_device_count() <= device || device < 0) return QB200_ERR_NO_DEVICE;
    if (cudaSetDevice(device) != cudaSuccess) { (void)cudaGetLastError(); return QB200_ERR_NO_DEVICE; }
    qb200_ctx *ctx = new qb200_ctx();
    ctx->device = device;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return QB200_ERR_CUDA; }
    ctx->own_stream = true;
    *out = ctx;
    return 0;
}

void qb200_destroy(qb200_ctx_t *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (DevBuf *b : {&ctx->d_raw, &ctx->d_codes, &ctx->d_pairs, &ctx->d_peq, &ctx->d_peqjobs, &ctx->d_bound, &ctx->d_hew,
                      &ctx->d_score, &ctx->d_status, &ctx->d_textlen, &ctx->d_cigoff, &ctx->d_cigar, &ctx->d_counters,
                      &ctx->d_scan_tmp, &ctx->d_leaves, &ctx->d_leafout, &ctx->d_pairleaves, &ctx->d_work, &ctx->d_bandout,
                      &ctx->d_matrix, &ctx->d_scores, &ctx->d_state, &ctx->d_ops, &ctx->d_ranges, &ctx->d_cls, &ctx->d_cutoff,
                      &ctx->d_plan_items, &ctx->d_plan_offs, &ctx->d_textbytes, &ctx->d_list_t, &ctx->d_list_w, &ctx->d_list_slow,
                      &ctx->d_gsize, &ctx->d_goff, &ctx->d_gB})
        b->release();
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int qb200_set_stream(qb200_ctx_t *ctx, void *cuda_stream)
{
    if (!ctx) return QB200_ERR_ARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
    return 0;
}

int qb200_set_workspace_limit(qb200_ctx_t *ctx, size_t bytes)
{
    if (!ctx || bytes < ((size_t)1 << 20)) return QB200_ERR_ARG;
    ctx->matrix_limit = bytes;
    return 0;
}

const char *qb200_last_error(qb200_ctx_t *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

void *qb200_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    return p;
}
void qb200_host_free(void *p) { if (p) cudaFreeHost(p); }

int qb200_upload(qb200_ctx_t *ctx, const qb200_batch_t *b)
{
    if (!ctx || !b || b->n_pairs < 0 || b->seqs_bytes < 0) return QB200_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    ctx->n_pairs = b->n_pairs; ctx->raw_bytes = b->seqs_bytes; ctx->d_raw_ext = nullptr;
    memset(&ctx->stats, 0, sizeof ctx->stats);
    int rc = build_pair_records(ctx, b->n_pairs, b->pattern_off, b->pattern_len, b->text_off, b->text_len, b->seqs_bytes);
    if (rc) return rc;
    const size_t padded = ((size_t)b->seqs_bytes + 15) / 16 * 16 + 32;
    CK(ctx->d_raw.reserve(padded));
    CK(cudaMemsetAsync(ctx->d_raw.as<char>() + (padded - 48), 0, 48, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_raw.p, b->seqs, (size_t)b->seqs_bytes, cudaMemcpyHostToDevice, ctx->stream));
    ctx->stats.h2d_bytes += b->seqs_bytes;
    return finish_upload(ctx);
}

int qb200_upload_device(qb200_ctx_t *ctx, const qb200_batch_t *b)
{
    if (!ctx || !b || b->n_pairs < 0 || b->seqs_bytes < 0) return QB200_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const i64 n = b->n_pairs;
    std::vector<int64_t> po((size_t)n), to((size_t)n);
    std::vector<int32_t> pl((size_t)n), tl((size_t)n);
    CK(cudaMemcpy(po.data(), b->pattern_off, (size_t)n * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(to.data(), b->text_off, (size_t)n * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(pl.data(), b->pattern_len, (size_t)n * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(tl.data(), b->text_len, (size_t)n * 4, cudaMemcpyDeviceToHost));
    ctx->n_pairs = n; ctx->raw_bytes = b->seqs_bytes;
    ctx->d_raw_ext = reinterpret_cast<const unsigned char *>(b->seqs);
    memset(&ctx->stats, 0, sizeof ctx->stats);
    int rc = build_pair_records(ctx, n, po.data(), pl.data(), to.data(), tl.data(), b->seqs_bytes);
    if (rc) return rc;
    return finish_upload(ctx);
}

// Host-driven slow path (WindowEd(L), band doubling, Hirschberg splits, WINDOWED, BANDED only_score): defined below.
static int run_slow_path(qb200_ctx *ctx, const quicked_params_t &prm, const std::vector<int> &slow_pairs,
                         i64 &n_leaves_total, i64 &ops_words_total, i64 &range_total, std::vector<PairLeaves> &slow_pl);

int qb200_run(qb200_ctx_t *ctx, const quicked_params_t *params)
{
    if (!ctx || !params) return QB200_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const quicked_params_t prm = *params;
    const i64 n = ctx->n_pairs;
    const i64 h2d = ctx->stats.h2d_bytes;
    memset(&ctx->stats, 0, sizeof ctx->stats);
    ctx->stats.h2d_bytes = h2d; ctx->stats.n_pairs = n; ctx->stats.cells = ctx->cells;
    ctx->ev_used = 0; ctx->ev_spans.clear();
    ctx->have_cigar = false; ctx->cigar_total = 0; ctx->unknown_algo = false;
    if (n == 0) { ctx->ran = true; return 0; }
    if (prm.algo != QUICKED && prm.algo != BANDED && prm.algo != WINDOWED && prm.algo != HIRSCHBERG) {
        ctx->unknown_algo = true;                                   // quicked.c:433: every pair -> QUICKED_UNKNOWN_ALGO
        ctx->ran = true;
        return 0;
    }
    const int ni = (int)n;
    const int nb256 = (ni + 255) / 256;
    cudaEvent_t ev_begin = new_event(ctx), ev_end = new_event(ctx);
    CK(cudaEventRecord(ev_begin, ctx->stream));
    CK(ctx->d_counters.reserve(256));
    CK(cudaMemsetAsync(ctx->d_counters.p, 0, 256, ctx->stream));
    CK(ctx->d_bound.reserve((size_t)n * 4));
    CK(ctx->d_hew.reserve((size_t)n * 4));
    CK(ctx->d_score.reserve((size_t)n * 4));
    CK(ctx->d_status.reserve((size_t)n * 4));
    CK(ctx->d_cls.reserve((size_t)n));
    CK(ctx->d_cutoff.reserve((size_t)n * 8));
    CK(ctx->d_plan_items.reserve((size_t)n * sizeof(PlanSum)));
    CK(ctx->d_plan_offs.reserve((size_t)(n + 1) * sizeof(PlanSum)));
    CK(ctx->d_pairleaves.reserve(sizeof(PairLeaves) * (size_t)n));
    CK(ctx->d_textbytes.reserve((size_t)n * 8));
    CK(ctx->d_cigoff.reserve((size_t)(n + 1) * 8));
    CK(ctx->d_list_t.reserve((size_t)n * 4));
    CK(ctx->d_list_w.reserve((size_t)n * 4));
    CK(ctx->d_list_slow.reserve((size_t)n * 4));
    if (!ctx->h_pinned) CK(cudaHostAlloc(&ctx->h_pinned, 4096, cudaHostAllocDefault));

    // ---- prepare: codes + forward match masks ----
    {
        Span sp(ctx, ST_PREP);
        const size_t padded = ((size_t)ctx->raw_bytes + 15) / 16 * 16;
        CK(ctx->d_codes.reserve(padded + 32));
        const i64 nvec = (i64)(padded / 16);
        if (ctx->d_raw_ext && ((uintptr_t)ctx->d_raw_ext & 15)) { ctx->err = "device character buffer must be 16-byte aligned"; return QB200_ERR_ARG; }
        if (ctx->d_raw_ext && (size_t)ctx->raw_bytes != padded) { ctx->err = "device character buffer size must be a multiple of 16"; return QB200_ERR_ARG; }
        if (nvec) {
            const int blocks = (int)std::min<i64>((nvec + 255) / 256, 148 * 16);
            k_encode<<<blocks, 256, 0, ctx->stream>>>(reinterpret_cast<const uint4 *>(ctx->raw()), ctx->d_codes.as<uint4>(), nvec);
            CK(cudaGetLastError());
            ctx->stats.kernel_launches++;
        }
        CK(ctx->d_peq.reserve((size_t)ctx->peq_words * 8 + 64));
        const int nj = (int)ctx->h_peqjobs.size();
        if (nj) {
            k_build_peq<<<(nj + 7) / 8, 256, 0, ctx->stream>>>(ctx->d_peqjobs.as<PeqJob>(), nj, ctx->d_codes.as<unsigned char>(), ctx->d_peq.as<u64>());
            CK(cudaGetLastError());
            ctx->stats.kernel_launches++;
        }
    }

    // ---- QUICKED stage 1: WindowEd(S) bound (quicked.c:178-199) ----
    if (prm.algo == QUICKED) {
        Span sp(ctx, ST_WS);
        const int T = 64;
        const size_t smem = (size_t)kWsSlots * T * 8;
        const int blocks = (int)((n + T - 1) / T);
        if (prm.force_scalar) {
            CK(cudaFuncSetAttribute(k_windowed21_score<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_windowed21_score<false><<<blocks, T, smem, ctx->stream>>>(ctx->d_pairs.as<PairRec>(), ni, ctx->d_codes.as<unsigned char>(), ctx->raw(),
                ctx->d_peq.as<u64>(), (int)prm.hew_threshold[0], ctx->d_bound.as<int>(), ctx->d_hew.as<int>(), ctx->d_counters.as<u64>());
        } else {
            CK(cudaFuncSetAttribute(k_windowed21_score<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_windowed21_score<true><<<blocks, T, smem, ctx->stream>>>(ctx->d_pairs.as<PairRec>(), ni, ctx->d_codes.as<unsigned char>(), ctx->raw(),
                ctx->d_peq.as<u64>(), (int)prm.hew_threshold[0], ctx->d_bound.as<int>(), ctx->d_hew.as<int>(), ctx->d_counters.as<u64>());
        }
        CK(cudaGetLastError());
        ctx->stats.kernel_launches++;
    }

    // ---- plan: classify every pair and lay out the pools with one scan ----
    RunPlan plan;
    {
        Span sp(ctx, ST_PLAN);
        PlanParams pp;
        pp.algo = (int)prm.algo; pp.bandwidth = prm.bandwidth; pp.hew_pct0 = prm.hew_percentage[0];
        pp.only_score = prm.only_score; pp.thread_band_max = kThreadBandMax;
        pp.ok_status = (prm.algo == HIRSCHBERG) ? QUICKED_OK : QUICKED_WIP;
        k_plan<<<nb256, 256, 0, ctx->stream>>>(ctx->d_pairs.as<PairRec>(), ni, pp, ctx->d_bound.as<int>(), ctx->d_hew.as<int>(),
                                               ctx->d_plan_items.as<PlanSum>(), ctx->d_cls.as<unsigned char>(), ctx->d_cutoff.as<i64>(),
                                               ctx->d_status.as<int>(), ctx->d_score.as<int>());
        CK(cudaGetLastError());
        size_t tmp = 0;
        const PlanSum zero = {0, 0, 0, 0, 0, 0, 0, 0};
        CK(cub::DeviceScan::ExclusiveScan(nullptr, tmp, ctx->d_plan_items.as<PlanSum>(), ctx->d_plan_offs.as<PlanSum>(), PlanAdd(), zero, ni, ctx->stream));
        CK(ctx->d_scan_tmp.reserve(tmp + 256));
        CK(cub::DeviceScan::ExclusiveScan(ctx->d_scan_tmp.p, tmp, ctx->d_plan_items.as<PlanSum>(), ctx->d_plan_offs.as<PlanSum>(), PlanAdd(), zero, ni, ctx->stream));
        k_plan_totals<<<1, 32, 0, ctx->stream>>>(ctx->d_plan_items.as<PlanSum>(), ctx->d_plan_offs.as<PlanSum>(), ni, ctx->d_plan_offs.as<PlanSum>() + n);
        CK(cudaGetLastError());
        ctx->stats.kernel_launches += 4;
        CK(cudaMemcpyAsync(ctx->h_pinned, ctx->d_plan_offs.as<PlanSum>() + n, sizeof(PlanSum), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        plan.tot = *reinterpret_cast<PlanSum *>(ctx->h_pinned);
    }
    const PlanSum &tot = plan.tot;

    // ---- slow path first (it appends its leaves after the fast ones and returns its pool usage) ----
    i64 n_leaves = tot.leaf, ops_words = tot.ops, range_ints = tot.rng;
    std::vector<PairLeaves> slow_pl;
    std::vector<int> slow_pairs;
    if (tot.slow > 0) {
        slow_pairs.resize((size_t)tot.slow);
    }
    CK(ctx->d_leaves.reserve(sizeof(BandTask) * (size_t)std::max<i64>(n_leaves, 1)));
    if (tot.leaf > 0 || tot.slow > 0) {
        k_build_leaves<<<nb256, 256, 0, ctx->stream>>>(ctx->d_pairs.as<PairRec>(), ni, ctx->d_cls.as<unsigned char>(), ctx->d_cutoff.as<i64>(),
                                                       ctx->d_plan_offs.as<PlanSum>(), ctx->d_leaves.as<BandTask>(), ctx->d_list_t.as<int>(),
                                                       ctx->d_list_w.as<int>(), ctx->d_list_slow.as<int>(), ctx->d_pairleaves.as<PairLeaves>());
        CK(cudaGetLastError());
        ctx->stats.kernel_launches++;
    }
    if (tot.slow > 0) {
        CK(cudaMemcpyAsync(slow_pairs.data(), ctx->d_list_slow.p, (size_t)tot.slow * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    // thread-kernel groups
    if (tot.t > 0) {
        plan.n_groups = (int)((tot.t + 31) / 32);
        CK(ctx->d_gsize.reserve((size_t)plan.n_groups * 8));
        CK(ctx->d_goff.reserve((size_t)(plan.n_groups + 1) * 8));
        CK(ctx->d_gB.reserve((size_t)plan.n_groups * 4));
        k_group_size<<<(plan.n_groups * 32 + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_list_t.as<int>(), (int)tot.t, ctx->d_leaves.as<BandTask>(),
                                                                                  ctx->d_gsize.as<i64>(), ctx->d_gB.as<int>());
        CK(cudaGetLastError());
        size_t tmp = 0;
        CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp, ctx->d_gsize.as<i64>(), ctx->d_goff.as<i64>(), plan.n_groups, ctx->stream));
        CK(ctx->d_scan_tmp.reserve(tmp + 256));
        CK(cub::DeviceScan::ExclusiveSum(ctx->d_scan_tmp.p, tmp, ctx->d_gsize.as<i64>(), ctx->d_goff.as<i64>(), plan.n_groups, ctx->stream));
        k_last_offset<<<1, 32, 0, ctx->stream>>>(ctx->d_goff.as<i64>(), ctx->d_gsize.as<i64>(), plan.n_groups, reinterpret_cast<i64 *>(ctx->d_counters.as<u64>() + 16),
                                                 ctx->d_goff.as<i64>() + plan.n_groups);
        k_group_assign<<<(int)((tot.t + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_list_t.as<int>(), (int)tot.t, ctx->d_leaves.as<BandTask>(),
                                                                             ctx->d_goff.as<i64>(), ctx->d_gB.as<int>(), tot.matw);
        CK(cudaGetLastError());
        ctx->stats.kernel_launches += 5;
        CK(cudaMemcpyAsync(ctx->h_pinned, ctx->d_counters.as<u64>() + 16, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        plan.mat_t = *reinterpret_cast<i64 *>(ctx->h_pinned);
    }

    // pools of the fast path
    CK(ctx->d_ops.reserve((size_t)std::max<i64>(ops_words, 1) * 4 + 16));
    CK(ctx->d_ranges.reserve((size_t)std::max<i64>(range_ints, 1) * 8 + 16));
    CK(ctx->d_leafout.reserve(sizeof(LeafOut) * (size_t)std::max<i64>(n_leaves, 1)));
    CK(ctx->d_bandout.reserve(sizeof(BandOut) * (size_t)std::max<i64>(n_leaves, 1)));
    if (tot.sc > 0) {
        CK(ctx->d_scores.reserve((size_t)tot.sc * 4 + 16));
        CK(cudaMemsetAsync(ctx->d_scores.p, 0, (size_t)tot.sc * 4, ctx->stream));
    }

    // ---- fast path: fill + traceback, chunked only if the traceback state exceeds the pool ----
    if (tot.leaf > 0) {
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        const i64 limit = (i64)(std::min<size_t>(ctx->matrix_limit, (size_t)((free_b + ctx->d_matrix.cap) * 0.85)) / 16);
        const i64 need = tot.matw + plan.mat_t;
        if (need <= limit) {
            CK(ctx->d_matrix.reserve((size_t)need * 16));
            {
                Span sp(ctx, ST_FILL);
                int rc = launch_thread_fill(ctx, ctx->d_list_t.as<int>(), 0, (int)tot.t, 0);
                if (!rc) rc = launch_banded<true>(ctx, 63u, ctx->d_leaves.as<BandTask>(), ctx->d_list_w.as<int>(), 0, (int)tot.w, 0);
                if (rc) return rc;
            }
            {
                Span sp(ctx, ST_TRACE);
                int rc = launch_traceback(ctx, nullptr, 0, (int)tot.leaf, 0);
                if (rc) return rc;
            }
            ctx->stats.matrix_bytes += need * 16;
        } else {
            CK(ctx->d_matrix.reserve((size_t)limit * 16));
            int *d_idx = reinterpret_cast<int *>(ctx->d_counters.as<u64>() + 20);
            i64 *d_off = reinterpret_cast<i64 *>(ctx->d_counters.as<u64>() + 22);
            // warp-kernel leaves
            for (int s0 = 0; s0 < (int)tot.w;) {
                k_chunk_end_leaves<<<1, 1, 0, ctx->stream>>>(ctx->d_leaves.as<BandTask>(), ctx->d_list_w.as<int>(), (int)tot.w, s0, limit, d_idx, d_off);
                CK(cudaMemcpyAsync(ctx->h_pinned, d_idx, 24, cudaMemcpyDeviceToHost, ctx->stream));
                CK(cudaStreamSynchronize(ctx->stream));
                const int s1 = *reinterpret_cast<int *>(ctx->h_pinned);
                const i64 sub = *reinterpret_cast<i64 *>(ctx->h_pinned + 16);
                { Span sp(ctx, ST_FILL); int rc = launch_banded<true>(ctx, 63u, ctx->d_leaves.as<BandTask>(), ctx->d_list_w.as<int>(), s0, s1 - s0, sub); if (rc) return rc; }
                { Span sp(ctx, ST_TRACE); int rc = launch_traceback(ctx, ctx->d_list_w.as<int>(), s0, s1 - s0, sub); if (rc) return rc; }
                s0 = s1;
            }
            // thread-kernel groups
            for (int g0 = 0; g0 < plan.n_groups;) {
                k_chunk_end_groups<<<1, 1, 0, ctx->stream>>>(ctx->d_goff.as<i64>(), ctx->d_gsize.as<i64>(), plan.n_groups, g0, limit, d_idx);
                CK(cudaMemcpyAsync(ctx->h_pinned, d_idx, 4, cudaMemcpyDeviceToHost, ctx->stream));
                CK(cudaMemcpyAsync(ctx->h_pinned + 16, ctx->d_goff.as<i64>() + g0, 8, cudaMemcpyDeviceToHost, ctx->stream));
                CK(cudaStreamSynchronize(ctx->stream));
                const int g1 = *reinterpret_cast<int *>(ctx->h_pinned);
                const i64 sub = tot.matw + *reinterpret_cast<i64 *>(ctx->h_pinned + 16);
                const int q0 = g0 * 32, q1 = (int)std::min<i64>(tot.t, (i64)g1 * 32);
                { Span sp(ctx, ST_FILL); int rc = launch_thread_fill(ctx, ctx->d_list_t.as<int>(), q0, q1 - q0, sub); if (rc) return rc; }
                { Span sp(ctx, ST_TRACE); int rc = launch_traceback(ctx, ctx->d_list_t.as<int>(), q0, q1 - q0, sub); if (rc) return rc; }
                g0 = g1;
            }
            ctx->stats.matrix_bytes += need * 16;
        }
        ctx->stats.leaves += tot.leaf;
    }

    // ---- slow path ----
    if (tot.slow > 0) {
        int rc = run_slow_path(ctx, prm, slow_pairs, n_leaves, ops_words, range_ints, slow_pl);
        if (rc) return rc;
    }

    // ---- scores + CIGAR text ----
    const bool want_cigar = !prm.only_score;
    {
        Span sp(ctx, ST_CIGAR);
        k_pair_finish<<<nb256, 256, 0, ctx->stream>>>(ctx->d_pairleaves.as<PairLeaves>(), ni, ctx->d_leafout.as<LeafOut>(), ctx->d_status.as<int>(),
                                                      ctx->d_score.as<int>(), ctx->d_textbytes.as<i64>(), want_cigar ? 1 : 0);
        CK(cudaGetLastError());
        ctx->stats.kernel_launches++;
        if (want_cigar) {
            if (ctx->multi_leaf_pairs) {
                CK(ctx->d_textlen.reserve((size_t)n * 4));
                k_cigar_text<false><<<(ni + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_pairleaves.as<PairLeaves>(), ni, ctx->d_leaves.as<BandTask>(),
                    ctx->d_leafout.as<LeafOut>(), ctx->d_ops.as<u32>(), ctx->d_textlen.as<int>(), nullptr, nullptr);
                k_merge_text_len<<<nb256, 256, 0, ctx->stream>>>(ctx->d_textlen.as<int>(), ctx->d_textbytes.as<i64>(), ni);
                CK(cudaGetLastError());
                ctx->stats.kernel_launches += 2;
            }
            size_t tmp = 0;
            CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp, ctx->d_textbytes.as<i64>(), ctx->d_cigoff.as<i64>(), ni, ctx->stream));
            CK(ctx->d_scan_tmp.reserve(tmp + 256));
            CK(cub::DeviceScan::ExclusiveSum(ctx->d_scan_tmp.p, tmp, ctx->d_textbytes.as<i64>(), ctx->d_cigoff.as<i64>(), ni, ctx->stream));
            k_last_offset<<<1, 32, 0, ctx->stream>>>(ctx->d_cigoff.as<i64>(), ctx->d_textbytes.as<i64>(), ni, reinterpret_cast<i64 *>(ctx->d_counters.as<u64>() + 24),
                                                     ctx->d_cigoff.as<i64>() + n);
            CK(cudaGetLastError());
            ctx->stats.kernel_launches += 3;
            CK(cudaMemcpyAsync(ctx->h_pinned, ctx->d_counters.as<u64>() + 24, 8, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            ctx->cigar_total = *reinterpret_cast<i64 *>(ctx->h_pinned);
            CK(ctx->d_cigar.reserve((size_t)ctx->cigar_total + 16));
            CK(cudaMemsetAsync(ctx->d_cigar.p, 0, (size_t)ctx->cigar_total, ctx->stream));
            if (n_leaves > 0) {
                k_cigar_text<true><<<(ni + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_pairleaves.as<PairLeaves>(), ni, ctx->d_leaves.as<BandTask>(),
                    ctx->d_leafout.as<LeafOut>(), ctx->d_ops.as<u32>(), nullptr, ctx->d_cigoff.as<i64>(), ctx->d_cigar.as<char>());
                CK(cudaGetLastError());
                ctx->stats.kernel_launches++;
            }
            ctx->have_cigar = true;
        }
    }
    CK(cudaEventRecord(ev_end, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));

    // ---- stats ----
    u64 counters[4];
    CK(cudaMemcpy(counters, ctx->d_counters.p, 32, cudaMemcpyDeviceToHost));
    ctx->stats.word_steps_windowed = (i64)counters[0];
    ctx->stats.word_steps_banded = (i64)counters[1];
    ctx->stats.word_steps = (i64)(counters[0] + counters[1]);
    float ms = 0;
    cudaEventElapsedTime(&ms, ev_begin, ev_end);
    ctx->stats.ms_total = ms;
    float st[ST_COUNT] = {0};
    for (auto &sp : ctx->ev_spans) { float t = 0; cudaEventElapsedTime(&t, sp.second.first, sp.second.second); st[sp.first] += t; }
    ctx->stats.ms_prepare = st[ST_PREP] + st[ST_PLAN]; ctx->stats.ms_windowed_s = st[ST_WS]; ctx->stats.ms_windowed_l = st[ST_WL];
    ctx->stats.ms_banded = st[ST_BANDED]; ctx->stats.ms_align_fill = st[ST_FILL]; ctx->stats.ms_align_trace = st[ST_TRACE];
    ctx->stats.ms_cigar = st[ST_CIGAR];
    ctx->ran = true;
    return 0;
}

static int run_slow_path(qb200_ctx *ctx, const quicked_params_t &prm, const std::vector<int> &slow_pairs,
                         i64 &n_leaves_total, i64 &ops_words_total, i64 &range_total, std::vector<PairLeaves> &slow_pl)
{
    // Not implemented yet: the pairs keep QUICKED_UNIMPLEMENTED so the caller sees a loud, per-pair error.
    (void)prm; (void)n_leaves_total; (void)ops_words_total; (void)range_total; (void)slow_pl;
    std::vector<int> st(slow_pairs.size(), (int)QUICKED_UNIMPLEMENTED);
    for (size_t q = 0; q < slow_pairs.size(); ++q)
        CK(cudaMemcpyAsync(ctx->d_status.as<int>() + slow_pairs[q], &st[q], 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->stats.pairs_stage2 += (i64)slow_pairs.size();
    return 0;
}

int qb200_download(qb200_ctx_t *ctx, qb200_results_t *res)
{
    if (!ctx || !res || !ctx->ran) return QB200_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const i64 n = ctx->n_pairs;
    res->cigar_bytes = 0;
    if (ctx->unknown_algo) {
        for (i64 i = 0; i < n; ++i) { if (res->score) res->score[i] = -1; if (res->status) res->status[i] = QUICKED_UNKNOWN_ALGO; }
        if (res->cigar_off) for (i64 i = 0; i <= n; ++i) res->cigar_off[i] = 0;
        return 0;
    }
    if (n == 0) { if (res->cigar_off) res->cigar_off[0] = 0; return 0; }
    if (res->score) CK(cudaMemcpyAsync(res->score, ctx->d_score.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (res->status) CK(cudaMemcpyAsync(res->status, ctx->d_status.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->stats.d2h_bytes += n * 8;
    int rc = 0;
    if (ctx->have_cigar && res->cigar_off) {
        CK(cudaMemcpyAsync(res->cigar_off, ctx->d_cigoff.p, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
        ctx->stats.d2h_bytes += (n + 1) * 8;
        res->cigar_bytes = ctx->cigar_total;
        if (!res->cigar || res->cigar_capacity < ctx->cigar_total) rc = QB200_ERR_CAPACITY;
        else {
            CK(cudaMemcpyAsync(res->cigar, ctx->d_cigar.p, (size_t)ctx->cigar_total, cudaMemcpyDeviceToHost, ctx->stream));
            ctx->stats.d2h_bytes += ctx->cigar_total;
        }
    } else if (res->cigar_off) {
        for (i64 i = 0; i <= n; ++i) res->cigar_off[i] = 0;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return rc;
}

int qb200_get_stats(qb200_ctx_t *ctx, qb200_stats_t *stats)
{
    if (!ctx || !stats) return QB200_ERR_ARG;
    *stats = ctx->stats;
    return 0;
}

int qb200_measure_int_peak(qb200_ctx_t *ctx, double *tera_ops_per_s)
{
    if (!ctx || !tera_ops_per_s) return QB200_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(ctx->d_counters.reserve(64));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    const int blocks = sms * 8, iters = 4096;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(a, ctx->stream));
        k_int_peak<<<blocks, 256, 0, ctx->stream>>>(ctx->d_counters.as<u32>() + 8, iters, 0x5bd1e995u + rep, 0x9e3779b9u, 0x85ebca6bu);
        CK(cudaEventRecord(b, ctx->stream));
        CK(cudaEventSynchronize(b));
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        const double ops = (double)blocks * 256.0 * iters * 8.0 * 8.0 * 2.0;
        if (rep > 0) best = std::max(best, ops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    *tera_ops_per_s = best;
    return 0;
}

int qb200_align_batch(qb200_ctx_t *ctx, const quicked_params_t *params, const qb200_batch_t *b, qb200_results_t *res)
{
    int rc = qb200_upload(ctx, b);
    if (rc) return rc;
    rc = qb200_run(ctx, params);
    if (rc) return rc;
    return qb200_download(ctx, res);
}

}  // extern "C"
