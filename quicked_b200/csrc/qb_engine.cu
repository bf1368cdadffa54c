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
#include "qb_banded.cuh"
#include "qb_common.cuh"
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

enum Stage { ST_PREP = 0, ST_WS, ST_WL, ST_BANDED, ST_FILL, ST_TRACE, ST_CIGAR, ST_COUNT };

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
int launch_banded_r(qb200_ctx *ctx, const BandTask *d_tasks, int n_tasks)
{
    if (n_tasks <= 0) return 0;
    const int bpw = BandedSmem<R>::kBytesPerWarp;
    int wpb = std::max(1, std::min(4, (200 * 1024) / bpw));
    const size_t smem = (size_t)wpb * bpw;
    auto kern = k_banded_warp<R, FULL>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int blocks = (n_tasks + wpb - 1) / wpb;
    kern<<<blocks, wpb * 32, smem, ctx->stream>>>(d_tasks, n_tasks, ctx->d_codes.as<unsigned char>(), ctx->d_peq.as<u64>(),
                                                   ctx->d_matrix.as<ulonglong2>(), ctx->d_scores.as<int>(),
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

template <bool FULL>
int launch_banded(qb200_ctx *ctx, int R, const BandTask *d_tasks, int n_tasks)
{
    switch (R) {
    case 1: return launch_banded_r<1, FULL>(ctx, d_tasks, n_tasks);
    case 2: return launch_banded_r<2, FULL>(ctx, d_tasks, n_tasks);
    case 4: return launch_banded_r<4, FULL>(ctx, d_tasks, n_tasks);
    case 8: return launch_banded_r<8, FULL>(ctx, d_tasks, n_tasks);
    case 16: return launch_banded_r<16, FULL>(ctx, d_tasks, n_tasks);
    case 32: return launch_banded_r<32, FULL>(ctx, d_tasks, n_tasks);
    }
    ctx->err = "band too tall for the implemented kernels";
    return QB200_ERR_ARG;
}

// ---- leaves: BandEd full matrix + traceback for a list of tasks (pair order), chunked by the matrix pool ----
// h_leaves[i].slot must equal i.  On return d_leaves / d_leafout hold all leaves and their results.
int run_leaves(qb200_ctx *ctx, std::vector<BandTask> &h_leaves)
{
    const size_t nl = h_leaves.size();
    if (!nl) return 0;
    // op regions + per-task geometry
    i64 ops_words = 0;
    std::vector<int> rounds(nl);
    std::vector<i64> mat_entries(nl), score_ints(nl);
    for (size_t i = 0; i < nl; ++i) {
        BandTask &t = h_leaves[i];
        const BandGeom g = band_geometry(t.m, t.n, t.cutoff);
        rounds[i] = rounds_for(g.Bc);
        if (!rounds[i]) { ctx->err = "leaf band of " + std::to_string(g.Bc) + " blocks exceeds 1024"; return QB200_ERR_ARG; }
        mat_entries[i] = (i64)(t.n + 1) * g.Bc;
        score_ints[i] = (i64)((t.m + 63) / 64) + g.Bc + 2;
        t.ops_cap = ((t.m + t.n + 15) / 16) * 16;
        t.ops_off = ops_words;
        ops_words += t.ops_cap / 16;
        t.slot = (int)i;
    }
    CK(ctx->d_ops.reserve((size_t)ops_words * 4 + 16));
    CK(ctx->d_leaves.reserve(sizeof(BandTask) * nl));
    CK(ctx->d_leafout.reserve(sizeof(LeafOut) * nl));
    CK(ctx->d_bandout.reserve(sizeof(BandOut) * nl));
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    const i64 limit_entries = (i64)(std::min<size_t>(ctx->matrix_limit, (size_t)((free_b + ctx->d_matrix.cap) * 0.85)) / 16);

    size_t i0 = 0;
    std::vector<BandTask> work;
    while (i0 < nl) {
        i64 ent = 0, sc = 0, rg = 0;
        size_t i1 = i0;
        while (i1 < nl && (i1 == i0 || ent + mat_entries[i1] <= limit_entries)) { ent += mat_entries[i1]; sc += score_ints[i1]; ++i1; }
        if (ent > limit_entries && i1 == i0 + 1 && (size_t)ent * 16 > free_b + ctx->d_matrix.cap) {
            ctx->err = "a single traceback matrix does not fit the device"; return QB200_ERR_OOM;
        }
        // assign pool offsets, group by rounds
        work.clear();
        work.reserve(i1 - i0);
        i64 mo = 0, so = 0;
        for (size_t i = i0; i < i1; ++i) {
            h_leaves[i].mat_off = mo; h_leaves[i].scores_off = so; h_leaves[i].range_off = rg;
            mo += mat_entries[i]; so += score_ints[i]; rg += h_leaves[i].n / 64 + 2;
        }
        int group_begin[7] = {0}, gi = 0;
        const int Rs[6] = {1, 2, 4, 8, 16, 32};
        for (int r = 0; r < 6; ++r) {
            group_begin[gi++] = (int)work.size();
            for (size_t i = i0; i < i1; ++i) if (rounds[i] == Rs[r]) work.push_back(h_leaves[i]);
        }
        group_begin[6] = (int)work.size();
        CK(ctx->d_matrix.reserve((size_t)ent * 16));
        CK(ctx->d_scores.reserve((size_t)sc * 4 + 16));
        CK(ctx->d_ranges.reserve((size_t)rg * 8 + 16));
        CK(ctx->d_work.reserve(sizeof(BandTask) * work.size()));
        CK(cudaMemcpyAsync(ctx->d_work.p, work.data(), sizeof(BandTask) * work.size(), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemsetAsync(ctx->d_scores.p, 0, (size_t)sc * 4, ctx->stream));
        {
            Span sp(ctx, ST_FILL);
            for (int r = 0; r < 6; ++r) {
                const int nb = group_begin[r + 1] - group_begin[r];
                int rc = launch_banded<true>(ctx, Rs[r], ctx->d_work.as<BandTask>() + group_begin[r], nb);
                if (rc) return rc;
            }
        }
        {
            Span sp(ctx, ST_TRACE);
            const int nt = (int)work.size();
            k_traceback_thread<<<(nt + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_work.as<BandTask>(), nt, ctx->raw(),
                                                                          ctx->d_matrix.as<ulonglong2>(), ctx->d_ranges.as<int2>(),
                                                                          ctx->d_ops.as<u32>(), ctx->d_leafout.as<LeafOut>());
            CK(cudaGetLastError());
            ctx->stats.kernel_launches++;
        }
        ctx->stats.matrix_bytes += ent * 16;
        // the work list is reused by the next chunk: wait for this one (also bounds the pool lifetime)
        CK(cudaStreamSynchronize(ctx->stream));
        i0 = i1;
    }
    ctx->stats.leaves += (i64)nl;
    CK(cudaMemcpyAsync(ctx->d_leaves.p, h_leaves.data(), sizeof(BandTask) * nl, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

// ---- CIGAR text + scores for pairs whose leaves are done ----
int emit_results(qb200_ctx *ctx, const std::vector<PairLeaves> &h_pl, bool want_cigar)
{
    const i64 n = ctx->n_pairs;
    Span sp(ctx, ST_CIGAR);
    CK(ctx->d_pairleaves.reserve(sizeof(PairLeaves) * (size_t)n));
    CK(cudaMemcpyAsync(ctx->d_pairleaves.p, h_pl.data(), sizeof(PairLeaves) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    const int blocks = (int)((n + 127) / 128);
    // pass 1: text bytes per pair (also needed for the offsets when only the score is wanted: cheap, skip then)
    CK(ctx->d_textlen.reserve((size_t)(n + 1) * 4));
    CK(ctx->d_cigoff.reserve((size_t)(n + 1) * 8));
    if (want_cigar) {
        k_cigar_text<false><<<blocks, 128, 0, ctx->stream>>>(ctx->d_pairleaves.as<PairLeaves>(), (int)n, ctx->d_leaves.as<BandTask>(),
                                                             ctx->d_leafout.as<LeafOut>(), ctx->d_ops.as<u32>(), ctx->d_textlen.as<int>(),
                                                             nullptr, nullptr);
        CK(cudaGetLastError());
        ctx->stats.kernel_launches++;
    }
    return 0;
}

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
                      &ctx->d_matrix, &ctx->d_scores, &ctx->d_state, &ctx->d_ops, &ctx->d_ranges})
        b->release();
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

int qb200_run(qb200_ctx_t *ctx, const quicked_params_t *params)
{
    if (!ctx || !params) return QB200_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const quicked_params_t prm = *params;
    const i64 n = ctx->n_pairs;
    // reset per-run stats (keep upload byte counts)
    const i64 h2d = ctx->stats.h2d_bytes;
    memset(&ctx->stats, 0, sizeof ctx->stats);
    ctx->stats.h2d_bytes = h2d; ctx->stats.n_pairs = n; ctx->stats.cells = ctx->cells;
    ctx->ev_used = 0; ctx->ev_spans.clear();
    ctx->h_score.assign((size_t)n, -1);
    ctx->h_status.assign((size_t)n, QUICKED_ERROR);
    ctx->have_cigar = false; ctx->cigar_total = 0;
    if (n == 0) { ctx->ran = true; return 0; }
    if (prm.algo != QUICKED && prm.algo != BANDED && prm.algo != WINDOWED && prm.algo != HIRSCHBERG) {
        std::fill(ctx->h_status.begin(), ctx->h_status.end(), (int)QUICKED_UNKNOWN_ALGO);   // quicked.c:433
        ctx->ran = true;
        return 0;
    }
    cudaEvent_t ev_begin = new_event(ctx), ev_end = new_event(ctx);
    CK(cudaEventRecord(ev_begin, ctx->stream));
    CK(ctx->d_counters.reserve(64));
    CK(cudaMemsetAsync(ctx->d_counters.p, 0, 64, ctx->stream));

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

    std::vector<i64> cutoff((size_t)n, 0);
    std::vector<char> valid((size_t)n, 0);
    for (i64 i = 0; i < n; ++i) {
        const PairRec &r = ctx->h_pairs[(size_t)i];
        valid[(size_t)i] = (r.m > 0 && r.n > 0);
        if (!valid[(size_t)i]) ctx->h_status[(size_t)i] = QUICKED_EMPTY_SEQUENCE;          // quicked.c:411-414
    }
    const bool want_cigar = !prm.only_score;
    int ok_status = QUICKED_WIP;

    if (prm.algo == QUICKED) {
        // ---- stage 1: WindowEd(S) bound (quicked.c:178-199) ----
        CK(ctx->d_bound.reserve((size_t)n * 4));
        CK(ctx->d_hew.reserve((size_t)n * 4));
        {
            Span sp(ctx, ST_WS);
            const int T = 64;
            const size_t smem = (size_t)kWsSlots * T * 8;
            const int blocks = (int)((n + T - 1) / T);
            if (prm.force_scalar) {
                CK(cudaFuncSetAttribute(k_windowed21_score<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                k_windowed21_score<false><<<blocks, T, smem, ctx->stream>>>(ctx->d_pairs.as<PairRec>(), (int)n, ctx->d_codes.as<unsigned char>(), ctx->raw(),
                    ctx->d_peq.as<u64>(), (int)prm.hew_threshold[0], ctx->d_bound.as<int>(), ctx->d_hew.as<int>(), ctx->d_counters.as<u64>());
            } else {
                CK(cudaFuncSetAttribute(k_windowed21_score<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                k_windowed21_score<true><<<blocks, T, smem, ctx->stream>>>(ctx->d_pairs.as<PairRec>(), (int)n, ctx->d_codes.as<unsigned char>(), ctx->raw(),
                    ctx->d_peq.as<u64>(), (int)prm.hew_threshold[0], ctx->d_bound.as<int>(), ctx->d_hew.as<int>(), ctx->d_counters.as<u64>());
            }
            CK(cudaGetLastError());
            ctx->stats.kernel_launches++;
        }
        std::vector<int> h_bound((size_t)n), h_hew((size_t)n);
        CK(cudaMemcpyAsync(h_bound.data(), ctx->d_bound.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(h_hew.data(), ctx->d_hew.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (i64 i = 0; i < n; ++i) {
            if (!valid[(size_t)i]) continue;
            const PairRec &r = ctx->h_pairs[(size_t)i];
            const unsigned maxlen = (unsigned)std::max(r.m, r.n);
            cutoff[(size_t)i] = h_bound[(size_t)i];
            if ((i64)h_hew[(size_t)i] * 64 > (i64)(maxlen * prm.hew_percentage[0] / 100)) {   // quicked.c:201-202
                ctx->stats.pairs_stage2++;
                ctx->h_status[(size_t)i] = QUICKED_UNIMPLEMENTED;   // stages 2-3 arrive with the generic WindowEd kernel
                valid[(size_t)i] = 0;
            }
        }
    } else if (prm.algo == BANDED || prm.algo == HIRSCHBERG) {
        for (i64 i = 0; i < n; ++i) {
            const PairRec &r = ctx->h_pairs[(size_t)i];
            cutoff[(size_t)i] = (i64)((unsigned)std::max(r.m, r.n) * prm.bandwidth / 100);        // quicked.c:64, :131
        }
        if (prm.algo == HIRSCHBERG) ok_status = QUICKED_OK;
    } else {
        for (i64 i = 0; i < n; ++i) if (valid[(size_t)i]) { ctx->h_status[(size_t)i] = QUICKED_UNIMPLEMENTED; valid[(size_t)i] = 0; }
    }

    // ---- alignment: leaves (no Hirschberg split yet) ----
    std::vector<BandTask> leaves;
    std::vector<PairLeaves> pl((size_t)n);
    leaves.reserve((size_t)n);
    for (i64 i = 0; i < n; ++i) {
        pl[(size_t)i].first_leaf = (i64)leaves.size(); pl[(size_t)i].n_leaves = 0; pl[(size_t)i].pad_ = 0;
        if (!valid[(size_t)i]) continue;
        const PairRec &r = ctx->h_pairs[(size_t)i];
        const BandGeom g = band_geometry(r.m, r.n, cutoff[(size_t)i]);
        const bool split = (prm.algo != BANDED) && ((unsigned long long)g.Bc * (unsigned long long)r.n * 16ull > (1ull << 24));
        if (split || (prm.algo == BANDED && prm.only_score)) {
            ctx->h_status[(size_t)i] = QUICKED_UNIMPLEMENTED;
            continue;
        }
        BandTask t{};
        t.p_off = r.p_off; t.t_off = r.t_off; t.m = r.m; t.n = r.n; t.rev = 0; t.finish = r.n;
        t.cutoff = cutoff[(size_t)i]; t.peq_off = r.peq_off; t.nbp = r.nbp; t.pair = (int)i;
        leaves.push_back(t);
        pl[(size_t)i].n_leaves = 1;
        ctx->h_status[(size_t)i] = ok_status;
    }
    int rc = run_leaves(ctx, leaves);
    if (rc) return rc;

    // ---- scores + CIGAR text ----
    if (!leaves.empty()) {
        std::vector<LeafOut> lo(leaves.size());
        rc = emit_results(ctx, pl, want_cigar);
        if (rc) return rc;
        CK(cudaMemcpyAsync(lo.data(), ctx->d_leafout.p, sizeof(LeafOut) * lo.size(), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (i64 i = 0; i < n; ++i) {
            const PairLeaves &p = pl[(size_t)i];
            if (!p.n_leaves) continue;
            int s = 0;
            for (int l = 0; l < p.n_leaves; ++l) s += lo[(size_t)p.first_leaf + l].cost;
            ctx->h_score[(size_t)i] = s;                                   // cigar_score_edit, quicked.c:54
        }
    }
    if (want_cigar) {
        // offsets: exclusive scan of (text_len + 1); pairs without leaves get an empty string
        Span sp(ctx, ST_CIGAR);
        std::vector<int> tl((size_t)n + 1, 0);
        if (!leaves.empty()) CK(cudaMemcpyAsync(tl.data(), ctx->d_textlen.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        std::vector<i64> off((size_t)n + 1);
        i64 acc = 0;
        for (i64 i = 0; i < n; ++i) { off[(size_t)i] = acc; acc += (pl[(size_t)i].n_leaves ? tl[(size_t)i] : 0) + 1; }
        off[(size_t)n] = acc;
        ctx->cigar_total = acc;
        CK(ctx->d_cigar.reserve((size_t)acc + 16));
        CK(cudaMemsetAsync(ctx->d_cigar.p, 0, (size_t)acc, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_cigoff.p, off.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        if (!leaves.empty()) {
            k_cigar_text<true><<<(int)((n + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_pairleaves.as<PairLeaves>(), (int)n, ctx->d_leaves.as<BandTask>(),
                ctx->d_leafout.as<LeafOut>(), ctx->d_ops.as<u32>(), nullptr, ctx->d_cigoff.as<i64>(), ctx->d_cigar.as<char>());
            CK(cudaGetLastError());
            ctx->stats.kernel_launches++;
        }
        ctx->have_cigar = true;
        ctx->h_cigoff_.swap(off);   // host copy of the offsets for qb200_download
    }
    CK(cudaEventRecord(ev_end, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));

    // ---- stats ----
    u64 counters[8];
    CK(cudaMemcpy(counters, ctx->d_counters.p, 64, cudaMemcpyDeviceToHost));
    ctx->stats.word_steps_windowed = (i64)counters[0];
    ctx->stats.word_steps_banded = (i64)counters[1];
    ctx->stats.word_steps = (i64)(counters[0] + counters[1]);
    float ms = 0;
    cudaEventElapsedTime(&ms, ev_begin, ev_end);
    ctx->stats.ms_total = ms;
    float st[ST_COUNT] = {0};
    for (auto &s : ctx->ev_spans) { float t = 0; cudaEventElapsedTime(&t, s.second.first, s.second.second); st[s.first] += t; }
    ctx->stats.ms_prepare = st[ST_PREP]; ctx->stats.ms_windowed_s = st[ST_WS]; ctx->stats.ms_windowed_l = st[ST_WL];
    ctx->stats.ms_banded = st[ST_BANDED]; ctx->stats.ms_align_fill = st[ST_FILL]; ctx->stats.ms_align_trace = st[ST_TRACE];
    ctx->stats.ms_cigar = st[ST_CIGAR];
    ctx->ran = true;
    return 0;
}

int qb200_download(qb200_ctx_t *ctx, qb200_results_t *res)
{
    if (!ctx || !res || !ctx->ran) return QB200_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const i64 n = ctx->n_pairs;
    if (res->score) memcpy(res->score, ctx->h_score.data(), (size_t)n * 4);
    if (res->status) memcpy(res->status, ctx->h_status.data(), (size_t)n * 4);
    ctx->stats.d2h_bytes += n * 8;
    res->cigar_bytes = 0;
    if (ctx->have_cigar && res->cigar_off) {
        memcpy(res->cigar_off, ctx->h_cigoff_.data(), (size_t)(n + 1) * 8);
        res->cigar_bytes = ctx->cigar_total;
        if (!res->cigar || res->cigar_capacity < ctx->cigar_total) return QB200_ERR_CAPACITY;
        CK(cudaMemcpyAsync(res->cigar, ctx->d_cigar.p, (size_t)ctx->cigar_total, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->stats.d2h_bytes += ctx->cigar_total;
    } else if (res->cigar_off) {
        for (i64 i = 0; i <= n; ++i) res->cigar_off[i] = 0;
    }
    return 0;
}

int qb200_get_stats(qb200_ctx_t *ctx, qb200_stats_t *stats)
{
    if (!ctx || !stats) return QB200_ERR_ARG;
    *stats = ctx->stats;
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
