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
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "../../include/quicked_b200.h"
#include <cub/device/device_scan.cuh>

#include "qb_banded.cuh"
#include "qb_common.cuh"
#include "qb_fused.cuh"
#include "qb_generate.cuh"
#include "qb_hirschberg.cuh"
#include "qb_plan.cuh"
#include "qb_prep.cuh"
#include "qb_traceback.cuh"
#include "qb_tiles.cuh"
#include "qb_tiletrace.cuh"
#include "qb_windowed.cuh"
#include "qb_wintile.cuh"

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
    cudaError_t grow_keep(size_t bytes, size_t keep, cudaStream_t st)   // like reserve() but the first `keep` bytes survive
    {
        if (bytes <= cap) return cudaSuccess;
        void *q = nullptr;
        const size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&q, want);
        if (e != cudaSuccess) return e;
        if (p && keep) { e = cudaMemcpyAsync(q, p, keep, cudaMemcpyDeviceToDevice, st); if (e == cudaSuccess) e = cudaStreamSynchronize(st); }
        if (p) cudaFree(p);
        p = q; cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

// Growable page-locked host array (contents are not kept across a growth).
template <class T> struct PinnedArray {
    T *p = nullptr;
    size_t cap = 0, n = 0;
    bool resize(size_t count)
    {
        if (count > cap) {
            if (p) cudaFreeHost(p);
            p = nullptr; cap = 0;
            const size_t want = count + count / 8 + 64;
            if (cudaHostAlloc(reinterpret_cast<void **>(&p), want * sizeof(T), cudaHostAllocDefault) != cudaSuccess) { (void)cudaGetLastError(); return false; }
            cap = want;
        }
        n = count;
        return true;
    }
    T *data() { return p; }
    size_t size() const { return n; }
    T &operator[](size_t i) { return p[i]; }
    const T &operator[](size_t i) const { return p[i]; }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = n = 0; }
};

constexpr i64 kMaxPairs = (1ll << 31) - (1ll << 20);   // pair / leaf indices, grid sizes and list offsets are 32-bit

enum Stage { ST_PREP = 0, ST_WS, ST_WL, ST_BANDED, ST_FILL, ST_TRACE, ST_CIGAR, ST_PLAN, ST_FUSED, ST_COUNT };

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
    PinnedArray<PairRec> h_pairs;          // pinned: its H2D copy must not hold the driver lock of a pageable copy
    PinnedArray<unsigned char> h_stage;    // pinned staging of results bound for pageable caller memory
    i64 peq_words = 0, cells = 0;
    DevBuf d_raw, d_codes, d_pairs, d_peq, d_peqjobs, d_pairodd;
    // per-run
    DevBuf d_bound, d_hew, d_score, d_status, d_textlen, d_cigoff, d_cigar, d_counters, d_scan_tmp;
    DevBuf d_leaves, d_leafout, d_pairleaves, d_work, d_bandout, d_matrix, d_scores, d_state, d_ops, d_ranges;
    DevBuf d_cls, d_cutoff, d_plan_items, d_plan_offs, d_textbytes, d_list_t, d_list_w, d_list_slow, d_gsize, d_goff, d_gB;
    unsigned char *h_pinned = nullptr;     // small pinned mailbox for totals
    DevBuf d_quad;                         // WindowEd(S) quadrant scratch
    DevBuf d_packed, d_excpos, d_excchr;   // 2-bit packed upload: the stream and its exception list
    DevBuf d_fmat, d_franges, d_done;      // fused fast path: per-resident-warp matrix slots, live ranges, per-pair done flags
    i64 ops_words_fused = 0;               // op words of all pairs (fused-path region of the op pool)
    int max_n = 0, max_m = 0;
    bool ws_compact_set[2] = {false, false};
    int ws_carve_set[2][2] = {{-1, -1}, {-1, -1}};   // shared-memory carve-out already requested for each WindowEd(S) kernel variant
    int sms = 0;                           // SM count of the device (queried once)
    DevBuf d_tclass, d_tctl, d_punt, d_gather, d_ttext, d_wintile;   // tile path: per-class task lists, counters, punted tasks
    bool use_tiles = true;
    int thread_band_max = 4;               // leaves with B_cigar <= this use the thread-per-leaf full-matrix kernels (0: everything through the tile kernels)
    DevBuf d_peq2, d_jobs2, d_tasks2, d_wintasks, d_winout, d_winscratch, d_split, d_splitout, d_splitscratch, d_scatter;
    bool unknown_algo = false, multi_leaf_pairs = false, tile_walks = false;   // tile_walks: some leaf's text length is still to be measured
    static constexpr int kWorkers = 8;
    qb200_ctx *child[kWorkers] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // workers of the pipelined qb200_align_batch
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
    if (!ctx->h_pairs.resize((size_t)n)) { ctx->err = "out of pinned host memory"; return QB200_ERR_OOM; }
    i64 words = 0, cells = 0, ops_words = 0;
    int max_n = 0, max_m = 0;
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
        r.ops_off = ops_words;
        if (r.m > 0 && r.n > 0) {
            ops_words += (r.m + r.n + 15) / 16;
            max_n = std::max(max_n, r.n); max_m = std::max(max_m, r.m);
            words += (i64)kPeqStride * r.nbp;
            cells += (i64)r.m * r.n;
        }
    }
    ctx->peq_words = words;
    ctx->cells = cells;
    ctx->ops_words_fused = ops_words;
    ctx->max_n = max_n; ctx->max_m = max_m;
    return 0;
}

int finish_upload(qb200_ctx *ctx)
{
    const i64 n = ctx->n_pairs;
    CK(ctx->d_pairs.reserve(sizeof(PairRec) * (size_t)std::max<i64>(n, 1)));
    CK(cudaMemcpyAsync(ctx->d_pairs.p, ctx->h_pairs.data(), sizeof(PairRec) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    ctx->stats.h2d_bytes += (i64)sizeof(PairRec) * n;
    CK(ctx->d_peqjobs.reserve(sizeof(PeqJob) * (size_t)std::max<i64>(n, 1)));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->ran = false;
    return 0;
}

template <int R, bool FULL>
int launch_banded_r(qb200_ctx *ctx, const BandTask *d_tasks, const int *d_list, int begin, int n_tasks, i64 mat_sub, const u64 *peq_base, int min_B)
{
    if (n_tasks <= 0) return 0;
    const int bpw = BandedSmem<R>::kBytesPerWarp;
    int wpb = std::max(1, std::min(4, (200 * 1024) / bpw));
    const size_t smem = (size_t)wpb * bpw;
    auto kern = k_banded_warp<R, FULL>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int blocks = (n_tasks + wpb - 1) / wpb;
    kern<<<blocks, wpb * 32, smem, ctx->stream>>>(d_tasks, d_list, begin, n_tasks, mat_sub, ctx->d_codes.as<unsigned char>(),
                                                   peq_base, ctx->d_matrix.as<ulonglong2>(), ctx->d_scores.as<int>(),
                                                   ctx->d_state.as<u64>(), ctx->d_ranges.as<int2>(), ctx->d_bandout.as<BandOut>(),
                                                   ctx->d_counters.as<u64>(), min_B);
    CK(cudaGetLastError());
    ctx->stats.kernel_launches++;
    return 0;
}

constexpr i64 kDynBandMax = 11000;     // blocks whose Pv/Mv (16 B each) fit one CTA's shared memory
int rounds_for(i64 B)                  // 64 = the dynamic (shared-memory resident) kernel
{
    for (int r : {1, 2, 4, 8, 16, 32}) if (B <= 32 * r) return r;
    return B <= kDynBandMax ? 64 : 0;
}

template <bool FULL>
int launch_banded_dyn(qb200_ctx *ctx, const BandTask *d_tasks, const int *d_list, int begin, int n_tasks, i64 mat_sub, const u64 *peq_base, int min_B)
{
    if (n_tasks <= 0) return 0;
    const int cap = (int)kDynBandMax + 2;
    const size_t smem = (size_t)cap * 16;
    auto kern = k_banded_warp_dyn<FULL>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<n_tasks, 32, smem, ctx->stream>>>(d_tasks, d_list, begin, n_tasks, mat_sub, ctx->d_codes.as<unsigned char>(), peq_base,
                                             ctx->d_matrix.as<ulonglong2>(), ctx->d_scores.as<int>(), ctx->d_state.as<u64>(),
                                             ctx->d_ranges.as<int2>(), ctx->d_bandout.as<BandOut>(), ctx->d_counters.as<u64>(), cap, min_B);
    CK(cudaGetLastError());
    ctx->stats.kernel_launches++;
    return 0;
}

// One launch per band-height class present in the list (each warp exits at once if its task belongs to another class).
template <bool FULL>
int launch_banded(qb200_ctx *ctx, unsigned r_mask, const BandTask *d_tasks, const int *d_list, int begin, int n_tasks, i64 mat_sub,
                  const u64 *peq_base = nullptr, int min_B = 0)
{
    if (!peq_base) peq_base = ctx->d_peq.as<u64>();
    int rc = 0;
    if (!rc && (r_mask & 1)) rc = launch_banded_r<1, FULL>(ctx, d_tasks, d_list, begin, n_tasks, mat_sub, peq_base, min_B);
    if (!rc && (r_mask & 2)) rc = launch_banded_r<2, FULL>(ctx, d_tasks, d_list, begin, n_tasks, mat_sub, peq_base, min_B);
    if (!rc && (r_mask & 4)) rc = launch_banded_r<4, FULL>(ctx, d_tasks, d_list, begin, n_tasks, mat_sub, peq_base, min_B);
    if (!rc && (r_mask & 8)) rc = launch_banded_r<8, FULL>(ctx, d_tasks, d_list, begin, n_tasks, mat_sub, peq_base, min_B);
    if (!rc && (r_mask & 16)) rc = launch_banded_r<16, FULL>(ctx, d_tasks, d_list, begin, n_tasks, mat_sub, peq_base, min_B);
    if (!rc && (r_mask & 32)) rc = launch_banded_r<32, FULL>(ctx, d_tasks, d_list, begin, n_tasks, mat_sub, peq_base, min_B);
    if (!rc && (r_mask & 64)) rc = launch_banded_dyn<FULL>(ctx, d_tasks, d_list, begin, n_tasks, mat_sub, peq_base, min_B);
    return rc;
}
constexpr int kTileBandMax = kTileMaxRing - 2;      // widest band (blocks) the tile kernels take

constexpr int kThreadBandMax = 4;

int launch_thread_fill(qb200_ctx *ctx, const int *d_list, int begin, int n_tasks, i64 mat_sub, const u64 *peq_base = nullptr, bool records = false)
{
    if (n_tasks <= 0) return 0;
    if (!peq_base) peq_base = ctx->d_peq.as<u64>();
    const int T = kThreadFillThreads;
    const size_t smem = (size_t)kThreadBandMax * kAlpha * T * 8;
    auto kern = records ? k_banded_thread<kThreadBandMax, true> : k_banded_thread<kThreadBandMax, false>;
    if (const char *e = getenv("QB200_FILL_CARVE")) cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e));
    kern<<<(n_tasks + T - 1) / T, T, smem, ctx->stream>>>(
        ctx->d_leaves.as<BandTask>(), d_list, begin, n_tasks, mat_sub, ctx->d_codes.as<unsigned char>(), peq_base,
        ctx->d_matrix.as<ulonglong2>(), ctx->d_ranges.as<int2>(), ctx->d_counters.as<u64>());
    CK(cudaGetLastError());
    ctx->stats.kernel_launches++;
    return 0;
}

int launch_traceback(qb200_ctx *ctx, const int *d_list, int begin, int n_tasks, i64 mat_sub, bool warp_layout = false, int min_B = 0)
{
    if (n_tasks <= 0) return 0;
    if (warp_layout) {      // leaves written by the warp kernel: cooperative, tile-prefetching walk
        k_traceback_warp<<<(n_tasks + kTraceWarpsPerCta - 1) / kTraceWarpsPerCta, 32 * kTraceWarpsPerCta, 0, ctx->stream>>>(
            ctx->d_leaves.as<BandTask>(), d_list, begin, n_tasks, mat_sub, ctx->raw(), ctx->d_matrix.as<ulonglong2>(),
            ctx->d_ranges.as<int2>(), ctx->d_ops.as<u32>(), ctx->d_leafout.as<LeafOut>(), min_B);
        CK(cudaGetLastError());
        ctx->stats.kernel_launches++;
        return 0;
    }
    auto kern = k_traceback_thread;
    if (const char *e = getenv("QB200_TRACE_CARVE")) cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e));
    kern<<<(n_tasks + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_leaves.as<BandTask>(), d_list, begin, n_tasks, mat_sub,
                                                                        ctx->raw(), ctx->d_matrix.as<ulonglong2>(), ctx->d_ranges.as<int2>(),
                                                                        ctx->d_ops.as<u32>(), ctx->d_leafout.as<LeafOut>());
    CK(cudaGetLastError());
    ctx->stats.kernel_launches++;
    return 0;
}

// ---- tile path (qb_tiles.cuh / qb_tiletrace.cuh) ----------------------------------------------------------------
constexpr int kTileClasses = 8;                     // ring sizes 8 << c: bands up to kTileMaxRing - 2 blocks
static int kTileBigClass = getenv("QB200_TILE_BIG_CLASS") ? atoi(getenv("QB200_TILE_BIG_CLASS")) : 5;   // classes from this one on run 256 compute lanes per CTA
struct TileCtl { int counts[kTileClasses]; int next[kTileClasses]; int punt_count; int pad_; unsigned long long tt_words; int pad2_[12]; };
inline bool tile_band_ok(i64 B) { return tile_ring_for(B) <= kTileMaxRing; }

template <bool FULL, int LANES>
int launch_tiles_class(qb200_ctx *ctx, const TilePools &P, int c, int cap, int n_bound)
{
    const int RB = 8 << c;
    const size_t fixed = tile_smem_bytes(RB, 0, LANES), per = tile_slot_arena_bytes(RB) + sizeof(TileSlot);
    size_t budget = (RB <= 128 ? 55 : 110) * 1024;           // <= 128 blocks: four CTAs per SM
    if (const char *e = getenv("QB200_TILE_SMEM_KB")) budget = (size_t)atoi(e) * 1024;
    int nslots = (int)((budget > fixed ? budget - fixed : 0) / per);
    nslots = std::max(1, std::min(32, nslots));
    if (!ctx->sms) { ctx->sms = 148; cudaDeviceGetAttribute(&ctx->sms, cudaDevAttrMultiProcessorCount, ctx->device); }
    {   // few tasks: no more slots per CTA than it takes to give every resident CTA its share (26 slots for 12.5 k tasks
        // left 111 of 592 CTAs without work: 5.6 vs 5.0 ms; 4 slots for the 512 half-passes of a Hirschberg level used
        // 128 of 296)
        int occ0 = (int)((227 * 1024) / (tile_smem_bytes(RB, nslots, LANES) + 1024));
        occ0 = std::max(1, std::min(occ0, 2048 / (LANES + 32)));
        const int share = (int)(((i64)n_bound + (i64)ctx->sms * occ0 - 1) / ((i64)ctx->sms * occ0));
        nslots = std::min(nslots, std::max(share, 1));
    }
    if (const char *e = getenv("QB200_TILE_SLOTS")) nslots = std::max(1, std::min(32, atoi(e)));
    const size_t smem = tile_smem_bytes(RB, nslots, LANES);
    auto kern = k_band_tiles<FULL, LANES>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = (int)((227 * 1024) / (smem + 1024));
    occ = std::max(1, std::min(occ, 2048 / (LANES + 32)));
    occ = std::min(occ, LANES >= 256 ? 2 : LANES >= 128 ? 4 : LANES >= 64 ? 6 : 8);       // the kernel's launch bound
    const int blocks = std::max(1, std::min(ctx->sms * occ, (n_bound + nslots - 1) / nslots));
    TileCtl *ctl = ctx->d_tctl.as<TileCtl>();
    TileLaunch Q;
    Q.list = ctx->d_tclass.as<int>() + (size_t)c * cap; Q.count = &ctl->counts[c]; Q.next = &ctl->next[c];
    Q.counters = ctx->d_counters.as<u64>(); Q.RB = RB; Q.nslots = nslots; Q.lanes = LANES;
    kern<<<blocks, LANES + 32, smem, ctx->stream>>>(P, Q);
    CK(cudaGetLastError());
    ctx->stats.kernel_launches++;
    return 0;
}

// BandEd fill of tasks[list[begin .. begin+n)] by the tile kernels, one launch per band-height class in class_mask.
// FULL: writes tile records (pool d_matrix, task.mat_off in 16-byte units) + live ranges; !FULL: score-only passes.
// Tasks the kernels give up on: FULL -> BandOut.pos_v = kTilePunted; !FULL -> appended to d_punt (count in d_tctl).
// Per-launch set-up of the tile kernels for tasks[list[begin .. begin+n)]: band-height class lists and, for every task,
// its aligned text in the tile-text pool (task.tt_off).  first: this is the first tile work of a fill phase — the text
// pool and the punt list start empty (several calls of one phase share both: thread-fill leaves, then the wider ones).
template <bool FULL>
int tile_prepare(qb200_ctx *ctx, BandTask *d_tasks, const int *d_list, int begin, int n, bool first, bool want_lists = true)
{
    if (n <= 0) return 0;
    if (want_lists) CK(ctx->d_tclass.reserve((size_t)kTileClasses * (size_t)n * 4));
    CK(ctx->d_tctl.reserve(sizeof(TileCtl)));
    if (first) {
        // tile-text pool: every task's text once more, aligned (the tasks of one phase are distinct (sub-)texts of the
        // batch, forward and reverse pass of a Hirschberg node at most; 80 bytes of rounding per task)
        CK(ctx->d_ttext.reserve(2 * (size_t)ctx->raw_bytes + (size_t)std::max<i64>(ctx->n_pairs, n) * 160 + 64));
        CK(cudaMemsetAsync(ctx->d_tctl.p, 0, sizeof(TileCtl), ctx->stream));
    } else CK(cudaMemsetAsync(ctx->d_tctl.p, 0, offsetof(TileCtl, punt_count), ctx->stream));
    k_tile_classes<FULL><<<(n + 255) / 256, 256, 0, ctx->stream>>>(d_tasks, d_list, begin, n, want_lists ? ctx->d_tclass.as<int>() : nullptr, n,
                                                                  ctx->d_tctl.as<TileCtl>()->counts, &ctx->d_tctl.as<TileCtl>()->tt_words);
    if (ctx->max_n > 1024)
        k_tile_text<FULL, 32><<<(int)(((i64)n * 32 + 255) / 256), 256, 0, ctx->stream>>>(d_tasks, d_list, begin, n, ctx->d_codes.as<unsigned char>(), ctx->d_ttext.as<u64>());
    else
        k_tile_text<FULL, 4><<<(int)(((i64)n * 4 + 255) / 256), 256, 0, ctx->stream>>>(d_tasks, d_list, begin, n, ctx->d_codes.as<unsigned char>(), ctx->d_ttext.as<u64>());
    CK(cudaGetLastError());
    ctx->stats.kernel_launches += 2;
    return 0;
}

// BandEd fill of tasks[list[begin .. begin+n)] by the tile kernels, one launch per band-height class in class_mask.
// FULL: writes tile records (pool d_matrix, task.mat_off in 16-byte units) + live ranges; !FULL: score-only passes.
// Tasks the kernels give up on: FULL -> BandOut.pos_v = kTilePunted; !FULL -> appended to d_punt (count in d_tctl).
template <bool FULL>
int launch_tiles(qb200_ctx *ctx, BandTask *d_tasks, const int *d_list, int begin, int n, i64 sub, const u64 *peq_base,
                 unsigned class_mask = 0xffu, bool first = true)
{
    if (n <= 0) return 0;
    CK(ctx->d_punt.reserve((size_t)std::max<i64>(ctx->n_pairs, n) * 4 + 16));
    { const int rc = tile_prepare<FULL>(ctx, d_tasks, d_list, begin, n, first); if (rc) return rc; }
    TilePools P;
    P.ttext = ctx->d_ttext.as<u64>();
    P.tasks = d_tasks; P.codes = ctx->d_codes.as<unsigned char>(); P.peq = peq_base; P.recs = ctx->d_matrix.as<TileRec>();
    P.ranges = ctx->d_ranges.as<int2>(); P.scores = ctx->d_scores.as<int>(); P.state = ctx->d_state.as<u64>();
    P.outs = ctx->d_bandout.as<BandOut>(); P.punt_list = ctx->d_punt.as<int>(); P.punt_count = &ctx->d_tctl.as<TileCtl>()->punt_count;
    P.rec_sub = sub;
    for (int c = 0; c < kTileClasses; ++c) {
        if (!((class_mask >> c) & 1u)) continue;
        int rc;
        // rings of 256+ blocks leave room for two CTAs per SM only (110 KB of slots each): 256 compute lanes per CTA keep
        // 18 warps on the SM instead of 10
        int lanes = c == 0 ? 32 : c == 1 ? 64 : c >= kTileBigClass ? 256 : 128;
        if (const char *e = getenv("QB200_TILE_LANES")) lanes = atoi(e);
        if (lanes == 32) rc = launch_tiles_class<FULL, 32>(ctx, P, c, n, n);
        else if (lanes == 256) rc = launch_tiles_class<FULL, 256>(ctx, P, c, n, n);
        else if (lanes == 128) rc = launch_tiles_class<FULL, 128>(ctx, P, c, n, n);
        else rc = launch_tiles_class<FULL, 64>(ctx, P, c, n, n);
        if (rc) return rc;
    }
    return 0;
}

inline unsigned tile_class_bit(i64 B)
{
    const int rb = tile_ring_for(B);
    int c = 0;
    while ((8 << c) < rb) ++c;
    return 1u << c;
}

// thread_fill: the records come from the thread fill, which never gives a leaf up (no BandOut marker to look at)
int launch_tile_traceback(qb200_ctx *ctx, const int *d_list, int begin, int n, i64 sub, const u64 *peq_base, bool thread_fill = false)
{
    if (n <= 0) return 0;
    ctx->tile_walks = true;
    CK(ctx->d_punt.reserve((size_t)std::max<i64>(ctx->n_pairs, n) * 4 + 16));
    static bool carve_set = false;
    if (!carve_set) { cudaFuncSetAttribute(k_traceback_tiles, cudaFuncAttributePreferredSharedMemoryCarveout, 100); carve_set = true; }
    k_traceback_tiles<<<(n + kTileTraceThreads - 1) / kTileTraceThreads, kTileTraceThreads, 0, ctx->stream>>>(
        ctx->d_leaves.as<BandTask>(), d_list, begin, n, sub, ctx->d_ttext.as<u64>(), ctx->raw(), peq_base, ctx->d_matrix.as<TileRec>(),
        ctx->d_ranges.as<int2>(), thread_fill ? nullptr : ctx->d_bandout.as<BandOut>(), ctx->d_ops.as<u32>(), ctx->d_leafout.as<LeafOut>(), ctx->d_punt.as<int>(),
        &ctx->d_tctl.as<TileCtl>()->punt_count);
    CK(cudaGetLastError());
    ctx->stats.kernel_launches++;
    return 0;
}

__global__ void k_gather_tasks(const BandTask *tasks, const int *ids, int n, BandTask *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = tasks[ids[i]];
}
__global__ void k_scatter_tasks(BandTask *tasks, const int *ids, int n, const BandTask *in)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) tasks[ids[i]] = in[i];
}

// Leaves the tile path punted on (walk outside the live band, empty band): redo them with the exact full-matrix
// kernels (warp fill in the reference's [column][word] layout + the thread walk, 2-bit ops).  Synchronises.
int rerun_punted_leaves(qb200_ctx *ctx, const u64 *peq_base)
{
    int cnt = 0;
    CK(cudaMemcpyAsync(ctx->h_pinned + 128, &ctx->d_tctl.as<TileCtl>()->punt_count, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cnt = *reinterpret_cast<int *>(ctx->h_pinned + 128);
    ctx->stats.leaves_punted += cnt;
    if (cnt <= 0) return 0;
    std::vector<BandTask> tk((size_t)cnt);
    CK(ctx->d_gather.reserve(sizeof(BandTask) * (size_t)cnt));
    k_gather_tasks<<<(cnt + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_leaves.as<BandTask>(), ctx->d_punt.as<int>(), cnt, ctx->d_gather.as<BandTask>());
    CK(cudaMemcpyAsync(tk.data(), ctx->d_gather.p, sizeof(BandTask) * (size_t)cnt, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    const i64 limit = (i64)(std::min<size_t>(ctx->matrix_limit, (size_t)((free_b + ctx->d_matrix.cap) * 0.85)) / 16);
    for (int q0 = 0; q0 < cnt;) {
        i64 ent = 0; int q1 = q0; unsigned mask = 0;
        while (q1 < cnt) {
            BandTask &t = tk[(size_t)q1];
            const BandGeom g = band_geometry(t.m, t.n, t.cutoff);
            const i64 e = (i64)(t.n + 1) * g.Bc;
            if (q1 > q0 && ent + e > limit) break;
            const int R = rounds_for(g.Bc);
            if (!R) { ctx->err = "leaf band of " + std::to_string(g.Bc) + " blocks exceeds the supported 11000"; return QB200_ERR_ARG; }
            mask |= (unsigned)R;
            t.mat_off = ent; t.mat_cs = (int)g.Bc; t.mat_ws = 1;
            ent += e; ++q1;
        }
        if ((size_t)ent * 16 > free_b + ctx->d_matrix.cap) { ctx->err = "a single traceback matrix does not fit the device"; return QB200_ERR_OOM; }
        CK(cudaMemcpyAsync(ctx->d_gather.as<BandTask>() + q0, tk.data() + q0, sizeof(BandTask) * (size_t)(q1 - q0), cudaMemcpyHostToDevice, ctx->stream));
        k_scatter_tasks<<<(q1 - q0 + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_leaves.as<BandTask>(), ctx->d_punt.as<int>() + q0, q1 - q0, ctx->d_gather.as<BandTask>() + q0);
        CK(ctx->d_matrix.reserve((size_t)ent * 16));
        { Span sp(ctx, ST_FILL); int rc = launch_banded<true>(ctx, mask, ctx->d_leaves.as<BandTask>(), ctx->d_punt.as<int>(), q0, q1 - q0, 0, peq_base); if (rc) return rc; }
        { Span sp(ctx, ST_TRACE); int rc = launch_traceback(ctx, ctx->d_punt.as<int>(), q0, q1 - q0, 0, false); if (rc) return rc; }
        ctx->stats.matrix_bytes += ent * 16;
        CK(cudaStreamSynchronize(ctx->stream));
        q0 = q1;
    }
    return 0;
}

// Largest e in (s, n] such that the entries of items s..e-1 fit `limit` (at least one item).  One thread, binary search.
// out[0] = e, ent[0] = entries of the chunk (can exceed `limit` when a single item does: the caller grows the pool or fails)
__global__ void k_chunk_end_groups(const i64 *goff, const i64 *gsize, int n, int s, i64 limit, int *out, i64 *ent)
{
    int lo = s + 1, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (goff[mid - 1] + gsize[mid - 1] - goff[s] <= limit) lo = mid; else hi = mid - 1;
    }
    out[0] = lo;
    ent[0] = goff[lo - 1] + gsize[lo - 1] - goff[s];
}
__device__ __forceinline__ i64 leaf_entries(const BandTask &t, int tiles)
{
    const bool tile = tiles && tile_ring_for(t.mat_cs) <= kTileMaxRing;          // mat_cs = band height of a warp-class leaf
    return tile ? 2 * (i64)((t.n + 63) / 64) * t.mat_cs : (i64)(t.n + 1) * t.mat_cs;
}
__global__ void k_chunk_end_leaves(const BandTask *leaves, const int *list, int n, int s, i64 limit, int *out, i64 *start_off, i64 *ent, int tiles)
{
    const i64 base = leaves[list[s]].mat_off;
    int lo = s + 1, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        const BandTask &t = leaves[list[mid - 1]];
        if (t.mat_off + leaf_entries(t, tiles) - base <= limit) lo = mid; else hi = mid - 1;
    }
    out[0] = lo; start_off[0] = base;
    const BandTask &t = leaves[list[lo - 1]];
    ent[0] = t.mat_off + leaf_entries(t, tiles) - base;
}

struct RunPlan {
    PlanSum tot;          // totals of the fast path
    i64 mat_t = 0;        // entries of the thread-kernel groups
    int n_groups = 0;
    unsigned tile_mask = 0;   // band-height classes of the tile leaves; bit 31: some warp-class leaf needs the full matrix
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
    if (const char *e = getenv("QB200_TILES")) ctx->use_tiles = atoi(e) != 0;
    if (const char *e = getenv("QB200_THREAD_BAND_MAX")) ctx->thread_band_max = std::max(0, std::min(atoi(e), (int)kThreadBandMax));
    if (!ctx->use_tiles) ctx->thread_band_max = kThreadBandMax;
    *out = ctx;
    return 0;
}

void qb200_destroy(qb200_ctx_t *ctx)
{
    if (!ctx) return;
    for (int k = 0; k < qb200_ctx::kWorkers; ++k) if (ctx->child[k]) { qb200_destroy(ctx->child[k]); ctx->child[k] = nullptr; }
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (DevBuf *b : {&ctx->d_raw, &ctx->d_codes, &ctx->d_pairs, &ctx->d_peq, &ctx->d_peqjobs, &ctx->d_pairodd, &ctx->d_bound, &ctx->d_hew,
                      &ctx->d_score, &ctx->d_status, &ctx->d_textlen, &ctx->d_cigoff, &ctx->d_cigar, &ctx->d_counters,
                      &ctx->d_scan_tmp, &ctx->d_leaves, &ctx->d_leafout, &ctx->d_pairleaves, &ctx->d_work, &ctx->d_bandout,
                      &ctx->d_matrix, &ctx->d_scores, &ctx->d_state, &ctx->d_ops, &ctx->d_ranges, &ctx->d_cls, &ctx->d_cutoff,
                      &ctx->d_plan_items, &ctx->d_plan_offs, &ctx->d_textbytes, &ctx->d_list_t, &ctx->d_list_w, &ctx->d_list_slow,
                      &ctx->d_gsize, &ctx->d_goff, &ctx->d_gB, &ctx->d_peq2, &ctx->d_jobs2, &ctx->d_tasks2, &ctx->d_wintasks,
                      &ctx->d_quad, &ctx->d_fmat, &ctx->d_franges, &ctx->d_done, &ctx->d_winout, &ctx->d_winscratch, &ctx->d_split, &ctx->d_splitout, &ctx->d_splitscratch, &ctx->d_scatter, &ctx->d_tclass, &ctx->d_tctl, &ctx->d_punt, &ctx->d_gather, &ctx->d_ttext, &ctx->d_wintile, &ctx->d_packed, &ctx->d_excpos, &ctx->d_excchr})
        b->release();
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    ctx->h_pairs.release();
    ctx->h_stage.release();
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
    if (b->n_pairs > kMaxPairs) { ctx->err = "a batch holds at most 2^31 - 2^20 pairs: split it"; return QB200_ERR_ARG; }
    if (b->n_pairs > 0 && (!b->pattern_off || !b->pattern_len || !b->text_off || !b->text_len || (b->seqs_bytes > 0 && !b->seqs))) return QB200_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    ctx->n_pairs = b->n_pairs; ctx->raw_bytes = b->seqs_bytes; ctx->d_raw_ext = nullptr;
    memset(&ctx->stats, 0, sizeof ctx->stats);
    // the characters go first: from pinned memory the copy runs while the host builds the pair records below
    const size_t padded = ((size_t)b->seqs_bytes + 15) / 16 * 16 + 48;       // >= 48: the zeroed tail never starts before the buffer
    CK(ctx->d_raw.reserve(padded));
    CK(cudaMemsetAsync(ctx->d_raw.as<char>() + (padded - 48), 0, 48, ctx->stream));
    if (b->seqs_bytes > 0) CK(cudaMemcpyAsync(ctx->d_raw.p, b->seqs, (size_t)b->seqs_bytes, cudaMemcpyHostToDevice, ctx->stream));
    int rc = build_pair_records(ctx, b->n_pairs, b->pattern_off, b->pattern_len, b->text_off, b->text_len, b->seqs_bytes);
    if (rc) { cudaStreamSynchronize(ctx->stream); return rc; }
    ctx->stats.h2d_bytes += b->seqs_bytes;
    return finish_upload(ctx);
}

// ---- 2-bit packed input ----
__global__ void __launch_bounds__(256) k_unpack2(const u32 *__restrict__ packed, uint4 *__restrict__ raw, i64 nvec)
{
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nvec) return;
    const u32 w = __ldg(packed + i);                      // 16 characters
    u32 o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        u32 v = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) v |= ((0x54474341u >> (8 * ((w >> (8 * k + 2 * j)) & 3u))) & 0xffu) << (8 * j);   // "ACGT"[code]
        o[k] = v;
    }
    raw[i] = make_uint4(o[0], o[1], o[2], o[3]);
}
__global__ void __launch_bounds__(256) k_patch_exceptions(const i64 *__restrict__ pos, const unsigned char *__restrict__ chr, i64 n, unsigned char *__restrict__ raw)
{
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) raw[pos[i]] = chr[i];
}

int64_t qb200_pack_batch(const qb200_batch_t *in, uint8_t *packed, int64_t *exc_pos, uint8_t *exc_chr, int64_t exc_cap, int threads)
{
    if (!in || !packed || in->seqs_bytes < 0 || in->n_pairs < 0 || exc_cap < 0) return QB200_ERR_ARG;
    const i64 n = in->n_pairs, nb = in->seqs_bytes, nq = (nb + 3) / 4;
    for (i64 i = 0; i < n; ++i) {
        const i64 po = in->pattern_off[i], to = in->text_off[i], m = in->pattern_len[i], t = in->text_len[i];
        if (m < 0 || t < 0 || po < 0 || to < 0 || po + m > nb || to + t > nb) return QB200_ERR_ARG;
    }
    int nt = threads > 0 ? threads : (int)std::max(1u, std::thread::hardware_concurrency());
    nt = (int)std::max<i64>(1, std::min<i64>(nt, nq / (1 << 16) + 1));
    // 1) the stream, four characters per byte (anything that is not ACGT packs as 0)
    {
        std::vector<std::thread> th;
        for (int k = 0; k < nt; ++k) th.emplace_back([=]() {
            i64 q0 = nq * k / nt, q1 = nq * (k + 1) / nt;
            q0 &= ~(i64)1; if (k + 1 < nt) q1 &= ~(i64)1;     // whole 8-character groups per thread
            const unsigned char *s = reinterpret_cast<const unsigned char *>(in->seqs);
            // eight characters per step: bits 1-2 of A C G T are 0 1 3 2 -> code = x ^ (x >> 1); the eight 2-bit codes are
            // then squeezed out of their bytes (bytes -> nibbles -> bytes -> 16 bits)
            for (; q0 + 2 <= q1 && 4 * q0 + 8 <= nb; q0 += 2) {
                uint64_t w;
                memcpy(&w, s + 4 * q0, 8);
                uint64_t x = (w >> 1) & 0x0303030303030303ull;
                x ^= (x >> 1) & 0x0101010101010101ull;
                x = (x | (x >> 6)) & 0x000f000f000f000full;
                x = (x | (x >> 12)) & 0x000000ff000000ffull;
                const unsigned v = (unsigned)((x | (x >> 24)) & 0xffffull);
                packed[q0] = (uint8_t)v; packed[q0 + 1] = (uint8_t)(v >> 8);
            }
            for (i64 q = q0; q < q1; ++q) {
                unsigned v = 0;
                const i64 lim = std::min<i64>(4, nb - 4 * q);
                for (i64 j = 0; j < lim; ++j) {
                    const unsigned c = s[4 * q + j];
                    // A 0x41 C 0x43 G 0x47 T 0x54: bits 1-2 of the byte are 0, 1, 3, 2
                    const unsigned code = (c >> 1) & 3u;
                    v |= (code ^ (code >> 1)) << (2 * j);
                }
                packed[q] = (uint8_t)v;
            }
        });
        for (auto &t : th) t.join();
    }
    // 2) the exceptions, pair by pair (pairs are scanned in index order; the list is sorted afterwards if the batch is not)
    std::vector<std::vector<std::pair<i64, uint8_t>>> found((size_t)nt);
    {
        std::vector<std::thread> th;
        for (int k = 0; k < nt; ++k) th.emplace_back([=, &found]() {
            const i64 i0 = n * k / nt, i1 = n * (k + 1) / nt;
            const unsigned char *s = reinterpret_cast<const unsigned char *>(in->seqs);
            auto &out = found[(size_t)k];
            auto scan = [&](i64 off, i64 len) {
                i64 j = off;
                const i64 end = off + len;
                // 0x80 in every byte of v that is zero, exactly (no borrow across bytes)
                auto zero_bytes = [](uint64_t v) { return ~(((v & 0x7f7f7f7f7f7f7f7full) + 0x7f7f7f7f7f7f7f7full) | v | 0x7f7f7f7f7f7f7f7full); };
                for (; j + 8 <= end; j += 8) {                // eight characters at a time: which of them are not A, C, G, T?
                    uint64_t w;
                    memcpy(&w, s + j, 8);
                    const uint64_t acgt = zero_bytes(w ^ 0x4141414141414141ull) | zero_bytes(w ^ 0x4343434343434343ull) |
                                          zero_bytes(w ^ 0x4747474747474747ull) | zero_bytes(w ^ 0x5454545454545454ull);
                    uint64_t other = ~acgt & 0x8080808080808080ull;
                    while (other) {
                        const int bidx = __builtin_ctzll(other) >> 3;
                        other &= other - 1;
                        out.emplace_back(j + bidx, s[j + bidx]);
                    }
                }
                for (; j < end; ++j) {
                    const unsigned c = s[j];
                    if (c != 'A' && c != 'C' && c != 'G' && c != 'T') out.emplace_back(j, (uint8_t)c);
                }
            };
            for (i64 i = i0; i < i1; ++i) { scan(in->pattern_off[i], in->pattern_len[i]); scan(in->text_off[i], in->text_len[i]); }
        });
        for (auto &t : th) t.join();
    }
    std::vector<std::pair<i64, uint8_t>> all;
    for (auto &f : found) all.insert(all.end(), f.begin(), f.end());
    if (!std::is_sorted(all.begin(), all.end())) std::sort(all.begin(), all.end());
    all.erase(std::unique(all.begin(), all.end()), all.end());          // sequences may overlap (shared texts)
    const i64 ne = (i64)all.size();
    if (ne > exc_cap || (ne > 0 && (!exc_pos || !exc_chr))) return -std::max<i64>(ne, 1);
    for (i64 k = 0; k < ne; ++k) {
        exc_pos[k] = all[(size_t)k].first; exc_chr[k] = all[(size_t)k].second;
        packed[exc_pos[k] >> 2] &= (uint8_t)~(3u << (2 * (exc_pos[k] & 3)));      // exceptions pack as 0
    }
    return ne;
}

int qb200_upload_packed(qb200_ctx_t *ctx, const qb200_packed_batch_t *b)
{
    if (!ctx || !b || b->n_pairs < 0 || b->n_chars < 0 || b->n_exc < 0) return QB200_ERR_ARG;
    if (b->n_pairs > kMaxPairs) { ctx->err = "a batch holds at most 2^31 - 2^20 pairs: split it"; return QB200_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    ctx->n_pairs = b->n_pairs; ctx->raw_bytes = b->n_chars; ctx->d_raw_ext = nullptr;
    memset(&ctx->stats, 0, sizeof ctx->stats);
    const size_t padded = ((size_t)b->n_chars + 15) / 16 * 16 + 48;
    const i64 nvec = (i64)((b->n_chars + 15) / 16);
    CK(ctx->d_raw.reserve(padded));
    CK(ctx->d_packed.reserve((size_t)nvec * 4 + 16));
    CK(cudaMemsetAsync(ctx->d_raw.as<char>() + (padded - 48), 0, 48, ctx->stream));
    const size_t pbytes = (size_t)(b->n_chars + 3) / 4;
    if (pbytes) {
        CK(cudaMemsetAsync(ctx->d_packed.as<char>() + (size_t)nvec * 4 - 4, 0, 4, ctx->stream));       // the last word may be partial
        CK(cudaMemcpyAsync(ctx->d_packed.p, b->packed, pbytes, cudaMemcpyHostToDevice, ctx->stream));
        k_unpack2<<<(unsigned)((nvec + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_packed.as<u32>(), ctx->d_raw.as<uint4>(), nvec);
    }
    for (i64 k = 0; k < b->n_exc; ++k)
        if (b->exc_pos[k] < 0 || b->exc_pos[k] >= b->n_chars) { cudaStreamSynchronize(ctx->stream); ctx->err = "exception position outside the stream"; return QB200_ERR_ARG; }
    if (b->n_exc) {
        CK(ctx->d_excpos.reserve((size_t)b->n_exc * 8));
        CK(ctx->d_excchr.reserve((size_t)b->n_exc));
        CK(cudaMemcpyAsync(ctx->d_excpos.p, b->exc_pos, (size_t)b->n_exc * 8, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_excchr.p, b->exc_chr, (size_t)b->n_exc, cudaMemcpyHostToDevice, ctx->stream));
        k_patch_exceptions<<<(unsigned)((b->n_exc + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_excpos.as<i64>(), ctx->d_excchr.as<unsigned char>(), b->n_exc,
                                                                                      ctx->d_raw.as<unsigned char>());
    }
    CK(cudaGetLastError());
    int rc = build_pair_records(ctx, b->n_pairs, b->pattern_off, b->pattern_len, b->text_off, b->text_len, b->n_chars);
    if (rc) { cudaStreamSynchronize(ctx->stream); return rc; }
    ctx->stats.h2d_bytes += (i64)pbytes + b->n_exc * 9;
    return finish_upload(ctx);
}

static int align_batch_pipelined(qb200_ctx *ctx, const quicked_params_t *params, const qb200_batch_t *b, qb200_results_t *res,
                                 const qb200_packed_batch_t *pk);

int qb200_align_batch_packed(qb200_ctx_t *ctx, const quicked_params_t *params, const qb200_packed_batch_t *b, qb200_results_t *res)
{
    if (!ctx || !params || !b || !res) return QB200_ERR_ARG;
    // big jobs are pipelined like qb200_align_batch's (the thresholds count characters, not packed bytes: what a sub-batch
    // costs on the device is the same)
    i64 min_pairs = 200000;
    if (const char *e = getenv("QB200_PIPELINE_MIN_PAIRS")) min_pairs = std::max<i64>(2, atoll(e));
    const bool big = b->n_pairs >= min_pairs || (b->n_pairs >= 4096 && b->n_chars >= ((i64)512 << 20));
    if (big && !getenv("QB200_NO_PIPELINE")) {
        for (i64 k = 1; k < b->n_exc; ++k) if (b->exc_pos[k - 1] >= b->exc_pos[k]) { ctx->err = "exception positions must ascend"; return QB200_ERR_ARG; }
        const qb200_batch_t view = {nullptr, b->n_chars, b->n_pairs, b->pattern_off, b->pattern_len, b->text_off, b->text_len};
        return align_batch_pipelined(ctx, params, &view, res, b);
    }
    int rc = qb200_upload_packed(ctx, b);
    if (!rc) rc = qb200_run(ctx, params);
    if (!rc) rc = qb200_download(ctx, res);
    return rc;
}

// ---- a batch born on the device: the seeded generator of qb200_generate_pairs_ex as a kernel ----
int qb200_generate_device(qb200_ctx_t *ctx, uint64_t seed, int64_t first_pair, int64_t n_pairs, int32_t length, double error,
                          int32_t indels_num, int32_t indels_len)
{
    if (!ctx || n_pairs < 0 || first_pair < 0 || length <= 0 || indels_num < 0 || indels_len < 0) return QB200_ERR_ARG;
    if (n_pairs > kMaxPairs) { ctx->err = "a batch holds at most 2^31 - 2^20 pairs: split it"; return QB200_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    GenParams g;
    g.seed = seed; g.first_pair = first_pair; g.length = length; g.indels_num = indels_num; g.indels_len = indels_len;
    g.num_errors = error >= 1.0 ? (int)error : (int)std::ceil((double)((float)length * (float)error));     // generate_dataset.c:370
    g.stride = 2 * (i64)length + g.num_errors + 2;
    const size_t smem = ((size_t)(length + g.num_errors + 16) / 4 + 8) * 4;
    if (smem > 200 * 1024) { ctx->err = "qb200_generate_device: reads of this length do not fit a CTA's shared memory (use qb200_generate_pairs_ex)"; return QB200_ERR_ARG; }
    const i64 bytes = n_pairs * g.stride;
    ctx->n_pairs = n_pairs; ctx->raw_bytes = (bytes + 15) / 16 * 16; ctx->d_raw_ext = nullptr;
    memset(&ctx->stats, 0, sizeof ctx->stats);
    const size_t padded = (size_t)ctx->raw_bytes + 48;
    CK(ctx->d_raw.reserve(padded));
    CK(cudaMemsetAsync(ctx->d_raw.p, 0, padded, ctx->stream));
    CK(ctx->d_bound.reserve((size_t)std::max<i64>(n_pairs, 1) * 4));             // borrowed: the pattern lengths
    if (n_pairs) {
        CK(cudaFuncSetAttribute(k_generate_pairs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (!ctx->sms) { ctx->sms = 148; cudaDeviceGetAttribute(&ctx->sms, cudaDevAttrMultiProcessorCount, ctx->device); }
        const int occ = (int)std::max<size_t>(1, std::min<size_t>(16, (200 * 1024) / (smem + 1024)));
        const int blocks = (int)std::min<i64>(n_pairs, (i64)ctx->sms * occ);
        k_generate_pairs<<<blocks, kGenThreads, smem, ctx->stream>>>(g, (int)n_pairs, ctx->d_raw.as<unsigned char>(), ctx->d_bound.as<int>());
        CK(cudaGetLastError());
    }
    // the pair records need the pattern lengths (4 bytes per pair come back; the characters stay in HBM)
    std::vector<int32_t> pl((size_t)n_pairs), tl((size_t)n_pairs, length);
    std::vector<int64_t> po((size_t)n_pairs), to((size_t)n_pairs);
    if (n_pairs) CK(cudaMemcpyAsync(pl.data(), ctx->d_bound.p, (size_t)n_pairs * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (i64 i = 0; i < n_pairs; ++i) { po[(size_t)i] = i * g.stride; to[(size_t)i] = i * g.stride + length + g.num_errors + 1; }
    int rc = build_pair_records(ctx, n_pairs, po.data(), pl.data(), to.data(), tl.data(), ctx->raw_bytes);
    if (rc) return rc;
    return finish_upload(ctx);
}

// The batch a context holds (uploaded, unpacked or generated), copied back to host arrays: seqs needs qb200_batch_bytes().
int64_t qb200_batch_bytes(qb200_ctx_t *ctx) { return ctx ? ctx->raw_bytes : QB200_ERR_ARG; }
int qb200_download_batch(qb200_ctx_t *ctx, char *seqs, int64_t *pattern_off, int32_t *pattern_len, int64_t *text_off, int32_t *text_len)
{
    if (!ctx) return QB200_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (seqs && ctx->raw_bytes) CK(cudaMemcpyAsync(seqs, ctx->raw(), (size_t)ctx->raw_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (i64 i = 0; i < ctx->n_pairs; ++i) {
        const PairRec &r = ctx->h_pairs[(size_t)i];
        if (pattern_off) pattern_off[i] = r.p_off;
        if (pattern_len) pattern_len[i] = r.m;
        if (text_off) text_off[i] = r.t_off;
        if (text_len) text_len[i] = r.n;
    }
    return 0;
}

int qb200_upload_device(qb200_ctx_t *ctx, const qb200_batch_t *b)
{
    if (!ctx || !b || b->n_pairs < 0 || b->seqs_bytes < 0) return QB200_ERR_ARG;
    if (b->n_pairs > kMaxPairs) { ctx->err = "a batch holds at most 2^31 - 2^20 pairs: split it"; return QB200_ERR_ARG; }
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

// QB200_TRACE=2: wall-clock checkpoints inside qb200_run (host-side stalls do not show in the CUDA-event stage times)
struct RunTrace {
    bool on; std::chrono::steady_clock::time_point t0, last;
    RunTrace() : on(getenv("QB200_TRACE") && atoi(getenv("QB200_TRACE")) >= 2), t0(std::chrono::steady_clock::now()), last(t0) {}
    void pt(const char *what)
    {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[qb200 run] %-18s +%8.3f ms (%.3f)\n", what, std::chrono::duration<double, std::milli>(now - last).count(),
                std::chrono::duration<double, std::milli>(now - t0).count());
        last = now;
    }
};

int qb200_run(qb200_ctx_t *ctx, const quicked_params_t *params)
{
    if (!ctx || !params) return QB200_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    RunTrace rt;
    const quicked_params_t prm = *params;
    const i64 n = ctx->n_pairs;
    const i64 h2d = ctx->stats.h2d_bytes;
    memset(&ctx->stats, 0, sizeof ctx->stats);
    ctx->stats.h2d_bytes = h2d; ctx->stats.n_pairs = n; ctx->stats.cells = ctx->cells;
    ctx->ev_used = 0; ctx->ev_spans.clear();
    ctx->have_cigar = false; ctx->cigar_total = 0; ctx->unknown_algo = false; ctx->multi_leaf_pairs = false; ctx->tile_walks = false;
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

    rt.pt("buffers");
    // ---- prepare: codes + forward match masks ----
    {
        Span sp(ctx, ST_PREP);
        const size_t padded = ((size_t)ctx->raw_bytes + 15) / 16 * 16;
        CK(ctx->d_codes.reserve(padded + 128));            // thread kernels read whole aligned 16-byte chunks, up to 48 B past a text
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
        CK(ctx->d_pairodd.reserve((size_t)n + 16));
        // forward match masks (+ the per-pair odd-character flags): one thread per pattern on big batches of short
        // patterns, one warp per pattern otherwise (job list derived on the device from the pair records)
        if ((ni >= 16384 || ctx->max_m >= 4096) && !getenv("QB200_PEQ_WARP")) {
            if (ctx->max_m >= 4096)                                  // long patterns: a warp each (lane = every 32nd block)
                k_build_peq_pairs<32><<<(unsigned)(((i64)ni * 32 + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_pairs.as<PairRec>(), ni, ctx->d_codes.as<unsigned char>(),
                                                                                                     ctx->d_peq.as<u64>(), ctx->d_pairodd.as<unsigned char>());
            else
                k_build_peq_pairs<1><<<(ni + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_pairs.as<PairRec>(), ni, ctx->d_codes.as<unsigned char>(),
                                                                              ctx->d_peq.as<u64>(), ctx->d_pairodd.as<unsigned char>());
            CK(cudaGetLastError());
            ctx->stats.kernel_launches++;
        } else if (ni) {
            k_make_peqjobs<<<nb256, 256, 0, ctx->stream>>>(ctx->d_pairs.as<PairRec>(), ni, ctx->d_peqjobs.as<PeqJob>());
            k_build_peq<<<(ni + 7) / 8, 256, 0, ctx->stream>>>(ctx->d_peqjobs.as<PeqJob>(), ni, ctx->d_codes.as<unsigned char>(), ctx->d_peq.as<u64>(),
                                                               ctx->d_pairodd.as<unsigned char>());
            CK(cudaGetLastError());
            ctx->stats.kernel_launches += 2;
        }
    }

    rt.pt("prepare launched");
    // ---- QUICKED fast path: one fused kernel (WindowEd(S) -> BandEd fill -> traceback) for narrow-band pairs ----
    // Measured on B200 (profiles/README.md): the fused kernel wins on small, launch-bound jobs (100 bp x 100 k pairs:
    // 0.70 vs 1.09 ms) and needs no 48 GB traceback pool; on big batches three specialised kernels are ~10 % faster,
    // and so they are on small batches of long reads (15 k pairs of 1 kbp: 4.1 ms fused vs ~2.5 ms).
    bool use_fused = (prm.algo == QUICKED) && ctx->raw_bytes < ((i64)64 << 20) && ctx->max_n <= 300;
    if (const char *e = getenv("QB200_FUSED")) use_fused = (prm.algo == QUICKED) && atoi(e) != 0;
    i64 leaf_base = 0, ops_base = 0;
    if (use_fused) {
        if (!ctx->sms) { ctx->sms = 148; cudaDeviceGetAttribute(&ctx->sms, cudaDevAttrMultiProcessorCount, ctx->device); }
        const int sms = ctx->sms;
        int fctas = kFusedCtasPerSm;
        if (const char *e = getenv("QB200_FUSED_CTAS")) fctas = std::max(1, std::min(atoi(e), (int)kFusedCtasPerSm));
        const int blocks = (int)std::min<i64>((n + kWsThreads - 1) / kWsThreads, (i64)sms * fctas);   // persistent: one wave
        const i64 nthr = (i64)blocks * kWsThreads, warps = nthr / 32;
        const i64 budget = (i64)std::min<size_t>(ctx->matrix_limit, (size_t)24 << 30);
        const i64 n_cap = budget / (warps * kFusedBandMax * 32 * 16) - 1;
        FusedParams fp;
        fp.hew_threshold0 = (int)prm.hew_threshold[0]; fp.hew_pct0 = prm.hew_percentage[0];
        fp.n_lim = (int)std::max<i64>(0, std::min<i64>(ctx->max_n, n_cap));
        fp.ok_status = QUICKED_WIP;
        fp.mat_warp_stride = (i64)(fp.n_lim + 1) * kFusedBandMax * 32;
        CK(ctx->d_quad.reserve((size_t)nthr * kWsQuadSlots * 8));
        CK(ctx->d_fmat.reserve((size_t)warps * (size_t)fp.mat_warp_stride * 16));
        CK(ctx->d_franges.reserve((size_t)(fp.n_lim / 64 + 2) * (size_t)nthr * 8));
        CK(ctx->d_done.reserve((size_t)n));
        CK(ctx->d_leaves.reserve(sizeof(BandTask) * (size_t)n));
        CK(ctx->d_leafout.reserve(sizeof(LeafOut) * (size_t)n));
        CK(ctx->d_ops.reserve((size_t)std::max<i64>(ctx->ops_words_fused, 1) * 4 + 16));
        {
            Span sp(ctx, ST_FUSED);
            if (prm.force_scalar)
                k_quicked_fused<false><<<blocks, kWsThreads, 0, ctx->stream>>>(ctx->d_pairs.as<PairRec>(), ni, ctx->d_codes.as<unsigned char>(), ctx->raw(),
                    ctx->d_peq.as<u64>(), fp, ctx->d_quad.as<u64>(), ctx->d_fmat.as<ulonglong2>(), ctx->d_franges.as<int2>(), ctx->d_ops.as<u32>(),
                    ctx->d_bound.as<int>(), ctx->d_hew.as<int>(), ctx->d_done.as<unsigned char>(), ctx->d_status.as<int>(), ctx->d_leaves.as<BandTask>(),
                    ctx->d_leafout.as<LeafOut>(), ctx->d_pairleaves.as<PairLeaves>(), ctx->d_counters.as<u64>());
            else
                k_quicked_fused<true><<<blocks, kWsThreads, 0, ctx->stream>>>(ctx->d_pairs.as<PairRec>(), ni, ctx->d_codes.as<unsigned char>(), ctx->raw(),
                    ctx->d_peq.as<u64>(), fp, ctx->d_quad.as<u64>(), ctx->d_fmat.as<ulonglong2>(), ctx->d_franges.as<int2>(), ctx->d_ops.as<u32>(),
                    ctx->d_bound.as<int>(), ctx->d_hew.as<int>(), ctx->d_done.as<unsigned char>(), ctx->d_status.as<int>(), ctx->d_leaves.as<BandTask>(),
                    ctx->d_leafout.as<LeafOut>(), ctx->d_pairleaves.as<PairLeaves>(), ctx->d_counters.as<u64>());
            CK(cudaGetLastError());
            ctx->stats.kernel_launches++;
        }
        ctx->stats.matrix_bytes += 0;
        leaf_base = n; ops_base = ctx->ops_words_fused;
    }
    // ---- WINDOWED with CIGAR: planned and run on the device (tasks, pseudo-leaves and leaf lists from the pair records;
    //      the tile kernel; pairs it leaves to the warp kernel — odd characters — take the host-driven path below) ----
    const bool win_fast = prm.algo == WINDOWED && !prm.only_score && ctx->use_tiles && prm.window_size >= 1 && prm.window_size <= 32 &&
                          prm.overlap_size < prm.window_size && !(prm.window_size == 2 && !prm.force_scalar) && !getenv("QB200_WIN_HOST");
    if (win_fast) {
        if (!ctx->sms) { ctx->sms = 148; cudaDeviceGetAttribute(&ctx->sms, cudaDevAttrMultiProcessorCount, ctx->device); }
        const int W = (int)prm.window_size;
        const int blocks = (int)std::min<i64>((n + kWtThreads - 1) / kWtThreads, (i64)ctx->sms * 4);
        const i64 nthr = (i64)blocks * kWtThreads, rec_tiles = (i64)W * W;
        CK(ctx->d_wintile.reserve((size_t)((2 * rec_tiles + W) * nthr) * 16 + 64));
        CK(ctx->d_wintasks.reserve(sizeof(WinTask) * (size_t)n));
        CK(ctx->d_winout.reserve(sizeof(WinOut) * (size_t)n));
        CK(ctx->d_done.reserve((size_t)n));
        CK(ctx->d_leaves.reserve(sizeof(BandTask) * (size_t)n));
        CK(ctx->d_leafout.reserve(sizeof(LeafOut) * (size_t)n));
        CK(ctx->d_ops.reserve((size_t)std::max<i64>(ctx->ops_words_fused, 1) * 4 + 16));
        Span sp(ctx, ST_WL);
        k_win_build<<<nb256, 256, 0, ctx->stream>>>(ctx->d_pairs.as<PairRec>(), ni, W, (int)prm.overlap_size, prm.force_scalar ? 0 : 1, ctx->d_wintasks.as<WinTask>(),
                                                    ctx->d_leaves.as<BandTask>(), ctx->d_pairleaves.as<PairLeaves>(), ctx->d_status.as<int>(), QUICKED_WIP);
        k_windowed_tiles<false><<<blocks, kWtThreads, 0, ctx->stream>>>(ctx->d_wintasks.as<WinTask>(), ni, ctx->d_codes.as<unsigned char>(), ctx->d_peq.as<u64>(),
            ctx->d_wintile.as<ulonglong2>(), rec_tiles, W, ctx->d_ops.as<u32>(), ctx->d_winout.as<WinOut>(), ctx->d_leafout.as<LeafOut>(), ctx->d_counters.as<u64>());
        k_win_done<<<nb256, 256, 0, ctx->stream>>>(ctx->d_pairs.as<PairRec>(), ni, ctx->d_winout.as<WinOut>(), ctx->d_done.as<unsigned char>(), ctx->d_counters.as<u64>());
        CK(cudaGetLastError());
        ctx->stats.kernel_launches += 3;
        ctx->tile_walks = true;                     // text lengths are measured by k_cigar_text
        leaf_base = n; ops_base = ctx->ops_words_fused;
    }
    const bool have_done = use_fused || win_fast;   // d_done marks the pairs that are finished already
    // ---- QUICKED stage 1: WindowEd(S) bound (quicked.c:178-199) ----
    if (prm.algo == QUICKED && !use_fused) {
        Span sp(ctx, ST_WS);
        if (!ctx->sms) { ctx->sms = 148; cudaDeviceGetAttribute(&ctx->sms, cudaDevAttrMultiProcessorCount, ctx->device); }
        const int sms = ctx->sms;
        int ctas = kWsResidentCtas;
        // one pair per thread and every pair takes about as long: a batch that fits five CTAs per SM in ONE wave runs
        // better that way than as a full wave plus a nearly empty one (100 k pairs: 10.7 vs 11.2 ms)
        if (n <= (i64)sms * 5 * kWsThreads * 11 / 10) ctas = 5;
        if (const char *e = getenv("QB200_WS_CTAS")) ctas = std::max(1, std::min(atoi(e), (int)kWsCtasPerSm));
        // Batches with texts of >= 128 characters have full windows: the SLIM kernel keeps their quadrants in shared
        // memory (43.5 KB per CTA).  Shorter reads only ever run non-full windows, whose quadrant lives in the L2
        // scratch: there the 10 KB kernel with most of the SM left as L1 is the faster one (C1: 4.2 vs 5.6 ms per 4 M).
        const bool slim = ctx->max_n >= 128;
        // COMPACT residency (704 pairs per SM, qb_windowed.cuh) when that turns two waves into one
        const i64 cap = (i64)sms * ctas * kWsThreads, cap_c = (i64)sms * 2 * kWsCompactThreads;
        bool compact = slim && n > cap && n <= cap_c;
        if (const char *e = getenv("QB200_WS_COMPACT")) compact = slim && atoi(e) != 0;
        const int T = compact ? kWsCompactThreads : kWsThreads;
        const int blocks = (int)std::min<i64>((n + T - 1) / T, compact ? (i64)sms * 2 : (i64)sms * ctas);   // persistent: one wave
        // the plain kernel: all pairs, or (after the COMPACT one) the pairs that one skipped
        const int blocks_p = compact ? (int)std::min<i64>((n + kWsThreads - 1) / kWsThreads, (i64)sms) : blocks;
        CK(ctx->d_quad.reserve((size_t)std::max<i64>((i64)blocks * T, (i64)blocks_p * kWsThreads) * kWsQuadSlots * 8));
        auto kern = prm.force_scalar ? (slim ? k_windowed21_score<false, true> : k_windowed21_score<false, false>)
                                     : (slim ? k_windowed21_score<true, true> : k_windowed21_score<true, false>);
        auto kern_c = prm.force_scalar ? k_windowed21_score<false, true, true> : k_windowed21_score<true, true, true>;
        const size_t smem = ws_smem_bytes(slim, kWsThreads, kAlpha), smem_c = ws_smem_bytes(true, kWsCompactThreads, 4);
        {   // carve out what the resident CTAs need (+1 KB per CTA of system use), the rest stays L1
            int carve = std::min(100, (ctas * (slim ? 45 : 11) * 100 + 227) / 228 + 1);
            if (const char *e = getenv("QB200_WS_CARVE")) carve = atoi(e);
            if (ctx->ws_carve_set[prm.force_scalar ? 1 : 0][slim ? 1 : 0] != carve) {      // once per kernel variant, not per run
                cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
                ctx->ws_carve_set[prm.force_scalar ? 1 : 0][slim ? 1 : 0] = carve;
            }
            if (compact && !ctx->ws_compact_set[prm.force_scalar ? 1 : 0]) {
                CK(cudaFuncSetAttribute(kern_c, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
                cudaFuncSetAttribute(kern_c, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
                ctx->ws_compact_set[prm.force_scalar ? 1 : 0] = true;
            }
        }
        if (compact) {
            kern_c<<<blocks, kWsCompactThreads, smem_c, ctx->stream>>>(ctx->d_pairs.as<PairRec>(), ni, ctx->d_codes.as<unsigned char>(), ctx->raw(),
                ctx->d_peq.as<u64>(), (int)prm.hew_threshold[0], ctx->d_bound.as<int>(), ctx->d_hew.as<int>(), ctx->d_counters.as<u64>(),
                ctx->d_quad.as<u64>(), ctx->d_pairodd.as<unsigned char>(), 0);
            ctx->stats.kernel_launches++;
        }
        kern<<<blocks_p, kWsThreads, smem, ctx->stream>>>(ctx->d_pairs.as<PairRec>(), ni, ctx->d_codes.as<unsigned char>(), ctx->raw(),
            ctx->d_peq.as<u64>(), (int)prm.hew_threshold[0], ctx->d_bound.as<int>(), ctx->d_hew.as<int>(), ctx->d_counters.as<u64>(),
            ctx->d_quad.as<u64>(), ctx->d_pairodd.as<unsigned char>(), compact ? 1 : 0);
        CK(cudaGetLastError());
        ctx->stats.kernel_launches++;
    }

    rt.pt("ws launched");
    // ---- plan: classify every pair and lay out the pools with one scan ----
    RunPlan plan;
    {
        Span sp(ctx, ST_PLAN);
        PlanParams pp;
        pp.algo = (int)prm.algo; pp.bandwidth = prm.bandwidth; pp.hew_pct0 = prm.hew_percentage[0];
        pp.only_score = prm.only_score; pp.thread_band_max = ctx->thread_band_max;
        pp.ok_status = (prm.algo == HIRSCHBERG) ? QUICKED_OK : QUICKED_WIP;
        pp.tiles = ctx->use_tiles ? 1 : 0;
        k_plan<<<nb256, 256, 0, ctx->stream>>>(ctx->d_pairs.as<PairRec>(), ni, pp, ctx->d_bound.as<int>(), ctx->d_hew.as<int>(),
                                               ctx->d_plan_items.as<PlanSum>(), ctx->d_cls.as<unsigned char>(), ctx->d_cutoff.as<i64>(),
                                               ctx->d_status.as<int>(), ctx->d_score.as<int>(), have_done ? ctx->d_done.as<unsigned char>() : nullptr,
                                               reinterpret_cast<unsigned *>(ctx->d_counters.as<u64>() + 28));
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
        CK(cudaMemcpyAsync(ctx->h_pinned + 192, ctx->d_counters.as<u64>() + 28, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        plan.tot = *reinterpret_cast<PlanSum *>(ctx->h_pinned);
        plan.tile_mask = *reinterpret_cast<unsigned *>(ctx->h_pinned + 192);
    }
    const PlanSum &tot = plan.tot;

    // ---- slow path first (it appends its leaves after the fast ones and returns its pool usage) ----
    i64 n_leaves = leaf_base + tot.leaf, ops_words = ops_base + tot.ops, range_ints = tot.rng;
    std::vector<PairLeaves> slow_pl;
    std::vector<int> slow_pairs;
    if (tot.slow > 0) {
        slow_pairs.resize((size_t)tot.slow);
    }
    CK(ctx->d_leaves.grow_keep(sizeof(BandTask) * (size_t)std::max<i64>(n_leaves, 1), sizeof(BandTask) * (size_t)leaf_base, ctx->stream));
    {   // every pair not finished by the fused kernel gets its leaf list (possibly empty) here
        k_build_leaves<<<nb256, 256, 0, ctx->stream>>>(ctx->d_pairs.as<PairRec>(), ni, ctx->d_cls.as<unsigned char>(), ctx->d_cutoff.as<i64>(),
                                                       ctx->d_plan_offs.as<PlanSum>(), ctx->d_leaves.as<BandTask>(), ctx->d_list_t.as<int>(),
                                                       ctx->d_list_w.as<int>(), ctx->d_list_slow.as<int>(), ctx->d_pairleaves.as<PairLeaves>(),
                                                       have_done ? ctx->d_done.as<unsigned char>() : nullptr, leaf_base, ops_base);
        CK(cudaGetLastError());
        ctx->stats.kernel_launches++;
    }
    if (tot.slow > 0) {
        CK(cudaMemcpyAsync(slow_pairs.data(), ctx->d_list_slow.p, (size_t)tot.slow * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    // thread-kernel leaves: with the tile path they write tile records like everybody else (qb_banded.cuh, REC) and need
    // no interleaved 32-leaf groups; the full-matrix mode (QB200_TILES=0) keeps the groups
    const bool trec = ctx->use_tiles;
    if (tot.t > 0 && !trec) {
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
    CK(ctx->d_ops.grow_keep((size_t)std::max<i64>(ops_words, 1) * 4 + 16, (size_t)ops_base * 4, ctx->stream));
    CK(ctx->d_ranges.reserve((size_t)std::max<i64>(range_ints, 1) * 8 + 16));
    CK(ctx->d_leafout.grow_keep(sizeof(LeafOut) * (size_t)std::max<i64>(n_leaves, 1), sizeof(LeafOut) * (size_t)leaf_base, ctx->stream));
    CK(ctx->d_bandout.reserve(sizeof(BandOut) * (size_t)std::max<i64>(n_leaves, 1)));
    if (tot.sc > 0) {
        CK(ctx->d_scores.reserve((size_t)tot.sc * 4 + 16));
        CK(cudaMemsetAsync(ctx->d_scores.p, 0, (size_t)tot.sc * 4, ctx->stream));
    }

    rt.pt("plan synced");
    // ---- fast path: fill + traceback, chunked only if the traceback state exceeds the pool ----
    if (tot.leaf > 0) {
        const i64 need = tot.matw + plan.mat_t;
        // cudaMemGetInfo is a slow, cross-process serialising driver call: only ask when the pool has to grow
        size_t free_b = 0, total_b = 0;
        if ((size_t)need * 16 > ctx->d_matrix.cap) CK(cudaMemGetInfo(&free_b, &total_b));
        const i64 limit = (i64)(std::min<size_t>(ctx->matrix_limit, std::max(ctx->d_matrix.cap, (size_t)((free_b + ctx->d_matrix.cap) * 0.85))) / 16);
        const bool tiles = ctx->use_tiles && (plan.tile_mask & 0xffu);
        const bool wide = !ctx->use_tiles || (plan.tile_mask >> 31);          // leaves for the full-matrix warp kernels
        const int min_B = ctx->use_tiles ? kTileBandMax + 1 : 0;
        BandTask *d_lv = ctx->d_leaves.as<BandTask>();
        const u64 *d_pq = ctx->d_peq.as<u64>();
        // fill + traceback of thread-class leaves [t0, t0+tn) and wider leaves [w0, w0+wn) whose traceback state sits in the pool
        auto fill_and_trace = [&](int t0, int tn, int w0, int wn, i64 sub_t, i64 sub_w) -> int {
            {
                Span sp(ctx, ST_FILL);
                int rc = 0;
                if (trec && tn > 0) rc = tile_prepare<true>(ctx, d_lv, ctx->d_list_t.as<int>(), t0, tn, true, false);
                if (!rc) rc = launch_thread_fill(ctx, ctx->d_list_t.as<int>(), t0, tn, sub_t, nullptr, trec);
                if (!rc && tiles && wn > 0) rc = launch_tiles<true>(ctx, d_lv, ctx->d_list_w.as<int>(), w0, wn, sub_w, d_pq, plan.tile_mask & 0xffu, !(trec && tn > 0));
                if (!rc && wide && wn > 0) rc = launch_banded<true>(ctx, 127u, d_lv, ctx->d_list_w.as<int>(), w0, wn, sub_w, nullptr, min_B);
                if (rc) return rc;
            }
            {
                Span sp(ctx, ST_TRACE);
                int rc = 0;
                if (trec) rc = launch_tile_traceback(ctx, ctx->d_list_t.as<int>(), t0, tn, sub_t, d_pq, true);
                else rc = launch_traceback(ctx, ctx->d_list_t.as<int>(), t0, tn, sub_t, false);
                if (!rc && tiles) rc = launch_tile_traceback(ctx, ctx->d_list_w.as<int>(), w0, wn, sub_w, d_pq);
                if (!rc && wide) rc = launch_traceback(ctx, ctx->d_list_w.as<int>(), w0, wn, sub_w, true, min_B);
                if (rc) return rc;
            }
            if ((tiles && wn > 0) || (trec && tn > 0)) return rerun_punted_leaves(ctx, d_pq);
            return 0;
        };
        if (need <= limit) {
            CK(ctx->d_matrix.reserve((size_t)need * 16));
            const int rc = fill_and_trace(0, (int)tot.t, 0, (int)tot.w, 0, 0);
            if (rc) return rc;
            ctx->stats.matrix_bytes += need * 16;
        } else {
            CK(ctx->d_matrix.reserve((size_t)limit * 16));
            int *d_idx = reinterpret_cast<int *>(ctx->d_counters.as<u64>() + 20);
            i64 *d_ent = reinterpret_cast<i64 *>(ctx->d_counters.as<u64>() + 21);
            i64 *d_off = reinterpret_cast<i64 *>(ctx->d_counters.as<u64>() + 22);
            // a single leaf / thread group can be larger than the pool limit: grow the pool for that chunk if the device
            // has the memory, fail with QB200_ERR_OOM otherwise (never launch a fill past the end of the pool)
            auto fit_chunk = [&](i64 ent) -> int {
                if ((size_t)ent * 16 <= ctx->d_matrix.cap) return 0;
                size_t fb = 0, tb = 0;
                if (cudaMemGetInfo(&fb, &tb) != cudaSuccess) { (void)cudaGetLastError(); fb = 0; }
                if ((size_t)ent * 16 > (size_t)((fb + ctx->d_matrix.cap) * 0.95)) {
                    ctx->err = "a single traceback matrix (" + std::to_string((long long)ent * 16) + " bytes) does not fit the workspace";
                    return QB200_ERR_OOM;
                }
                cudaError_t e = ctx->d_matrix.reserve((size_t)ent * 16);
                if (e != cudaSuccess) { (void)cudaGetLastError(); ctx->err = "out of device memory for a single traceback matrix"; return QB200_ERR_OOM; }
                return 0;
            };
            // leaves whose state is laid out leaf by leaf: the wider ones, and with tile records the thread-class ones too
            for (int pass = 0; pass < 2; ++pass) {
                const bool thr = pass == 1;
                if (thr && !trec) break;
                const int *lst = thr ? ctx->d_list_t.as<int>() : ctx->d_list_w.as<int>();
                const int cnt = (int)(thr ? tot.t : tot.w);
                for (int s0 = 0; s0 < cnt;) {
                    k_chunk_end_leaves<<<1, 1, 0, ctx->stream>>>(d_lv, lst, cnt, s0, limit, d_idx, d_off, d_ent, ctx->use_tiles ? 1 : 0);
                    CK(cudaMemcpyAsync(ctx->h_pinned, d_idx, 24, cudaMemcpyDeviceToHost, ctx->stream));
                    CK(cudaStreamSynchronize(ctx->stream));
                    const int s1 = *reinterpret_cast<int *>(ctx->h_pinned);
                    const i64 sub = *reinterpret_cast<i64 *>(ctx->h_pinned + 16);
                    { const int rc = fit_chunk(*reinterpret_cast<i64 *>(ctx->h_pinned + 8)); if (rc) return rc; }
                    const int rc = thr ? fill_and_trace(s0, s1 - s0, 0, 0, sub, 0) : fill_and_trace(0, 0, s0, s1 - s0, 0, sub);
                    if (rc) return rc;
                    s0 = s1;
                }
            }
            // thread-kernel groups (full-matrix mode)
            for (int g0 = 0; g0 < plan.n_groups;) {
                k_chunk_end_groups<<<1, 1, 0, ctx->stream>>>(ctx->d_goff.as<i64>(), ctx->d_gsize.as<i64>(), plan.n_groups, g0, limit, d_idx, d_ent);
                CK(cudaMemcpyAsync(ctx->h_pinned, d_idx, 16, cudaMemcpyDeviceToHost, ctx->stream));
                CK(cudaMemcpyAsync(ctx->h_pinned + 16, ctx->d_goff.as<i64>() + g0, 8, cudaMemcpyDeviceToHost, ctx->stream));
                CK(cudaStreamSynchronize(ctx->stream));
                const int g1 = *reinterpret_cast<int *>(ctx->h_pinned);
                { const int rc = fit_chunk(*reinterpret_cast<i64 *>(ctx->h_pinned + 8)); if (rc) return rc; }
                const i64 sub = tot.matw + *reinterpret_cast<i64 *>(ctx->h_pinned + 16);
                const int q0 = g0 * 32, q1 = (int)std::min<i64>(tot.t, (i64)g1 * 32);
                const int rc = fill_and_trace(q0, q1 - q0, 0, 0, sub, 0);
                if (rc) return rc;
                g0 = g1;
            }
            ctx->stats.matrix_bytes += need * 16;
        }
        ctx->stats.leaves += tot.leaf;
    }

    rt.pt("fill+trace launched");
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
            // pairs of >= kLongOps op slots (and multi-leaf pairs) get a warp each for the text passes
            const bool long_pairs = !getenv("QB200_TEXT_THREAD") && (ctx->multi_leaf_pairs || (i64)ctx->max_m + ctx->max_n >= kLongOps);
            if (ctx->multi_leaf_pairs || ctx->tile_walks) {
                CK(ctx->d_textlen.reserve((size_t)n * 4));
                k_cigar_text<false><<<(ni + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_pairleaves.as<PairLeaves>(), ni, ctx->d_leaves.as<BandTask>(),
                    ctx->d_leafout.as<LeafOut>(), ctx->d_ops.as<u32>(), ctx->d_textlen.as<int>(), nullptr, nullptr, long_pairs ? 1 : 0);
                if (long_pairs) {
                    k_cigar_text_warp<false><<<(ni + 3) / 4, 128, 0, ctx->stream>>>(ctx->d_pairleaves.as<PairLeaves>(), ni, ctx->d_leaves.as<BandTask>(),
                        ctx->d_leafout.as<LeafOut>(), ctx->d_ops.as<u32>(), ctx->d_textlen.as<int>(), nullptr, nullptr);
                    ctx->stats.kernel_launches++;
                }
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
                    ctx->d_leafout.as<LeafOut>(), ctx->d_ops.as<u32>(), nullptr, ctx->d_cigoff.as<i64>(), ctx->d_cigar.as<char>(), long_pairs ? 1 : 0);
                if (long_pairs) {
                    k_cigar_text_warp<true><<<(ni + 3) / 4, 128, 0, ctx->stream>>>(ctx->d_pairleaves.as<PairLeaves>(), ni, ctx->d_leaves.as<BandTask>(),
                        ctx->d_leafout.as<LeafOut>(), ctx->d_ops.as<u32>(), nullptr, ctx->d_cigoff.as<i64>(), ctx->d_cigar.as<char>());
                    ctx->stats.kernel_launches++;
                }
                CK(cudaGetLastError());
                ctx->stats.kernel_launches++;
            }
            ctx->have_cigar = true;
        }
    }
    rt.pt("text launched");
    CK(cudaEventRecord(ev_end, ctx->stream));
    // word-step counters: through the pinned scratch on this stream (a synchronous cudaMemcpy on the legacy stream
    // queued behind the other pipeline workers' copies: 3 ms per call)
    CK(cudaMemcpyAsync(ctx->h_pinned + 64, ctx->d_counters.p, 32, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    rt.pt("final sync");

    // ---- stats ----
    u64 counters[4];
    memcpy(counters, ctx->h_pinned + 64, 32);
    if (getenv("QB200_TILE_DEBUG")) {
        u64 c[16];
        cudaMemcpy(c, ctx->d_counters.p, 128, cudaMemcpyDeviceToHost);
        fprintf(stderr, "[qb200 tiles] sched cycles %llu of %llu (%.1f%%), passes %llu, %.0f cycles/pass, %.0f sched cycles/pass\n", (unsigned long long)c[8],
                (unsigned long long)c[9], c[9] ? 100.0 * c[8] / c[9] : 0.0, (unsigned long long)c[10], c[10] ? (double)c[9] / c[10] : 0.0,
                c[10] ? (double)c[8] / c[10] : 0.0);
    }
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
    ctx->stats.ms_fused = st[ST_FUSED];
    ctx->stats.pairs_fused = (i64)counters[2];
    ctx->stats.leaves += (i64)counters[2];
    ctx->ran = true;
    return 0;
}

namespace {

struct HNode { int pair; i64 p_off, t_off; int m, n; i64 cutoff; };

// Build match-mask tables for `jobs` in the slow-path pool (d_peq2); fills job.peq_off.
int build_tables(qb200_ctx *ctx, std::vector<PeqJob> &jobs)
{
    if (jobs.empty()) return 0;
    i64 words = 0;
    for (auto &j : jobs) { j.peq_off = words; words += (i64)kPeqStride * ((j.m + 63) / 64 + 2); }
    CK(ctx->d_peq2.reserve((size_t)words * 8 + 64));
    CK(ctx->d_jobs2.reserve(sizeof(PeqJob) * jobs.size()));
    CK(cudaMemcpyAsync(ctx->d_jobs2.p, jobs.data(), sizeof(PeqJob) * jobs.size(), cudaMemcpyHostToDevice, ctx->stream));
    const int nj = (int)jobs.size();
    k_build_peq<<<(nj + 7) / 8, 256, 0, ctx->stream>>>(ctx->d_jobs2.as<PeqJob>(), nj, ctx->d_codes.as<unsigned char>(), ctx->d_peq2.as<u64>(), nullptr);
    CK(cudaGetLastError());
    ctx->stats.kernel_launches++;
    return 0;
}

// Score-only BandEd passes for host-built tasks (tables in d_peq2).  Leaves the exported state / scores / BandOut
// resident for a following combine.  tasks[i].slot = i.
int run_score_tasks(qb200_ctx *ctx, std::vector<BandTask> &tasks, std::vector<BandOut> &outs, int stage)
{
    const size_t nt = tasks.size();
    outs.resize(nt);
    if (!nt) return 0;
    i64 st = 0, sc = 0;
    unsigned mask = 0, tmask = 0;          // band-height classes for the sweep kernels / the tile kernels
    for (size_t i = 0; i < nt; ++i) {
        BandTask &t = tasks[i];
        const BandGeom g = band_geometry(t.m, t.n, t.cutoff);
        if (ctx->use_tiles && tile_band_ok(g.Bs)) tmask |= tile_class_bit(g.Bs);
        else {
            const int R = rounds_for(g.Bs);
            if (!R) { ctx->err = "score-only band of " + std::to_string(g.Bs) + " blocks exceeds the supported 11000"; return QB200_ERR_ARG; }
            mask |= (unsigned)R;
        }
        t.slot = (int)i; t.state_off = st; t.scores_off = sc;
        st += 2 * g.Bs; sc += (i64)((t.m + 63) / 64) + g.Bs + 2;
    }
    CK(ctx->d_state.reserve((size_t)st * 8 + 16));
    CK(ctx->d_scores.reserve((size_t)sc * 4 + 16));
    CK(cudaMemsetAsync(ctx->d_scores.p, 0, (size_t)sc * 4, ctx->stream));
    CK(ctx->d_bandout.reserve(sizeof(BandOut) * nt));
    CK(ctx->d_tasks2.reserve(sizeof(BandTask) * nt));
    CK(cudaMemcpyAsync(ctx->d_tasks2.p, tasks.data(), sizeof(BandTask) * nt, cudaMemcpyHostToDevice, ctx->stream));
    {
        Span sp(ctx, stage);
        int rc = 0;
        if (tmask) rc = launch_tiles<false>(ctx, ctx->d_tasks2.as<BandTask>(), nullptr, 0, (int)nt, 0, ctx->d_peq2.as<u64>(), tmask);
        if (!rc && mask) rc = launch_banded<false>(ctx, mask, ctx->d_tasks2.as<BandTask>(), nullptr, 0, (int)nt, 0, ctx->d_peq2.as<u64>(),
                                                   ctx->use_tiles ? kTileBandMax + 1 : 0);
        if (rc) return rc;
    }
    if (tmask) {        // passes the tile kernels gave up on (band ran empty): the sweep kernels redo them
        CK(cudaMemcpyAsync(ctx->h_pinned + 128, &ctx->d_tctl.as<TileCtl>()->punt_count, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        const int cnt = *reinterpret_cast<int *>(ctx->h_pinned + 128);
        if (cnt > 0) {
            Span sp(ctx, stage);
            int rc = launch_banded<false>(ctx, 127u, ctx->d_tasks2.as<BandTask>(), ctx->d_punt.as<int>(), 0, cnt, 0, ctx->d_peq2.as<u64>());
            if (rc) return rc;
        }
    }
    CK(cudaMemcpyAsync(outs.data(), ctx->d_bandout.p, sizeof(BandOut) * nt, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int run_win_tasks(qb200_ctx *ctx, std::vector<WinTask> &tasks, std::vector<WinOut> &outs, int stage, const u64 *peq_base = nullptr)
{
    if (!peq_base) peq_base = ctx->d_peq2.as<u64>();
    const size_t nt = tasks.size();
    outs.resize(nt);
    if (!nt) return 0;
    // Tile kernel first (one pair per thread, qb_wintile.cuh); the warp kernel then redoes the tasks it flagged: pairs with
    // characters outside "ACGTN" and 2-word windows with the SSE quirks.  The warp kernel stores whole windows, so its
    // scratch is only reserved for a bounded number of tasks at a time.
    i64 scr = 0;
    int Wmax = 1;
    for (size_t i = 0; i < nt; ++i) { tasks[i].slot = (int)i; tasks[i].scratch_off = scr; scr += (i64)(64 * tasks[i].W + 3) * tasks[i].W; Wmax = std::max(Wmax, tasks[i].W); }
    CK(ctx->d_wintasks.reserve(sizeof(WinTask) * nt));
    CK(ctx->d_winout.reserve(sizeof(WinOut) * nt));
    CK(cudaMemcpyAsync(ctx->d_wintasks.p, tasks.data(), sizeof(WinTask) * nt, cudaMemcpyHostToDevice, ctx->stream));
    const bool tiles = ctx->use_tiles && Wmax <= 32;
    if (tiles) {
        if (!ctx->sms) { ctx->sms = 148; cudaDeviceGetAttribute(&ctx->sms, cudaDevAttrMultiProcessorCount, ctx->device); }
        const int blocks = (int)std::min<i64>(((i64)nt + kWtThreads - 1) / kWtThreads, (i64)ctx->sms * 4);
        const i64 nthr = (i64)blocks * kWtThreads, rec_tiles = (i64)Wmax * Wmax;
        CK(ctx->d_wintile.reserve((size_t)((2 * rec_tiles + Wmax) * nthr) * 16 + 64));
        Span sp(ctx, stage);
        if (tasks[0].score_only)
            k_windowed_tiles<true><<<blocks, kWtThreads, 0, ctx->stream>>>(ctx->d_wintasks.as<WinTask>(), (int)nt, ctx->d_codes.as<unsigned char>(), peq_base,
                ctx->d_wintile.as<ulonglong2>(), rec_tiles, Wmax, ctx->d_ops.as<u32>(), ctx->d_winout.as<WinOut>(), ctx->d_leafout.as<LeafOut>(), ctx->d_counters.as<u64>());
        else
            k_windowed_tiles<false><<<blocks, kWtThreads, 0, ctx->stream>>>(ctx->d_wintasks.as<WinTask>(), (int)nt, ctx->d_codes.as<unsigned char>(), peq_base,
                ctx->d_wintile.as<ulonglong2>(), rec_tiles, Wmax, ctx->d_ops.as<u32>(), ctx->d_winout.as<WinOut>(), ctx->d_leafout.as<LeafOut>(), ctx->d_counters.as<u64>());
        CK(cudaGetLastError());
        ctx->stats.kernel_launches++;
        ctx->tile_walks = true;                     // their text length is measured by k_cigar_text
    }
    {
        // how many tasks does the warp kernel have to redo?  (none on ACGTN data with W > 2)
        size_t redo = nt;
        if (tiles) {
            CK(cudaMemcpyAsync(outs.data(), ctx->d_winout.p, sizeof(WinOut) * nt, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            redo = 0;
            for (size_t i = 0; i < nt; ++i) redo += outs[i].hew == kWinPunted;
        }
        if (redo) {
            CK(ctx->d_winscratch.reserve((size_t)scr * 16 + 64));
            Span sp(ctx, stage);
            k_windowed_warp<<<(int)((nt + 3) / 4), 128, 0, ctx->stream>>>(ctx->d_wintasks.as<WinTask>(), (int)nt, ctx->d_codes.as<unsigned char>(), ctx->raw(),
                                                                           peq_base, ctx->d_winscratch.as<ulonglong2>(), ctx->d_ops.as<u32>(),
                                                                           ctx->d_winout.as<WinOut>(), ctx->d_leafout.as<LeafOut>(), ctx->d_counters.as<u64>(), tiles ? 1 : 0);
            CK(cudaGetLastError());
            ctx->stats.kernel_launches++;
        }
    }
    CK(cudaMemcpyAsync(outs.data(), ctx->d_winout.p, sizeof(WinOut) * nt, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// Full-matrix fill + traceback of host-built leaves that already sit in d_leaves[L0 .. L0+leaves.size()).
int run_leaves_host(qb200_ctx *ctx, std::vector<BandTask> &leaves, i64 L0)
{
    const size_t nl = leaves.size();
    if (!nl) return 0;
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    const i64 limit = (i64)(std::min<size_t>(ctx->matrix_limit, (size_t)((free_b + ctx->d_matrix.cap) * 0.85)) / 16);
    std::vector<int> Bc(nl), list_t, list_w, list_x;      // thread kernel / full-matrix warp kernel / tile kernels
    i64 rg = 0, sc = 0;
    unsigned xmask = 0;
    for (size_t i = 0; i < nl; ++i) {
        BandTask &t = leaves[i];
        const BandGeom g = band_geometry(t.m, t.n, t.cutoff);
        Bc[i] = (int)g.Bc;
        if (g.Bc > ctx->thread_band_max && !rounds_for(g.Bc)) { ctx->err = "leaf band of " + std::to_string(g.Bc) + " blocks exceeds the supported 11000"; return QB200_ERR_ARG; }
        if (g.Bc <= ctx->thread_band_max) list_t.push_back((int)i);
        else if (ctx->use_tiles && tile_band_ok(g.Bc)) { list_x.push_back((int)i); xmask |= tile_class_bit(g.Bc); }
        else list_w.push_back((int)i);
        t.range_off = rg; rg += t.n / 64 + 2;
        t.scores_off = sc; if (g.Bc > ctx->thread_band_max || ctx->use_tiles) sc += (i64)((t.m + 63) / 64) + g.Bc + 2;
    }
    const bool trec = ctx->use_tiles;       // thread-class leaves write tile records too (qb_banded.cuh, REC)
    CK(ctx->d_punt.reserve((size_t)std::max<i64>(ctx->n_pairs, (i64)nl) * 4 + 16));
    // chunk plans: (kind: 1 thread kernel, 0 warp kernel, 2 tile kernels; list begin, list end, entries)
    struct Chunk { int thr, q0, q1; i64 ent; };
    std::vector<Chunk> chunks;
    for (size_t q0 = 0; q0 < list_x.size();) {                        // tile leaves: one 32-byte record per tile
        i64 ent = 0; size_t q1 = q0;
        while (q1 < list_x.size()) {
            BandTask &t = leaves[list_x[q1]];
            const i64 e = 2 * (i64)((t.n + 63) / 64) * Bc[list_x[q1]];
            if (q1 > q0 && ent + e > limit) break;
            t.mat_off = ent; t.mat_cs = Bc[list_x[q1]]; t.mat_ws = 1;
            ent += e; ++q1;
        }
        if ((size_t)ent * 16 > free_b + ctx->d_matrix.cap) { ctx->err = "the tile records of a single leaf do not fit the device"; return QB200_ERR_OOM; }
        chunks.push_back({2, (int)q0, (int)q1, ent});
        q0 = q1;
    }
    for (size_t q0 = 0; trec && q0 < list_t.size();) {              // thread-class leaves in tile-record mode
        i64 ent = 0; size_t q1 = q0;
        while (q1 < list_t.size()) {
            BandTask &t = leaves[list_t[q1]];
            const i64 e = 2 * (i64)((t.n + 63) / 64) * Bc[list_t[q1]];
            if (q1 > q0 && ent + e > limit) break;
            t.mat_off = ent; t.mat_cs = Bc[list_t[q1]]; t.mat_ws = 1;
            ent += e; ++q1;
        }
        chunks.push_back({3, (int)q0, (int)q1, ent});
        q0 = q1;
    }
    for (size_t g0 = 0; !trec && g0 < list_t.size();) {             // thread-kernel groups of 32 (full-matrix mode)
        i64 ent = 0; size_t q0 = g0;
        while (g0 < list_t.size()) {
            const size_t g1 = std::min(list_t.size(), g0 + 32);
            int Bg = 1, nmax = 1;
            for (size_t q = g0; q < g1; ++q) { Bg = std::max(Bg, Bc[list_t[q]]); nmax = std::max(nmax, leaves[list_t[q]].n); }
            const i64 gent = (i64)(nmax + 1) * Bg * 32;
            if (g0 > q0 && ent + gent > limit) break;
            for (size_t q = g0; q < g1; ++q) { BandTask &t = leaves[list_t[q]]; t.mat_off = ent + (i64)(q - g0); t.mat_cs = Bg * 32; t.mat_ws = 32; }
            ent += gent; g0 = g1;
        }
        chunks.push_back({1, (int)q0, (int)g0, ent});
    }
    for (size_t q0 = 0; q0 < list_w.size();) {                        // warp-kernel leaves
        i64 ent = 0; size_t q1 = q0;
        while (q1 < list_w.size()) {
            BandTask &t = leaves[list_w[q1]];
            const i64 e = (i64)(t.n + 1) * Bc[list_w[q1]];
            if (q1 > q0 && ent + e > limit) break;
            t.mat_off = ent; t.mat_cs = Bc[list_w[q1]]; t.mat_ws = 1;
            ent += e; ++q1;
        }
        if ((size_t)ent * 16 > free_b + ctx->d_matrix.cap) { ctx->err = "a single traceback matrix does not fit the device"; return QB200_ERR_OOM; }
        chunks.push_back({0, (int)q0, (int)q1, ent});
        q0 = q1;
    }
    for (int &v : list_t) v += (int)L0;
    for (int &v : list_w) v += (int)L0;
    for (int &v : list_x) v += (int)L0;
    CK(cudaMemcpyAsync(ctx->d_leaves.as<BandTask>() + L0, leaves.data(), sizeof(BandTask) * nl, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx->d_list_t.reserve(list_t.size() * 4 + 16));
    CK(ctx->d_list_w.reserve((list_w.size() + list_x.size()) * 4 + 16));
    CK(cudaMemcpyAsync(ctx->d_list_t.p, list_t.data(), list_t.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_list_w.p, list_w.data(), list_w.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    int *d_list_x = ctx->d_list_w.as<int>() + list_w.size();
    CK(cudaMemcpyAsync(d_list_x, list_x.data(), list_x.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx->d_ranges.reserve((size_t)rg * 8 + 16));
    CK(ctx->d_scores.reserve((size_t)sc * 4 + 16));
    CK(cudaMemsetAsync(ctx->d_scores.p, 0, (size_t)sc * 4 + 16, ctx->stream));
    for (const Chunk &c : chunks) {
        CK(ctx->d_matrix.reserve((size_t)c.ent * 16));
        if (c.thr == 3) {
            {
                Span sp(ctx, ST_FILL);
                int rc = tile_prepare<true>(ctx, ctx->d_leaves.as<BandTask>(), ctx->d_list_t.as<int>(), c.q0, c.q1 - c.q0, true, false);
                if (!rc) rc = launch_thread_fill(ctx, ctx->d_list_t.as<int>(), c.q0, c.q1 - c.q0, 0, ctx->d_peq2.as<u64>(), true);
                if (rc) return rc;
            }
            { Span sp(ctx, ST_TRACE); int rc = launch_tile_traceback(ctx, ctx->d_list_t.as<int>(), c.q0, c.q1 - c.q0, 0, ctx->d_peq2.as<u64>(), true); if (rc) return rc; }
            ctx->stats.matrix_bytes += c.ent * 16;
            int rc = rerun_punted_leaves(ctx, ctx->d_peq2.as<u64>());       // synchronises
            if (rc) return rc;
            continue;
        }
        if (c.thr == 2) {
            { Span sp(ctx, ST_FILL); int rc = launch_tiles<true>(ctx, ctx->d_leaves.as<BandTask>(), d_list_x, c.q0, c.q1 - c.q0, 0, ctx->d_peq2.as<u64>(), xmask); if (rc) return rc; }
            { Span sp(ctx, ST_TRACE); int rc = launch_tile_traceback(ctx, d_list_x, c.q0, c.q1 - c.q0, 0, ctx->d_peq2.as<u64>()); if (rc) return rc; }
            ctx->stats.matrix_bytes += c.ent * 16;
            int rc = rerun_punted_leaves(ctx, ctx->d_peq2.as<u64>());       // synchronises
            if (rc) return rc;
            continue;
        }
        const int *lst = c.thr ? ctx->d_list_t.as<int>() : ctx->d_list_w.as<int>();
        {
            Span sp(ctx, ST_FILL);
            int rc = c.thr ? launch_thread_fill(ctx, lst, c.q0, c.q1 - c.q0, 0, ctx->d_peq2.as<u64>())
                           : launch_banded<true>(ctx, 127u, ctx->d_leaves.as<BandTask>(), lst, c.q0, c.q1 - c.q0, 0, ctx->d_peq2.as<u64>());
            if (rc) return rc;
        }
        {
            Span sp(ctx, ST_TRACE);
            int rc = launch_traceback(ctx, lst, c.q0, c.q1 - c.q0, 0, !c.thr);
            if (rc) return rc;
        }
        ctx->stats.matrix_bytes += c.ent * 16;
        CK(cudaStreamSynchronize(ctx->stream));
    }
    ctx->stats.leaves += (i64)nl;
    return 0;
}

}  // namespace

static int run_slow_path(qb200_ctx *ctx, const quicked_params_t &prm, const std::vector<int> &slow_pairs,
                         i64 &n_leaves_total, i64 &ops_words_total, i64 &range_total, std::vector<PairLeaves> &slow_pl)
{
    (void)range_total;
    const size_t ns = slow_pairs.size();
    const i64 n = ctx->n_pairs;
    std::vector<SlowResult> res(ns);
    for (size_t q = 0; q < ns; ++q) {
        res[q].pair = slow_pairs[q]; res[q].status = QUICKED_ERROR; res[q].score = -1; res[q].set_score = 0;
        res[q].pl.first_leaf = 0; res[q].pl.n_leaves = 0; res[q].pl.pad_ = 0;
    }
    const int W = (int)prm.window_size, O = (int)prm.overlap_size;
    const bool sse = !prm.force_scalar;
    std::vector<i64> cutoff(ns, 0);
    std::vector<char> align(ns, 0);        // goes on to the Hirschberg stage
    const int ok_status = (prm.algo == HIRSCHBERG) ? QUICKED_OK : QUICKED_WIP;
    const i64 L0 = n_leaves_total;
    std::vector<BandTask> leaves;          // appended leaves (incl. WINDOWED pseudo-leaves), global index L0 + k
    i64 ops_words = ops_words_total;

    auto pair_of = [&](size_t q) -> const PairRec & { return ctx->h_pairs[(size_t)slow_pairs[q]]; };

    if (prm.algo == WINDOWED) {                                             // run_windowed, quicked.c:91-123
        std::vector<WinTask> wt;
        std::vector<size_t> who;
        wt.reserve(ns); who.reserve(ns); leaves.reserve(ns);
        for (size_t q = 0; q < ns; ++q) {
            const PairRec &r = pair_of(q);
            if (W < 1 || W > 32 || O < 0 || O >= W) { res[q].status = QUICKED_UNIMPLEMENTED; continue; }
            WinTask t{};
            t.p_off = r.p_off; t.t_off = r.t_off; t.m = r.m; t.n = r.n; t.rev = 0; t.W = W; t.O = O; t.hew_threshold = 0;
            t.sse = sse; t.score_only = prm.only_score; t.nbp = (r.m + 63) / 64 + 2;
            if (!prm.only_score) {
                BandTask lf{};
                lf.p_off = r.p_off; lf.t_off = r.t_off; lf.m = r.m; lf.n = r.n; lf.pair = slow_pairs[q];
                lf.ops_cap = ((r.m + r.n + 15) / 16) * 16; lf.ops_off = ops_words; ops_words += lf.ops_cap / 16;
                lf.slot = (int)(L0 + (i64)leaves.size());
                t.ops_off = lf.ops_off; t.ops_cap = lf.ops_cap; t.leaf_slot = lf.slot;
                res[q].pl.first_leaf = lf.slot; res[q].pl.n_leaves = 1;
                leaves.push_back(lf);
            }
            wt.push_back(t); who.push_back(q);
        }
        // forward patterns: the match masks of the prepare stage serve as they are (no second table build)
        for (size_t k = 0; k < wt.size(); ++k) wt[k].peq_off = pair_of(who[k]).peq_off;
        int rc = 0;
        CK(ctx->d_ops.grow_keep((size_t)std::max<i64>(ops_words, 1) * 4 + 16, (size_t)ops_words_total * 4, ctx->stream));
        CK(ctx->d_leaves.grow_keep(sizeof(BandTask) * (size_t)std::max<i64>(L0 + (i64)leaves.size(), 1), sizeof(BandTask) * (size_t)L0, ctx->stream));
        CK(ctx->d_leafout.grow_keep(sizeof(LeafOut) * (size_t)std::max<i64>(L0 + (i64)leaves.size(), 1), sizeof(LeafOut) * (size_t)L0, ctx->stream));
        if (!leaves.empty())
            CK(cudaMemcpyAsync(ctx->d_leaves.as<BandTask>() + L0, leaves.data(), sizeof(BandTask) * leaves.size(), cudaMemcpyHostToDevice, ctx->stream));
        std::vector<WinOut> wo;
        rc = run_win_tasks(ctx, wt, wo, ST_WL, ctx->d_peq.as<u64>());
        if (rc) return rc;
        for (size_t k = 0; k < wt.size(); ++k) {
            const size_t q = who[k];
            res[q].status = QUICKED_WIP;
            if (prm.only_score) { res[q].score = wo[k].score; res[q].set_score = 1; }
        }
        n_leaves_total = L0 + (i64)leaves.size();
        ops_words_total = ops_words;
    } else if (prm.algo == BANDED) {                                        // BANDED only_score (run_banded with SCORE_ONLY)
        std::vector<PeqJob> jobs;
        std::vector<BandTask> bt;
        for (size_t q = 0; q < ns; ++q) {
            const PairRec &r = pair_of(q);
            PeqJob j; j.src_off = r.p_off; j.m = r.m; j.rev = 0; j.peq_off = 0;
            jobs.push_back(j);
            BandTask t{};
            t.p_off = r.p_off; t.t_off = r.t_off; t.m = r.m; t.n = r.n; t.rev = 0; t.finish = r.n;
            t.cutoff = (i64)((unsigned)std::max(r.m, r.n) * prm.bandwidth / 100);
            t.nbp = (r.m + 63) / 64 + 2; t.pair = slow_pairs[q];
            bt.push_back(t);
        }
        int rc = build_tables(ctx, jobs);
        if (rc) return rc;
        for (size_t k = 0; k < bt.size(); ++k) bt[k].peq_off = jobs[k].peq_off;
        std::vector<BandOut> bo;
        rc = run_score_tasks(ctx, bt, bo, ST_BANDED);
        if (rc) return rc;
        for (size_t q = 0; q < ns; ++q) { res[q].status = QUICKED_WIP; res[q].score = bo[q].score; res[q].set_score = 1; }
        ctx->stats.banded_tries += (i64)ns;
    } else {
        // ---------------- QUICKED stages 2-3 (quicked.c:201-280) / HIRSCHBERG cutoffs ----------------
        if (prm.algo == QUICKED) {
            std::vector<int> hb((size_t)n), hh((size_t)n);
            CK(cudaMemcpyAsync(hb.data(), ctx->d_bound.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaMemcpyAsync(hh.data(), ctx->d_hew.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            std::vector<size_t> s2;
            for (size_t q = 0; q < ns; ++q) {
                const PairRec &r = pair_of(q);
                const unsigned maxlen = (unsigned)std::max(r.m, r.n);
                cutoff[q] = hb[(size_t)slow_pairs[q]];
                align[q] = 1;
                if ((i64)hh[(size_t)slow_pairs[q]] * 64 > (i64)(maxlen * prm.hew_percentage[0] / 100)) s2.push_back(q);
            }
            ctx->stats.pairs_stage2 += (i64)s2.size();
            if (!s2.empty() && (W < 1 || W > 32 || O < 0 || O >= W)) {
                for (size_t q : s2) { res[q].status = QUICKED_UNIMPLEMENTED; align[q] = 0; }
                s2.clear();
            }
            std::vector<size_t> s3;
            if (!s2.empty()) {
                std::vector<PeqJob> jobs;
                std::vector<WinTask> wt;
                for (size_t q : s2) {
                    const PairRec &r = pair_of(q);
                    for (int rev = 0; rev < 2; ++rev) {
                        PeqJob j; j.src_off = r.p_off; j.m = r.m; j.rev = rev; j.peq_off = 0;
                        jobs.push_back(j);
                        WinTask t{};
                        t.p_off = r.p_off; t.t_off = r.t_off; t.m = r.m; t.n = r.n; t.rev = rev; t.W = W; t.O = O;
                        t.hew_threshold = (int)prm.hew_threshold[1]; t.sse = sse; t.score_only = 1; t.nbp = (r.m + 63) / 64 + 2;
                        wt.push_back(t);
                    }
                }
                int rc = build_tables(ctx, jobs);
                if (rc) return rc;
                for (size_t k = 0; k < wt.size(); ++k) wt[k].peq_off = jobs[k].peq_off;
                std::vector<WinOut> wo;
                rc = run_win_tasks(ctx, wt, wo, ST_WL);
                if (rc) return rc;
                for (size_t k = 0; k < s2.size(); ++k) {
                    const size_t q = s2[k];
                    const PairRec &r = pair_of(q);
                    const unsigned maxlen = (unsigned)std::max(r.m, r.n);
                    i64 score = wo[2 * k].score;                                   // quicked.c:213-230
                    unsigned long long hewv = (unsigned long long)wo[2 * k].hew;
                    score = std::min<i64>(score, wo[2 * k + 1].score);
                    if (score >= wo[2 * k + 1].score) hewv = (unsigned long long)wo[2 * k + 1].hew;
                    cutoff[q] = score;
                    if (hewv * 64ull * (unsigned long long)(prm.window_size - prm.overlap_size) >
                        (unsigned long long)(maxlen * prm.hew_percentage[1] / 100)) s3.push_back(q);   // :237-238
                }
            }
            ctx->stats.pairs_stage3 += (i64)s3.size();
            if (!s3.empty()) {                                                     // band doubling, quicked.c:240-278
                std::vector<PeqJob> jobs;
                for (size_t q : s3) {
                    const PairRec &r = pair_of(q);
                    PeqJob j; j.src_off = r.p_off; j.m = r.m; j.rev = 0; j.peq_off = 0;
                    jobs.push_back(j);
                    const i64 bw = (i64)((unsigned)std::max(r.m, r.n) * prm.bandwidth / 100);
                    cutoff[q] = std::min<i64>(bw, cutoff[q]);                      // :246
                }
                int rc = build_tables(ctx, jobs);
                if (rc) return rc;
                std::vector<size_t> active(s3.size());
                for (size_t k = 0; k < s3.size(); ++k) active[k] = k;
                while (!active.empty()) {
                    std::vector<BandTask> bt;
                    for (size_t k : active) {
                        const size_t q = s3[k];
                        const PairRec &r = pair_of(q);
                        BandTask t{};
                        t.p_off = r.p_off; t.t_off = r.t_off; t.m = r.m; t.n = r.n; t.rev = 0; t.finish = r.n;
                        t.cutoff = cutoff[q]; t.peq_off = jobs[k].peq_off; t.nbp = (r.m + 63) / 64 + 2; t.pair = slow_pairs[q];
                        bt.push_back(t);
                    }
                    std::vector<BandOut> bo;
                    rc = run_score_tasks(ctx, bt, bo, ST_BANDED);
                    if (rc) return rc;
                    ctx->stats.banded_tries += (i64)bt.size();
                    std::vector<size_t> next;
                    for (size_t a = 0; a < active.size(); ++a) {
                        const size_t q = s3[active[a]];
                        const PairRec &r = pair_of(q);
                        const i64 nw = bo[a].score, maxlen = std::max(r.m, r.n);
                        if ((nw > maxlen / 4 && cutoff[q] * 3 / 2 < nw) || nw < 0) { cutoff[q] *= 2; next.push_back(active[a]); }   // :260-263
                        else cutoff[q] = nw;                                        // :278
                    }
                    active.swap(next);
                }
            }
        } else {   // HIRSCHBERG
            for (size_t q = 0; q < ns; ++q) {
                const PairRec &r = pair_of(q);
                cutoff[q] = (i64)((unsigned)std::max(r.m, r.n) * prm.bandwidth / 100);   // quicked.c:131
                align[q] = 1;
            }
        }

        // ---------------- Hirschberg: level-synchronous split tree (bpm_hirschberg.c:33-270) ----------------
        std::vector<HNode> cur, leaf_nodes;
        std::vector<char> bad(ns, 0);
        std::vector<i64> fail_t(ns, -1);          // largest t_off of a non-converging node (DFS-first failure)
        std::vector<int> q_of_pair_local;         // node.pair holds the LOCAL slow index q
        for (size_t q = 0; q < ns; ++q) {
            if (!align[q]) continue;
            const PairRec &r = pair_of(q);
            cur.push_back({(int)q, r.p_off, r.t_off, r.m, r.n, cutoff[q]});
        }
        while (!cur.empty()) {
            std::vector<HNode> splits;
            for (const HNode &nd : cur) {
                const BandGeom g = band_geometry(nd.m, nd.n, nd.cutoff);
                if ((unsigned long long)g.Bc * (unsigned long long)nd.n * 16ull > (1ull << 24)) {
                    if (!rounds_for(g.Bs)) { bad[(size_t)nd.pair] = 2; continue; }
                    splits.push_back(nd);
                } else leaf_nodes.push_back(nd);
            }
            cur.clear();
            if (splits.empty()) break;
            ctx->stats.hirschberg_splits += (i64)splits.size();
            std::vector<PeqJob> jobs;
            std::vector<BandTask> bt;
            for (const HNode &nd : splits) {
                const int n_l = (nd.n + 1) / 2;
                for (int rev = 0; rev < 2; ++rev) {
                    PeqJob j; j.src_off = nd.p_off; j.m = nd.m; j.rev = rev; j.peq_off = 0;
                    jobs.push_back(j);
                    BandTask t{};
                    t.p_off = nd.p_off; t.t_off = nd.t_off; t.m = nd.m; t.n = nd.n; t.rev = rev;
                    t.finish = rev ? nd.n - n_l : n_l; t.cutoff = nd.cutoff; t.nbp = (nd.m + 63) / 64 + 2; t.pair = nd.pair;
                    bt.push_back(t);
                }
            }
            int rc = build_tables(ctx, jobs);
            if (rc) return rc;
            for (size_t k = 0; k < bt.size(); ++k) bt[k].peq_off = jobs[k].peq_off;
            std::vector<BandOut> bo;
            rc = run_score_tasks(ctx, bt, bo, ST_FILL);
            if (rc) return rc;
            std::vector<SplitTask> stv(splits.size());
            i64 scr = 0;
            for (size_t k = 0; k < splits.size(); ++k) {
                const BandGeom g = band_geometry(splits[k].m, splits[k].n, splits[k].cutoff);
                SplitTask &s = stv[k];
                s.m = splits[k].m; s.n = splits[k].n; s.cutoff = splits[k].cutoff;
                s.fwd_slot = (int)(2 * k); s.rev_slot = (int)(2 * k + 1);
                s.fwd_state = bt[2 * k].state_off; s.rev_state = bt[2 * k + 1].state_off;
                s.fwd_scores = bt[2 * k].scores_off; s.rev_scores = bt[2 * k + 1].scores_off;
                s.scratch_off = scr; scr += 2 * (64 * g.Bs + 8);
            }
            CK(ctx->d_split.reserve(sizeof(SplitTask) * stv.size()));
            CK(ctx->d_splitout.reserve(sizeof(SplitOut) * stv.size()));
            CK(ctx->d_splitscratch.reserve((size_t)scr * 4 + 64));
            CK(cudaMemcpyAsync(ctx->d_split.p, stv.data(), sizeof(SplitTask) * stv.size(), cudaMemcpyHostToDevice, ctx->stream));
            k_hirschberg_combine<<<(int)((stv.size() + 63) / 64), 64, 0, ctx->stream>>>(ctx->d_split.as<SplitTask>(), (int)stv.size(), ctx->d_bandout.as<BandOut>(),
                ctx->d_state.as<u64>(), ctx->d_scores.as<int>(), ctx->d_splitscratch.as<int>(), ctx->d_splitout.as<SplitOut>());
            CK(cudaGetLastError());
            ctx->stats.kernel_launches++;
            std::vector<SplitOut> so(stv.size());
            CK(cudaMemcpyAsync(so.data(), ctx->d_splitout.p, sizeof(SplitOut) * so.size(), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            for (size_t k = 0; k < splits.size(); ++k) {
                const HNode &nd = splits[k];
                if (so[k].status == -2) { fail_t[(size_t)nd.pair] = std::max(fail_t[(size_t)nd.pair], nd.t_off); continue; }
                if (so[k].status != 0) { bad[(size_t)nd.pair] = 1; continue; }
                const int n_l = (nd.n + 1) / 2, m_l = so[k].m_l;
                cur.push_back({nd.pair, nd.p_off + m_l, nd.t_off + n_l, nd.m - m_l, nd.n - n_l, so[k].score_r});   // right first (:212-222)
                cur.push_back({nd.pair, nd.p_off, nd.t_off, m_l, n_l, so[k].score_l});
            }
        }
        // leaves, left to right per pair; after a non-converging node only the leaves to its right were emitted
        std::stable_sort(leaf_nodes.begin(), leaf_nodes.end(), [](const HNode &a, const HNode &b) {
            return a.pair != b.pair ? a.pair < b.pair : a.t_off < b.t_off; });
        std::vector<PeqJob> jobs;
        for (const HNode &nd : leaf_nodes) {
            const size_t q = (size_t)nd.pair;
            if (bad[q]) continue;
            if (fail_t[q] >= 0 && nd.t_off <= fail_t[q]) continue;
            PeqJob j; j.src_off = nd.p_off; j.m = nd.m; j.rev = 0; j.peq_off = 0;
            jobs.push_back(j);
            BandTask lf{};
            lf.p_off = nd.p_off; lf.t_off = nd.t_off; lf.m = nd.m; lf.n = nd.n; lf.rev = 0; lf.finish = nd.n; lf.cutoff = nd.cutoff;
            lf.nbp = (nd.m + 63) / 64 + 2; lf.pair = slow_pairs[q];
            lf.ops_cap = ((nd.m + nd.n + 15) / 16) * 16; lf.ops_off = ops_words;
            {   // 2-bit ops (thread walk, tile walk); u32 runs, worst case, for the full-matrix warp walk
                const i64 bc = band_geometry(nd.m, nd.n, nd.cutoff).Bc;
                ops_words += (bc <= ctx->thread_band_max || (ctx->use_tiles && tile_band_ok(bc))) ? lf.ops_cap / 16 : lf.ops_cap;
            }
            lf.slot = (int)(L0 + (i64)leaves.size());
            if (res[q].pl.n_leaves == 0) res[q].pl.first_leaf = lf.slot;
            res[q].pl.n_leaves++;
            if (res[q].pl.n_leaves > 1) ctx->multi_leaf_pairs = true;
            leaves.push_back(lf);
        }
        for (size_t q = 0; q < ns; ++q) {
            if (!align[q]) continue;
            if (bad[q] == 2) res[q].status = QUICKED_UNIMPLEMENTED;
            else if (bad[q]) res[q].status = QUICKED_ERROR;
            else if (fail_t[q] >= 0 && prm.algo == HIRSCHBERG) res[q].status = QUICKED_FAIL_NON_CONVERGENCE;   // quicked.c:160
            else res[q].status = ok_status;                                                                   // QUICKED ignores it (:290)
            if (!bad[q] && res[q].pl.n_leaves == 0) { res[q].score = 0; res[q].set_score = 1; }               // empty op list
        }
        if (L0 + (i64)leaves.size() > kMaxPairs) { ctx->err = "more than 2^31 alignment leaves in one batch: split it"; return QB200_ERR_ARG; }
        int rc = build_tables(ctx, jobs);
        if (rc) return rc;
        for (size_t k = 0; k < leaves.size(); ++k) leaves[k].peq_off = jobs[k].peq_off;
        CK(ctx->d_ops.grow_keep((size_t)std::max<i64>(ops_words, 1) * 4 + 16, (size_t)ops_words_total * 4, ctx->stream));
        CK(ctx->d_leaves.grow_keep(sizeof(BandTask) * (size_t)std::max<i64>(L0 + (i64)leaves.size(), 1), sizeof(BandTask) * (size_t)L0, ctx->stream));
        CK(ctx->d_leafout.grow_keep(sizeof(LeafOut) * (size_t)std::max<i64>(L0 + (i64)leaves.size(), 1), sizeof(LeafOut) * (size_t)L0, ctx->stream));
        CK(ctx->d_bandout.reserve(sizeof(BandOut) * (size_t)std::max<i64>(L0 + (i64)leaves.size(), 1)));
        rc = run_leaves_host(ctx, leaves, L0);
        if (rc) return rc;
        n_leaves_total = L0 + (i64)leaves.size();
        ops_words_total = ops_words;
    }
    // scatter per-pair results
    CK(ctx->d_scatter.reserve(sizeof(SlowResult) * ns));
    CK(cudaMemcpyAsync(ctx->d_scatter.p, res.data(), sizeof(SlowResult) * ns, cudaMemcpyHostToDevice, ctx->stream));
    k_scatter_slow<<<(int)((ns + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_scatter.as<SlowResult>(), (int)ns, ctx->d_pairleaves.as<PairLeaves>(),
                                                                      ctx->d_status.as<int>(), ctx->d_score.as<int>());
    CK(cudaGetLastError());
    ctx->stats.kernel_launches++;
    CK(cudaStreamSynchronize(ctx->stream));     // `res` is on the host stack
    slow_pl.clear();
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
    // Device-to-host copies into pageable memory are staged by the driver under a lock that also stalls the other
    // pipeline workers' launches, so everything goes through page-locked memory: the caller's buffer when it is
    // (qb200_host_alloc / cudaHostRegister), this context's staging buffer otherwise.
    auto pinned = [](const void *p) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
        return a.type == cudaMemoryTypeHost;
    };
    const bool with_cigar = ctx->have_cigar && res->cigar_off;
    int rc = 0;
    if (with_cigar) {
        res->cigar_bytes = ctx->cigar_total;
        if (!res->cigar || res->cigar_capacity < ctx->cigar_total) rc = QB200_ERR_CAPACITY;
    }
    const bool copy_text = with_cigar && rc == 0 && ctx->cigar_total > 0;
    const bool stage_text = copy_text && !pinned(res->cigar);
    const size_t small = (size_t)n * 8 + (size_t)(n + 1) * 8;
    if (!ctx->h_stage.resize(small + (stage_text ? (size_t)ctx->cigar_total : 0))) { ctx->err = "out of pinned host memory"; return QB200_ERR_OOM; }
    unsigned char *st = ctx->h_stage.data();
    int32_t *st_score = reinterpret_cast<int32_t *>(st), *st_status = st_score + n;
    int64_t *st_off = reinterpret_cast<int64_t *>(st + (size_t)n * 8);
    if (res->score) CK(cudaMemcpyAsync(st_score, ctx->d_score.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (res->status) CK(cudaMemcpyAsync(st_status, ctx->d_status.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->stats.d2h_bytes += n * 8;
    if (with_cigar) {
        CK(cudaMemcpyAsync(st_off, ctx->d_cigoff.p, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
        ctx->stats.d2h_bytes += (n + 1) * 8;
        if (copy_text) {
            CK(cudaMemcpyAsync(stage_text ? reinterpret_cast<char *>(st + small) : res->cigar, ctx->d_cigar.p, (size_t)ctx->cigar_total,
                               cudaMemcpyDeviceToHost, ctx->stream));
            ctx->stats.d2h_bytes += ctx->cigar_total;
        }
    }
    CK(cudaStreamSynchronize(ctx->stream));
    if (res->score) memcpy(res->score, st_score, (size_t)n * 4);
    if (res->status) memcpy(res->status, st_status, (size_t)n * 4);
    if (with_cigar) {
        memcpy(res->cigar_off, st_off, (size_t)(n + 1) * 8);
        if (stage_text) memcpy(res->cigar, st + small, (size_t)ctx->cigar_total);
    } else if (res->cigar_off) {
        for (i64 i = 0; i <= n; ++i) res->cigar_off[i] = 0;
    }
    return rc;
}

int qb200_get_stats(qb200_ctx_t *ctx, qb200_stats_t *stats)
{
    if (!ctx || !stats) return QB200_ERR_ARG;
    *stats = ctx->stats;
    return 0;
}

int qb200_get_bounds(qb200_ctx_t *ctx, int32_t *bound, int32_t *high_error_windows, int64_t n)
{
    if (!ctx || n != ctx->n_pairs) return QB200_ERR_ARG;
    if (n == 0) return 0;
    if (ctx->d_bound.cap < (size_t)n * 4 || ctx->d_hew.cap < (size_t)n * 4) { ctx->err = "no QUICKED run on this batch yet"; return QB200_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    if (bound) CK(cudaMemcpyAsync(bound, ctx->d_bound.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (high_error_windows) CK(cudaMemcpyAsync(high_error_windows, ctx->d_hew.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
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

// Pipelined host-in / host-out path: the batch is cut into sub-batches of about one resident wave of the WindowEd
// kernel and pushed through a 3-stage software pipeline over a ring of worker contexts (own stream, own pools):
//   uploader thread  : host prep + H2D of sub-batch k      (PCIe host->device busy back to back)
//   compute thread   : kernels of sub-batch k-1            (the GPU runs one sub-batch at a time, at full occupancy)
//   downloader thread: D2H of sub-batch k-2 into the caller's buffers (PCIe device->host, full duplex with the upload)
// CIGAR strings stay packed in input order: the compute thread knows the text bytes of every earlier sub-batch.
// pk != nullptr: the characters come as a 2-bit packed stream (b then only carries the offset / length arrays and
// seqs_bytes = the stream's character count); sub-batches start on a packed byte (4 characters) and take their slice of
// the exception list.
static int align_batch_pipelined(qb200_ctx *ctx, const quicked_params_t *params, const qb200_batch_t *b, qb200_results_t *res,
                                 const qb200_packed_batch_t *pk = nullptr)
{
    const i64 n = b->n_pairs;
    if (n > kMaxPairs) { ctx->err = "a batch holds at most 2^31 - 2^20 pairs: split it"; return QB200_ERR_ARG; }
    // the sub-batches below are byte ranges [min offset, max end) of the caller's buffer with offsets rebased to them, so
    // a pair that lies outside the buffer must be caught HERE (build_pair_records would only see the rebased offsets)
    for (i64 i = 0; i < n; ++i) {
        const i64 p0 = b->pattern_off[i], t0 = b->text_off[i], m = b->pattern_len[i], tn = b->text_len[i];
        if (m < 0 || tn < 0 || p0 < 0 || t0 < 0 || p0 + m > b->seqs_bytes || t0 + tn > b->seqs_bytes) {
            ctx->err = "pair " + std::to_string(i) + ": offsets/lengths outside the packed buffer";
            return QB200_ERR_ARG;
        }
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    i64 sub = (i64)sms * kWsResidentCtas * kWsThreads * 2;
    // long reads: keep a sub-batch under ~768 MB of characters (its codes, match masks and traceback pool scale with it)
    if (n > 0 && b->seqs_bytes / n > 0) sub = std::min(sub, std::max<i64>(1024, ((i64)768 << 20) / (b->seqs_bytes / n)));
    if (const char *e = getenv("QB200_SUB_PAIRS")) sub = std::max<i64>(1024, atoll(e));
    // Sub-batch boundaries: full-size sub-batches in the middle, a ramp of smaller ones at both ends so that the
    // first upload (nothing to overlap with) and the last compute + download (nothing left to overlap) are short.
    std::vector<i64> cut{0};
    {
        const bool ramp = !getenv("QB200_NO_RAMP") && n >= 4 * sub;
        const i64 head[2] = {sub / 4, sub / 2}, tail[2] = {sub / 2, sub / 4};
        i64 tail_total = ramp ? tail[0] + tail[1] : 0;
        if (ramp) for (i64 h : head) cut.push_back(cut.back() + h);
        const i64 mid = n - tail_total - cut.back();                      // split evenly: no tiny remainder sub-batch
        const i64 parts = std::max<i64>(1, (mid + sub - 1) / sub), base0 = cut.back();
        for (i64 q = 1; q <= parts && mid > 0; ++q) cut.push_back(base0 + mid * q / parts);
        if (ramp) for (i64 t : tail) cut.push_back(cut.back() + t);
    }
    const int S = (int)cut.size() - 1;
    int NC = 2;                                               // compute threads: two sub-batches' kernels interleave on the
    if (const char *e = getenv("QB200_COMPUTE_THREADS")) NC = std::max(1, std::min(atoi(e), 4));   // GPU and fill each other's sync gaps
    int NW = NC + 2;                                          // ring slots: one uploading, NC computing, one downloading
    if (const char *e = getenv("QB200_WORKERS")) NW = std::max(NC + 1, std::min(atoi(e), (int)qb200_ctx::kWorkers));
    const bool trace = getenv("QB200_TRACE") != nullptr;
    for (int k = 0; k < NW; ++k) {
        if (!ctx->child[k]) {
            int rc = qb200_create(&ctx->child[k], ctx->device);
            if (rc) return rc;
        }
        ctx->child[k]->matrix_limit = ctx->matrix_limit / NW;
    }
    std::mutex mu;
    std::condition_variable cv;
    int uploaded = 0, computed = 0, downloaded = 0;          // sub-batches that finished each stage
    int failed = 0;
    std::vector<i64> text_base((size_t)S + 1, 0);             // first CIGAR byte of each sub-batch (prefix of totals)
    const bool want_cigar = !params->only_score && res->cigar_off;
    bool capacity_short = false;
    qb200_stats_t acc;
    memset(&acc, 0, sizeof acc);
    auto t_ms = [](std::chrono::steady_clock::time_point x) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - x).count(); };

    auto fail = [&](int rc, qb200_ctx *c) {
        std::lock_guard<std::mutex> lk(mu);
        if (!failed) { failed = rc; ctx->err = c->err; }
        cv.notify_all();
    };

    std::thread uploader([&] {
        cudaSetDevice(ctx->device);
        std::vector<int64_t> po, to, epos;
        for (int k = 0; k < S; ++k) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return failed || k < downloaded + NW; });        // ring slot k % NW is free again
                if (failed) return;
            }
            const auto t0 = std::chrono::steady_clock::now();
            qb200_ctx *c = ctx->child[k % NW];
            const i64 i0 = cut[(size_t)k], i1 = cut[(size_t)k + 1], cnt = i1 - i0;
            i64 lo = b->seqs_bytes, hi = 0;                   // byte range of this sub-batch in the caller's packed buffer
            for (i64 i = i0; i < i1; ++i) {
                lo = std::min<i64>(lo, std::min<i64>(b->pattern_off[i], b->text_off[i]));
                hi = std::max<i64>(hi, std::max<i64>(b->pattern_off[i] + b->pattern_len[i], b->text_off[i] + b->text_len[i]));
            }
            if (hi < lo) { lo = 0; hi = 0; }
            if (pk) lo &= ~(i64)3;                            // a packed sub-batch starts on a byte of the stream
            po.resize((size_t)cnt); to.resize((size_t)cnt);
            for (i64 i = 0; i < cnt; ++i) { po[(size_t)i] = b->pattern_off[i0 + i] - lo; to[(size_t)i] = b->text_off[i0 + i] - lo; }
            int rc;
            if (pk) {
                const int64_t *e0 = std::lower_bound(pk->exc_pos, pk->exc_pos + pk->n_exc, lo), *e1 = std::lower_bound(e0, pk->exc_pos + pk->n_exc, hi);
                epos.assign(e0, e1);
                for (auto &x : epos) x -= lo;
                qb200_packed_batch_t sp = {pk->packed + lo / 4, hi - lo, cnt, po.data(), b->pattern_len + i0, to.data(), b->text_len + i0,
                                           epos.data(), pk->exc_chr + (e0 - pk->exc_pos), (int64_t)epos.size()};
                rc = qb200_upload_packed(c, &sp);
            } else {
                qb200_batch_t sb = {b->seqs + lo, hi - lo, cnt, po.data(), b->pattern_len + i0, to.data(), b->text_len + i0};
                rc = qb200_upload(c, &sb);
            }
            if (rc) { fail(rc, c); return; }
            if (trace) fprintf(stderr, "[qb200 pipeline] sub %d uploaded in %.2f ms\n", k, t_ms(t0));
            { std::lock_guard<std::mutex> lk(mu); uploaded = k + 1; }
            cv.notify_all();
        }
    });
    auto compute_loop = [&](int tid) {
        cudaSetDevice(ctx->device);
        for (int k = tid; k < S; k += NC) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return failed || k < uploaded; });
                if (failed) return;
            }
            const auto t0 = std::chrono::steady_clock::now();
            qb200_ctx *c = ctx->child[k % NW];
            const int rc = qb200_run(c, params);
            if (rc) { fail(rc, c); return; }
            if (trace) fprintf(stderr, "[qb200 pipeline] sub %d computed in %.2f ms (gpu %.2f: prep %.2f ws %.2f fused %.2f fill %.2f trace %.2f cigar %.2f)\n", k, t_ms(t0), c->stats.ms_total,
                               c->stats.ms_prepare, c->stats.ms_windowed_s, c->stats.ms_fused, c->stats.ms_align_fill, c->stats.ms_align_trace, c->stats.ms_cigar);
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return failed || computed == k; });              // publish in order: text offsets are cumulative
                if (failed) return;
                text_base[(size_t)k + 1] = text_base[(size_t)k] + ((want_cigar && c->have_cigar) ? c->cigar_total : 0);
                computed = k + 1;
            }
            cv.notify_all();
        }
    };
    std::vector<std::thread> computers;
    for (int t = 0; t < NC; ++t) computers.emplace_back(compute_loop, t);
    std::thread downloader([&] {
        cudaSetDevice(ctx->device);
        std::vector<int64_t> loc_off;
        for (int k = 0; k < S; ++k) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return failed || k < computed; });
                if (failed) return;
            }
            const auto t0 = std::chrono::steady_clock::now();
            qb200_ctx *c = ctx->child[k % NW];
            const i64 i0 = cut[(size_t)k], i1 = cut[(size_t)k + 1], cnt = i1 - i0;
            const i64 base = text_base[(size_t)k], need = text_base[(size_t)k + 1] - base;
            if (res->cigar_off) loc_off.resize((size_t)cnt + 1);
            qb200_results_t sr;
            sr.score = res->score ? res->score + i0 : nullptr;
            sr.status = res->status ? res->status + i0 : nullptr;
            sr.cigar_off = res->cigar_off ? loc_off.data() : nullptr;
            const bool fits = want_cigar && res->cigar && base + need <= res->cigar_capacity;
            sr.cigar = fits ? res->cigar + base : nullptr;
            sr.cigar_capacity = fits ? res->cigar_capacity - base : 0;
            sr.cigar_bytes = 0;
            int rc = qb200_download(c, &sr);
            if (rc == QB200_ERR_CAPACITY) { capacity_short = true; rc = 0; }
            if (rc) { fail(rc, c); return; }
            if (res->cigar_off) for (i64 i = 0; i < cnt; ++i) res->cigar_off[i0 + i] = (want_cigar ? loc_off[(size_t)i] + base : 0);   // local -> global
            const qb200_stats_t &st = c->stats;
            acc.n_pairs += st.n_pairs; acc.kernel_launches += st.kernel_launches; acc.word_steps += st.word_steps;
            acc.word_steps_windowed += st.word_steps_windowed; acc.word_steps_banded += st.word_steps_banded; acc.cells += st.cells;
            acc.h2d_bytes += st.h2d_bytes; acc.d2h_bytes += st.d2h_bytes; acc.pairs_stage2 += st.pairs_stage2; acc.pairs_stage3 += st.pairs_stage3;
            acc.banded_tries += st.banded_tries; acc.hirschberg_splits += st.hirschberg_splits; acc.leaves += st.leaves;
            acc.ms_total += st.ms_total; acc.ms_prepare += st.ms_prepare; acc.ms_windowed_s += st.ms_windowed_s; acc.ms_windowed_l += st.ms_windowed_l;
            acc.ms_banded += st.ms_banded; acc.ms_align_fill += st.ms_align_fill; acc.ms_align_trace += st.ms_align_trace; acc.ms_cigar += st.ms_cigar;
            acc.matrix_bytes += st.matrix_bytes; acc.ms_fused += st.ms_fused; acc.pairs_fused += st.pairs_fused; acc.leaves_punted += st.leaves_punted;
            if (trace) fprintf(stderr, "[qb200 pipeline] sub %d downloaded in %.2f ms\n", k, t_ms(t0));
            { std::lock_guard<std::mutex> lk(mu); downloaded = k + 1; }
            cv.notify_all();
        }
    });
    uploader.join();
    for (auto &t : computers) t.join();
    downloader.join();
    if (failed) return failed;
    const i64 total = text_base[(size_t)S];
    if (res->cigar_off) res->cigar_off[n] = want_cigar ? total : 0;
    res->cigar_bytes = want_cigar ? total : 0;
    ctx->stats = acc;
    ctx->ran = false;       // results live in the caller's buffers, not in this context
    return capacity_short ? QB200_ERR_CAPACITY : 0;
}

int qb200_align_batch(qb200_ctx_t *ctx, const quicked_params_t *params, const qb200_batch_t *b, qb200_results_t *res)
{
    if (!ctx || !params || !b || !res) return QB200_ERR_ARG;
    // big jobs (many pairs, or many characters: 100 k pairs of 10 kbp are 2 GB) overlap H2D, kernels and D2H
    i64 min_pairs = 200000;
    if (const char *e = getenv("QB200_PIPELINE_MIN_PAIRS")) min_pairs = std::max<i64>(2, atoll(e));
    const bool big = b->n_pairs >= min_pairs || (b->n_pairs >= 4096 && b->seqs_bytes >= ((i64)512 << 20));
    if (big && !getenv("QB200_NO_PIPELINE")) return align_batch_pipelined(ctx, params, b, res);
    int rc = qb200_upload(ctx, b);
    if (rc) return rc;
    rc = qb200_run(ctx, params);
    if (rc) return rc;
    return qb200_download(ctx, res);
}

}  // extern "C"
