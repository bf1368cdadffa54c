// qb_plan.cuh — device-side planning of a batch: which pairs are plain leaves (the overwhelmingly common case on
// generate_dataset inputs: WindowEd(S) bound accepted, no Hirschberg split), where their traceback state, op
// words and results live, and which pairs need the host-driven slow path (WindowEd(L) / band doubling / splits).
//
// This replaces the per-pair control flow of run_quicked / run_banded / run_hirschberg (reference
// quicked.c:58-306) for the fast path; thresholds are evaluated with the reference's integer arithmetic.
#pragma once
#include "qb_common.cuh"
#include "qb_traceback.cuh"
#include "qb_tiles.cuh"

namespace qb {

enum PairClass { CLS_NONE = 0, CLS_T = 1, CLS_W = 2, CLS_SLOW = 3 };

// Running sums over pairs (one exclusive scan lays out every pool).
struct PlanSum {
    i64 leaf;   // fast leaves
    i64 t, w;   // thread-kernel / warp-kernel leaves
    i64 ops;    // u32 words of 2-bit ops
    i64 rng;    // int2 live-range entries
    i64 sc;     // int32 running scores (warp kernel)
    i64 matw;   // matrix entries of warp-kernel leaves
    i64 slow;   // pairs for the host-driven path
};
struct PlanAdd {
    __host__ __device__ PlanSum operator()(const PlanSum &a, const PlanSum &b) const
    {
        PlanSum r;
        r.leaf = a.leaf + b.leaf; r.t = a.t + b.t; r.w = a.w + b.w; r.ops = a.ops + b.ops;
        r.rng = a.rng + b.rng; r.sc = a.sc + b.sc; r.matw = a.matw + b.matw; r.slow = a.slow + b.slow;
        return r;
    }
};

struct PlanParams {
    int algo;               // quicked_algo_t
    unsigned bandwidth, hew_pct0;
    int only_score;
    int thread_band_max;    // B_cigar <= this -> thread kernel
    int ok_status;          // QUICKED_WIP or QUICKED_OK (HIRSCHBERG)
    int tiles;              // wider bands: tile records + tile traceback (qb_tiles.cuh) instead of the full matrix
};

__device__ __forceinline__ int rounds_for_dev(i64 B)
{
    return B <= 32 ? 1 : B <= 64 ? 2 : B <= 128 ? 4 : B <= 256 ? 8 : B <= 512 ? 16 : B <= 1024 ? 32 : B <= 11000 ? 64 : 0;
}

__global__ void __launch_bounds__(256)
k_plan(const PairRec *__restrict__ pairs, int n, PlanParams pp, const int *__restrict__ bound,
       const int *__restrict__ hew, PlanSum *__restrict__ items, unsigned char *__restrict__ cls,
       i64 *__restrict__ cutoff, int *__restrict__ status, int *__restrict__ score, const unsigned char *__restrict__ done,
       unsigned *__restrict__ tile_class_mask)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const PairRec r = pairs[i];
    PlanSum it = {0, 0, 0, 0, 0, 0, 0, 0};
    if (done && done[i]) {                                          // finished by the fused kernel
        items[i] = it; cls[i] = (unsigned char)CLS_NONE; cutoff[i] = 0; score[i] = -1;
        return;
    }
    int c = CLS_NONE, st = -1 /* QUICKED_ERROR */;
    i64 cut = 0;
    if (r.m <= 0 || r.n <= 0) {
        st = -4;                                                    // QUICKED_EMPTY_SEQUENCE, quicked.c:411-414
    } else {
        const unsigned maxlen = (unsigned)max(r.m, r.n);
        bool slow = false;
        if (pp.algo == 0) {                                         // QUICKED
            cut = bound[i];
            slow = (i64)hew[i] * 64 > (i64)(maxlen * pp.hew_pct0 / 100);   // quicked.c:201-202
        } else if (pp.algo == 2 || pp.algo == 3) {                  // BANDED / HIRSCHBERG: quicked.c:64, :131
            cut = (i64)(maxlen * pp.bandwidth / 100);
            slow = (pp.algo == 2 && pp.only_score);
        } else {
            slow = true;                                            // WINDOWED
        }
        const BandGeom g = band_geometry(r.m, r.n, cut);
        if (!slow && pp.algo != 2 && (unsigned long long)g.Bc * (unsigned long long)r.n * 16ull > (1ull << 24))
            slow = true;                                            // Hirschberg split, bpm_hirschberg.c:63-65
        if (slow) { c = CLS_SLOW; it.slow = 1; }
        else {
            const bool thr = g.Bc <= pp.thread_band_max;
            if (!thr && rounds_for_dev(g.Bc) == 0) { st = -10; }    // band taller than the implemented kernels
            else {
                c = thr ? CLS_T : CLS_W;
                st = pp.ok_status;
                it.leaf = 1; it.t = thr; it.w = !thr;
                const bool tile = !thr && pp.tiles && tile_ring_for(g.Bc) <= kTileMaxRing;
                // 2-bit ops (thread walk, tile walk) / u32 runs, worst case (warp walk)
                it.ops = (thr || tile) ? (r.m + r.n + 15) / 16 : (r.m + r.n + 15) / 16 * 16;
                it.rng = r.n / 64 + 2;
                if (thr && pp.tiles) {                  // thread fill in tile-record mode: same records, same walk as the wider bands
                    it.sc = (r.m + 63) / 64 + g.Bc + 2;            // (scores only if the exact kernels have to redo the leaf)
                    it.matw = 2 * (i64)((r.n + 63) / 64) * g.Bc;
                }
                if (!thr) {
                    it.sc = (r.m + 63) / 64 + g.Bc + 2;
                    // traceback state in 16-byte units: one 32-byte record per tile, or the reference's whole matrix
                    it.matw = tile ? 2 * (i64)((r.n + 63) / 64) * g.Bc : (i64)(r.n + 1) * g.Bc;
                    if (tile) {
                        int cb = 0;
                        while ((8 << cb) < tile_ring_for(g.Bc)) ++cb;
                        atomicOr(tile_class_mask, 1u << cb);
                    } else atomicOr(tile_class_mask, 1u << 31);      // some leaf needs the full-matrix warp kernels
                }
            }
        }
    }
    items[i] = it; cls[i] = (unsigned char)c; cutoff[i] = cut; status[i] = st; score[i] = -1;
}

// totals[0] = sum over all pairs (exclusive prefix of the last pair + its own item)
__global__ void k_plan_totals(const PlanSum *__restrict__ items, const PlanSum *__restrict__ offs, int n, PlanSum *totals)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) totals[0] = PlanAdd()(offs[n - 1], items[n - 1]);
}

__global__ void __launch_bounds__(256)
k_build_leaves(const PairRec *__restrict__ pairs, int n, const unsigned char *__restrict__ cls,
               const i64 *__restrict__ cutoff, const PlanSum *__restrict__ offs, BandTask *__restrict__ leaves,
               int *__restrict__ list_t, int *__restrict__ list_w, int *__restrict__ list_slow,
               PairLeaves *__restrict__ pl, const unsigned char *__restrict__ done, i64 leaf_base, i64 ops_base)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (done && done[i]) return;                                    // the fused kernel already wrote this pair's records
    const int c = cls[i];
    PlanSum o = offs[i];
    o.leaf += leaf_base; o.ops += ops_base;
    PairLeaves p; p.first_leaf = o.leaf; p.n_leaves = 0; p.pad_ = 0;
    if (c == CLS_T || c == CLS_W) {
        const PairRec r = pairs[i];
        BandTask t;
        t.p_off = r.p_off; t.t_off = r.t_off; t.m = r.m; t.n = r.n; t.rev = 0; t.finish = r.n;
        t.cutoff = cutoff[i]; t.peq_off = r.peq_off; t.nbp = r.nbp; t.pair = i;
        t.mat_off = o.matw; t.mat_ws = 1;
        { const BandGeom g = band_geometry(r.m, r.n, t.cutoff); t.mat_cs = (int)g.Bc; }     // band height (thread-kernel groups overwrite it)
        t.scores_off = o.sc; t.state_off = 0; t.ops_off = o.ops; t.range_off = o.rng;
        t.ops_cap = ((r.m + r.n + 15) / 16) * 16; t.slot = (int)o.leaf;
        leaves[o.leaf] = t;
        if (c == CLS_T) list_t[o.t] = (int)o.leaf; else list_w[o.w] = (int)o.leaf;
        p.n_leaves = 1;
    } else if (c == CLS_SLOW) {
        list_slow[o.slow] = i;
    }
    pl[i] = p;
}

// One warp per group of 32 thread-kernel leaves: the group's interleaved matrix needs max(B) words x (max(n)+1) columns.
__global__ void __launch_bounds__(256)
k_group_size(const int *__restrict__ list_t, int n_t, const BandTask *__restrict__ leaves, i64 *__restrict__ gsize,
             int *__restrict__ gB)
{
    const int g = (int)(((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (g * 32 >= n_t) return;
    int B = 1, nn = 1;
    if (g * 32 + lane < n_t) {
        const BandTask &t = leaves[list_t[g * 32 + lane]];
        B = (int)band_geometry(t.m, t.n, t.cutoff).Bc; nn = t.n;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { B = max(B, __shfl_xor_sync(kFull, B, o)); nn = max(nn, __shfl_xor_sync(kFull, nn, o)); }
    if (lane == 0) { gsize[g] = (i64)(nn + 1) * B * 32; gB[g] = B; }
}

__global__ void __launch_bounds__(256)
k_group_assign(const int *__restrict__ list_t, int n_t, BandTask *__restrict__ leaves, const i64 *__restrict__ goff,
               const int *__restrict__ gB, i64 base)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_t) return;
    BandTask &t = leaves[list_t[q]];
    t.mat_off = base + goff[q >> 5] + (q & 31);
    t.mat_cs = gB[q >> 5] * 32;
    t.mat_ws = 32;
}

// Scores and text lengths of finished pairs.  Single-leaf pairs take both from the traceback's LeafOut.
__global__ void __launch_bounds__(256)
k_pair_finish(const PairLeaves *__restrict__ pl, int n, const LeafOut *__restrict__ lo, const int *__restrict__ status,
              int *__restrict__ score, i64 *__restrict__ text_bytes, int want_cigar)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const PairLeaves p = pl[i];
    i64 tb = 1;                                            // the NUL
    if (p.n_leaves > 0 && status[i] >= -2) {
        int s = 0;
        for (int l = 0; l < p.n_leaves; ++l) s += lo[p.first_leaf + l].cost;
        score[i] = s;                                      // cigar_score_edit, quicked.c:54
        if (p.n_leaves == 1 && lo[p.first_leaf].text_len >= 0) tb += lo[p.first_leaf].text_len;
        else tb = -1;                                      // measured by k_cigar_text<false> (several leaves, or a tile walk)
    }
    if (want_cigar) text_bytes[i] = tb;
}

__global__ void __launch_bounds__(256)
k_merge_text_len(const int *__restrict__ text_len, i64 *__restrict__ text_bytes, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && text_bytes[i] < 0) text_bytes[i] = (i64)text_len[i] + 1;
}

__global__ void k_last_offset(const i64 *__restrict__ offs, const i64 *__restrict__ vals, int n, i64 *out, i64 *offs_n)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) { const i64 t = offs[n - 1] + vals[n - 1]; out[0] = t; offs_n[0] = t; }
}

// Results of the host-driven path scattered back into the per-pair arrays.
struct SlowResult { int pair; int status; int score; int set_score; PairLeaves pl; };
__global__ void __launch_bounds__(256)
k_scatter_slow(const SlowResult *__restrict__ res, int n, PairLeaves *__restrict__ pl, int *__restrict__ status, int *__restrict__ score)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const SlowResult r = res[q];
    pl[r.pair] = r.pl; status[r.pair] = r.status;
    if (r.set_score) score[r.pair] = r.score;
}

}  // namespace qb
