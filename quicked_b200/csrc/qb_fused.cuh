// qb_fused.cuh — the QUICKED fast path for short / medium pairs in ONE persistent kernel.
//
// For a pair whose WindowEd(S) bound is accepted (quicked.c:201-202 not taken), needs no Hirschberg split
// (bpm_hirschberg.c:63-65) and whose band is at most 4 blocks tall — every pair of the 100 bp and 1 kbp configs —
// the whole reference schedule (quicked.c:163-306) is one thread's job:
//     WindowEd(S) bound  ->  BandEd full-matrix fill  ->  traceback  ->  2-bit ops + score
// Fusing the three stages keeps a thread's traceback state in a FIXED per-resident-warp slot (6 GB for 1 kbp pairs
// instead of a 48 GB pool for a million pairs, no planning pass, no chunking), and lets the warps of an SM be in
// different stages at the same time: the integer-bound fills of some warps hide the memory-latency-bound walks of
// others, which three back-to-back kernels cannot do.
// Pairs that do not qualify are left untouched (done[i] = 0) for the planned path (warp kernels / host-driven stages).
#pragma once
#include "qb_banded.cuh"
#include "qb_common.cuh"
#include "qb_traceback.cuh"
#include "qb_windowed.cuh"

namespace qb {

constexpr int kFusedCtasPerSm = 5;     // 640 resident threads per SM (register budget 102)
constexpr int kFusedBandMax = 4;

struct FusedParams {
    int hew_threshold0;
    unsigned hew_pct0;
    int n_lim;            // longest text the per-warp matrix slot can hold
    int ok_status;
    i64 mat_warp_stride;  // entries per resident-warp slot = (n_lim+1) * kFusedBandMax * 32
};

template <bool SSE>
__global__ void __launch_bounds__(kWsThreads, kFusedCtasPerSm)
k_quicked_fused(const PairRec *__restrict__ pairs, int n_pairs, const unsigned char *__restrict__ codes,
                const unsigned char *__restrict__ raw, const u64 *__restrict__ peq, FusedParams fp, u64 *__restrict__ quad,
                ulonglong2 *__restrict__ matrix, int2 *__restrict__ ranges, u32 *__restrict__ ops_pool, int *__restrict__ bound,
                int *__restrict__ hew_out, unsigned char *__restrict__ done, int *__restrict__ status, BandTask *__restrict__ leaves,
                LeafOut *__restrict__ leaf_out, PairLeaves *__restrict__ pl, u64 *__restrict__ counters)
{
    __shared__ u64 s_mem[kFusedBandMax * kAlpha * kWsThreads];      // WindowEd uses the first 10 slots per thread, the fill all 20
    constexpr int T = kWsThreads;
    static_assert(kWsThreads == kThreadFillThreads, "banded_thread_fill is compiled for this CTA size");
    const int t = threadIdx.x, lane = t & 31;
    u64 *s_thr = s_mem + t;
    const i64 gtid = (i64)blockIdx.x * T + t, nthr = (i64)gridDim.x * T;
    u64 *qpv = quad + gtid, *qmv = quad + 65 * nthr + gtid;
    ulonglong2 *mat = matrix + (gtid >> 5) * fp.mat_warp_stride + lane;   // [column][word][lane] slot of this warp
    int2 *rng = ranges + gtid;                                           // [block][thread]
    const int hew_lim = 64 * fp.hew_threshold0 / 100;
    u64 ws_w = 0, ws_b = 0, n_done = 0;
    for (i64 i = gtid; i < n_pairs; i += nthr) {
        const PairRec pr = pairs[i];
        if (pr.m <= 0 || pr.n <= 0) { done[i] = 0; continue; }
        int score = 0, hew = 0;
        ws21_pair<SSE, false>(pr, codes, raw, peq, s_thr, qpv, qmv, nthr, nullptr, false, hew_lim, score, hew, ws_w);
        bound[i] = score; hew_out[i] = hew;
        const unsigned maxlen = (unsigned)max(pr.m, pr.n);
        const BandGeom g = band_geometry(pr.m, pr.n, score);
        const bool eligible = !((i64)hew * 64 > (i64)(maxlen * fp.hew_pct0 / 100)) &&                    // quicked.c:201-202
                              !((unsigned long long)g.Bc * (unsigned long long)pr.n * 16ull > (1ull << 24)) &&   // bpm_hirschberg.c:63-65
                              g.Bc <= kFusedBandMax && pr.n <= fp.n_lim;
        if (!eligible) { done[i] = 0; continue; }
        banded_thread_fill<kFusedBandMax>(pr.m, pr.n, score, 0, peq + pr.peq_off, pr.nbp, codes + pr.t_off, mat,
                                          (i64)kFusedBandMax * 32, 32, rng, nthr, s_thr, T, ws_b);
        const int ops_cap = ((pr.m + pr.n + 15) / 16) * 16;
        LeafOut o;
        traceback_walk_thread(pr.m, pr.n, score, mat, (i64)kFusedBandMax * 32, 32, rng, nthr, raw + pr.p_off, raw + pr.t_off,
                              ops_pool + pr.ops_off, ops_cap, o);
        leaf_out[i] = o;
        BandTask lf;                       // only what the CIGAR-text pass reads
        lf.p_off = pr.p_off; lf.t_off = pr.t_off; lf.m = pr.m; lf.n = pr.n; lf.rev = 0; lf.finish = pr.n; lf.cutoff = score;
        lf.peq_off = pr.peq_off; lf.nbp = pr.nbp; lf.pair = (int)i; lf.mat_off = 0; lf.mat_cs = 0; lf.mat_ws = 0;
        lf.scores_off = 0; lf.state_off = 0; lf.ops_off = pr.ops_off; lf.range_off = 0; lf.ops_cap = ops_cap; lf.slot = (int)i;
        leaves[i] = lf;
        PairLeaves p; p.first_leaf = i; p.n_leaves = 1; p.pad_ = 0;
        pl[i] = p;
        status[i] = fp.ok_status;
        done[i] = 1;
        ++n_done;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ws_w += __shfl_down_sync(kFull, ws_w, o); ws_b += __shfl_down_sync(kFull, ws_b, o); n_done += __shfl_down_sync(kFull, n_done, o);
    }
    if (lane == 0) {
        if (ws_w) atomicAdd(&counters[0], ws_w);
        if (ws_b) atomicAdd(&counters[1], ws_b);
        if (n_done) atomicAdd(&counters[2], n_done);
    }
}

}  // namespace qb
