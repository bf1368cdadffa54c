// qb_hirschberg.cuh — the combine step of one Hirschberg split (reference bpm_hirschberg.c:103-200).
//
// The two half-passes (forward over the first ceil(n/2) columns, reverse over the rest) are ordinary score-only
// BandEd tasks and run CONCURRENTLY as separate warps of one k_banded_warp launch; this kernel joins them on the
// middle column: prefix sums of the +/-1 vertical deltas over the overlap of the two bands, argmin of fwd + rev
// with the reference's tie-breaking (first minimum scanning upward, strict '<'), and the exact sub-scores that
// become the children's cutoffs.  One thread per split: the scan is <= 64*B+2 cells and splits are few.
#pragma once
#include "qb_common.cuh"

namespace qb {

struct SplitTask {
    int m, n;                  // the node's sub-problem
    i64 cutoff;
    int fwd_slot, rev_slot;    // BandOut slots of the two passes
    i64 fwd_state, rev_state;  // u64 index: Pv[Bs] then Mv[Bs] of each pass
    i64 fwd_scores, rev_scores;
    i64 scratch_off;           // int32 index: 2 * (64*Bs + 8) ints
};
struct SplitOut {
    int status;                // 0 ok, -2 QUICKED_FAIL_NON_CONVERGENCE
    int m_l;                   // pattern rows given to the left child
    i64 score_l, score_r;      // children's cutoffs
};

__global__ void __launch_bounds__(64)
k_hirschberg_combine(const SplitTask *__restrict__ tasks, int n_tasks, const BandOut *__restrict__ bo,
                     const u64 *__restrict__ state, const int *__restrict__ scores_pool, int *__restrict__ scratch,
                     SplitOut *__restrict__ outs)
{
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n_tasks) return;
    const SplitTask tk = tasks[id];
    const BandGeom g = band_geometry(tk.m, tk.n, tk.cutoff);
    const i64 m = tk.m;
    const i64 n_l = (tk.n + 1) / 2, n_r = tk.n - n_l;                                  // :68-69
    const int Bs = (int)g.Bs;
    const BandOut f = bo[tk.fwd_slot], r = bo[tk.rev_slot];
    const u64 *pv = state + tk.fwd_state, *mv = pv + Bs, *pvr = state + tk.rev_state, *mvr = pvr + Bs;
    const int *sf = scores_pool + tk.fwd_scores, *sr = scores_pool + tk.rev_scores;
    const i64 org = n_l < g.prolog * 64 ? 0 : n_l / 64 - g.prolog;                     // :103-104 (clamped like the reference)
    const i64 org_r = n_r < g.prolog * 64 ? 0 : n_r / 64 - g.prolog;
    const i64 lo_f = (i64)f.first * 64 + 63 + org * 64;                                // :110-113
    const i64 lo_r = (m - 1) - ((i64)r.last * 64 + 63 + org_r * 64);
    const i64 hi_f = (i64)f.last * 64 + 63 + org * 64;
    const i64 hi_r = (m - 1) - ((i64)r.first * 64 + 63 + org_r * 64);
    SplitOut o; o.status = 0; o.m_l = 0; o.score_l = 0; o.score_r = 0;
    if (lo_f > hi_r || lo_r > hi_f) { o.status = -2; outs[id] = o; return; }           // :116-122
    i64 cell0, start, top, top_r;
    if (lo_f > lo_r) { cell0 = (i64)f.first * 64 + 63; start = lo_f; } else { cell0 = lo_r - org * 64; start = lo_r; }      // :125-134
    if (hi_f < hi_r) { top = (i64)f.last * 64 + 63; top_r = (m - 1) - hi_f - org_r * 64; }                                 // :137-146
    else { top = hi_r - org * 64; top_r = (i64)r.first * 64 + 63; }
    const i64 ncell = top - cell0 + 2;                                                 // :147
    int *cs = scratch + tk.scratch_off, *csr = cs + (64 * Bs + 8);
    const i64 cap = 64 * (i64)Bs + 6;
    if (ncell < 1 || ncell > cap || cell0 < 0 || top_r < 0) { o.status = -2; outs[id] = o; return; }   // outside what the reference can index
    cs[0] = 0; csr[0] = 0;
    for (i64 i = 0; i < ncell; ++i) {                                                  // :152-167
        const i64 c = cell0 + i, cr = top_r + i;
        const i64 wc = c >> 6, wr = cr >> 6;
        const u64 a = wc < Bs ? pv[wc] : 0, b = wc < Bs ? mv[wc] : 0, ar = wr < Bs ? pvr[wr] : 0, br = wr < Bs ? mvr[wr] : 0;
        cs[i + 1] = cs[i] + (int)((a >> (c & 63)) & 1) - (int)((b >> (c & 63)) & 1);
        csr[i + 1] = csr[i] + (int)((ar >> (cr & 63)) & 1) - (int)((br >> (cr & 63)) & 1);
    }
    i64 best = 0, best_score = (i64)csr[ncell - 1] + cs[0];                            // :170-180
    for (i64 i = 1; i < ncell; ++i) {
        const i64 s = (i64)csr[ncell - 1 - i] + cs[i];
        if (s < best_score) { best = i; best_score = s; }
    }
    const i64 m_l = start + best, m_r = m - m_l;                                       // :183-184
    const i64 ref_l = ceil_div(m_l, 64) - (ncell < best + 64);                         // :194-196
    const i64 sp_l = ref_l * 64 - (cell0 + org * 64);
    const i64 ref_r = ceil_div(m_r, 64) - (best < 64);                                 // :198-200
    const i64 sp_r = ref_r * 64 - (top_r + org_r * 64);
    if (m_l <= 0 || m_r <= 0 || sp_l < 0 || sp_l > ncell || sp_r < 0 || sp_r > ncell || ref_l < 1 || ref_r < 1) {
        o.status = -1; outs[id] = o; return;                                           // the reference indexes out of bounds here
    }
    o.m_l = (int)m_l;
    o.score_l = (i64)cs[best] - cs[sp_l] + sf[ref_l - 1];
    o.score_r = (i64)csr[ncell - 1 - best] - csr[sp_r] + sr[ref_r - 1];
    outs[id] = o;
}

}  // namespace qb
