// qb_wintile.cuh — WindowEd for any window / overlap (W <= 32 words), score-only or CIGAR mode, as TILES.
// Reference: windowed_compute (bpm_windowed.c:563-628) = per window windowed_compute_window (:202-280, fill) +
// windowed_backtrace / _score_only (:448-561, walk); callers run_windowed (quicked.c:91-123) and the WindowEd(L)
// stage of QUICKED (quicked.c:204-233).
//
// WindowEd is a chain: the walk of one window decides where the next window sits, so a pair offers no parallelism
// beyond one window and the throughput comes from pairs in flight — ONE PAIR PER THREAD, like the WindowEd(S) kernel.
// What a thread keeps is what makes that affordable for 9-word windows (the reference stores the whole window,
// (64W+2) x W words of Pv and Mv = 83 KB at W = 9, bpm_windowed.c:143):
//   * the fill goes through the window tile by tile (64 rows x 64 columns, the inner loop of the BandEd tile kernel:
//     block state in registers, the 64 horizontal carries between vertically adjacent tiles as two shift registers);
//     per tile it stores a 32-byte record (block state at the tile's first column + its 64 carry-ins) in a
//     [tile][thread] scratch (coalesced; W*W records = 2.6 KB per thread at W = 9, L2-resident);
//   * the walk recomputes only the tiles it visits from their records (about 2 per 64 columns) and keeps, per column,
//     the decision for the 16 rows around its diagonal as two bit planes in one u32 of shared memory — score-only
//     mode tests D, I, then the diagonal (bpm_windowed.c:532-552); CIGAR mode tests M (equal characters) first, then
//     D, I, X (:476-496).
// Equal characters are read off the match masks; a pair with a character outside "ACGTN" (equal codes, possibly
// different bytes) is left to k_windowed_warp, which compares raw bytes like the reference — and so are 2-word windows
// with the SSE quirks.  The kernel flags those tasks (WinOut.hew = kWinPunted) and the warp kernel picks them up.
#pragma once
#include "qb_tiles.cuh"
#include "qb_tiletrace.cuh"
#include "qb_windowed.cuh"

namespace qb {

constexpr int kWtThreads = 128;
constexpr int kWinPunted = -2147483647 - 1;
constexpr int kTraceHalfW = 8;                // rows kept on each side of the walk's diagonal

// bits [lo, lo+16) of x (zero outside 0..63), -80 < lo < 64
__device__ __forceinline__ u32 win_slice16(u64 x, int lo)
{
    u64 y;
    if (lo >= 0) y = x >> lo;
    else y = (-lo < 64) ? (x << (-lo)) : 0ull;
    return (u32)y & 0xffffu;
}

template <bool SCORE_ONLY>
__global__ void __launch_bounds__(kWtThreads, 4)
k_windowed_tiles(const WinTask *__restrict__ tasks, int n_tasks, const unsigned char *__restrict__ codes,
                 const u64 *__restrict__ peq, ulonglong2 *__restrict__ scratch, i64 rec_tiles, int Wmax,
                 u32 *__restrict__ ops_pool, WinOut *__restrict__ outs, LeafOut *__restrict__ leaf_outs,
                 u64 *__restrict__ counters)
{
    constexpr int T = kWtThreads;
    __shared__ u64 s_eq[kAlpha * T];          // the current block's window-aligned match masks
    __shared__ u64 s_txt[8 * T];              // 64 text codes of the current column block
    __shared__ u32 s_planes[64 * T];          // walk decisions of the tile being walked
    const int t = threadIdx.x;
    u64 *eq = s_eq + t, *txt = s_txt + t;
    u32 *planes = s_planes + t;
    const i64 gtid = (i64)blockIdx.x * T + t, nthr = (i64)gridDim.x * T;
    // scratch: [2 * rec_tiles][nthr] record halves, then [Wmax][nthr] block states
    ulonglong2 *rec_a = scratch + gtid;                                   // (pv0, mv0) of tile q at rec_a[q * nthr]
    ulonglong2 *rec_b = scratch + rec_tiles * nthr + gtid;                // carry-ins (p0,p1 | m0,m1)
    ulonglong2 *wst = scratch + 2 * rec_tiles * nthr + gtid;              // block b at wst[b * nthr]
    u64 ws = 0;
    for (i64 i = gtid; i < n_tasks; i += nthr) {
        const WinTask tk = tasks[i];
        const int W = tk.W, O = tk.O;
        if (tk.m <= 0 || tk.n <= 0) continue;                             // a pair with an empty side has no task (k_win_build)
        if (W > Wmax || (tk.sse && W == 2)) { WinOut wo; wo.score = 0; wo.hew = kWinPunted; outs[tk.slot] = wo; continue; }
        const u64 *pq = peq + tk.peq_off;
        const unsigned char *tc = codes + tk.t_off;
        ShiftWriter ow;
        if (!SCORE_ONLY) ow.init(ops_pool + tk.ops_off, tk.ops_cap);
        int cv = tk.m - 1, ch = tk.n - 1, score = 0, hew = 0;
        const int hew_lim = (W - O) * 64 * tk.hew_threshold / 100;
        bool odd = false;
        // 64 codes of window column block (first column h_first, nc columns) -> txt; returns the odd-character mask
        auto stage_text = [&](int h_first, int nc) -> u64 {
            u64 od = 0;
#pragma unroll 1
            for (int c8 = 0; c8 < 8; ++c8) {
                u64 w = 0;
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const int s = 8 * c8 + b, col = h_first + s;
                    unsigned cd = 4;
                    if (s < nc && col < tk.n) cd = tk.rev ? tc[tk.n - 1 - col] : tc[col];
                    od |= (u64)((cd >> 3) & 1u) << s;
                    w |= (u64)(cd & 7u) << (8 * b);
                }
                txt[c8 * T] = w;
            }
            return od;
        };
        while (cv >= 0 && ch >= 0 && !odd) {
            const int v0 = max(cv - 64 * W + 1, 0), h0 = max(ch - 64 * W + 1, 0);
            const int words = ((cv - v0) >> 6) + 1, cols = ch - h0 + 1, K = (cols + 63) >> 6;
            const unsigned sh = v0 & 63;
            const int blk0 = v0 >> 6;
            auto stage_eq = [&](int b) -> u64 {                          // window-aligned masks of block b (bpm_windowed.c:237-244)
                const u64 *q0 = pq + (i64)(blk0 + b) * kPeqStride, *q1 = q0 + kPeqStride;
#pragma unroll
                for (int c = 0; c < kAlpha; ++c) eq[c * T] = funnel_r(q0[c], q1[c], sh);
                return funnel_r(q0[kAlpha], q1[kAlpha], sh);             // rows holding a character outside "ACGTN"
            };
            // ---- fill: column block by column block, top to bottom ----
            for (int b = 0; b < words; ++b) wst[b * nthr] = make_ulonglong2(h0 == 0 ? ~0ull : 0ull, 0ull);   // :225-229
            for (int k = 0; k < K && !odd; ++k) {
                const int nc = min(64, cols - 64 * k);
                if (stage_text(h0 + 64 * k, nc)) odd = true;
                TileCarry c;
                c.p0 = c.p1 = (v0 == 0) ? 0xffffffffu : 0u;              // top of the window: Hin = +1 only on pattern row 0 (:247-252)
                c.m0 = c.m1 = 0u;
                for (int b = 0; b < words; ++b) {
                    const ulonglong2 st = wst[b * nthr];
                    u64 pv = st.x, mv = st.y;
                    const int q = k * W + b;
                    rec_a[q * nthr] = st;
                    rec_b[q * nthr] = make_ulonglong2(((u64)c.p1 << 32) | c.p0, ((u64)c.m1 << 32) | c.m0);
                    if (stage_eq(b)) odd = true;
                    if (nc == 64) tile_fill64<T, T>(pv, mv, c, eq, T, txt, txt[0], txt[T]);
                    else {
                        TileCarry o; o.p0 = o.p1 = o.m0 = o.m1 = 0;
                        for (int s = 0; s < nc; ++s) {
                            const u32 code = (u32)(txt[(s >> 3) * T] >> (8 * (s & 7))) & 7u;
                            const u32 hp_in = ((s < 32 ? c.p0 : c.p1) >> (31 - (s & 31))) & 1u, hm_in = ((s < 32 ? c.m0 : c.m1) >> (31 - (s & 31))) & 1u;
                            u32 hp_out, hm_out;
                            myers_step(eq[code * T], pv, mv, hp_in, hm_in, hp_out, hm_out);
                            if (s < 32) { o.p0 |= hp_out << (31 - s); o.m0 |= hm_out << (31 - s); }
                            else { o.p1 |= hp_out << (63 - s); o.m1 |= hm_out << (63 - s); }
                        }
                        c = o;
                    }
                    wst[b * nthr] = make_ulonglong2(pv, mv);
                }
            }
            if (odd) break;
            ws += (u64)words * cols;
            // ---- walk ----
            const int v_stop = max(cv - 64 * (W - O) + 1, 0), h_stop = max(ch - 64 * (W - O) + 1, 0);
            int v = cv, h = ch, cost = 0;
            int eq_b = -1, txt_k = K - 1;                                // txt still holds the last column block
            while (v >= v_stop && h >= h_stop) {
                const int r = v - v0, cc = h - h0;
                const int b = r >> 6, k = cc >> 6, r0 = r & 63, s0 = cc & 63;
                if (b != eq_b) { stage_eq(b); eq_b = b; }
                if (k != txt_k) { stage_text(h0 + 64 * k, min(64, cols - 64 * k)); txt_k = k; }
                const ulonglong2 ra = rec_a[(k * W + b) * nthr], rb = rec_b[(k * W + b) * nthr];
                u64 pv = ra.x, mv = ra.y;
                const int lo0 = r0 - s0 - kTraceHalfW;                   // lowest slice row at column 0
                // all 64 columns in groups of eight, every column with its planes (qb_tiletrace.cuh: a trip count or a
                // "planes needed?" test that depends on the pair makes the lanes of a warp take turns); columns past the
                // window hold code 4 and are never read by the walk
#pragma unroll 1
                for (int s8 = 0; s8 < 64; s8 += 8) {
                    u64 cw = txt[(s8 >> 3) * T];
                    u32 wp = (u32)(s8 < 32 ? rb.x : rb.x >> 32) << (s8 & 31), wm = (u32)(s8 < 32 ? rb.y : rb.y >> 32) << (s8 & 31);
                    u32 *pl = planes + s8 * T;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const u64 e = eq[(unsigned)(cw & 7u) * T];
                        cw >>= 8;
                        const u64 mv_old = mv;
                        QB_TRACE_STEP(e);
                        u64 pa, pb;
                        if (SCORE_ONLY) { pa = pv | ~(mv_old | e); pb = ~pv & (mv_old | ~e); }      // D (1,0)  I (0,1)  X (1,1)  M (0,0)
                        else { pa = ~e & (pv | ~mv_old); pb = ~e & ~pv; }                          // M first: (0,0) whenever the characters match
                        const int lo = trace_slice_lo(lo0, s8 + i);
                        pl[i * T] = __byte_perm((u32)(pa >> lo), (u32)(pb >> lo), 0x5410);
                    }
                }
                int rr = r0, s = s0, qd = kTraceHalfW;                   // qd: row - (diagonal - kTraceHalfW), the unclamped slice index
                const int r_stop = v_stop - v0 - 64 * b, s_stop = h_stop - h0 - 64 * k;   // walk limits in tile coordinates
                const int r_lim = max(r_stop, 0), s_lim = max(s_stop, 0);
                while (rr >= r_lim && s >= s_lim && (unsigned)qd <= 15u) {
                    const u32 t2 = (planes[s * T] >> (rr - trace_slice_lo(lo0, s))) & 0x10001u;
                    const u32 a = t2 & 1u, bb = t2 >> 16;
                    const int op = (int)(((a ^ bb) << 1) | a);             // (1,0) D = 3, (0,1) I = 2, (1,1) X = 1, (0,0) M = 0
                    if (SCORE_ONLY) cost += (int)(a | bb);
                    else ow.emit(op);
                    qd += (int)bb - (int)a;
                    rr -= (int)(1u - (bb & ~a));
                    s -= (int)(1u - (a & ~bb));
                }
                v = v0 + 64 * b + rr;
                h = h0 + 64 * k + s;
            }
            if (SCORE_ONLY) { if (cost > hew_lim) ++hew; score += cost; }
            cv = v; ch = h;
        }
        WinOut wo;
        if (odd) { wo.score = 0; wo.hew = kWinPunted; outs[tk.slot] = wo; continue; }
        if (SCORE_ONLY) {
            if (ch >= 0) score += ch + 1;                                // bpm_windowed.c:599-607
            if (cv >= 0) score += cv + 1;
        } else {
            for (int h = ch; h >= 0; --h) ow.emit(OP_I);                 // :608-627
            for (int v = cv; v >= 0; --v) ow.emit(OP_D);
            ow.finish();
            LeafOut o;
            o.n_ops = tk.ops_cap - ow.pos; o.cost = ow.cost; o.text_len = -1; o.fmt = 0; o.pad_ = 0;
            leaf_outs[tk.leaf_slot] = o;
            score = ow.cost;
        }
        wo.score = score; wo.hew = hew;
        outs[tk.slot] = wo;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ws += __shfl_down_sync(kFull, ws, o);
    if ((t & 31) == 0 && ws) atomicAdd(&counters[0], ws);
}

// The WINDOWED algorithm planned on the device (reference run_windowed, quicked.c:91-123): one task, one pseudo-leaf and the
// leaf list of every pair, straight from its PairRec; pair i owns leaf slot i and the op words PairRec.ops_off.
__global__ void __launch_bounds__(256)
k_win_build(const PairRec *__restrict__ pairs, int n, int W, int O, int sse, WinTask *__restrict__ tasks, BandTask *__restrict__ leaves,
            PairLeaves *__restrict__ pl, int *__restrict__ status, int ok_status)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const PairRec r = pairs[i];
    WinTask t;
    t.p_off = r.p_off; t.t_off = r.t_off; t.m = r.m; t.n = r.n; t.rev = 0; t.W = W; t.O = O; t.hew_threshold = 0;
    t.sse = sse; t.score_only = 0; t.peq_off = r.peq_off; t.nbp = r.nbp; t.slot = i; t.scratch_off = 0;
    t.ops_off = r.ops_off; t.ops_cap = ((r.m + r.n + 15) / 16) * 16; t.leaf_slot = i;
    tasks[i] = t;
    PairLeaves p; p.first_leaf = i; p.n_leaves = 0; p.pad_ = 0;
    if (r.m > 0 && r.n > 0) {
        BandTask lf;
        memset(&lf, 0, sizeof lf);
        lf.p_off = r.p_off; lf.t_off = r.t_off; lf.m = r.m; lf.n = r.n; lf.pair = i;
        lf.ops_off = t.ops_off; lf.ops_cap = t.ops_cap; lf.slot = i;
        leaves[i] = lf;
        p.n_leaves = 1;
        status[i] = ok_status;
    }
    pl[i] = p;
}
// done[i] = 1 for the pairs k_windowed_tiles finished; the others (empty sides, tasks it flagged for the warp kernel) go
// through the planner to the host-driven path
__global__ void __launch_bounds__(256)
k_win_done(const PairRec *__restrict__ pairs, int n, const WinOut *__restrict__ outs, unsigned char *__restrict__ done, u64 *__restrict__ counters)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool ok = pairs[i].m > 0 && pairs[i].n > 0 && outs[i].hew != kWinPunted;
    done[i] = ok ? 1 : 0;
    if (ok) atomicAdd(&counters[2], 1ull);
}

}  // namespace qb
