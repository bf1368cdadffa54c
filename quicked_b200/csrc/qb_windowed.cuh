// qb_windowed.cuh — WindowEd kernels.
//
// k_windowed21_score: the QUICKED stage-1 bound estimator, WindowEd(S) = 2-word windows overlapping by 1 word,
// score only (reference quicked.c:178-199 -> bpm_windowed.c:563-628 with W=2, O=1, SCORE_ONLY).
//
// Mapping: ONE PAIR PER THREAD, persistent grid.  A 2-word window offers only a 2-way wavefront, so throughput comes
// from pairs in flight.  Only what the walk can read is kept: the bottom-right 64x64 quadrant of the window (the
// reference keeps all 130 columns x 2 words, bpm_windowed.c:143) — for full windows as 64 u32 of walk decisions in
// shared memory (SLIM, below), otherwise as 65 columns of one funnel-shifted 64-row slice of Pv and Mv in an L2/HBM
// scratch.  The 10 window-aligned match masks live in shared memory ([slot][thread], conflict-free).
//
// SSE=true reproduces the observable behaviour of windowed_compute_window_sse (bpm_windowed.c:283-445), which is
// what the reference runs on x86 unless force_scalar is set (dispatch :577); SSE=false follows the scalar
// windowed_compute_window (:202-280).  The differences are listed in SURVEY.md App. A.4 and DESIGN.md.
#pragma once
#include "qb_common.cuh"
#include "qb_traceback.cuh"
#include "qb_tiles.cuh"

namespace qb {

constexpr int kWsThreads = 128;                 // threads per CTA of the WindowEd(S) kernel
constexpr int kWsCtasPerSm = 6;                 // launch bound: register budget 85 per thread
constexpr int kWsResidentCtas = 4;              // CTAs per SM actually launched (SLIM: 43.5 KB of shared memory each; measured best)
constexpr int kWsQuadSlots = 65 * 2;            // u64 slots per thread in the quadrant scratch ([slot][thread] layout)

// SLIM: instead of the 64-row words of Pv and Mv, column s (1..64) of the quadrant keeps the walk's DECISION for the
// 16 rows around its diagonal (see ws21_pair), two bit planes packed into one u32 of shared memory at sq[s * T]:
//   plane a (bits 0..15)  = Pv[s] | (~Mv[s-1] & ~Eq[s])        plane b (bits 16..31) = ~Pv[s] & (Mv[s-1] | ~Eq[s])
//   (a,b) = (1,0) deletion, (0,1) insertion, (1,1) mismatch, (0,0) codes match   — the order of bpm_windowed.c:520-545
// for rows s-9..s+6, read at slice index d = row - jp + 9 when the walk stands in column jp = s.
__device__ __forceinline__ u32 slice16(u64 x, int lo)       // bits [lo, lo+16) of x, zero outside 0..63; -16 <= lo < 64
{
    const u64 y = lo >= 0 ? (x >> lo) : (x << (-lo));
    return (u32)y & 0xffffu;
}
__device__ __forceinline__ u32 slim_entry(u64 pv, u64 mv_prev, u64 eq, int s)
{
    const u64 a = pv | ~(mv_prev | eq), b = ~pv & (mv_prev | ~eq);
    return slice16(a, s - 9) | (slice16(b, s - 9) << 16);
}

// Eight columns of a FULL window (2 words x 128 columns; the walk's quadrant is word 1 of columns 64..128), the case
// of all but the last one or two windows of a pair.  STORE 0: nothing is kept (columns 0..63), 1: Pv/Mv of word 1 go
// to the L2/HBM scratch, 2: the walk decisions go to shared memory (SLIM).  w0/w1: the eight columns' code bytes.
// The body is kept this small on purpose: the kernel's hot loop has to stay inside the 32 KB instruction cache.
template <bool SSE, int STORE, int T = kWsThreads, int NC = kAlpha>
__device__ __forceinline__ void ws_full_group8(u32 w0, u32 w1, int c0, u32 top_in, const u64 *weq,
                                               u64 &pv0, u64 &mv0, u64 &pv1, u64 &mv1, u64 &pv1_prev, u64 &mv1_prev,
                                               u64 *qpv, u64 *qmv, i64 nthr, u32 *sq)
{
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int c = c0 + k;
        const u32 code = ((k < 4 ? w0 : w1) >> (8 * (k & 3))) & 7u;
        u32 hp_in0 = top_in;
        if (SSE) hp_in0 = (c == 0) ? top_in : ((c == 1) ? 1u : (u32)((k & 1) ^ 1));   // c0 is a multiple of 8: parity(c) = parity(k)
        // word 0 (MHin = 0, PHin = hp_in0), then word 1 with word 0's carries taken straight from the top bits of its
        // Ph / Mh: the adder takes MHin through the carry flag and the two shifts are funnel shifts (the forms of the
        // BandEd tile step, qb_tiles.cuh; bit-identical to two BPM_ADVANCE_BLOCKs, bpm_commons.h:49-68)
        u32 ph0_hi, mh0_hi;
        {
            const u64 eq0 = weq[code * T];
            const u64 xv = eq0 | mv0;
            const u64 xh = (((eq0 & pv0) + pv0) ^ pv0) | eq0;
            const u64 ph = mv0 | ~(xh | pv0), mh = pv0 & xh;
            ph0_hi = (u32)(ph >> 32); mh0_hi = (u32)(mh >> 32);
            const u64 ph_s = (ph << 1) | (u64)hp_in0, mh_s = mh << 1;
            pv0 = mh_s | ~(xv | ph_s);
            mv0 = ph_s & xv;
        }
        if (SSE && c == 127) { pv1_prev = pv1; mv1_prev = mv1; }
        const u64 eq1 = weq[(NC + code) * T], mv1_before = mv1;
        {
            const u64 xv = eq1 | mv1;
            const u64 xh = (add_with_top_bit(eq1 & pv1, pv1, mh0_hi) ^ pv1) | eq1;
            const u64 ph = mv1 | ~(xh | pv1), mh = pv1 & xh;
            const u32 phl = (u32)ph, phh = (u32)(ph >> 32), mhl = (u32)mh, mhh = (u32)(mh >> 32);
            const u64 ph_s = ((u64)fsl32(phl, phh, 1) << 32) | (u64)fsl32(ph0_hi, phl, 1);
            const u64 mh_s = ((u64)fsl32(mhl, mhh, 1) << 32) | (u64)fsl32(mh0_hi, mhl, 1);
            pv1 = mh_s | ~(xv | ph_s);
            mv1 = ph_s & xv;
        }
        if (STORE == 2) sq[(c - 63) * T] = slim_entry(pv1, mv1_before, eq1, c - 63);
        if (STORE == 1) {
            qpv[(i64)(c - 63) * nthr] = pv1;
            qmv[(i64)(c - 63) * nthr] = mv1;
        }
    }
}

// WindowEd(S) of ONE pair by one thread (see k_windowed21_score).  weq: this thread's 10 shared-memory slots
// (stride kWsThreads); qpv/qmv: this thread's quadrant scratch (slot stride nthr).
//
// SLIM: full windows (all but the last one or two of a pair) keep their quadrant in shared memory as 64 u32 of walk
// decisions (slim_entry): the walk starts on the diagonal (row 63 of column 64) and only an insertion or a deletion
// moves it off, so it stays within the 16 rows kept per column unless the 64-column window holds 8 more of one than
// of the other (about 1 window in 10^4 at 10 % error).  If it does leave the slice, the window is simply recomputed
// with the full 64-row quadrant in the L2/HBM scratch (the path non-full windows always take), so the result never
// depends on the slice width.  The walk then touches global memory only for characters outside "ACGTN".
//
// T: threads per CTA (the stride of the shared-memory slots).  NC: match-mask rows kept per window word — kAlpha, or 4
// for pairs whose TEXT holds only A, C, G, T (k_build_peq_pairs flags the others): their columns never look the code-4
// row up, except the SSE look-ahead past the end of the text, which takes it from a register.  64 + 260 bytes of shared
// memory per pair instead of 80 + 260 is what lets 704 pairs share an SM (k_windowed21_score, COMPACT).
template <bool SSE, bool SLIM, int T = kWsThreads, int NC = kAlpha>
__device__ __forceinline__ void ws21_pair(const PairRec &pr, const unsigned char *__restrict__ codes,
                                          const unsigned char *__restrict__ raw, const u64 *__restrict__ peq, u64 *weq,
                                          u64 *qpv, u64 *qmv, i64 nthr, u32 *sq, bool slim_ok, int hew_lim, int &score_out,
                                          int &hew_out, u64 &ws)
{
        int score = 0, hew = 0;
        int cv = pr.m - 1, ch = pr.n - 1;                    // corner (pos_v,pos_h), bpm_windowed.c:148-149
        if (pr.m > 0 && pr.n > 0) {
            const unsigned char *tc = codes + pr.t_off;
            const unsigned char *traw = raw + pr.t_off, *praw = raw + pr.p_off;
            const u64 *pq = peq + pr.peq_off;
            bool wide = false;                               // this window left the slim slice: redo it with the full quadrant
            while (cv >= 0 && ch >= 0) {
                // The lanes of a warp run their FULL windows first and their last, non-full ones (other code: the quadrant in
                // the L2 scratch, a walk that loads per step) together: pairs differ by a window or two, and a lane that
                // entered its tail alone made the whole warp sit through that tail's load latencies — up to 32 times per warp
                // (ncu on 12.5 k pairs: 19 % of the kernel's stall samples on the tail walk's raw-byte compare).
                {
                    const bool want_full = cv >= 127 && ch >= 127;
                    const unsigned others_full = __ballot_sync(__activemask(), want_full);
                    if (!want_full && others_full) continue;
                }
                // ---- window geometry (bpm_windowed.c:219-232) ----
                const int v0 = max(cv - 127, 0), h0 = max(ch - 127, 0);
                const int words = ((cv - v0) >> 6) + 1, cols = ch - h0 + 1;
                const int v_stop = max(cv - 63, 0), h_stop = max(ch - 63, 0);
                const int cs = h_stop - h0;                  // first stored column index the walk can touch
                const unsigned r0 = (unsigned)(v_stop - v0); // first row of the 64-row slice, 0..64
                // ---- match masks re-aligned to the window origin (bpm_windowed.c:237-244) ----
                u64 eq0_code4;                               // word 0's row of code 4 (only read when NC < kAlpha)
                {
                    const unsigned sh = v0 & 63;
                    const int blk0 = v0 >> 6;
                    u64 a[kAlpha], b[kAlpha], c2[kAlpha];
                    {
                        // three consecutive 48-byte blocks: nine 16-byte loads from five DRAM sectors
                        const ulonglong2 *q = reinterpret_cast<const ulonglong2 *>(pq + (i64)blk0 * kPeqStride);
                        const ulonglong2 z = make_ulonglong2(0, 0);
                        const ulonglong2 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2), q3 = __ldg(q + 3), q4 = __ldg(q + 4),
                                         q5 = __ldg(q + 5);
                        const ulonglong2 q6 = (words == 2) ? __ldg(q + 6) : z, q7 = (words == 2) ? __ldg(q + 7) : z,
                                         q8 = (words == 2) ? __ldg(q + 8) : z;
                        a[0] = q0.x; a[1] = q0.y; a[2] = q1.x; a[3] = q1.y; a[4] = q2.x;
                        b[0] = q3.x; b[1] = q3.y; b[2] = q4.x; b[3] = q4.y; b[4] = q5.x;
                        c2[0] = q6.x; c2[1] = q6.y; c2[2] = q7.x; c2[3] = q7.y; c2[4] = q8.x;
                    }
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        weq[c * T] = funnel_r(a[c], b[c], sh);
                        weq[(NC + c) * T] = (words == 2) ? funnel_r(b[c], c2[c], sh) : 0ull;
                    }
                    eq0_code4 = funnel_r(a[kAlpha - 1], b[kAlpha - 1], sh);
                }
                u64 pv0 = (h0 == 0) ? ~0ull : 0ull, pv1 = pv0, mv0 = 0, mv1 = 0;   // :225-229
                const u32 top_in = (v0 == 0);                                      // :247-252
                u64 pv1_prev = 0, mv1_prev = 0;
                if (cs == 0) {
                    qpv[0] = (r0 >= 64) ? pv1 : funnel_r(pv0, pv1, r0);
                    qmv[0] = 0;
                }
                const bool full = (words == 2 && cols == 128 && cv >= 127);      // => r0 == 64, cs == 64
                const bool slim = SLIM && slim_ok && full && !wide;
                // the window's codes come in as aligned 16-byte chunks (one load per 16 columns, one chunk ahead in
                // flight) and are realigned in registers; the flat code buffer is readable 48 B past its end
                const int csh = (int)((unsigned long long)(tc + h0) & 15ull);
                const uint4 *cvec = reinterpret_cast<const uint4 *>(tc + h0 - csh);
                // cpre: the chunk after cnxt, loaded a whole rotation before it is moved (a register move of a value still
                // in flight waits for it: with the load issued in the rotation that consumed it, every 16 columns stalled
                // on an L2 round trip when the warp had the SM sub-partition to itself)
                uint4 ccur = __ldg(cvec), cnxt = __ldg(cvec + 1), cpre = __ldg(cvec + 2);
                u32 code_last = 4;
                if (full) {
                    uint4 r = make_uint4(0, 0, 0, 0);
#pragma unroll 1
                    for (int q = 0; q < 16; ++q) {           // sixteen groups of eight columns
                        if (!(q & 1)) {
                            r = realign16(ccur, cnxt, csh);
                            ccur = cnxt; cnxt = cpre;
                            cpre = __ldg(cvec + min((q >> 1) + 3, 8));
                        }
                        const u32 w0 = (q & 1) ? r.z : r.x, w1 = (q & 1) ? r.w : r.y;
                        if (q < 8) {
                            ws_full_group8<SSE, 0, T, NC>(w0, w1, q * 8, top_in, weq, pv0, mv0, pv1, mv1, pv1_prev, mv1_prev, qpv, qmv, nthr, sq);
                            if (q == 7 && !slim) { qpv[0] = pv1; qmv[0] = mv1; }     // stored column 0: the state after window column 63
                        } else if (slim) {
                            ws_full_group8<SSE, 2, T, NC>(w0, w1, q * 8, top_in, weq, pv0, mv0, pv1, mv1, pv1_prev, mv1_prev, qpv, qmv, nthr, sq);
                        } else {
                            ws_full_group8<SSE, 1, T, NC>(w0, w1, q * 8, top_in, weq, pv0, mv0, pv1, mv1, pv1_prev, mv1_prev, qpv, qmv, nthr, sq);
                        }
                    }
                    code_last = r.w >> 24;
                }
                {
                    uint4 r = make_uint4(0, 0, 0, 0);
#pragma unroll 1
                    for (int c0 = full ? cols : 0; c0 < cols; c0 += 8) {
                        if (!(c0 & 8)) {
                            r = realign16(ccur, cnxt, csh);
                            ccur = cnxt; cnxt = cpre;
                            cpre = __ldg(cvec + min((c0 >> 4) + 3, ((cols - 1) >> 4) + 2));     // never past what the loop reads anyway
                        }
                        const u32 w0 = (c0 & 8) ? r.z : r.x, w1 = (c0 & 8) ? r.w : r.y;
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const int c = c0 + k;
                            if (c < cols) {
                                const int code = (int)(((k < 4 ? w0 : w1) >> (8 * (k & 3))) & 7u);
                                if (c == cols - 1) code_last = (u32)code;
                                u32 hp_in0 = top_in;
                                if (SSE && c > 0) hp_in0 = (c == 1) | ((c & 1) ^ 1);       // :348,:393,:424
                                u32 hp, hm, o1, o2;
                                myers_step(weq[code * T], pv0, mv0, hp_in0, 0u, hp, hm);
                                if (words == 2) {
                                    if (SSE && c == cols - 1) { pv1_prev = pv1; mv1_prev = mv1; }
                                    if (SSE && cols == 1) { hp = 0; hm = 0; }              // uninitialised carry in the reference
                                    myers_step(weq[(NC + code) * T], pv1, mv1, hp, hm, o1, o2);
                                }
                                if (c + 1 >= cs) {
                                    const i64 s = (i64)(c + 1 - cs) * nthr;
                                    qpv[s] = (r0 >= 64) ? pv1 : funnel_r(pv0, pv1, r0);
                                    qmv[s] = (r0 >= 64) ? mv1 : funnel_r(mv0, mv1, r0);
                                }
                            }
                        }
                    }
                }
                if (SSE && words == 2 && !(cols & 1)) {
                    // look-ahead column of word 0 (:361) and the redone last column of word 1 (:428-444)
                    const int ti = h0 + cols;
                    const int code_la = (ti < pr.n) ? (int)(tc[ti] & 7) : 4;
                    u64 lpv = pv0, lmv = mv0;
                    u32 hpL, hmL, o1, o2;
                    myers_step((NC < kAlpha && code_la >= NC) ? eq0_code4 : weq[code_la * T], lpv, lmv, 1u, 0u, hpL, hmL);
                    pv1 = pv1_prev; mv1 = mv1_prev;
                    const u64 eq1 = weq[(NC + (code_last & 7u)) * T];
                    myers_step(eq1, pv1, mv1, hpL, hmL, o1, o2);
                    if (slim) sq[64 * T] = slim_entry(pv1, mv1_prev, eq1, 64);
                    else {
                        const i64 s = (i64)(cols - cs) * nthr;
                        qpv[s] = (r0 >= 64) ? pv1 : funnel_r(pv0, pv1, r0);
                        qmv[s] = (r0 >= 64) ? mv1 : funnel_r(mv0, mv1, r0);
                    }
                }
                // ---- walk back through the non-overlapping 64 rows/columns (bpm_windowed.c:504-561): D, I, M, X ----
                // The next column's words are prefetched one column ahead, and the raw-byte compare of a diagonal
                // step (it only decides the cost, never the path) is consumed one step later.
                int v = cv, h = ch, cost = 0;
                int jp = h - h_stop + 1;                     // stored column index of Pv for the current h
                u32 pend_t = 0, pend_p = 0;
                if (slim) {
                    // a diagonal step costs 1 iff the raw characters differ (bpm_windowed.c:540); slim pairs hold only
                    // "ACGTN", for which that is the same as "the codes differ": no global memory in this loop
                    u32 w = sq[64 * T], wn = sq[63 * T];
                    int d = 8;                               // slice index of the current cell: row - jp + 9
                    while (v >= v_stop && jp >= 1 && (unsigned)d < 16u) {
                        const u32 a = (w >> d) & 1u, b = (w >> (16 + d)) & 1u;
                        cost += (int)(a | b);
                        const int del = (int)(a & ~b), ins = (int)(b & ~a);
                        v -= 1 - ins;                        // a deletion or a diagonal step consumes a pattern row
                        d += ins - del;
                        if (!del) {                          // an insertion or a diagonal step consumes a text column
                            --h; --jp;
                            w = wn;
                            wn = sq[max(jp - 1, 1) * T];
                        }
                    }
                    if ((unsigned)d >= 16u) { wide = true; continue; }     // left the slice: same window again, full quadrant
                } else {
                    u64 dp = qpv[(i64)jp * nthr], im = qmv[(i64)(jp - 1) * nthr];
                    u64 dpn = 0, imn = 0;
                    if (jp >= 2) { dpn = qpv[(i64)(jp - 1) * nthr]; imn = qmv[(i64)(jp - 2) * nthr]; }
                    while (v >= v_stop && jp >= 1) {
                        const int bit = v - v_stop;
                        cost += (pend_t != pend_p);
                        pend_t = pend_p = 0;
                        if ((dp >> bit) & 1ull) { ++cost; --v; }
                        else {
                            if ((im >> bit) & 1ull) ++cost;
                            else { pend_t = traw[h]; pend_p = praw[v]; --v; }
                            --h; --jp;
                            dp = dpn; im = imn;
                            if (jp >= 2) { dpn = qpv[(i64)(jp - 1) * nthr]; imn = qmv[(i64)(jp - 2) * nthr]; }
                        }
                    }
                }
                wide = false;
                ws += (u64)(words * cols);
                cost += (pend_t != pend_p);
                if (cost > hew_lim) ++hew;
                score += cost;
                cv = v; ch = h;
            }
            if (ch >= 0) score += ch + 1;      // bpm_windowed.c:599-607
            if (cv >= 0) score += cv + 1;
        }
        score_out = score; hew_out = hew;
}

// COMPACT (SLIM only): CTAs of kWsCompactThreads threads, two per SM, four match-mask rows per word — 704 pairs per SM
// instead of 640.  A pair is a chain of ~n/64 windows, so a batch that overflows the resident threads by a few pairs
// pays a whole second chain for them: 100 k pairs of 10 kbp (BASELINE configs[2]) are 94 720 + 5 280 with 128-thread
// CTAs (10.7 ms) and one wave here.  Pairs flagged by the table builder (a character outside "ACGT" in the text, or
// outside "ACGTN" anywhere) are skipped by the COMPACT kernel and done by a launch of the plain one with only_flagged.
constexpr int kWsCompactThreads = 352;
__host__ __device__ constexpr size_t ws_smem_bytes(bool slim, int T, int NC) { return (size_t)T * (2 * NC * 8 + (slim ? 65 * 4 : 0)) + (slim ? 0 : 16); }

template <bool SSE, bool SLIM, bool COMPACT = false>
__global__ void __launch_bounds__(COMPACT ? kWsCompactThreads : kWsThreads, COMPACT ? 2 : kWsCtasPerSm)
k_windowed21_score(const PairRec *__restrict__ pairs, int n_pairs, const unsigned char *__restrict__ codes,
                   const unsigned char *__restrict__ raw, const u64 *__restrict__ peq, int hew_threshold,
                   int *__restrict__ bound, int *__restrict__ hew_out, u64 *__restrict__ counters,
                   u64 *__restrict__ quad, const unsigned char *__restrict__ pair_odd, int only_flagged)
{
    constexpr int T = COMPACT ? kWsCompactThreads : kWsThreads, NC = COMPACT ? 4 : kAlpha;
    extern __shared__ __align__(16) unsigned char ws_smem[];
    u64 *s_weq = reinterpret_cast<u64 *>(ws_smem);                         // [2 * NC][T] window-aligned match masks
    u32 *s_slim = reinterpret_cast<u32 *>(ws_smem + (size_t)2 * NC * T * 8);  // SLIM: [65][T] slim quadrant of full windows (see ws21_pair)
    const int t = threadIdx.x;
    u64 *weq = s_weq + t;
    const i64 gtid = (i64)blockIdx.x * T + t, nthr = (i64)gridDim.x * T;
    // Quadrant scratch of non-full and redone windows: [slot][resident thread] in HBM/L2 (reused window after window).
    u64 *qpv = quad + gtid;                                  // slot s at qpv[s * nthr]
    u64 *qmv = quad + 65 * nthr + gtid;
    const int hew_lim = 64 * hew_threshold / 100;            // (W-O)*64*thr/100, bpm_windowed.c:555
    u64 ws = 0;
    for (i64 i = gtid; i < n_pairs; i += nthr) {
        const bool flagged = pair_odd[i] != 0;
        if (COMPACT ? flagged : (only_flagged && !flagged)) continue;
        const PairRec pr = pairs[i];
        int score = 0, hew = 0;
        ws21_pair<SSE, SLIM, T, NC>(pr, codes, raw, peq, weq, qpv, qmv, nthr, s_slim + (SLIM ? t : 0), SLIM && !flagged, hew_lim, score, hew, ws);
        bound[i] = score;
        hew_out[i] = hew;
    }
    // device-side word-step count (roofline numerator), one atomic per warp
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ws += __shfl_down_sync(kFull, ws, o);
    if ((t & 31) == 0 && ws) atomicAdd(&counters[0], ws);
}

// ----------------------------------------------------------------------------------------------------------------
// k_windowed_warp: WindowEd for ANY window / overlap (W <= 32 words), score-only or CIGAR mode, forward or
// reversed sequences.  ONE PAIR PER WARP, one window word per lane, all lanes on the same column with the same
// ballot carry-lookahead as k_banded_warp (the window has no band bookkeeping: reference bpm_windowed.c:202-280).
// The filled window (64W+2 columns x W words of (Pv,Mv)) goes to a per-warp scratch in HBM/L2 exactly as the
// reference stores it (bpm_windowed.c:143); lane 0 then walks it (bpm_windowed.c:448-561).
// Used by: QUICKED stage 2 (WindowEd(L) forward and reverse, quicked.c:204-233), the WINDOWED algorithm
// (quicked.c:91-123).  For W == 2 and !force_scalar the SSE variant's quirks are reproduced (see k_windowed21_score).
struct WinTask {
    i64 p_off, t_off;      // pattern / text start in the packed buffer (forward coordinates)
    int m, n;
    int rev;               // 1: run on the reversed sequences
    int W, O;
    int hew_threshold;
    int sse;               // emulate windowed_compute_window_sse (only when W == 2)
    int score_only;
    i64 peq_off; int nbp;  // match masks of the (possibly reversed) pattern
    int slot;              // output index
    i64 scratch_off;       // ulonglong2 index of this warp's window scratch ((64W+3)*W entries)
    i64 ops_off; int ops_cap;   // CIGAR mode: op region
    int leaf_slot;              // CIGAR mode: index of the pseudo-leaf's LeafOut
};
struct WinOut { int score; int hew; };

__global__ void __launch_bounds__(128)
k_windowed_warp(const WinTask *__restrict__ tasks, int n_tasks, const unsigned char *__restrict__ codes,
                const unsigned char *__restrict__ raw, const u64 *__restrict__ peq, ulonglong2 *__restrict__ scratch,
                u32 *__restrict__ ops_pool, WinOut *__restrict__ outs, LeafOut *__restrict__ leaf_outs,
                u64 *__restrict__ counters, int only_punted)
{
    const int lane = threadIdx.x & 31;
    const int id = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (id >= n_tasks) return;
    const WinTask tk = tasks[id];
    // only_punted: the tile kernel (qb_wintile.cuh) went first and flagged the tasks it leaves to this one
    if (only_punted && outs[tk.slot].hew != -2147483647 - 1) return;
    const int W = tk.W, O = tk.O;
    const u64 *pq = peq + tk.peq_off;
    const unsigned char *tc = codes + tk.t_off, *traw = raw + tk.t_off, *praw = raw + tk.p_off;
    ulonglong2 *win = scratch + tk.scratch_off;                 // [column][word]
    const bool sse = tk.sse && W == 2;
    OpWriter ow;
    if (!tk.score_only && lane == 0) ow.init(ops_pool + tk.ops_off, tk.ops_cap);
    int cv = tk.m - 1, ch = tk.n - 1, score = 0, hew = 0;
    u64 ws = 0;
    const int hew_lim = (W - O) * 64 * tk.hew_threshold / 100;
    while (cv >= 0 && ch >= 0) {
        const int v0 = max(cv - 64 * W + 1, 0), h0 = max(ch - 64 * W + 1, 0);
        const int words = ((cv - v0) >> 6) + 1, cols = ch - h0 + 1;
        const bool act = lane < words;
        // window-aligned match masks of this lane's word (bpm_windowed.c:237-244)
        u64 eqw[kAlpha];
        {
            const unsigned sh = v0 & 63;
            const int blk = (v0 >> 6) + lane;
#pragma unroll
            for (int c = 0; c < kAlpha; ++c) {
                u64 e = 0;
                if (act) e = funnel_r(pq[(i64)blk * kPeqStride + c], pq[(i64)(blk + 1) * kPeqStride + c], sh);
                eqw[c] = e;
            }
        }
        u64 pv = (h0 == 0) ? ~0ull : 0ull, mv = 0;
        if (lane < W) win[lane] = make_ulonglong2(pv, 0ull);   // column 0 (bpm_windowed.c:225-229)
        const u32 top_in = (v0 == 0);
        u64 pv_prev = 0, mv_prev = 0;
        const int ncol = (sse && words == 2 && !(cols & 1)) ? cols + 1 : cols;   // + SSE look-ahead column
        for (int c = 0; c < ncol; ++c) {
            const bool la = c >= cols;                           // look-ahead pass: lane 0 = column `cols`, lane 1 redoes column cols-1
            int col_t = h0 + c;
            if (la && lane == 1) col_t = h0 + cols - 1;
            int code = 4;
            if (col_t < tk.n) code = (tk.rev ? tc[tk.n - 1 - col_t] : tc[col_t]) & 7;
            if (sse && c == cols - 1) { pv_prev = pv; mv_prev = mv; }
            if (la && lane == 1) { pv = pv_prev; mv = mv_prev; }
            u64 eq = eqw[0];
#pragma unroll
            for (int k = 1; k < kAlpha; ++k) if (code == k) eq = eqw[k];
            u32 hp_top = top_in;
            if (sse && c > 0) hp_top = (c == 1) | ((c & 1) ^ 1);
            const u64 a = eq & pv, s = a + pv;
            const u32 G = __ballot_sync(kFull, act && (s < a));
            const u32 P = __ballot_sync(kFull, act && (s == ~0ull));
            const u32 X = G | P;
            u32 carries = (X + G) ^ X ^ G;                       // carry into lane l (top MHin = 0)
            u32 my_c = (carries >> lane) & 1u;
            if (sse && cols == 1 && lane == 1) my_c = 0;         // uninitialised carry in the reference (single-column window)
            const u64 xh = ((s + my_c) ^ pv) | eq;
            u64 ph = mv | ~(xh | pv);
            u64 mh = pv & xh;
            const u32 HP = __ballot_sync(kFull, (ph >> 63) != 0);
            u32 hp_in = lane ? ((HP >> (lane - 1)) & 1u) : hp_top;
            if (sse && cols == 1 && lane == 1) hp_in = 0;
            ph = (ph << 1) | (u64)hp_in;
            mh = (mh << 1) | (u64)my_c;
            const u64 xv = eq | mv;
            pv = mh | ~(xv | ph);
            mv = ph & xv;
            if (!la) { if (act) win[(i64)(c + 1) * W + lane] = make_ulonglong2(pv, mv); }
            else if (lane == 1) win[(i64)cols * W + 1] = make_ulonglong2(pv, mv);
        }
        ws += (u64)words * cols;
        __syncwarp();
        // ---- walk (lane 0): D, I, M, X when score-only; M(raw) first, then D, I, X in CIGAR mode ----
        int v = cv, h = ch;
        if (lane == 0) {
            const int v_stop = max(cv - 64 * (W - O) + 1, 0), h_stop = max(ch - 64 * (W - O) + 1, 0);
            int cost = 0;
            while (v >= v_stop && h >= h_stop) {
                const int word = (v - v0) >> 6, bit = (v - v0) & 63;
                const u64 dp = win[(i64)(h - h0 + 1) * W + word].x, im = win[(i64)(h - h0) * W + word].y;
                const bool del = (dp >> bit) & 1ull, ins = (im >> bit) & 1ull;
                const unsigned char tch = tk.rev ? traw[tk.n - 1 - h] : traw[h], pch = tk.rev ? praw[tk.m - 1 - v] : praw[v];
                const bool same = tch == pch;
                if (tk.score_only) {
                    if (del) { ++cost; --v; }
                    else if (ins) { ++cost; --h; }
                    else { cost += !same; --h; --v; }
                } else {
                    if (same) { ow.emit(OP_M); --h; --v; }
                    else if (del) { ow.emit(OP_D); --v; }
                    else if (ins) { ow.emit(OP_I); --h; }
                    else { ow.emit(OP_X); --h; --v; }
                }
            }
            if (tk.score_only) { if (cost > hew_lim) ++hew; score += cost; }
        }
        cv = __shfl_sync(kFull, v, 0);
        ch = __shfl_sync(kFull, h, 0);
        __syncwarp();
    }
    if (lane == 0) {
        if (tk.score_only) {
            if (ch >= 0) score += ch + 1;                        // bpm_windowed.c:599-607
            if (cv >= 0) score += cv + 1;
        } else {
            for (int h = ch; h >= 0; --h) ow.emit(OP_I);         // :608-627
            for (int v = cv; v >= 0; --v) ow.emit(OP_D);
            ow.finish();
            LeafOut o;
            o.n_ops = tk.ops_cap - ow.pos; o.cost = ow.cost; o.text_len = ow.text_len; o.fmt = 0; o.pad_ = 0;
            leaf_outs[tk.leaf_slot] = o;
            score = ow.cost;
        }
        WinOut wo; wo.score = score; wo.hew = hew;
        outs[tk.slot] = wo;
        atomicAdd(&counters[0], ws);
    }
}

}  // namespace qb
