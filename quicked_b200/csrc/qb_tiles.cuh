// qb_tiles.cuh — BandEd as a TILE DATAFLOW: the B200-first form of the reference's banded bit-parallel kernel
// (reference bpm_banded.c:199-316 full matrix, :791-964 score-only; band bookkeeping :264-301 / :889-922).
//
// Unit of work = one TILE: one 64-row pattern block x one 64-column text block (64 word-steps), run by ONE LANE with
// the block's Pv/Mv in registers.  The only coupling between vertically adjacent tiles is the pair of horizontal
// carries per column (PHout/MHout of BPM_ADVANCE_BLOCK, bpm_commons.h:60-61): 64 columns of them are two 64-bit
// words.  They travel through a shift register: the carry-in of step s sits at the top bit, leaves into the Ph/Mh
// shift (one funnel shift that the 64-bit shift needed anyway) and the carry-out of the same step enters at the
// bottom, so after 64 steps the register holds the tile's carry-outs in the order the tile below consumes them.
// Running scores (reference scores[], bpm_banded.c:260) are popcounts of those words, once per tile.
//
// Tiles of a task form a wavefront: tile (block b, column block k) runs at round base[k] + b, exactly one round after
// the tile above it (b-1, k) and after its own predecessor (b, k-1).  A persistent CTA keeps `nslots` tasks resident
// and a scheduler warp (lane = task slot) packs the ready tiles of ALL slots onto the CTA's compute lanes every
// round, so lanes are never tied to one task: the half-empty warps of the one-task-per-warp kernel (≈16 live blocks
// of a 29-block band at 10 kbp) disappear, and a single long pair spreads over the whole CTA (one pair per CTA at
// 100 kbp and beyond).  The reference's band decisions (lower cut / prolog widening, new bottom block, upper cut /
// pattern-end clamp) are taken by the slot's scheduler lane between rounds, in the reference's order, as soon as the
// tiles they read have completed (`sched_advance`).
//
// FULL mode does not store the (n+1) x B matrix of the reference (bpm_banded.c:139-140): it stores one 32-byte
// RECORD per tile — the block's (Pv,Mv) at the tile's first column and the 64 carry-in pairs — from which the
// traceback kernel (qb_tiletrace.cuh) recomputes just the tiles the walk visits.  16 B per word-step become 0.5 B.
//
// Everything that decides results is __host__ __device__ so that tests/emu (tools/tile_emu.cu) runs the identical
// logic on the CPU against the oracle.
#pragma once
#include "qb_common.cuh"

namespace qb {

#define QB_HD __host__ __device__ __forceinline__

QB_HD u32 fsl32(u32 lo, u32 hi, u32 s)          // high word of (hi:lo) << s, 0 < s < 32
{
#ifdef __CUDA_ARCH__
    return __funnelshift_l(lo, hi, s);
#else
    return (hi << s) | (lo >> (32 - s));
#endif
}
QB_HD int popc32(u32 x)
{
#ifdef __CUDA_ARCH__
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
QB_HD int popc64(u64 x)
{
#ifdef __CUDA_ARCH__
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}

// 64 carry pairs of a tile edge.  p0/m0 hold steps 0..31, p1/m1 steps 32..63, step s at bit 31-(s&31).
struct TileCarry { u32 p0, p1, m0, m1; };

// One tile record of FULL mode (32 bytes): state of the block when the tile starts + its carry-ins.
struct TileRec { u64 pv0, mv0; TileCarry cin; };

// Ring size of a launch: RB block slots = RK column-block slots (power of two >= B+2).
QB_HD int tile_ring_for(i64 B)
{
    int r = 8;
    while (r < B + 2) r <<= 1;
    return r;
}
constexpr int kTileMaxRing = 1024;             // bands up to 1022 blocks; taller ones use the shared-memory sweep kernel
constexpr int kTilePunted = -2147483647 - 1;   // BandOut.pos_v of a FULL-mode task the tile fill gave up on

// Per-slot bookkeeping.  The scheduler lane keeps its slot in registers; compute lanes read the shared-memory copy
// (static fields written when the task is loaded, rho / kb refreshed every pass).
struct TileSlot {
    int task;                 // index into the task array, -1: empty
    int m, n, ncols, K, nshift;
    int nblk, mmod, clamp, prolog, B, rev, nbp, out_slot;
    int fin, kcut;            // band geometry (fits 32 bits: longer sequences are punted to the sweep kernels)
    i64 peq_off, t_off, rec_off, scores_off, state_off, range_off, tt_off;
    // dynamic
    int rho;                  // round being run / to run next
    int kt, kb;               // top[] / bot[] decided for column blocks <= kt / <= kb
    int kmin, kmax;           // tiles of round rho: column blocks kmin..kmax, block = rho - base[k]
    int cnt;                  // number of those tiles
    int state;                // 0 running, 1 finished, 2 punt (task handed to the exact fallback kernels)
    int koff;                 // tiles of the round already dispatched (a round may be spread over several passes)
    // ring values the scheduler needs every pass, mirrored in registers
    int c_top_t, c_base_t;                        // top[kt], base[kt]
    int c_top_b, c_bot_b, c_base_b, c_top_b1;     // top[kb], bot[kb], base[kb], top[kb+1] (valid once kt > kb)
    u64 ws;                   // word-steps of the task (reference schedule)
};

// Views of one slot's rings inside the CTA's arena.
struct TileRings {
    u64 *pv, *mv;             // [RB] state of the live blocks
    TileCarry *carry;         // [RB] carry-outs of a block's latest tile (read by the tile below BEFORE the mid-pass barrier)
    int *sc;                  // [2][RB] by column-block parity: running score after column block k
    int *top, *bot, *base;    // [RB] by column block
    int RB;
};
QB_HD size_t tile_slot_arena_bytes(int RB) { return (size_t)RB * (8 + 8 + 16 + 8 + 12); }
QB_HD TileRings tile_rings(unsigned char *arena, int RB)
{
    TileRings r;
    r.RB = RB;
    r.pv = reinterpret_cast<u64 *>(arena);
    r.mv = r.pv + RB;
    r.carry = reinterpret_cast<TileCarry *>(r.mv + RB);
    r.sc = reinterpret_cast<int *>(r.carry + RB);
    r.top = r.sc + 2 * RB;
    r.bot = r.top + RB;
    r.base = r.bot + RB;
    return r;
}

// Global pools a launch works on.
struct TilePools {
    const BandTask *tasks;
    const unsigned char *codes;     // base codes of the packed batch (bit 3 = odd-character flag, masked here)
    const u64 *peq;
    const u64 *ttext;               // per-task text codes, 8 per u64 in column order (reversed passes stored reversed), see k_tile_text
    TileRec *recs;                  // FULL: records, index rec_off + k*B + j  (j = band-relative word)
    int2 *ranges;                   // FULL: live range (first,last) per column block
    int *scores;                    // per-task running scores by absolute block (reference scores[])
    u64 *state;                     // score-only: exported final column, Pv[B] then Mv[B]
    BandOut *outs;
    int *punt_list; int *punt_count;
    i64 rec_sub;                    // subtracted from every task's mat_off (chunked pools)
};

// ---- slot set-up from a task --------------------------------------------------------------------------------------
template <bool FULL>
QB_HD void tile_slot_load(TileSlot &S, const TileRings &R, const BandTask &tk, int task_index, const TilePools &P)
{
    const BandGeom g = band_geometry(tk.m, tk.n, tk.cutoff);
    S.task = task_index;
    S.m = tk.m; S.n = tk.n; S.ncols = FULL ? tk.n : tk.finish;
    S.K = (S.ncols + 63) >> 6; S.nshift = S.ncols >> 6;
    S.nblk = (tk.m + 63) >> 6; S.mmod = tk.m & 63;
    S.clamp = FULL ? S.nblk - 1 : S.nblk;                           // bpm_banded.c:295 vs :917
    S.prolog = (int)g.prolog; S.B = (int)(FULL ? g.Bc : g.Bs);
    S.rev = tk.rev; S.nbp = tk.nbp; S.out_slot = tk.slot;
    S.fin = (int)g.fin; S.kcut = (int)g.k;
    S.peq_off = tk.peq_off; S.t_off = tk.t_off; S.rec_off = (tk.mat_off - P.rec_sub) / 2; S.scores_off = tk.scores_off;
    S.state_off = tk.state_off; S.range_off = tk.range_off; S.tt_off = tk.tt_off;
    S.rho = 0; S.kt = 0; S.kb = 0; S.kmin = 0; S.kmax = -1; S.cnt = 0; S.state = 0; S.koff = 0; S.ws = 0;
    R.top[0] = 0;                                                   // first + pos_v = prolog - prolog (bpm_banded.c:222-225)
    R.bot[0] = S.B - 1 - S.prolog;
    R.base[0] = 0;
    S.c_top_t = 0; S.c_base_t = 0; S.c_top_b = 0; S.c_bot_b = S.B - 1 - S.prolog; S.c_base_b = 0; S.c_top_b1 = 0;
    // 32-bit bookkeeping: |fin|, k and 64 * blocks must stay below 2^30
    if (g.k >= (1ll << 30) || g.fin >= (1ll << 30) || g.fin <= -(1ll << 30) || tk.m >= (1 << 30) || tk.n >= (1 << 30)) S.state = 2;
    int *gs = P.scores + S.scores_off;
    for (int j = 0; j < S.B; ++j) gs[j] = 64 * (j + 1);             // bpm_banded.c:180-197
    if (FULL) P.ranges[S.range_off] = make_int2(S.prolog, S.B - 1);
}

// ---- the reference's band decisions, taken as early as the data allows -------------------------------------------
// Column block k -> k+1 (bpm_banded.c:264-301 / :889-922), in absolute block numbers (top = first + pos_v,
// bot = last + pos_v, pos_v = k - prolog):
//   D1: cut_lo  -> top[k+1] in {top, top+1, top+2};   opens column block k+1 (base[k+1])
//   D2: new bottom block (score + 64), cut_hi / clamp -> bot[k+1] in {bot, bot+1}
// Tile (b, k) runs at round base[k] + b, so "tile complete" is base[k] + b < rho.  Each call takes at most ONE D1 and
// ONE D2 in straight-line code (the lanes of the scheduler warp stay converged); one per round is all a task can need:
// consecutive column blocks become due at least one round apart.
QB_HD void tile_try_d1(TileSlot &S, const TileRings &R)
{
    const int RBm = R.RB - 1;
    if (!(S.kt < S.nshift && S.kt - S.kb < R.RB - 4)) return;
    const int k = S.kt, top = S.c_top_t;
    const int first = top - (k - S.prolog);
    const int lo = S.c_bot_b, hi = lo + (k - S.kb);                  // bot[] never decreases and grows by <= 1 per block
    const int guard = (top + 2 < lo) ? 1 : (top + 2 >= hi) ? 0 : -1;   // first + 2 < last  <=>  top + 2 < bot[k]
    if (guard < 0) return;
    bool cut = false;
    if (guard && S.fin > 64 * (first + 1)) {
        if (S.c_base_t + top + 1 >= S.rho) return;                   // tile (top+1, k) has not completed yet
        cut = R.sc[(k & 1) * R.RB + ((top + 1) & RBm)] + (S.fin - 64 * (first + 1)) > S.kcut;
    }
    int ntop = top + 1;
    if (cut && k >= S.prolog) ++ntop;
    else if (!cut && k < S.prolog) --ntop;
    const int b0 = S.c_base_t + 1, b1 = S.rho - ntop;
    const int nbase = b0 > b1 ? b0 : b1;
    R.top[(k + 1) & RBm] = ntop;
    R.base[(k + 1) & RBm] = nbase;
    if (S.kb == k) S.c_top_b1 = ntop;
    S.c_top_t = ntop; S.c_base_t = nbase;
    S.kt = k + 1;
}

template <bool FULL>
QB_HD void tile_try_d2(TileSlot &S, const TileRings &R, const TilePools &P)
{
    const int RBm = R.RB - 1;
    if (!(S.kb < S.nshift && S.kt >= S.kb + 1)) return;
    const int k = S.kb, top = S.c_top_b, bot = S.c_bot_b;
    if (bot < top) { S.state = 2; return; }                          // empty band: the reference reads stale scores here
    if (S.c_base_b + bot >= S.rho) return;                           // tile (bot, k) has not completed yet
    const int pos_v = k - S.prolog;
    const int first1 = S.c_top_b1 - 1 - pos_v, last = bot - pos_v;
    const int nbs = R.sc[(k & 1) * R.RB + (bot & RBm)] + 64;         // scores[nb] = scores[nb-1] + 64
    R.sc[(k & 1) * R.RB + ((bot + 1) & RBm)] = nbs;
    R.pv[(bot + 1) & RBm] = ~0ull; R.mv[(bot + 1) & RBm] = 0ull;     // the block entering at the bottom (exported if the pass ends here)
    P.scores[S.scores_off + bot + 1] = nbs;
    const bool cut_hi = (first1 + 2 < last) && (64 * (last - 1) > S.fin) &&
                        (R.sc[(k & 1) * R.RB + ((bot - 1) & RBm)] + (64 * (last - 1) - S.fin) > S.kcut);
    const int nbot = (cut_hi || bot >= S.clamp) ? bot : bot + 1;
    R.bot[(k + 1) & RBm] = nbot;
    S.ws += (u64)(bot - top + 1) * 64;
    if (FULL) P.ranges[S.range_off + k + 1] = make_int2(S.c_top_b1 - (pos_v + 1), nbot - (pos_v + 1));
    // mirrors for kb = k + 1
    S.c_top_b = S.c_top_b1; S.c_bot_b = nbot;
    S.c_base_b = (S.kt == k + 1) ? S.c_base_t : R.base[(k + 1) & RBm];
    S.c_top_b1 = (S.kt == k + 2) ? S.c_top_t : (S.kt > k + 2 ? R.top[(k + 2) & RBm] : 0);
    S.kb = k + 1;
}

// Decisions + the tiles of round S.rho (column blocks kmin..kmax); returns their count.  Finished tasks: S.state = 1.
template <bool FULL>
QB_HD int tile_plan_round(TileSlot &S, const TileRings &R, const TilePools &P)
{
    const int RBm = R.RB - 1;
    tile_try_d1(S, R);
    tile_try_d2<FULL>(S, R, P);
    tile_try_d1(S, R);                                               // a new bot[kb] may have settled D1's guard
    if (S.state) { S.cnt = 0; return 0; }
    const int kopen = (S.kt < S.K - 1) ? S.kt : S.K - 1;
    if (S.kmax + 1 <= kopen) {                                       // at most one column block becomes due per round
        const int kk = S.kmax + 1;
        int bx, tx;
        if (kk == S.kt) { bx = S.c_base_t; tx = S.c_top_t; } else { bx = R.base[kk & RBm]; tx = R.top[kk & RBm]; }
        if (S.rho - bx >= tx) S.kmax = kk;
    }
    if (S.kmin <= S.kmax) {                                          // ... and at most one completes
        const int km = S.kmin;
        const int bm = (km == S.kb) ? S.c_base_b : (km == S.kt) ? S.c_base_t : R.base[km & RBm];
        const int lim = (km >= S.kb) ? S.c_bot_b : R.bot[km & RBm];  // undecided column blocks: the last decided bottom bounds theirs
        if (S.rho - bm > lim) {
            if (km > S.kb) { S.state = 2; S.cnt = 0; return 0; }    // bottom of an undecided column block: cannot happen (checked by the emulator)
            ++S.kmin;
        }
    }
    S.cnt = S.kmax - S.kmin + 1;
    if (S.cnt <= 0) {
        S.cnt = 0;
        if (S.kmin > S.K - 1 && S.kt == S.nshift && S.kb == S.nshift) S.state = 1;
    }
    return S.cnt;
}

// ---- results of a finished task (bpm_banded.c:952-963 and the band state Hirschberg reads) -------------------------
template <bool FULL>
QB_HD void tile_slot_finish(TileSlot &S, const TileRings &R, const TilePools &P)
{
    const int RBm = R.RB - 1;
    const int pos_v = S.nshift - S.prolog;
    const int top = R.top[S.nshift & RBm], bot = R.bot[S.nshift & RBm];
    if (S.ncols & 63) S.ws += (u64)(bot - top + 1 > 0 ? bot - top + 1 : 0) * (u64)(S.ncols & 63);
    int *gs = P.scores + S.scores_off;
    BandOut o;
    const int sfin = gs[S.nblk - 1];
    o.score = S.mmod ? sfin - (64 - S.mmod) : sfin;
    o.first = top - pos_v; o.last = bot - pos_v; o.pos_v = pos_v;
    P.outs[S.out_slot] = o;
    if (!FULL) {
        u64 *st = P.state + S.state_off;
        for (int j = 0; j < S.B; ++j) {
            const int blk = j + pos_v;
            const bool act = (j >= o.first && j <= o.last && blk >= 0);
            st[j] = act ? R.pv[blk & RBm] : 0ull;
            st[S.B + j] = act ? R.mv[blk & RBm] : 0ull;
        }
    }
}

// ---- one tile ------------------------------------------------------------------------------------------------------
// (a + b + (c >> 31)): on the device the top bit of c goes through the carry flag (add.cc c,c), 3 instructions in all
QB_HD u64 add_with_top_bit(u64 a, u64 b, u32 c)
{
#ifdef __CUDA_ARCH__
    u32 lo, hi;
    asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %6, %6;\n\taddc.cc.u32 %0, %2, %4;\n\taddc.u32 %1, %3, %5;\n\t}"
        : "=r"(lo), "=r"(hi)
        : "r"((u32)a), "r"((u32)(a >> 32)), "r"((u32)b), "r"((u32)(b >> 32)), "r"(c));
    return ((u64)hi << 32) | lo;
#else
    return a + b + (u64)(c >> 31);
#endif
}

#define QB_TILE_STEP(EQ)                                                                                    \
    {                                                                                                       \
        const u64 eq_ = (EQ);                                                                               \
        const u64 xv_ = eq_ | mv;                                                                           \
        const u64 xh_ = (add_with_top_bit(eq_ & pv, pv, wm) ^ pv) | eq_;                                    \
        const u64 ph_ = mv | ~(xh_ | pv);                                                                   \
        const u64 mh_ = pv & xh_;                                                                           \
        const u32 phl_ = (u32)ph_, phh_ = (u32)(ph_ >> 32), mhl_ = (u32)mh_, mhh_ = (u32)(mh_ >> 32);       \
        const u64 ph2_ = ((u64)fsl32(phl_, phh_, 1) << 32) | (u64)fsl32(wp, phl_, 1);                       \
        const u64 mh2_ = ((u64)fsl32(mhl_, mhh_, 1) << 32) | (u64)fsl32(wm, mhl_, 1);                       \
        wp = fsl32(phh_, wp, 1);                                                                            \
        wm = fsl32(mhh_, wm, 1);                                                                            \
        pv = mh2_ | ~(xv_ | ph2_);                                                                          \
        mv = ph2_ & xv_;                                                                                    \
    }

// byte i (0..3) of w, zero-extended
QB_HD u32 byte_of(u32 w, int i)
{
#ifdef __CUDA_ARCH__
    return __byte_perm(w, 0, 0x4440 + i);
#else
    return (w >> (8 * i)) & 0xffu;
#endif
}

// Full tile, carry-out at bit 63.  eq: this lane's five match masks, eq[code * EQS] (EQS = 0: runtime stride eqs);
// tt: the tile's 64 text codes, 8 per u64 (low byte = first column), chunk c at tt[c * TS] — TS = 1: the aligned
// tile-text pool in global memory, otherwise a shared-memory staging area; w/wn: its first two chunks, loaded by the
// caller long before (two chunks stay in flight: an L2 round trip outlasts 8 word-steps).
template <int EQS, int TS = 1>
QB_HD void tile_fill64(u64 &pv, u64 &mv, TileCarry &c, const u64 *eq, int eqs, const u64 *tt, u64 w, u64 wn)
{
    const int st = EQS ? EQS : eqs;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        u32 wp = half ? c.p1 : c.p0, wm = half ? c.m1 : c.m0;
#pragma unroll 1
        for (int it = 0; it < 4; ++it) {
            const u32 w0 = (u32)w, w1 = (u32)(w >> 32);
            w = wn;
#ifdef __CUDA_ARCH__
            wn = (TS == 1) ? __ldg(tt + ((half * 4 + it + 2) & 7)) : tt[((half * 4 + it + 2) & 7) * TS];
#else
            wn = tt[((half * 4 + it + 2) & 7) * TS];
#endif
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const u32 code = byte_of(i < 4 ? w0 : w1, i & 3);
                QB_TILE_STEP(eq[code * st]);
            }
        }
        if (half) { c.p1 = wp; c.m1 = wm; } else { c.p0 = wp; c.m0 = wm; }
    }
}

// Any tile: nc <= 64 columns, carry-out taken at bit `ob` (level_mask of the last pattern block, bpm_commons.h:60-61).
// The carry-outs of a short tile are left-aligned like a full one's, so the tile below reads them the same way.
QB_HD void tile_fill_any(u64 &pv, u64 &mv, TileCarry &c, const u64 *eq, int eqs, const unsigned char *text, int n, int rev,
                         int col0, int nc, int ob)
{
    u32 in_p[2] = {c.p0, c.p1}, in_m[2] = {c.m0, c.m1}, out_p[2] = {0, 0}, out_m[2] = {0, 0};
    for (int s = 0; s < nc; ++s) {
        const int col = col0 + s;
        const int code = (int)(text[rev ? n - 1 - col : col] & 7u);
        const u32 hp_in = (in_p[s >> 5] >> (31 - (s & 31))) & 1u, hm_in = (in_m[s >> 5] >> (31 - (s & 31))) & 1u;
        u32 hp_out, hm_out;
        myers_step_at(eq[code * eqs], pv, mv, hp_in, hm_in, ob, hp_out, hm_out);
        out_p[s >> 5] |= hp_out << (31 - (s & 31));
        out_m[s >> 5] |= hm_out << (31 - (s & 31));
    }
    c.p0 = out_p[0]; c.p1 = out_p[1]; c.m0 = out_m[0]; c.m1 = out_m[1];
}

// A tile in two halves around the CTA's mid-pass barrier.  tile_begin reads everything another tile of the same pass may
// overwrite (the carry-outs of the block above: that block's next tile runs in this very pass) and issues the tile's
// global loads; tile_end computes and publishes.
struct TileIn {
    int k, b, nc;
    u64 pv, mv;
    int sprev;
    TileCarry c;
    u64 w0, w1;               // first two text chunks
    u64 e[kAlpha];            // match masks of the block
    bool lastblk, below;
};

template <bool FULL>
QB_HD void tile_begin(const TileSlot &S, const TileRings &R, int k, const TilePools &P, TileIn &t)
{
    const int RBm = R.RB - 1;
    const int b = S.rho - R.base[k & RBm];
    t.k = k; t.b = b;
    t.nc = (S.ncols - 64 * k < 64) ? S.ncols - 64 * k : 64;
    if (b == R.top[k & RBm]) { t.c.p0 = t.c.p1 = 0xffffffffu; t.c.m0 = t.c.m1 = 0u; }  // top of the band: PHin = 1, MHin = 0 (bpm_banded.c:238)
    else t.c = R.carry[(b - 1) & RBm];
    {
        const u64 *tt = P.ttext + S.tt_off + 8 * (i64)k;
#ifdef __CUDA_ARCH__
        t.w0 = __ldg(tt); t.w1 = __ldg(tt + 1);
#else
        t.w0 = tt[0]; t.w1 = tt[1];
#endif
        const u64 *q = P.peq + S.peq_off + (i64)b * kPeqStride;      // zero past the table: no match
        const bool in = b < S.nbp;
#pragma unroll
        for (int cc = 0; cc < kAlpha; ++cc) {
#ifdef __CUDA_ARCH__
            t.e[cc] = in ? __ldg(q + cc) : 0ull;
#else
            t.e[cc] = in ? q[cc] : 0ull;
#endif
        }
    }
    // first tile of this block: reset / new bottom block (with bot[k-1] still undecided the block is an old one: only
    // blocks up to the last decided bottom run ahead of the decisions)
    const bool fresh = (k == 0) || (k - 1 <= S.kb && b > R.bot[(k - 1) & RBm]);
    if (fresh) {
        t.pv = ~0ull; t.mv = 0ull;
        t.sprev = (k == 0) ? 64 * (b + 1) : R.sc[((k - 1) & 1) * R.RB + (b & RBm)];
    } else {
        t.pv = R.pv[b & RBm]; t.mv = R.mv[b & RBm];
        t.sprev = R.sc[((k - 1) & 1) * R.RB + (b & RBm)];
    }
    t.lastblk = (b == S.nblk - 1) && S.mmod;                        // carry-out / score below bit 63 (level_mask)
    const int hb = (k <= S.kb) ? R.bot[k & RBm] : R.bot[S.kb & RBm];
    t.below = t.lastblk && (b < hb || k > S.kb);                    // a block below consumes this tile's carries
}

// eq: the lane's match-mask slots (stride EQS, or eqs when EQS = 0).
template <bool FULL, int EQS>
QB_HD void tile_end(const TileSlot &S, const TileRings &R, TileIn &t, u64 *eq, int eqs, const TilePools &P)
{
    const int RBm = R.RB - 1;
    const int k = t.k, b = t.b;
    u64 pv = t.pv, mv = t.mv;
    TileCarry c = t.c;
    if (FULL) {
        TileRec *rec = P.recs + S.rec_off + (i64)k * S.B + (b - (k - S.prolog));
        rec->pv0 = pv; rec->mv0 = mv; rec->cin = c;
    }
#pragma unroll
    for (int cc = 0; cc < kAlpha; ++cc) eq[cc * (EQS ? EQS : eqs)] = t.e[cc];
    int delta;
    if (t.nc == 64 && !t.below) {
        const int adj0 = t.lastblk ? popc64(pv >> S.mmod) - popc64(mv >> S.mmod) : 0;
        tile_fill64<EQS>(pv, mv, c, eq, eqs, P.ttext + S.tt_off + 8 * (i64)k, t.w0, t.w1);
        delta = popc32(c.p0) + popc32(c.p1) - popc32(c.m0) - popc32(c.m1);
        // score of the last pattern block follows row m-1, not row 63 of the block: D[m-1] = D[63] - sum of the vertical
        // deltas of the padding rows, which the block's own Pv/Mv hold
        if (t.lastblk) delta += adj0 - (popc64(pv >> S.mmod) - popc64(mv >> S.mmod));
    } else {
        tile_fill_any(pv, mv, c, eq, EQS ? EQS : eqs, P.codes + S.t_off, S.n, S.rev, 64 * k, t.nc, t.lastblk ? S.mmod - 1 : 63);
        delta = popc32(c.p0) + popc32(c.p1) - popc32(c.m0) - popc32(c.m1);
    }
    R.pv[b & RBm] = pv; R.mv[b & RBm] = mv;
    R.carry[b & RBm] = c;
    const int sc = t.sprev + delta;
    R.sc[(k & 1) * R.RB + (b & RBm)] = sc;
    P.scores[S.scores_off + b] = sc;
}

}  // namespace qb

#ifdef __CUDACC__
namespace qb {

// ---- the kernel ----------------------------------------------------------------------------------------------------
// One persistent CTA = LANES compute threads + one scheduler warp (the last).  Per pass:
//   scheduler warp: lane s owns task slot s (its bookkeeping in registers) — band decisions, the slot's next round, a new
//                   task into a free slot; then the warp packs the ready tiles of all slots onto the compute lanes
//                   (work-conserving: a round may be split over passes)                       -> __syncthreads
//   compute warps : tile_begin (inputs another tile of the pass may overwrite, global loads)  -> __syncthreads
//                   tile_end (64 word-steps per lane, results)                                -> __syncthreads
struct TileLaunch {
    const int *list;      // task ids handled by this launch (one band-height class)
    const int *count;     // how many (device memory: the list is built on the device)
    int *next;            // work counter, zeroed before the launch
    u64 *counters;        // counters[1] += word-steps of finished tasks
    int RB, nslots, lanes;
};

__host__ __device__ inline size_t tile_smem_bytes(int RB, int nslots, int lanes)
{
    size_t b = 16 + (size_t)lanes * 4;                       // control words + the plan
    b = (b + 15) & ~(size_t)15;
    b += (size_t)nslots * sizeof(TileSlot);
    b = (b + 15) & ~(size_t)15;
    b += (size_t)kAlpha * lanes * 8;                        // match masks per lane
    b += (size_t)nslots * tile_slot_arena_bytes(RB);
    return b;
}

template <bool FULL, int LANES>
__global__ void __launch_bounds__(LANES + 32, LANES >= 256 ? 2 : LANES >= 128 ? 4 : LANES >= 64 ? 6 : 8) k_band_tiles(TilePools P, TileLaunch Q)
{
    extern __shared__ __align__(16) unsigned char tile_smem[];
    constexpr int lanes = LANES;
    const int nslots = Q.nslots, RB = Q.RB;
    int *ctl = reinterpret_cast<int *>(tile_smem);
    u32 *plan = reinterpret_cast<u32 *>(tile_smem + 16);
    size_t off = (16 + (size_t)lanes * 4 + 15) & ~(size_t)15;
    TileSlot *slots = reinterpret_cast<TileSlot *>(tile_smem + off);
    off = (off + (size_t)nslots * sizeof(TileSlot) + 15) & ~(size_t)15;
    u64 *s_eq = reinterpret_cast<u64 *>(tile_smem + off);
    off += (size_t)kAlpha * lanes * 8;
    unsigned char *arenas = tile_smem + off;
    const size_t ab = tile_slot_arena_bytes(RB);

    const int tid = threadIdx.x, lane = tid & 31;
    const bool sched = tid >= lanes;
    if (sched && lane < nslots) slots[lane].task = -1;
    if (tid == 0) ctl[0] = 0;
    __syncthreads();
    const int n_tasks = sched ? *Q.count : 0;
    bool exhausted = false;
    const bool mine = sched && lane < nslots;
    const TileRings R = tile_rings(arenas + (size_t)(mine ? lane : 0) * ab, RB);
    TileSlot S;                                            // scheduler lanes: the slot, in registers
    S.task = -1; S.cnt = 0; S.koff = 0; S.rho = 0; S.kb = 0; S.state = 0;
    long long t_sched = 0, t_all0 = clock64(), n_pass = 0;

    for (unsigned pass = 0;; ++pass) {
        if (sched) {
            const long long t_s0 = clock64();
            int want = 0, kstart = 0;
            if (mine) {
                bool need_plan = false;
                if (S.task >= 0 && S.koff >= S.cnt) { ++S.rho; need_plan = true; }
                for (;;) {
                    if (S.task < 0) {
                        if (exhausted) break;
                        const int idx = atomicAdd(Q.next, 1);
                        if (idx >= n_tasks) { exhausted = true; break; }
                        const int ti = Q.list[idx];
                        tile_slot_load<FULL>(S, R, P.tasks[ti], ti, P);
                        slots[lane] = S;
                        need_plan = true;
                    }
                    if (!need_plan) break;
                    if (!S.state) tile_plan_round<FULL>(S, R, P);
                    S.koff = 0;
                    if (S.state == 1) {
                        tile_slot_finish<FULL>(S, R, P);
                        atomicAdd(&Q.counters[1], S.ws);
                        S.task = -1;
                        continue;
                    }
                    if (S.state == 2) {
                        // FULL: the traceback kernel sees the marker and hands the leaf to the exact kernels itself
                        if (FULL) P.outs[S.out_slot].pos_v = kTilePunted;
                        else P.punt_list[atomicAdd(P.punt_count, 1)] = S.task;
                        S.task = -1;
                        continue;
                    }
                    if (S.cnt == 0) { ++S.rho; continue; }              // no tile is due in this round
                    break;
                }
                if (S.task >= 0) {
                    want = S.cnt - S.koff; kstart = S.kmin + S.koff;
                    slots[lane].rho = S.rho; slots[lane].kb = S.kb;
                }
            }
            // pack: exclusive scan of `want` in an order that rotates every pass (no slot waits forever)
            const int rot = (int)(pass & 31u);
            const int x = __shfl_sync(kFull, want, (lane + rot) & 31);       // x of lane v = want of slot (v + rot) & 31
            int incl = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += y; }
            const int excl = __shfl_sync(kFull, incl - x, (lane - rot) & 31);  // back to slot order
            int take = lanes - excl;
            take = take < 0 ? 0 : (take > want ? want : take);
            if (take > 0) {
                S.koff += take;
                for (int i = 0; i < take; ++i) plan[excl + i] = ((u32)lane << 24) | (u32)(kstart + i);
            }
            const int total = __shfl_sync(kFull, incl, 31);
            for (int i = (total < lanes ? total : lanes) + lane; i < lanes; i += 32) plan[i] = 0xffffffffu;
            const bool idle = !mine || (S.task < 0 && exhausted);
            if (__all_sync(kFull, idle) && lane == 0) ctl[0] = 1;
            t_sched += clock64() - t_s0; ++n_pass;
        }
        __syncthreads();
        if (ctl[0]) break;
        TileIn t;
        int s = -1;
        if (!sched) {
            const u32 e = plan[tid];
            if (e != 0xffffffffu) {
                s = (int)(e >> 24);
                tile_begin<FULL>(slots[s], tile_rings(arenas + (size_t)s * ab, RB), (int)(e & 0xffffffu), P, t);
            }
        }
        __syncthreads();
        if (s >= 0) tile_end<FULL, LANES>(slots[s], tile_rings(arenas + (size_t)s * ab, RB), t, s_eq + tid, lanes, P);
        __syncthreads();
    }
    if (tid == lanes) {                                   // scheduler lane 0: cycles spent scheduling / in all / passes
        atomicAdd(&Q.counters[8], (u64)t_sched);
        atomicAdd(&Q.counters[9], (u64)(clock64() - t_all0));
        atomicAdd(&Q.counters[10], (u64)n_pass);
    }
}

// Task ids of `list` sorted into band-height classes (ring size 8 << c): out[c * cap + i], counts[c] (skipped when out is
// null: the thread fill needs no lists); every task also gets its place in the tile-text pool (tt_words: running total in
// u64 words).  One atomic per warp for the pool, one per warp and class for the lists.
template <bool FULL>
__global__ void __launch_bounds__(256) k_tile_classes(BandTask *__restrict__ tasks, const int *__restrict__ list, int begin, int n,
                                                      int *__restrict__ out, int cap, int *__restrict__ counts,
                                                      unsigned long long *__restrict__ tt_words)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    int ti = -1, c = -1;
    unsigned words = 0;
    if (i < n) {
        ti = list ? list[begin + i] : begin + i;
        const BandTask &t = tasks[ti];
        const BandGeom g = band_geometry(t.m, t.n, t.cutoff);
        const int rb = tile_ring_for(FULL ? g.Bc : g.Bs);
        if (rb <= kTileMaxRing) {                        // taller bands: the shared-memory sweep kernel (qb_banded.cuh)
            c = 0;
            while ((8 << c) < rb) ++c;
            const int ncols = FULL ? t.n : t.finish;
            words = (unsigned)((ncols + 63) / 64 * 8 + 1);   // +1: the odd-character flag word
        }
    }
    // text pool: warp-wide exclusive scan of the word counts, one atomic for the warp
    unsigned incl = words;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += y; }
    const unsigned total = __shfl_sync(kFull, incl, 31);
    unsigned long long base = 0;
    if (lane == 0 && total) base = atomicAdd(tt_words, (unsigned long long)total);
    base = __shfl_sync(kFull, base, 0);
    if (c >= 0) tasks[ti].tt_off = (i64)(base + incl - words);
    if (!out) return;
    // class lists: lanes of the same class share one atomic
#pragma unroll 1
    for (int cc = 0; cc < 8; ++cc) {
        const unsigned m = __ballot_sync(kFull, c == cc);
        if (!m) continue;
        int pos = 0;
        if (lane == __ffs(m) - 1) pos = atomicAdd(&counts[cc], __popc(m));
        pos = __shfl_sync(kFull, pos, __ffs(m) - 1);
        if (c == cc) out[(size_t)cc * cap + pos + __popc(m & ((1u << lane) - 1u))] = ti;
    }
}

// Tile-text pool: the text codes of a task, masked to the 3 code bits, in COLUMN order (a reversed pass reads its text
// backwards), 8 columns per u64 and 8-byte aligned, so a tile's 64 codes are eight aligned loads with no realignment
// in the fill's inner loop.  G lanes per task (32 for long texts, 4 for short reads); columns past the pass
// (ncols..64*ceil) are code 4.
template <bool FULL, int G>
__global__ void __launch_bounds__(256) k_tile_text(const BandTask *__restrict__ tasks, const int *__restrict__ list, int begin, int n,
                                                   const unsigned char *__restrict__ codes, u64 *__restrict__ ttext)
{
    const i64 gt = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const int wi = (int)(gt / G), gl = (int)(gt % G);
    const bool live = wi < n;
    const BandTask *t = live ? &tasks[list ? list[begin + wi] : begin + wi] : nullptr;
    bool ok = live;
    if (live) {
        const BandGeom g = band_geometry(t->m, t->n, t->cutoff);
        ok = tile_ring_for(FULL ? g.Bc : g.Bs) <= kTileMaxRing;
    }
    u32 odd = 0;
    int nw = 0;
    u64 *dst = nullptr;
    if (ok) {
        const int ncols = FULL ? t->n : t->finish;
        nw = (ncols + 63) / 64 * 8;
        const unsigned char *src = codes + t->t_off;
        dst = ttext + t->tt_off;
        for (int w = gl; w < nw; w += G) {
            u64 v = 0;
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const int col = 8 * w + b;
                const unsigned raw = col < ncols ? (unsigned)src[t->rev ? t->n - 1 - col : col] : 4u;
                odd |= raw & kCodeOdd;
                v |= (u64)(raw & 7u) << (8 * b);
            }
            dst[w] = v;
        }
    }
    // last word of the task's slot: does the text hold a character outside "ACGTN"? (the tile traceback then compares
    // raw bytes on diagonal steps, reference bpm_banded.c:1012)
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) odd |= __shfl_xor_sync(kFull, odd, o);
    if (ok && gl == 0) dst[nw] = odd ? 1ull : 0ull;
}

}  // namespace qb
#endif
