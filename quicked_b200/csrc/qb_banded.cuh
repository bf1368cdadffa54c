// qb_banded.cuh — BandEd: banded bit-parallel (Myers) edit distance with the reference's sliding, self-cutting band.
//
// k_banded_warp<R, FULL>: ONE TASK PER WARP, one 64-row block per lane (R blocks per lane for bands up to 32*R
// blocks).  All lanes work on the SAME text column, so the match-mask fetch is a conflict-free broadcast-indexed
// LDS and the band bookkeeping is warp-uniform.  The vertical dependency between the blocks of a column is
//   (a) the 64-bit adder carry of Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq crossing block boundaries, and
//   (b) bit 63 of Ph / Mh shifting into bit 0 of the next block.
// The reference resolves it serially, block after block (bpm_banded.c:238-261: PHin/MHin = previous PHout/MHout,
// with MHin folded into Eq).  Here (a) is resolved for all 32 lanes at once by a carry-lookahead over two ballots
// (generate = local add overflowed, propagate = local sum is all ones): the adder carry into a block equals the
// reference's MHin of that block (carry-out of bit 63 of (Eq&Pv)+Pv is Pv63 & Xh63 = Mh63), and (b) is one more
// ballot of Ph bit 63.  The update is therefore bit-identical to BPM_ADVANCE_BLOCK (bpm_commons.h:49-68) on every
// block of the pattern.
//
// Between two band shifts (64 columns) a lane owns fixed blocks, so Pv/Mv/score stay in registers; shared memory
// is only the re-mapping medium at the shift (slots indexed by absolute block, so the reference's "move every word
// down by one" (bpm_banded.c:279-287 / :902-910) costs nothing) and holds the lane's 5 match masks of the block.
//
// FULL=false: score-only pass up to column `finish` (reference bpm_banded.c:791-964 == AVX2 :349-788); exports the
//             final column (Pv, Mv per band word), per-block scores and lower/higher block for Hirschberg.
// FULL=true : every column is stored for the traceback (reference bpm_banded.c:199-316), as 16-byte (Pv,Mv)
//             entries [column][band word] — one coalesced 512-byte store per warp-column.
#pragma once
#include "qb_common.cuh"

namespace qb {

template <int R>
struct BandedSmem {
    static constexpr int kCap = 32 * R + 2;                       // slots (absolute block mod kCap)
    static constexpr int kBytesPerWarp = kCap * 16 + R * kAlpha * 32 * 8 + 64;
};

template <int R, bool FULL>
__global__ void __launch_bounds__(128)
k_banded_warp(const BandTask *__restrict__ tasks, const int *__restrict__ list, int begin, int n_tasks, i64 mat_sub,
              const unsigned char *__restrict__ codes, const u64 *__restrict__ peq, ulonglong2 *__restrict__ matrix,
              int *__restrict__ scores_pool, u64 *__restrict__ state_pool, int2 *__restrict__ range_pool,
              BandOut *__restrict__ outs, u64 *__restrict__ counters, int min_B)
{
    constexpr int kCap = BandedSmem<R>::kCap;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int task_id = blockIdx.x * (blockDim.x >> 5) + wib;
    if (task_id >= n_tasks) return;
    unsigned char *base = smem_raw + (size_t)wib * BandedSmem<R>::kBytesPerWarp;
    u64 *s_pv = reinterpret_cast<u64 *>(base);
    u64 *s_mv = s_pv + kCap;
    u64 *s_eq = s_mv + kCap;                                        // [R][5][32]
    unsigned char *s_txt = reinterpret_cast<unsigned char *>(s_eq + R * kAlpha * 32);

    BandTask tk = tasks[list ? list[begin + task_id] : begin + task_id];
    tk.mat_off -= mat_sub;
    const BandGeom g = band_geometry(tk.m, tk.n, tk.cutoff);
    const int B = (int)(FULL ? g.Bc : g.Bs);
    {   // tasks of other band heights are handled by the launch of their own R (one list, one launch per R)
        const int need = B <= 32 ? 1 : B <= 64 ? 2 : B <= 128 ? 4 : B <= 256 ? 8 : B <= 512 ? 16 : B <= 1024 ? 32 : 64;
        if (need != R || B < min_B) return;                          // min_B: narrower bands belong to the tile kernels
    }
    const int nblk = (tk.m + 63) >> 6, mmod = tk.m & 63;
    const int clamp = FULL ? nblk - 1 : nblk;                       // bpm_banded.c:295 vs :917
    const int prolog = (int)g.prolog;
    const i64 fin = g.fin, kcut = g.k;
    const u64 *pq = peq + tk.peq_off;
    int *scores = scores_pool + tk.scores_off;
    const unsigned char *tcodes = codes + tk.t_off;
    const int ncols = FULL ? tk.n : tk.finish;

    int first = prolog, last = B - 1, pos_v = -prolog, pos_h = 0;  // bpm_banded.c:222-225
    // FULL: live range of every 64-column block, so the traceback can treat never-written cells as 0 (what a fresh
    // arena holds in the reference) instead of reading stale pool memory
    int2 *ranges = FULL ? range_pool + tk.range_off : nullptr;
    if (FULL && lane == 0) ranges[0] = make_int2(first, last);
    // ---- reset (bpm_banded.c:180-197): Pv = ~0, Mv = 0, scores[i] = 64(i+1) ----
    for (int j = lane; j < B; j += 32) {
        scores[j] = 64 * (j + 1);
        const int blk = j + pos_v;
        if (blk >= 0) { s_pv[blk % kCap] = ~0ull; s_mv[blk % kCap] = 0ull; }
        if (FULL) matrix[tk.mat_off + j] = make_ulonglong2(~0ull, 0ull);   // warp layout: mat_cs == B, mat_ws == 1
    }
    __syncwarp();

    u64 ws = 0;
    u64 pv[R], mv[R];
    int sc[R], ob[R];
    for (int col0 = 0; col0 < ncols; col0 += 64) {
        const int nc = min(64, ncols - col0);
        // ---- load the 64 columns' codes and this lane's blocks ----
        for (int c = lane; c < 64; c += 32) {
            const int col = col0 + c;
            unsigned char code = 4;
            if (col < tk.n) code = (tk.rev ? tcodes[tk.n - 1 - col] : tcodes[col]) & 7;
            s_txt[c] = code;
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int j = first + 32 * r + lane;
            const bool act = j <= last;
            const int blk = j + pos_v;
            pv[r] = act ? s_pv[blk % kCap] : 0ull;
            mv[r] = act ? s_mv[blk % kCap] : 0ull;
            sc[r] = act ? scores[blk] : 0;
            ob[r] = (blk == nblk - 1 && mmod) ? mmod - 1 : 63;      // level_mask, bpm_banded.c:88-102
#pragma unroll
            for (int c = 0; c < kAlpha; ++c)
                s_eq[(r * kAlpha + c) * 32 + lane] = (act && blk < tk.nbp) ? pq[(i64)blk * kPeqStride + c] : 0ull;   // past the table: no match
        }
        __syncwarp();
        const int live = last - first + 1;
        // does any live lane hold the last pattern block with a carry-out below bit 63 (level_mask)?  warp-uniform
        bool any_masked = false;
#pragma unroll
        for (int r = 0; r < R; ++r) any_masked |= (first + 32 * r + lane <= last) && ob[r] != 63;
        any_masked = __any_sync(kFull, any_masked);
        // running store pointer of this lane's word in the current column (entries; +B per column)
        ulonglong2 *col_ptr = FULL ? matrix + tk.mat_off + (i64)(col0 + 1) * B + first + lane : nullptr;
        // ---- the column loop ----
        for (int c = 0; c < nc; ++c) {
            const int code = s_txt[c];
            u32 cin = 0, hp_carry = 1;                               // top of the band: PHin = 1, MHin = 0 (:238)
            const bool store_col = FULL && !(c == 63);               // the 64th column is stored after the shift
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int j = first + 32 * r + lane;
                const bool act = j <= last;
                const u64 eq = s_eq[(r * kAlpha + code) * 32 + lane];
                const u64 a = eq & pv[r];
                const u64 s = a + pv[r];
                const u32 G = __ballot_sync(kFull, act && (s < a));
                const u32 P = __ballot_sync(kFull, act && (s == ~0ull));
                const u32 X = G | P;
                const u64 sum = (u64)X + (u64)G + (u64)cin;
                const u32 carries = (u32)sum ^ X ^ G;                // bit l = carry into lane l
                const u32 my_c = (carries >> lane) & 1u;             // == the reference's MHin of this block
                const u64 xh = ((s + my_c) ^ pv[r]) | eq;
                u64 ph = mv[r] | ~(xh | pv[r]);
                u64 mh = pv[r] & xh;
                const u32 HP = __ballot_sync(kFull, (ph >> 63) != 0);
                const u32 hp_in = lane ? ((HP >> (lane - 1)) & 1u) : hp_carry;
                sc[r] += (int)(ph >> 63) - (int)(mh >> 63);                            // :260, carry-out at bit 63 ...
                if (any_masked && ob[r] != 63)                                         // ... except the last pattern block (level_mask)
                    sc[r] += ((int)((ph >> ob[r]) & 1ull) - (int)((mh >> ob[r]) & 1ull)) - ((int)(ph >> 63) - (int)(mh >> 63));
                ph = (ph << 1) | (u64)hp_in;
                mh = (mh << 1) | (u64)my_c;
                const u64 xv = eq | mv[r];
                pv[r] = mh | ~(xv | ph);
                mv[r] = ph & xv;
                if (store_col && act) col_ptr[32 * r] = make_ulonglong2(pv[r], mv[r]);
                cin = (u32)(sum >> 32);
                hp_carry = HP >> 31;
            }
            if (FULL) col_ptr += B;
        }
        ws += (u64)live * nc;
        // ---- write the lane's blocks back ----
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int j = first + 32 * r + lane;
            if (j <= last) {
                const int blk = j + pos_v;
                s_pv[blk % kCap] = pv[r]; s_mv[blk % kCap] = mv[r];
                scores[blk] = sc[r];
            }
        }
        __syncwarp();
        if (nc < 64) break;                                          // tail columns: no shift (:925-951)
        // ---- end of a 64-column block (bpm_banded.c:264-301 / :889-922), warp-uniform ----
        {
            const bool cut_lo = (first + 2 < last) && (fin > 64 * (i64)(first + 1)) &&
                                ((i64)scores[first + pos_v + 1] + (fin - 64 * (i64)(first + 1)) > kcut);
            const int first_old = first;
            if (cut_lo && pos_h >= prolog) ++first;
            else if (!cut_lo && pos_h < prolog) --first;
            const int nb = last + pos_v + 1;                          // the block entering at the bottom
            __syncwarp();
            if (lane == 0) {
                s_pv[nb % kCap] = ~0ull; s_mv[nb % kCap] = 0ull;
                scores[nb] = scores[nb - 1] + 64;
            }
            __syncwarp();
            if (FULL) {    // column col0+64 is stored in the coordinates of the next block of columns (:279-287)
                // when the top block is cut, the reference's in-place shift leaves the pre-shift word at index
                // first_old; a too-narrow band can make the traceback read it, so keep it identical
                if (first > first_old && lane == 0) {
                    const int blk = first_old + pos_v;
                    matrix[tk.mat_off + (i64)(col0 + 64) * B + first_old] = make_ulonglong2(s_pv[blk % kCap], s_mv[blk % kCap]);
                }
                for (int j = first + lane; j <= last; j += 32) {
                    const int blk = j + pos_v + 1;
                    matrix[tk.mat_off + (i64)(col0 + 64) * B + j] = make_ulonglong2(s_pv[blk % kCap], s_mv[blk % kCap]);
                }
            }
            const bool cut_hi = (first + 2 < last) && (64 * (i64)(last - 1) > fin) &&
                                ((i64)scores[last + pos_v - 1] + (64 * (i64)(last - 1) - fin) > kcut);
            if (cut_hi || (pos_v + last >= clamp)) --last;
            ++pos_v; ++pos_h;
            if (FULL && lane == 0) ranges[pos_h] = make_int2(first, last);
        }
        __syncwarp();
    }
    // ---- results ----
    if (!FULL) {
        u64 *st = state_pool + tk.state_off;                         // Pv[B] then Mv[B], band-relative
        for (int j = lane; j < B; j += 32) {
            const int blk = j + pos_v;
            const bool act = (j >= first && j <= last && blk >= 0);
            st[j] = act ? s_pv[blk % kCap] : 0ull;
            st[B + j] = act ? s_mv[blk % kCap] : 0ull;
        }
    }
    if (lane == 0) {
        BandOut o;
        const int sfin = scores[nblk - 1];                            // bpm_banded.c:952-961
        o.score = mmod ? sfin - (64 - mmod) : sfin;
        o.first = first; o.last = last; o.pos_v = pos_v;
        outs[tk.slot] = o;
        atomicAdd(&counters[1], ws);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// k_banded_warp_dyn<FULL>: the same algorithm for bands taller than 1024 blocks (e.g. the stage-3 pass of a 500 kbp
// ONT pair at 15 % bandwidth: 1194 blocks).  One warp per CTA; Pv / Mv / scores of the whole band stay in shared
// memory (20 bytes per block, up to ~11 000 blocks) and the warp sweeps the band in rounds of 32 blocks per column.
// Correctness path for rare, very long pairs: throughput comes from the R-templated kernel above.
template <bool FULL>
__global__ void __launch_bounds__(32)
k_banded_warp_dyn(const BandTask *__restrict__ tasks, const int *__restrict__ list, int begin, int n_tasks, i64 mat_sub,
                  const unsigned char *__restrict__ codes, const u64 *__restrict__ peq, ulonglong2 *__restrict__ matrix,
                  int *__restrict__ scores_pool, u64 *__restrict__ state_pool, int2 *__restrict__ range_pool,
                  BandOut *__restrict__ outs, u64 *__restrict__ counters, int cap, int min_B)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    const int task_id = blockIdx.x;
    if (task_id >= n_tasks) return;
    BandTask tk = tasks[list ? list[begin + task_id] : begin + task_id];
    tk.mat_off -= mat_sub;
    const BandGeom g = band_geometry(tk.m, tk.n, tk.cutoff);
    const int B = (int)(FULL ? g.Bc : g.Bs);
    if (B <= 1024 || B < min_B) return;                              // handled by the R-templated launches / the tile kernels
    u64 *s_pv = reinterpret_cast<u64 *>(smem_raw);
    u64 *s_mv = s_pv + cap;
    const int nblk = (tk.m + 63) >> 6, mmod = tk.m & 63;
    const int clamp = FULL ? nblk - 1 : nblk;
    const int prolog = (int)g.prolog;
    const i64 fin = g.fin, kcut = g.k;
    const u64 *pq = peq + tk.peq_off;
    int *scores = scores_pool + tk.scores_off;
    const unsigned char *tcodes = codes + tk.t_off;
    const int ncols = FULL ? tk.n : tk.finish;
    int first = prolog, last = B - 1, pos_v = -prolog, pos_h = 0;
    int2 *ranges = FULL ? range_pool + tk.range_off : nullptr;
    if (FULL && lane == 0) ranges[0] = make_int2(first, last);
    for (int j = lane; j < B; j += 32) {
        scores[j] = 64 * (j + 1);
        const int blk = j + pos_v;
        if (blk >= 0) { s_pv[blk % cap] = ~0ull; s_mv[blk % cap] = 0ull; }
        if (FULL) matrix[tk.mat_off + j] = make_ulonglong2(~0ull, 0ull);
    }
    __syncwarp();
    u64 ws = 0;
    for (int col = 0; col < ncols; ++col) {
        unsigned char code = (tk.rev ? tcodes[tk.n - 1 - col] : tcodes[col]) & 7;
        u32 cin = 0, hp_carry = 1;
        const bool store_col = FULL && ((col & 63) != 63);
        for (int j0 = first; j0 <= last; j0 += 32) {
            const int j = j0 + lane;
            const bool act = j <= last;
            const int blk = j + pos_v, slot = act ? blk % cap : 0;
            const u64 eq = (act && blk < tk.nbp) ? pq[(i64)blk * kPeqStride + code] : 0ull;
            u64 pv = act ? s_pv[slot] : 0ull, mv = act ? s_mv[slot] : 0ull;
            const int ob = (blk == nblk - 1 && mmod) ? mmod - 1 : 63;
            const u64 a = eq & pv, s = a + pv;
            const u32 G = __ballot_sync(kFull, act && (s < a));
            const u32 P = __ballot_sync(kFull, act && (s == ~0ull));
            const u32 X = G | P;
            const u64 sum = (u64)X + (u64)G + (u64)cin;
            const u32 carries = (u32)sum ^ X ^ G;
            const u32 my_c = (carries >> lane) & 1u;
            const u64 xh = ((s + my_c) ^ pv) | eq;
            u64 ph = mv | ~(xh | pv);
            u64 mh = pv & xh;
            const u32 HP = __ballot_sync(kFull, (ph >> 63) != 0);
            const u32 hp_in = lane ? ((HP >> (lane - 1)) & 1u) : hp_carry;
            const int dsc = (int)((ph >> ob) & 1ull) - (int)((mh >> ob) & 1ull);
            ph = (ph << 1) | (u64)hp_in;
            mh = (mh << 1) | (u64)my_c;
            const u64 xv = eq | mv;
            pv = mh | ~(xv | ph);
            mv = ph & xv;
            if (act) {
                s_pv[slot] = pv; s_mv[slot] = mv; scores[blk] += dsc;
                if (store_col) matrix[tk.mat_off + (i64)(col + 1) * B + j] = make_ulonglong2(pv, mv);
            }
            cin = (u32)(sum >> 32);
            hp_carry = HP >> 31;
        }
        ws += (u64)max(last - first + 1, 0);
        __syncwarp();
        if ((col & 63) != 63) continue;
        {   // end of a 64-column block
            const int col0 = col - 63;
            const bool cut_lo = (first + 2 < last) && (fin > 64 * (i64)(first + 1)) &&
                                ((i64)scores[first + pos_v + 1] + (fin - 64 * (i64)(first + 1)) > kcut);
            const int first_old = first;
            if (cut_lo && pos_h >= prolog) ++first;
            else if (!cut_lo && pos_h < prolog) --first;
            const int nb = last + pos_v + 1;
            __syncwarp();
            if (lane == 0) { s_pv[nb % cap] = ~0ull; s_mv[nb % cap] = 0ull; scores[nb] = scores[nb - 1] + 64; }
            __syncwarp();
            if (FULL) {
                if (first > first_old && lane == 0) {
                    const int blk = first_old + pos_v;
                    matrix[tk.mat_off + (i64)(col0 + 64) * B + first_old] = make_ulonglong2(s_pv[blk % cap], s_mv[blk % cap]);
                }
                for (int j = first + lane; j <= last; j += 32) {
                    const int blk = j + pos_v + 1;
                    matrix[tk.mat_off + (i64)(col0 + 64) * B + j] = make_ulonglong2(s_pv[blk % cap], s_mv[blk % cap]);
                }
            }
            const bool cut_hi = (first + 2 < last) && (64 * (i64)(last - 1) > fin) &&
                                ((i64)scores[last + pos_v - 1] + (64 * (i64)(last - 1) - fin) > kcut);
            if (cut_hi || (pos_v + last >= clamp)) --last;
            ++pos_v; ++pos_h;
            if (FULL && lane == 0) ranges[pos_h] = make_int2(first, last);
        }
        __syncwarp();
    }
    if (!FULL) {
        u64 *st = state_pool + tk.state_off;
        for (int j = lane; j < B; j += 32) {
            const int blk = j + pos_v;
            const bool act = (j >= first && j <= last && blk >= 0);
            st[j] = act ? s_pv[blk % cap] : 0ull;
            st[B + j] = act ? s_mv[blk % cap] : 0ull;
        }
    }
    __syncwarp();
    if (lane == 0) {
        BandOut o;
        const int sfin = scores[nblk - 1];
        o.score = mmod ? sfin - (64 - mmod) : sfin;
        o.first = first; o.last = last; o.pos_v = pos_v;
        outs[tk.slot] = o;
        atomicAdd(&counters[1], ws);
    }
}

// Full-matrix BandEd of ONE leaf by one thread (see k_banded_thread).  mat: first entry of this leaf, column stride
// cs / word stride wsd (entries); ranges: live range per 64-column block, stride rstride; s_eq: this thread's
// BMAX*5 shared-memory slots, stride T.
constexpr int kThreadLeafWordStride = 32;   // entries between consecutive band words of a thread-kernel leaf (= lanes per group)
constexpr int kThreadFillThreads = 128;     // threads per CTA of every kernel that calls banded_thread_fill
// REC: instead of the reference's matrix (one (Pv,Mv) entry per word-step, bpm_banded.c:139-140) the fill writes the
// TILE RECORDS of qb_tiles.cuh — per 64 columns and live block the block's state at the first column and the 64
// carry-in pairs it received, 32 bytes — which is all the tile traceback (qb_tiletrace.cuh) needs: `mat` then points at
// the leaf's records (TileRec[k * B + band word], as ulonglong2 pairs).  16 B per word-step become 0.5 B.
template <int BMAX, bool REC = false>
__device__ __forceinline__ void banded_thread_fill(int m, int n, i64 cutoff, int rev, const u64 *__restrict__ pq, int nbp,
                                                   const unsigned char *__restrict__ tcodes, ulonglong2 *mat, i64 cs, i64 wsd_,
                                                   int2 *ranges, i64 rstride, u64 *s_eq, int T_, u64 &ws)
{
        // Both strides are fixed by the callers (32 interleaved leaves per group, 128 threads per CTA); as compile-time
        // constants they fold into the LDS / STG immediates instead of costing ~7 integer instructions per word-step.
        constexpr i64 wsd = kThreadLeafWordStride;
        constexpr int T = kThreadFillThreads;
        (void)wsd_; (void)T_;
        const BandGeom g = band_geometry(m, n, cutoff);
        const int B = (int)g.Bc, prolog = (int)g.prolog;
        const int nblk = (m + 63) >> 6, mmod = m & 63, clamp = nblk - 1;
        const i64 fin = g.fin, kcut = g.k;
        int first = prolog, last = B - 1, pos_v = -prolog, pos_h = 0;
        u64 pv[BMAX], mv[BMAX];
        int sc[BMAX];
        u64 cwp[REC ? BMAX : 1], cwm[REC ? BMAX : 1];     // REC: the carry-ins of each block over the current 64 columns, first column in the top bit
#pragma unroll
        for (int j = 0; j < BMAX; ++j) {
            pv[j] = ~0ull; mv[j] = 0ull; sc[j] = 64 * (j + pos_v + 1);            // scores[blk] = 64(blk+1)
            if (!REC && j < B) mat[j * wsd] = make_ulonglong2(~0ull, 0ull);       // column 0
        }
        ranges[0] = make_int2(first, last);
        // forward text: aligned 16-byte view of the codes (the flat buffer keeps >= 48 readable bytes past its end)
        const int csh = (int)((unsigned long long)tcodes & 15ull);
        const uint4 *cvec = reinterpret_cast<const uint4 *>(tcodes - csh);
        uint4 ccur = make_uint4(0, 0, 0, 0), cnxt = ccur;
        if (!rev) { ccur = __ldg(cvec); cnxt = __ldg(cvec + 1); }
        int st_lo = 0, st_hi = -1;                       // band slots whose match masks are staged in s_eq
        for (int col0 = 0; col0 < n; col0 += 64) {
            const int nc = min(64, n - col0);
            // match masks of the live blocks: the band moved down one block, so the staged masks move up one slot
            // and only blocks that were not staged before are fetched (48 B = three 16-byte loads each)
            if (col0 > 0) {
#pragma unroll
                for (int j = 0; j < BMAX - 1; ++j)
#pragma unroll
                    for (int c = 0; c < kAlpha; ++c) s_eq[(j * kAlpha + c) * T] = s_eq[((j + 1) * kAlpha + c) * T];
                --st_lo; --st_hi;
            }
#pragma unroll
            for (int j = 0; j < BMAX; ++j) {
                const int blk = j + pos_v;
                if (j >= first && j <= last && (j < st_lo || j > st_hi)) {
                    ulonglong2 q0 = make_ulonglong2(0, 0), q1 = q0, q2 = q0;
                    if (blk < nbp) {
                        const ulonglong2 *q = reinterpret_cast<const ulonglong2 *>(pq + (i64)blk * kPeqStride);
                        q0 = __ldg(q); q1 = __ldg(q + 1); q2 = __ldg(q + 2);
                    }
                    s_eq[(j * kAlpha + 0) * T] = q0.x; s_eq[(j * kAlpha + 1) * T] = q0.y; s_eq[(j * kAlpha + 2) * T] = q1.x;
                    s_eq[(j * kAlpha + 3) * T] = q1.y; s_eq[(j * kAlpha + 4) * T] = q2.x;
                }
            }
            st_lo = (st_lo > st_hi) ? first : min(st_lo, first);
            st_hi = max(st_hi, last);
            // band index of the last pattern block when its carry-out sits below bit 63 (level_mask, bpm_banded.c:88-102)
            const int jl = mmod ? (nblk - 1 - pos_v) : -1;
            if (REC) {       // the state every live block starts this column block with
                ulonglong2 *rec = mat + 2 * ((i64)(col0 >> 6) * B);
#pragma unroll
                for (int j = 0; j < BMAX; ++j) {
                    if (j >= first && j <= last) rec[2 * j] = make_ulonglong2(pv[j], mv[j]);
                    cwp[j] = 0ull; cwm[j] = 0ull;
                }
            }
            // eight columns per loop body (the body has to stay well inside the 32 KB instruction cache); their codes
            // come from one aligned 16-byte load per 16 columns, realigned in registers, two chunks ahead in flight
            uint4 cr = make_uint4(0, 0, 0, 0);
#pragma unroll 1
            for (int c0 = 0; c0 < nc; c0 += 8) {
                u32 w0, w1;
                if (rev) {
                    w0 = w1 = 0;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int col = col0 + c0 + k;
                        const u32 c = (c0 + k < nc) ? (u32)tcodes[n - 1 - col] : 4u;
                        if (k < 4) w0 |= c << (8 * k); else w1 |= c << (8 * (k - 4));
                    }
                } else {
                    if (!(c0 & 8)) {
                        const uint4 c2 = __ldg(cvec + (((col0 + c0) >> 4) + 2));
                        cr = realign16(ccur, cnxt, csh);
                        ccur = cnxt; cnxt = c2;
                    }
                    w0 = (c0 & 8) ? cr.z : cr.x; w1 = (c0 & 8) ? cr.w : cr.y;
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (c0 + k < nc) {
                        const int code = (int)(((k < 4 ? w0 : w1) >> (8 * (k & 3))) & 7u);
                        ulonglong2 *dst = REC ? nullptr : mat + (i64)(col0 + c0 + k + 1) * cs;
                        u32 hp = 1, hm = 0;
#pragma unroll
                        for (int j = 0; j < BMAX; ++j) {
                            if (j >= first && j <= last) {
                                u32 hpo, hmo;
                                u64 phr, mhr;
                                if (REC) { cwp[j] = (cwp[j] << 1) | (u64)hp; cwm[j] = (cwm[j] << 1) | (u64)hm; }
                                myers_step_hv(s_eq[(j * kAlpha + code) * T], pv[j], mv[j], hp, hm, hpo, hmo, phr, mhr);
                                int d = (int)hpo - (int)hmo;
                                if (j == jl) d = (int)((phr >> (mmod - 1)) & 1ull) - (int)((mhr >> (mmod - 1)) & 1ull);
                                sc[j] += d;
                                hp = hpo; hm = hmo;
                                if (!REC) dst[j * wsd] = make_ulonglong2(pv[j], mv[j]);
                            }
                        }
                    }
                }
            }
            ws += (u64)max(last - first + 1, 0) * nc;
            if (REC) {       // the carry-ins of the column block, left-aligned when it is a short one (TileCarry: steps 0..31 | 32..63)
                ulonglong2 *rec = mat + 2 * ((i64)(col0 >> 6) * B);
#pragma unroll
                for (int j = 0; j < BMAX; ++j)
                    if (j >= first && j <= last) {
                        const u64 p = nc < 64 ? cwp[j] << (64 - nc) : cwp[j], q = nc < 64 ? cwm[j] << (64 - nc) : cwm[j];
                        rec[2 * j + 1] = make_ulonglong2((p >> 32) | (p << 32), (q >> 32) | (q << 32));
                    }
            }
            if (nc < 64) break;
            // ---- end of a 64-column block (bpm_banded.c:264-301) ----
            int s_f1 = 0, s_l1 = 0, s_l = 0;            // scores[first+1], scores[last-1], scores[last] (band-relative)
#pragma unroll
            for (int j = 0; j < BMAX; ++j) {
                if (j == first + 1) s_f1 = sc[j];
                if (j == last - 1) s_l1 = sc[j];
                if (j == last) s_l = sc[j];
            }
            const bool cut_lo = (first + 2 < last) && (fin > 64 * (i64)(first + 1)) && ((i64)s_f1 + (fin - 64 * (i64)(first + 1)) > kcut);
            if (cut_lo && pos_h >= prolog) ++first;
            else if (!cut_lo && pos_h < prolog) --first;
            // shift every word up by one band index; the block entering at the bottom starts at Pv = ~0, Mv = 0
#pragma unroll
            for (int j = 0; j < BMAX - 1; ++j) { pv[j] = pv[j + 1]; mv[j] = mv[j + 1]; sc[j] = sc[j + 1]; }
#pragma unroll
            for (int j = 0; j < BMAX; ++j)
                if (j == last) { pv[j] = ~0ull; mv[j] = 0ull; sc[j] = s_l + 64; }
            // column col0+64 re-stored in the next block's coordinates (the pre-shift store above stays underneath,
            // exactly like the reference's in-place shift, bpm_banded.c:279-287)
            if (!REC) {
                ulonglong2 *dst = mat + (i64)(col0 + 64) * cs;
#pragma unroll
                for (int j = 0; j < BMAX; ++j)
                    if (j >= first && j <= last) dst[j * wsd] = make_ulonglong2(pv[j], mv[j]);
            }
            // scores[last-1] in the OLD coordinates is sc[last-2] after the shift; s_l1 was read before it
            const bool cut_hi = (first + 2 < last) && (64 * (i64)(last - 1) > fin) && ((i64)s_l1 + (64 * (i64)(last - 1) - fin) > kcut);
            if (cut_hi || (pos_v + last >= clamp)) --last;
            ++pos_v; ++pos_h;
            ranges[(i64)pos_h * rstride] = make_int2(first, last);
        }
}

// ------------------------------------------------------------------------------------------------------------------
// k_banded_thread<BMAX>: full-matrix BandEd for NARROW bands (B_cigar <= BMAX blocks), ONE LEAF PER THREAD.
//
// At 100 bp - 1 kbp and <= 10 % error the band is 3 blocks tall: a warp-per-pair mapping would idle 29 lanes, so
// here each thread carries its whole band (Pv, Mv, running scores) in registers and walks the blocks of a column
// serially exactly like the reference's inner loop (bpm_banded.c:238-261) — no cross-lane traffic at all.
// The 32 leaves of a warp are interleaved in the matrix: entry (column c, band word w, lane l) lives at
// group_base + (c*Bg + w)*32 + l, so every (Pv,Mv) store of the warp is one coalesced 512-byte line group.
// The lane's 5 match masks per live block sit in shared memory ([slot][thread], conflict-free); at every 64-column
// band shift they move up one slot and only the block that enters the band is fetched (3 x 16 B from the
// [block][6] table); the per-column fetch is an LDS indexed by the column's code.
template <int BMAX, bool REC = false>
__global__ void __launch_bounds__(128, 5)
k_banded_thread(const BandTask *__restrict__ tasks, const int *__restrict__ list, int begin, int n_tasks, i64 mat_sub,
                const unsigned char *__restrict__ codes, const u64 *__restrict__ peq, ulonglong2 *__restrict__ matrix,
                int2 *__restrict__ range_pool, u64 *__restrict__ counters)
{
    extern __shared__ u64 s_eq_all[];                       // [BMAX*5][blockDim.x]
    const int T = blockDim.x, tid = threadIdx.x;
    u64 *s_eq = s_eq_all + tid;
    const int id = blockIdx.x * T + tid;
    u64 ws = 0;
    if (id < n_tasks) {
        BandTask tk = tasks[list ? list[begin + id] : begin + id];
        tk.mat_off -= mat_sub;
        banded_thread_fill<BMAX, REC>(tk.m, tk.n, tk.cutoff, tk.rev, peq + tk.peq_off, tk.nbp, codes + tk.t_off, matrix + tk.mat_off,
                                 tk.mat_cs, tk.mat_ws, range_pool + tk.range_off, 1, s_eq, T, ws);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ws += __shfl_down_sync(kFull, ws, o);
    if ((tid & 31) == 0 && ws) atomicAdd(&counters[1], ws);
}

}  // namespace qb
