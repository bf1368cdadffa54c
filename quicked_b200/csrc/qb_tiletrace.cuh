// qb_tiletrace.cuh — BandEd traceback over TILE RECORDS (reference walk: bpm_banded.c:967-1036).
//
// The reference walks a stored (n+1) x B matrix of (Pv,Mv) words (bpm_banded.c:139-140).  The tile fill
// (qb_tiles.cuh) stores, per 64x64 tile, only the block's (Pv,Mv) at the tile's first column and the 64 carry-in
// pairs (32 bytes).  A tile's 64 columns are a pure function of that record, the block's five match masks and 64
// text codes, so the walk recomputes exactly the tiles it visits — about 1.5 per 64 columns instead of the band's
// whole height — and never reads more than 32 bytes of traceback state per tile from HBM.
//
// ONE LEAF PER THREAD.  For the tile the walk is in, the thread recomputes the columns from the tile's first up to the
// walk's column and keeps, per column, the walk's DECISION for the 16 rows around the walk's diagonal as two bit planes
// in one u32 of shared memory (the walk can only leave its diagonal through an insertion or a deletion):
//     a = Pv[c+1] | ~(Mv[c] | Eq[c]),  b = ~Pv[c+1] & (Mv[c] | ~Eq[c])
//     (a,b) = (1,0) deletion, (0,1) insertion, (1,1) mismatch, (0,0) match
// which is the reference's test order: D if Pv[col h+1] bit v, else I if Mv[col h] bit v, else M/X (:1002-1023).
// M/X come from the match masks; a cell whose row or column holds a character outside "ACGTN" (where equal codes do
// not imply equal bytes) compares the RAW bytes like the reference (:1012).  When the walk drifts out of the 16-row
// slice the tile is recomputed around the walk's current cell.
//
// The walk PUNTS (returns non-zero, nothing written is used) as soon as it would read a cell outside the live band
// of its column block, or the first-row word of a block the lower cut has just dropped: there the reference reads
// stale or never-written words (too-narrow bands) and the exact full-matrix kernels
// (qb_banded.cuh + qb_traceback.cuh) redo the leaf.
#pragma once
#include "qb_tiles.cuh"
#include "qb_traceback.cuh"

namespace qb {

constexpr int kTraceCols = 64;        // plane words per thread (one per tile column)
constexpr int kTraceHalf = 8;         // rows kept on each side of the walk's diagonal

// One column of the recompute: the Myers block update of qb_tiles.cuh without the carry-outs (the record holds the tile's
// carry-ins, top bit first in wp / wm).
#define QB_TRACE_STEP(EQ)                                                                                   \
    {                                                                                                       \
        const u64 xv_ = (EQ) | mv;                                                                          \
        const u64 xh_ = (add_with_top_bit((EQ) & pv, pv, wm) ^ pv) | (EQ);                                   \
        const u64 ph_ = mv | ~(xh_ | pv);                                                                   \
        const u64 mh_ = pv & xh_;                                                                           \
        const u32 phl_ = (u32)ph_, phh_ = (u32)(ph_ >> 32), mhl_ = (u32)mh_, mhh_ = (u32)(mh_ >> 32);       \
        const u64 ph2_ = ((u64)fsl32(phl_, phh_, 1) << 32) | (u64)fsl32(wp, phl_, 1);                       \
        const u64 mh2_ = ((u64)fsl32(mhl_, mhh_, 1) << 32) | (u64)fsl32(wm, mhl_, 1);                       \
        wp <<= 1; wm <<= 1;                                                                                 \
        pv = mh2_ | ~(xv_ | ph2_);                                                                          \
        mv = ph2_ & xv_;                                                                                    \
    }

// first row of the 16-row slice kept for tile column s: the walk's diagonal - kTraceHalf, clamped into the block
QB_HD int trace_slice_lo(int lo0, int s)
{
    const int lo = lo0 + s;
    return lo < 0 ? 0 : (lo > 48 ? 48 : lo);
}

// planes[s * ps]: decision planes of tile column s; eq[c * eqs]: the block's match masks; ttext: the task's aligned text
// codes (tile-text pool, 8 per u64; its last word flags a text with characters outside "ACGTN").
QB_HD int tile_traceback(const BandTask &tk, const TileRec *recs, const int2 *ranges, const u64 *ttext,
                         const unsigned char *raw, const u64 *peq_pool, u32 *ops, u32 *planes, int ps, u64 *eq, int eqs,
                         LeafOut &o)
{
    const BandGeom g = band_geometry(tk.m, tk.n, tk.cutoff);
    const int B = (int)g.Bc, prolog = (int)g.prolog;
    const int nshift = tk.n >> 6;
    const u64 *pq = peq_pool + tk.peq_off;
    const u64 *tt = ttext + tk.tt_off;
    const bool text_odd = tt[(tk.n + 63) / 64 * 8] != 0;
    const unsigned char *praw = raw + tk.p_off, *traw = raw + tk.t_off;
    ShiftWriter w; w.init(ops, tk.ops_cap);
    int h = tk.n - 1, v = tk.m - 1;
    int eq_block = -1;
    u64 rowodd = 0;
    while (v >= 0 && h >= 0) {
        const int kb = h >> 6, b = v >> 6;
        const int j = b - (kb - prolog);
        const int2 rg = ranges[kb];
        if (j < rg.x || j > rg.y) return 1;                          // outside the live band of this column block
        const int s0 = h & 63, r0 = v & 63;
        if (s0 == 63) {
            // Pv[col 64(kb+1)] lives in the NEXT block's coordinates (stored after the shift, bpm_banded.c:279-287):
            // word j-1 there, which the reference only wrote from the new `first` on
            if (kb + 1 > nshift || j - 1 < ranges[kb + 1].x) return 2;
        }
        // the tile's record and text first: one round trip for everything the recompute reads
        const TileRec rec = recs[(i64)kb * B + j];
        const u64 *tw = tt + 8 * (i64)kb;
        u64 cw = tw[0];
        if (b != eq_block) {
#pragma unroll
            for (int c = 0; c < kAlpha; ++c) eq[c * eqs] = pq[(i64)b * kPeqStride + c];
            rowodd = pq[(i64)b * kPeqStride + kAlpha];
            eq_block = b;
        }
        // ---- recompute tile (j, kb) ----
        u64 pv = rec.pv0, mv = rec.mv0;
        const int lo0 = r0 - s0 - kTraceHalf;                        // lowest slice row at column 0 (may be negative)
        // ALL 64 columns, every column with its planes, whatever s0 is: the lanes of a warp walk different leaves, and a
        // trip count or a "planes needed?" test that depends on the leaf makes them take turns (measured: 15 of 32 lanes
        // active on average with both, and the kernel is bound by warp instructions on the integer pipe).  Columns past s0
        // cost nothing extra that way; past the text they hold code 4 (k_tile_text) and are never read by the walk.
#pragma unroll 1
        for (int s8 = 0; s8 < 64; s8 += 8) {
            const u64 cnext = tw[(s8 >> 3) + 1];                     // the next group's codes (the pool has a word past every tile)
            u32 wp = (s8 < 32 ? rec.cin.p0 : rec.cin.p1) << (s8 & 31), wm = (s8 < 32 ? rec.cin.m0 : rec.cin.m1) << (s8 & 31);
            u32 *pl = planes + s8 * ps;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const u64 e = eq[(unsigned)(cw & 7u) * eqs];
                cw >>= 8;
                const u64 mv_old = mv;
                QB_TRACE_STEP(e);
                const u64 pa = pv | ~(mv_old | e), pb = ~pv & (mv_old | ~e);
                const int lo = trace_slice_lo(lo0, s8 + i);
                const u32 sa = (u32)(pa >> lo), sb = (u32)(pb >> lo);
#ifdef __CUDA_ARCH__
                pl[i * ps] = __byte_perm(sa, sb, 0x5410);
#else
                pl[i * ps] = (sa & 0xffffu) | (sb << 16);
#endif
            }
            cw = cnext;
        }
        // ---- walk inside the tile ----
        int r = r0, s = s0, q = kTraceHalf;                          // q: row - (diagonal - kTraceHalf), the unclamped slice index
        if (!(text_odd || rowodd != 0)) {
            // until the walk leaves the tile (top / left edge) or drifts out of the slice
            while ((r | s) >= 0 && (unsigned)q <= 15u) {
                const u32 t2 = (planes[s * ps] >> (r - trace_slice_lo(lo0, s))) & 0x10001u;
                const u32 a = t2 & 1u, bb = t2 >> 16;
                w.emit((int)(((a ^ bb) << 1) | a));                  // (1,0) D = 3, (0,1) I = 2, (1,1) X = 1, (0,0) M = 0
                q += (int)bb - (int)a;
                r -= (int)(1u - (bb & ~a));                          // all but I consume a pattern row
                s -= (int)(1u - (a & ~bb));                          // all but D consume a text column
            }
        } else {
            while ((r | s) >= 0 && (unsigned)q <= 15u) {
                const u32 t2 = (planes[s * ps] >> (r - trace_slice_lo(lo0, s))) & 0x10001u;
                const u32 a = t2 & 1u, bb = t2 >> 16;
                int op = (int)(((a ^ bb) << 1) | a);
                if (a == bb && (text_odd || ((rowodd >> r) & 1ull)))  // odd character: raw bytes decide (bpm_banded.c:1012)
                    op = (traw[64 * kb + s] == praw[64 * b + r]) ? OP_M : OP_X;
                w.emit(op);
                q += (int)bb - (int)a;
                r -= (op != OP_I) ? 1 : 0;
                s -= (op != OP_D) ? 1 : 0;
            }
        }
        v = 64 * b + r;
        h = 64 * kb + s;
    }
    while (h >= 0) { w.emit(OP_I); --h; }
    while (v >= 0) { w.emit(OP_D); --v; }
    w.finish();
    o.n_ops = tk.ops_cap - w.pos; o.cost = w.cost; o.text_len = -1; o.fmt = 0; o.pad_ = 0;      // text length: k_cigar_text<false>
    return 0;
}

#ifdef __CUDACC__
// One leaf per thread; leaves that punt are appended to punt_list for the exact kernels.
// 64 threads: 19 KB of shared memory per CTA, eleven CTAs = 704 leaves per SM.  100 k leaves (BASELINE configs[2]) then are
// ONE wave; with 128-thread CTAs (five per SM) 42 of 782 CTAs formed a second wave that doubled the kernel's time.
constexpr int kTileTraceThreads = 64;
__global__ void __launch_bounds__(kTileTraceThreads)
k_traceback_tiles(const BandTask *__restrict__ tasks, const int *__restrict__ list, int begin, int n_tasks, i64 rec_sub,
                  const u64 *__restrict__ ttext, const unsigned char *__restrict__ raw, const u64 *__restrict__ peq,
                  const TileRec *__restrict__ recs, const int2 *__restrict__ range_pool, const BandOut *__restrict__ fill_out,
                  u32 *__restrict__ ops_pool, LeafOut *__restrict__ outs, int *__restrict__ punt_list, int *__restrict__ punt_count)
{
    __shared__ u32 s_planes[kTraceCols * kTileTraceThreads];
    __shared__ u64 s_eq[kAlpha * kTileTraceThreads];
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n_tasks) return;
    const int ti = list ? list[begin + id] : begin + id;
    const BandTask tk = tasks[ti];
    if (tile_ring_for(band_geometry(tk.m, tk.n, tk.cutoff).Bc) > kTileMaxRing) return;   // full-matrix leaf (warp kernels)
    LeafOut o;
    int rc = 3;
    if (!fill_out || fill_out[tk.slot].pos_v != kTilePunted)
        rc = tile_traceback(tk, recs + (tk.mat_off - rec_sub) / 2, range_pool + tk.range_off, ttext, raw, peq, ops_pool + tk.ops_off,
                            s_planes + threadIdx.x, kTileTraceThreads, s_eq + threadIdx.x, kTileTraceThreads, o);
    if (rc) { punt_list[atomicAdd(punt_count, 1)] = ti; return; }
    outs[tk.slot] = o;
}
#endif

}  // namespace qb
