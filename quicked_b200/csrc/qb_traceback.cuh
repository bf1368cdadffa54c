// qb_traceback.cuh — BandEd traceback (reference bpm_banded.c:967-1036) and CIGAR text emission
// (reference cigar.c:453-488 "%d%c" run-length text with matches printed, cigar.c:274-289 edit score).
//
// The walk reproduces the reference's tie-breaking exactly: from (v,h) = (m-1,n-1),
//   D if bit v of Pv[column h+1];  else I if bit v of Mv[column h];  else M / X by RAW byte compare;
// leftovers become I... then D....  Column c of the stored matrix holds the state after c text columns in the band
// coordinates of column block c/64.
//
// Ops are emitted right-to-left as 2-bit codes packed 16 per u32 into the tail of the task's op region (the
// reference writes one char per op into a right-aligned buffer, cigar.h:33-47).  While walking we also count the
// edit cost and the byte length of the run-length text, so single-leaf pairs need no second pass over the ops.
#pragma once
#include "qb_common.cuh"

namespace qb {

struct OpWriter {
    u32 *words;     // op region of the task
    int pos;        // next op goes to pos-1
    u32 acc;
    int cost, text_len, cur_op, cur_len;
    __host__ __device__ __forceinline__ void init(u32 *w, int cap) { words = w; pos = cap; acc = 0; cost = 0; text_len = 0; cur_op = -1; cur_len = 0; }
    __host__ __device__ __forceinline__ void emit(int op)
    {
        --pos;
        acc |= (u32)op << (2 * (pos & 15));
        if ((pos & 15) == 0) { words[pos >> 4] = acc; acc = 0; }
        cost += (op != OP_M);
        if (op == cur_op) ++cur_len;
        else {
            if (cur_len) text_len += dec_digits((unsigned)cur_len) + 1;
            cur_op = op; cur_len = 1;
        }
    }
    __host__ __device__ __forceinline__ void finish()
    {
        if (pos & 15) words[pos >> 4] = acc;
        if (cur_len) text_len += dec_digits((unsigned)cur_len) + 1;
    }
};

// The tile walk's writer: 2-bit ops and the edit cost only.  The run-length bookkeeping of OpWriter (~8 instructions on
// the walk's one-op-per-iteration critical path) moves to k_cigar_text's measuring pass, which finds run boundaries
// 16 ops at a time with one XOR (LeafOut.text_len = -1 asks for it).
struct LeanWriter {
    u32 *words;
    int pos, cost;
    u32 acc;
    __host__ __device__ __forceinline__ void init(u32 *w, int cap) { words = w; pos = cap; acc = 0; cost = 0; }
    __host__ __device__ __forceinline__ void emit(int op)
    {
        --pos;
        acc |= (u32)op << (2 * (pos & 15));
        if ((pos & 15) == 0) { words[pos >> 4] = acc; acc = 0; }
        cost += (op != OP_M);
    }
    __host__ __device__ __forceinline__ void finish() { if (pos & 15) words[pos >> 4] = acc; }
};

// LeanWriter as a shift register: an op costs one multiply-add (acc = 4 * acc + op: the first op of a word ends up in its
// top field), the word leaves when its 16 ops are in, and the edit cost is counted per stored word (non-zero 2-bit
// fields), not per op.
struct ShiftWriter {
    u32 *words;
    int pos, cost;
    u32 acc;
    __host__ __device__ __forceinline__ static int nonzero_fields(u32 a)
    {
        const u32 x = (a | (a >> 1)) & 0x55555555u;
#ifdef __CUDA_ARCH__
        return __popc(x);
#else
        return __builtin_popcount(x);
#endif
    }
    __host__ __device__ __forceinline__ void init(u32 *w, int cap) { words = w; pos = cap; acc = 0; cost = 0; }
    __host__ __device__ __forceinline__ void emit(int op)
    {
        acc = acc * 4u + (u32)op;
        --pos;
        if ((pos & 15) == 0) { words[pos >> 4] = acc; cost += nonzero_fields(acc); acc = 0; }
    }
    __host__ __device__ __forceinline__ void finish()
    {
        if (pos & 15) { words[pos >> 4] = acc << (2 * (pos & 15)); cost += nonzero_fields(acc); }
    }
};

// Was band word `w` of stored column `c` ever written by the fill?  Column c (state after c text columns) was written
// with the live range of column block (c-1)/64; a column with c%64==0 is stored after the shift, i.e. with the next
// block's first and the previous block's last (bpm_banded.c:279-287).  Never-written cells read as 0, which is what
// the reference finds there when its arena is fresh (and what the oracle defines).
__device__ __forceinline__ bool cell_written(const int2 *ranges, int B, int c, int w)
{
    if (w < 0 || w >= B) return false;
    if (c == 0) return true;
    const int kb = (c - 1) >> 6;
    const int2 r = ranges[kb];
    const int lo = (c & 63) ? r.x : min(r.x, ranges[kb + 1].x);   // a cut top word keeps its pre-shift value
    return w >= lo && w <= r.y;
}

// The walk of ONE leaf by one thread (see k_traceback_thread).  Tight version for narrow bands: 32-bit cell keys,
// ONE emit point per step (the three outcomes D / I / diagonal are selected, not branched, so divergent lanes do not
// serialise three copies of the op writer), and the live range of the current column block cached in registers.
__device__ __forceinline__ void traceback_walk_thread(int m, int n, i64 cutoff, const ulonglong2 *mat, i64 cs, i64 wsd,
                                                      const int2 *ranges, i64 rstride, const unsigned char *__restrict__ praw,
                                                      const unsigned char *__restrict__ traw, u32 *ops, int ops_cap, LeafOut &o)
{
    const BandGeom g = band_geometry(m, n, cutoff);
    const int B = (int)g.Bc, prolog = (int)g.prolog;
    OpWriter w; w.init(ops, ops_cap);
    int h = n - 1, v = m - 1;
    // cached live ranges of column blocks kb_c and kb_c + 1
    int kb_c = -2; int2 rg0 = make_int2(0, -1), rg1 = make_int2(0, -1);
    // cached entries: key = column * B + word (the reference's flat index; < 2^31 for the narrow bands walked here)
    int keyR = -1, keyL = -1;
    ulonglong2 eR = make_ulonglong2(0, 0), eL = make_ulonglong2(0, 0);
    auto fetch = [&](int c, int wd) -> ulonglong2 {
        // never-written cells read as 0 (see cell_written); the live range of column c comes from block (c-1)/64
        if ((unsigned)c > (unsigned)n || (unsigned)wd >= (unsigned)B) return make_ulonglong2(0, 0);
        if (c > 0) {
            const int kb = (c - 1) >> 6;
            if (kb != kb_c) {
                if (kb == kb_c + 1) rg0 = rg1; else rg0 = ranges[(i64)kb * rstride];
                rg1 = ranges[(i64)(kb + 1) * rstride];
                kb_c = kb;
            }
            const int lo = (c & 63) ? rg0.x : min(rg0.x, rg1.x);
            if (wd < lo || wd > rg0.y) return make_ulonglong2(0, 0);
        }
        return mat[(i64)c * cs + (i64)wd * wsd];
    };
    // eP: the entry the walk will most likely need next (column h-1, the word of row v-1), fetched one step early so
    // that two dependent HBM round trips are in flight per thread instead of one (a second look-ahead entry was
    // measured slower: 11.6 vs 10.6 ms per 1M pairs at 1 kbp)
    int keyP = -1;
    ulonglong2 eP = make_ulonglong2(0, 0);
    while (v >= 0 && h >= 0) {
        const int ev = v - 64 * ((h >> 6) - prolog);
        const int evr = v - 64 * (((h + 1) >> 6) - prolog);
        // word numbers with C truncation; a row outside the band's coordinates makes the reference index the flat
        // [column][word] array across column boundaries (only possible when the band is too narrow)
        const int wr = evr / 64, wl = ev / 64;
        const int fR = (h + 1) * B + wr, fL = h * B + wl;
        if (fR != keyR) {
            if (fR == keyL) eR = eL;
            else if (fR == keyP) eR = eP;
            else if ((unsigned)wr < (unsigned)B) eR = fetch(h + 1, wr);
            else eR = fR >= 0 ? fetch(fR / B, fR % B) : make_ulonglong2(0, 0);
            keyR = fR;
        }
        if (fL != keyL) {
            if (fL == keyP) eL = eP;
            else if ((unsigned)wl < (unsigned)B) eL = fetch(h, wl);
            else eL = fL >= 0 ? fetch(fL / B, fL % B) : make_ulonglong2(0, 0);
            keyL = fL;
        }
        // the raw characters of this cell and the next column's entry: independent of eR / eL, so all in flight together
        const u32 ct = traw[h], cp = praw[v];
        if (h > 0 && v > 0) {
            const int wp = ((v - 1) - 64 * (((h - 1) >> 6) - prolog)) / 64;
            const int fP = (h - 1) * B + wp;
            if (fP != keyP && (unsigned)wp < (unsigned)B) { eP = fetch(h - 1, wp); keyP = fP; }
        }
        const bool isD = (eR.x >> (evr & 63)) & 1ull;
        const bool isI = (eL.y >> (ev & 63)) & 1ull;
        int op = isD ? OP_D : OP_I;
        if (!isD && !isI) op = (ct == cp) ? OP_M : OP_X;     // RAW byte compare (bpm_banded.c:1012)
        w.emit(op);
        v -= (isD || !isI) ? 1 : 0;          // D and diagonal consume a pattern row
        h -= isD ? 0 : 1;                    // I and diagonal consume a text column
    }
    while (h >= 0) { w.emit(OP_I); --h; }
    while (v >= 0) { w.emit(OP_D); --v; }
    w.finish();
    o.n_ops = ops_cap - w.pos; o.cost = w.cost; o.text_len = w.text_len;
    o.fmt = 0; o.pad_ = 0;
}

// One leaf per thread.  Works on both matrix layouts (warp kernel: [column][word]; thread kernel: 32 leaves
// interleaved) through the task's column / word strides.  Per step the walk needs bit v of Pv[column h+1] and of
// Mv[column h]; (Pv,Mv) of one (column, word) is a single 16-byte entry, so the entry fetched for Mv at column h
// is reused for Pv when the walk moves to column h-1, and the live ranges are cached per 64-column block:
// about one 16-byte load per visited column instead of four dependent loads per step.
// 56 registers, 9 CTAs per SM.  Trading registers for occupancy does not pay: 48 / 40 / 32 registers (10 / 12 / 14
// CTAs) spill the cached entries and take 16 / 29 / 65 ms instead of 10.5 per 1 M pairs of 1 kbp.
__global__ void __launch_bounds__(128, 9)
k_traceback_thread(const BandTask *__restrict__ tasks, const int *__restrict__ list, int begin, int n_tasks, i64 mat_sub,
                   const unsigned char *__restrict__ raw, const ulonglong2 *__restrict__ matrix,
                   const int2 *__restrict__ range_pool, u32 *__restrict__ ops_pool, LeafOut *__restrict__ outs)
{
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n_tasks) return;
    BandTask tk = tasks[list ? list[begin + id] : begin + id];
    tk.mat_off -= mat_sub;
    LeafOut o;
    traceback_walk_thread(tk.m, tk.n, tk.cutoff, matrix + tk.mat_off, tk.mat_cs, tk.mat_ws, range_pool + tk.range_off, 1,
                          raw + tk.p_off, raw + tk.t_off, ops_pool + tk.ops_off, tk.ops_cap, o);
    outs[tk.slot] = o;
}

// One leaf per WARP, for leaves written by the warp kernel ([column][band word] layout, long pairs).
// The reference's walk is a 1-bit-per-step dependent chain (bpm_banded.c:987-1023).  Here it is word-parallel:
//   * the 32 lanes prefetch, in ONE HBM round trip, a tile of the traceback state around the current cell
//     (32 columns x the 2 blocks covering rows v-64..v) into shared memory;
//   * lane l then PROBES the cell l steps down the diagonal, (v-l, h-l): its D bit, its I bit and its raw-byte
//     compare.  One ballot finds the first cell that is not a plain diagonal move; all the M/X ops before it are
//     emitted at once (one coalesced byte store), then the single D or I of that cell, and the probe restarts there.
// With ~17 % error a probe advances ~6 steps for the instruction cost of ~1.5 serial steps.  The tile is only a
// cache in front of the same fetch rules as k_traceback_thread (live ranges, flat-index quirks of too-narrow
// bands are walked one serial step at a time), so results are identical.  The output is already run-length
// encoded: u32 runs (len<<2 | op) written right to left (LeafOut.fmt = 1), with the exact text length.
constexpr int kTraceWarpsPerCta = 4;

__global__ void __launch_bounds__(32 * kTraceWarpsPerCta)
k_traceback_warp(const BandTask *__restrict__ tasks, const int *__restrict__ list, int begin, int n_tasks, i64 mat_sub,
                 const unsigned char *__restrict__ raw, const ulonglong2 *__restrict__ matrix,
                 const int2 *__restrict__ range_pool, u32 *__restrict__ ops_pool, LeafOut *__restrict__ outs, int min_B)
{
    __shared__ ulonglong2 s_tile[kTraceWarpsPerCta][32][2];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int id = blockIdx.x * kTraceWarpsPerCta + wib;
    if (id >= n_tasks) return;
    BandTask tk = tasks[list ? list[begin + id] : begin + id];
    tk.mat_off -= mat_sub;
    const BandGeom g = band_geometry(tk.m, tk.n, tk.cutoff);
    const int B = (int)g.Bc, prolog = (int)g.prolog;
    if (B < min_B) return;                   // narrower bands were walked by the tile traceback
    const ulonglong2 *mat = matrix + tk.mat_off;
    const int2 *ranges = range_pool + tk.range_off;
    const unsigned char *praw = raw + tk.p_off, *traw = raw + tk.t_off;
    const i64 cs = tk.mat_cs, wsd = tk.mat_ws;
    ulonglong2 (*tile)[2] = s_tile[wib];
    u32 *runs = ops_pool + tk.ops_off;       // u32 runs (len<<2 | op), written right to left; capacity ops_cap words
    int rpos = tk.ops_cap;                   // next run goes to rpos-1
    int cur_op = -1, cur_len = 0, text_len = 0;      // open run (warp-uniform)
    auto emit_run = [&](int op, int len) {
        if (op == cur_op) { cur_len += len; return; }
        if (cur_len) { if (lane == 0) runs[rpos - 1] = ((u32)cur_len << 2) | (u32)cur_op; --rpos; text_len += dec_digits((unsigned)cur_len) + 1; }
        cur_op = op; cur_len = len;
    };
    auto fetch = [&](int c, int wd) -> ulonglong2 {          // same rules as k_traceback_thread::fetch
        if (c < 0 || c > tk.n || wd < 0 || wd >= B) return make_ulonglong2(0, 0);
        if (c > 0) {
            const int kb = (c - 1) >> 6;
            const int2 r0 = ranges[kb];
            const int lo = (c & 63) ? r0.x : min(r0.x, ranges[kb + 1].x);
            if (wd < lo || wd > r0.y) return make_ulonglong2(0, 0);
        }
        return mat[(i64)c * cs + (i64)wd * wsd];
    };
    int h = tk.n - 1, v = tk.m - 1;          // warp-uniform
    int cost = 0;
    while (v >= 0 && h >= 0) {
        // ---- (re)load the tile: columns c_hi-31..c_hi, absolute blocks bA-1, bA ----
        const int c_hi = h + 1, bA = v >> 6;
        {
            const int c = c_hi - lane;
#pragma unroll
            for (int sidx = 0; sidx < 2; ++sidx) {
                const int b = bA - 1 + sidx;
                ulonglong2 e = make_ulonglong2(0, 0);
                if (c >= 0 && b >= 0) e = fetch(c, b - ((c >> 6) - prolog));
                tile[lane][sidx] = e;
            }
        }
        __syncwarp();
        // ---- probe along the diagonal until the walk leaves the tile ----
        while (v >= 0 && h >= 0) {
            const int vl = v - lane, hl = h - lane;
            bool inside = vl >= 0 && hl >= 0;
            const int ev = vl - 64 * ((hl >> 6) - prolog);
            const int evr = vl - 64 * (((hl + 1) >> 6) - prolog);
            const int wr = evr >> 6, wl = ev >> 6;
            const int tl = c_hi - hl, br = (vl >> 6) - (bA - 1);
            const bool regular = ev >= 0 && evr >= 0 && wr < B && wl < B;
            const bool in_tile = tl <= 31 && br >= 0;
            bool isD = false, isI = false, mm = false;
            if (inside && regular && in_tile) {
                const ulonglong2 eR = tile[tl - 1][br], eL = tile[tl][br];
                isD = (eR.x >> (evr & 63)) & 1ull;
                isI = (eL.y >> (ev & 63)) & 1ull;
                mm = traw[hl] != praw[vl];
            }
            const u32 stop = __ballot_sync(kFull, !inside || !regular || !in_tile || isD || isI);
            const u32 mmask = __ballot_sync(kFull, mm);
            const int nd = stop ? (__ffs(stop) - 1) : 32;        // leading plain-diagonal cells
            {   // the nd diagonal ops, lane 0's first: split the mismatch mask into runs of equal bits
                u32 bits = mmask & (nd >= 32 ? kFull : ((1u << nd) - 1u));
                cost += __popc(bits);
                int left = nd;
                while (left > 0) {
                    const int x = bits & 1u;
                    const u32 t = x ? ~bits : bits;                  // run of the low bit's value
                    int len = t ? (__ffs(t) - 1) : 32;
                    len = min(len, left);
                    emit_run(x ? OP_X : OP_M, len);
                    bits = (len >= 32) ? 0u : (bits >> len);
                    left -= len;
                }
            }
            v -= nd; h -= nd;
            if (nd == 32) continue;
            // lane nd holds the first non-diagonal cell: broadcast why it stopped
            const u32 why = __shfl_sync(kFull, (u32)(isD ? 1u : isI ? 2u : (!inside ? 4u : (!in_tile ? 8u : 16u))), nd);
            if (why == 1u) { emit_run(OP_D, 1); ++cost; --v; }
            else if (why == 2u) { emit_run(OP_I, 1); ++cost; --h; }
            else if (why == 4u) break;                           // ran off the matrix: leftovers below
            else if (why == 8u) break;                           // left the tile: reload it
            else {
                // too-narrow band: this cell follows the reference's flat-index quirks; one serial step (all lanes
                // compute the same thing, lane 0 writes)
                const int e0 = v - 64 * ((h >> 6) - prolog), e1 = v - 64 * (((h + 1) >> 6) - prolog);
                const int w1 = e1 / 64, w0 = e0 / 64;
                const i64 fR = (i64)(h + 1) * B + w1, fL = (i64)h * B + w0;
                const ulonglong2 eR = (w1 >= 0 && w1 < B) ? fetch(h + 1, w1) : (fR >= 0 ? fetch((int)(fR / B), (int)(fR % B)) : make_ulonglong2(0, 0));
                const ulonglong2 eL = (w0 >= 0 && w0 < B) ? fetch(h, w0) : (fL >= 0 ? fetch((int)(fL / B), (int)(fL % B)) : make_ulonglong2(0, 0));
                int op;
                if ((eR.x >> (e1 & 63)) & 1ull) { op = OP_D; --v; }
                else if ((eL.y >> (e0 & 63)) & 1ull) { op = OP_I; --h; }
                else { op = traw[h] == praw[v] ? OP_M : OP_X; --h; --v; }
                emit_run(op, 1); cost += (op != OP_M);
            }
        }
        __syncwarp();
    }
    // leftovers: I... then D... (bpm_banded.c:1025-1034)
    if (h >= 0) { emit_run(OP_I, h + 1); cost += h + 1; }
    if (v >= 0) { emit_run(OP_D, v + 1); cost += v + 1; }
    emit_run(-1, 0);                                                       // flush the open run
    if (lane == 0) {
        LeafOut o;
        o.n_ops = tk.ops_cap - rpos; o.cost = cost; o.text_len = text_len; o.fmt = 1; o.pad_ = 0;
        outs[tk.slot] = o;
    }
}

// Per-pair list of leaves (left to right).  Single-leaf pairs: n_leaves == 1.
struct PairLeaves {
    i64 first_leaf;   // index into the global leaf arrays (BandTask / LeafOut), leaves of a pair are contiguous
    int n_leaves;
    int pad_;
};

// Text sink of k_cigar_text: characters are collected in a 64-bit register and leave as aligned 8-byte stores (the
// pair's text starts at an arbitrary byte, so the first and last few characters go out one by one).  One byte store
// per character made the kernel store-transaction bound (3.2 ms per 1M pairs at 1 kbp for 0.3 GB of text).
struct TextSink {
    char *base;        // 8-byte aligned address of the word being filled
    u64 acc;
    int n;             // bytes of the current word already taken (including the `skip` bytes in front of the text)
    int skip;          // bytes of the first word that lie before the text
    int total;
    __device__ __forceinline__ void init(char *dst)
    {
        skip = (int)((unsigned long long)dst & 7ull);
        base = dst - skip; acc = 0; n = skip; total = 0;
    }
    __device__ __forceinline__ void flush_word()
    {
        if (skip) { for (int k = skip; k < 8; ++k) base[k] = (char)(acc >> (8 * k)); skip = 0; }
        else *reinterpret_cast<u64 *>(base) = acc;
        base += 8;
    }
    // chunk: `len` (1..8) characters, first character in the low byte
    __device__ __forceinline__ void put(u64 chunk, int len)
    {
        total += len;
        acc |= chunk << (8 * n);
        if (n + len >= 8) {
            flush_word();
            acc = n ? (chunk >> (8 * (8 - n))) : 0ull;
            n = n + len - 8;
        } else n += len;
    }
    __device__ __forceinline__ void put_run(int len, int op)
    {
        unsigned x = (unsigned)len;
        const u64 opc = (u64)(unsigned char)"MXID"[op];
        if (x < 10u) { put((u64)('0' + x) | (opc << 8), 2); return; }                 // most runs are this short
        if (x < 100u) { put((u64)('0' + x / 10u) | ((u64)('0' + x % 10u) << 8) | (opc << 16), 3); return; }
        const int nd = dec_digits(x);
        if (nd <= 7) {
            u64 chunk = (u64)(unsigned char)"MXID"[op] << (8 * nd);
            for (int k = nd - 1; k >= 0; --k) { chunk |= (u64)('0' + x % 10u) << (8 * k); x /= 10u; }
            put(chunk, nd + 1);
        } else {                                            // >= 10^7: digits first (most significant first), then the op
            unsigned p10 = 1; for (int k = 1; k < nd; ++k) p10 *= 10u;
            for (int k = 0; k < nd; ++k) { put((u64)('0' + (x / p10) % 10u), 1); p10 /= 10u; }
            put((u64)(unsigned char)"MXID"[op], 1);
        }
    }
    __device__ __forceinline__ void finish()                // the terminating NUL, then the partial word
    {
        put(0ull, 1); --total;
        for (int k = skip; k < n; ++k) base[k] = (char)(acc >> (8 * k));
    }
    __device__ __forceinline__ void finish_raw()            // the partial word only (a piece inside a longer text)
    {
        for (int k = skip; k < n; ++k) base[k] = (char)(acc >> (8 * k));
    }
};

// Pairs whose text the warp kernel (k_cigar_text_warp) writes: several leaves (Hirschberg pairs are long by
// construction) or one leaf of >= kLongOps op slots.
constexpr int kLongOps = 4096;
__device__ __forceinline__ bool pair_is_long(const PairLeaves &p, const BandTask *__restrict__ leaves)
{
    return p.n_leaves > 1 || (p.n_leaves == 1 && leaves[p.first_leaf].ops_cap >= kLongOps);
}

// Pass over a pair's ops, merging runs across leaf boundaries.  WRITE=false: only measure (text bytes without NUL,
// and the edit cost); WRITE=true: write the text at cigar + cigar_off[pair] and NUL-terminate.
template <bool WRITE>
__global__ void __launch_bounds__(128)
k_cigar_text(const PairLeaves *__restrict__ pl, int n_pairs, const BandTask *__restrict__ leaves,
             const LeafOut *__restrict__ leaf_out, const u32 *__restrict__ ops_pool, int *__restrict__ text_len,
             const i64 *__restrict__ cigar_off, char *__restrict__ cigar, int skip_long)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    const PairLeaves p = pl[i];
    if (skip_long && pair_is_long(p, leaves)) return;      // k_cigar_text_warp's
    if (!WRITE && (p.n_leaves < 1 || (p.n_leaves == 1 && leaf_out[p.first_leaf].text_len >= 0))) return;   // the walk already measured it
    if (WRITE && p.n_leaves == 0) return;        // empty string: the buffer is pre-zeroed
    TextSink sink;
    if (WRITE) sink.init(cigar + cigar_off[i]);
    int out = 0, cur_op = -1, cur_len = 0;
    auto close_run = [&]() {
        if (!cur_len) return;
        if (WRITE) sink.put_run(cur_len, cur_op);
        else out += dec_digits((unsigned)cur_len) + 1;
    };
    for (int l = 0; l < p.n_leaves; ++l) {
        const BandTask tk = leaves[p.first_leaf + l];
        const LeafOut lo = leaf_out[p.first_leaf + l];
        const int n_ops = lo.n_ops;
        const u32 *words = ops_pool + tk.ops_off;
        int pos = tk.ops_cap - n_ops;
        const int end = tk.ops_cap;
        if (lo.fmt == 1) {                           // u32 runs (warp traceback)
            for (; pos < end; ++pos) {
                const u32 r = words[pos];
                const int op = (int)(r & 3u), len = (int)(r >> 2);
                if (op == cur_op) cur_len += len;
                else { close_run(); cur_op = op; cur_len = len; }
            }
            continue;
        }
        // 16 ops per word: the positions where the op changes come from one XOR against the word shifted by one op,
        // so the (divergent) run bookkeeping runs once per run, not once per op
        while (pos < end) {
            const u32 wv = words[pos >> 4];
            const int lo = pos & 15, hi = min(end - (pos & ~15), 16);
            u32 prevv = wv << 2;
            prevv = (prevv & ~(3u << (2 * lo))) | ((u32)(cur_op & 3) << (2 * lo));
            const u32 diff = wv ^ prevv;
            u32 chg = (diff | (diff >> 1)) & 0x55555555u;
            if (cur_op < 0) chg |= 1u << (2 * lo);
            chg &= (hi == 16 ? 0xffffffffu : ((1u << (2 * hi)) - 1u)) & ~((1u << (2 * lo)) - 1u);
            int start = lo;
            while (chg) {
                const int k = (__ffs((int)chg) - 1) >> 1;
                chg &= chg - 1;
                cur_len += k - start;
                close_run();
                cur_op = (int)((wv >> (2 * k)) & 3u); cur_len = 0; start = k;
            }
            cur_len += hi - start;
            pos = (pos & ~15) + hi;
        }
    }
    close_run();
    if (WRITE) sink.finish();
    else text_len[i] = out;
}

// The same pass with ONE WARP PER PAIR, for long pairs (pair_is_long): a 10 kbp pair has ~12 k ops, a 100 kbp pair
// ~120 k, and one thread walking them alone is a chain of that length (3 ms per 100 k pairs of 10 kbp, 21 ms per 2 k
// pairs of 100 kbp).  Here every lane takes one 16-op word of a 512-op batch, finds the run starts inside its word with
// the XOR trick above, and CLOSES the runs that end in its word: the run's first op comes from a max-scan of the
// lanes' last run starts (the open run of the previous batch / leaf rides along as a negative coordinate).  An
// exclusive scan of the lanes' text bytes gives every lane its place in the output.
template <bool WRITE>
__global__ void __launch_bounds__(128)
k_cigar_text_warp(const PairLeaves *__restrict__ pl, int n_pairs, const BandTask *__restrict__ leaves,
                  const LeafOut *__restrict__ leaf_out, const u32 *__restrict__ ops_pool, int *__restrict__ text_len,
                  const i64 *__restrict__ cigar_off, char *__restrict__ cigar)
{
    __shared__ __align__(16) unsigned char s_stage[WRITE ? 4 : 1][WRITE ? 1056 : 16];   // <= 2 bytes per op of a 512-op batch + one long run
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n_pairs) return;
    const PairLeaves p = pl[i];
    if (!pair_is_long(p, leaves)) return;
    char *dst = WRITE ? cigar + cigar_off[i] : nullptr;
    constexpr int NONE = -0x40000000;
    int out = 0;                           // text bytes so far (warp-uniform)
    int cur_op = -1, cur_len = 0;          // the open run (warp-uniform)
    for (int l = 0; l < p.n_leaves; ++l) {
        const BandTask &tk = leaves[p.first_leaf + l];
        const LeafOut lo = leaf_out[p.first_leaf + l];
        const u32 *words = ops_pool + tk.ops_off;
        const int end = tk.ops_cap, pos0 = end - lo.n_ops;
        if (lo.n_ops <= 0) continue;
        if (lo.fmt == 1) {                                   // u32 runs (full-matrix warp traceback): rare, walked uniformly
            for (int pos = pos0; pos < end; ++pos) {
                const u32 r = words[pos];
                const int op = (int)(r & 3u), len = (int)(r >> 2);
                if (op == cur_op) { cur_len += len; continue; }
                if (cur_op >= 0) {
                    if (WRITE && lane == 0) { TextSink sk; sk.init(dst + out); sk.put_run(cur_len, cur_op); sk.finish_raw(); }
                    out += dec_digits((unsigned)cur_len) + 1;
                }
                cur_op = op; cur_len = len;
            }
            continue;
        }
        const int wfirst = pos0 >> 4, wlast = (end - 1) >> 4;
        for (int w0 = wfirst; w0 <= wlast; w0 += 32) {
            const int widx = w0 + lane;
            const bool have = widx <= wlast;
            const u32 wv = have ? __ldg(words + widx) : 0u;
            const int lo_k = (widx == wfirst) ? (pos0 & 15) : 0;
            const int hi_k = have ? (widx == wlast ? ((end - 1) & 15) + 1 : 16) : 0;
            // the op in front of this word's first op: the open run's op for the first word of the batch, else the
            // previous lane's top op
            const u32 up = __shfl_up_sync(kFull, wv >> 30, 1);
            const int prev_op = (lane == 0) ? cur_op : (int)up;
            u32 prevv = wv << 2;
            prevv = (prevv & ~(3u << (2 * lo_k))) | ((u32)(prev_op & 3) << (2 * lo_k));
            const u32 diff = wv ^ prevv;
            u32 chg = (diff | (diff >> 1)) & 0x55555555u;
            if (prev_op < 0) chg |= 1u << (2 * lo_k);
            chg &= (hi_k == 16 ? 0xffffffffu : ((1u << (2 * hi_k)) - 1u)) & ~((1u << (2 * lo_k)) - 1u);
            // coordinates: op k of lane's word sits at 16 * lane + k; the open run started at first - cur_len
            const int first = (w0 == wfirst) ? (pos0 & 15) : 0;
            const int open_start = (cur_op >= 0) ? first - cur_len : NONE;
            const int ls = chg ? 16 * lane + ((31 - __clz((int)chg)) >> 1) : NONE;
            int incl = ls;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl = max(incl, y); }
            int prev = __shfl_up_sync(kFull, incl, 1);
            if (lane == 0) prev = NONE;
            prev = max(prev, open_start);
            // pass 1: text bytes of the runs that end in this word
            int bytes = 0;
            {
                int pv = prev; u32 c = chg;
                while (c) {
                    const int k = (__ffs((int)c) - 1) >> 1;
                    c &= c - 1;
                    const int st = 16 * lane + k;
                    if (pv != NONE) bytes += dec_digits((unsigned)(st - pv)) + 1;
                    pv = st;
                }
            }
            int bi = bytes;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(kFull, bi, o); if (lane >= o) bi += y; }
            const int total = __shfl_sync(kFull, bi, 31);
            if (WRITE && total) {
                // the batch's text is staged in shared memory and leaves as aligned 4-byte words (one byte store per
                // character is store-transaction bound); stage[j] maps to the global byte (dst + out - mis) + j
                unsigned char *stg = s_stage[threadIdx.x >> 5];
                const int mis = (int)((unsigned long long)(dst + out) & 3ull);
                if (bytes) {
                    unsigned char *q = stg + mis + (bi - bytes);
                    int pv = prev; u32 c = chg;
                    while (c) {
                        const int k = (__ffs((int)c) - 1) >> 1;
                        c &= c - 1;
                        const int st = 16 * lane + k;
                        if (pv != NONE) {
                            unsigned x = (unsigned)(st - pv);
                            const unsigned char opc = (unsigned char)"MXID"[k > lo_k ? (int)((wv >> (2 * (k - 1))) & 3u) : prev_op];
                            if (x < 10u) { q[0] = (unsigned char)('0' + x); q[1] = opc; q += 2; }      // most runs are this short
                            else {
                                const int nd = dec_digits(x);
                                for (int d = nd - 1; d >= 0; --d) { q[d] = (unsigned char)('0' + x % 10u); x /= 10u; }
                                q[nd] = opc;
                                q += nd + 1;
                            }
                        }
                        pv = st;
                    }
                }
                __syncwarp();
                char *g = dst + out - mis;
                const int nb = mis + total;
                const int wlo = mis ? 1 : 0, whi = nb >> 2;               // the whole words [wlo, whi) of the span
                for (int wq = wlo + lane; wq < whi; wq += 32) reinterpret_cast<u32 *>(g)[wq] = reinterpret_cast<const u32 *>(stg)[wq];
                if (lane < 4) {                                           // its first and last few bytes
                    if (mis && lane >= mis && lane < nb) g[lane] = (char)stg[lane];
                    const int kt = 4 * whi + lane;
                    if (kt >= mis && kt < nb) g[kt] = (char)stg[kt];
                }
                __syncwarp();
            }
            out += total;
            // the open run after this batch
            const int nw = min(32, wlast - w0 + 1);
            const int hi_last = (w0 + nw - 1 == wlast) ? ((end - 1) & 15) + 1 : 16;
            const int last_start = max(__shfl_sync(kFull, incl, 31), open_start);
            cur_op = (int)((__shfl_sync(kFull, wv, nw - 1) >> (2 * (hi_last - 1))) & 3u);
            cur_len = 16 * (nw - 1) + hi_last - last_start;
        }
    }
    if (cur_op >= 0) {
        if (WRITE && lane == 0) { TextSink sk; sk.init(dst + out); sk.put_run(cur_len, cur_op); sk.finish_raw(); }
        out += dec_digits((unsigned)cur_len) + 1;
    }
    if (!WRITE && lane == 0) text_len[i] = out;
}

}  // namespace qb
