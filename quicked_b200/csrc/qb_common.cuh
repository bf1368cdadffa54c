// qb_common.cuh — shared host/device definitions of the B200-native QuickEd path.
//
// Vocabulary follows the reference: pattern = rows (bit-packed, 64 rows per "block"/word), text = columns,
// "word-step" = one Myers block update (64 rows x 1 column), band = the sliding set of live blocks of a column.
// Citations are file:line under the reference tree.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace qb {

typedef unsigned long long u64;
typedef long long i64;
typedef unsigned int u32;

constexpr int kAlpha = 5;            // A C G T other (reference bpm_commons.h:31)
constexpr u32 kFull = 0xffffffffu;

// ---- per-pair / per-task records living in HBM -------------------------------------------------------

// One pair of the uploaded batch.
struct PairRec {
    i64 p_off, t_off;     // offsets of pattern / text in the packed character buffer
    int m, n;             // lengths
    i64 peq_off;          // u64 index of the forward match-mask table in the PEQ pool ([nbp][kPeqStride] layout)
    int nbp;              // blocks in that table = ceil(m/64) + 2 (two all-zero blocks appended)
    int pad_;
    i64 ops_off;          // u32 index of this pair's 2-bit op words in the fused-path region of the op pool
};

// One BandEd work item: a score-only pass or a full-matrix leaf over a sub-rectangle of a pair.
struct BandTask {
    i64 p_off, t_off;     // sub-pattern / sub-text start (forward coordinates in the packed buffer)
    int m, n;             // sub-problem lengths
    int rev;              // 1: run on the reversed sub-sequences (Hirschberg reverse pass)
    int finish;           // score-only: number of text columns to process
    i64 cutoff;           // bound handed to the band geometry
    i64 peq_off;          // match masks of this sub-pattern ([nbp][kPeqStride])
    int nbp;
    int pair;             // owning pair index
    i64 mat_off;          // full mode: first 16-byte (Pv,Mv) entry of this task in the matrix pool
    int mat_cs, mat_ws;   // full mode: entry (column c, band word w) lives at mat_off + c*mat_cs + w*mat_ws
    i64 scores_off;       // int32 index: per-block running scores (zero-initialised)
    i64 state_off;        // score-only: u64 index where the final Pv[B] then Mv[B] are exported
    i64 ops_off;          // leaf: u32 index of the 2-bit op words region
    i64 range_off;        // leaf: int2 index of the per-column-block live range (first,last) of the band
    int ops_cap;          // leaf: capacity in ops (= m+n rounded up to 16)
    int slot;             // free for the scheduler (e.g. node id in the Hirschberg tree)
    i64 tt_off;           // tile kernels: u64 index of this task's aligned text codes in the tile-text pool (set on the device)
};

// Result of a BandEd pass.
struct BandOut {
    int score;            // band score (reference cigar->score of the pass)
    int first, last;      // lower_block / higher_block exported for Hirschberg (bpm_banded.c:962-963)
    int pos_v;            // band origin (block) at the end
};

// Result of a leaf traceback.
struct LeafOut {
    int n_ops;            // ops emitted (they occupy the tail of the task's ops region)
    int cost;             // X+I+D count
    int text_len;         // bytes of the RLE text of this leaf alone (without NUL)
    int fmt;              // 0: ops are 2-bit codes, 16 per u32 (thread walk); 1: u32 runs (len<<2 | op) (warp walk)
    int pad_;
};

// ---- BandEd geometry: reference bpm_banded.c:121-135 (allocate), :359-361/:801-803 (score-only height) ----
struct BandGeom {
    i64 k;        // effective cutoff = max(|n-m|+1, cutoff, 65)
    i64 d;        // m - n
    i64 rel, prolog, Bc, Bs, fin;
};

__host__ __device__ inline i64 ceil_div(i64 a, i64 b) { return (a + b - 1) / b; }

__host__ __device__ inline BandGeom band_geometry(i64 m, i64 n, i64 cutoff)
{
    BandGeom g;
    i64 kend = (n > m ? n - m : m - n) + 1;
    g.k = kend > cutoff ? kend : cutoff;
    if (g.k < 65) g.k = 65;
    g.d = m - n;
    const i64 ad = g.d >= 0 ? g.d : -g.d;
    g.rel = ceil_div(g.k - ad, 2);
    if (g.d >= 0) {
        g.prolog = ceil_div(g.rel, 64);
        g.Bc = ceil_div(g.rel + g.d, 64) + 1 + g.prolog;
    } else {
        g.prolog = ceil_div(g.rel - g.d, 64);
        g.Bc = ceil_div(g.rel, 64) + 1 + g.prolog;
    }
    g.Bs = ceil_div(g.k, 64) + 1;
    g.fin = g.prolog * 64 + g.d;
    return g;
}

// ---- device helpers ------------------------------------------------------------------------------------
#ifdef __CUDACC__

// reference dna_text.c:41-46: A/a 0, C/c 1, G/g 2, T/t 3, everything else 4
__device__ __forceinline__ int enc_base(unsigned c)
{
    const unsigned u = c & 0xdfu;                 // fold case (only meaningful for letters)
    const bool letter = ((c | 0x20u) - 'a') < 26u;
    int code = 4;
    if (letter) {
        code = (u == 'A') ? 0 : (u == 'C') ? 1 : (u == 'G') ? 2 : (u == 'T') ? 3 : 4;
    }
    return code;
}

// Stored code byte = enc_base | kCodeOdd when the character is not one of "ACGTN" (upper case): on the others the
// encoding is injective, so for two unflagged characters "same code" and "same raw byte" are the same statement —
// the WindowEd walk then prices a diagonal step from the match mask instead of loading both raw bytes.
constexpr unsigned kCodeOdd = 8u;
__device__ __forceinline__ unsigned enc_stored(unsigned c)
{
    const bool plain = (c == 'A') | (c == 'C') | (c == 'G') | (c == 'T') | (c == 'N');
    return (unsigned)enc_base(c) | (plain ? 0u : kCodeOdd);
}

// One Myers block update with bit-63 carry-out (reference bpm_commons.h:82-101).
__host__ __device__ __forceinline__ void myers_step(u64 eq, u64 &pv, u64 &mv, u32 hp_in, u32 hm_in, u32 &hp_out, u32 &hm_out)
{
    const u64 xv = eq | mv;
    const u64 eqh = eq | (u64)hm_in;
    const u64 xh = (((eqh & pv) + pv) ^ pv) | eqh;
    u64 ph = mv | ~(xh | pv);
    u64 mh = pv & xh;
    hp_out = (u32)(ph >> 63);
    hm_out = (u32)(mh >> 63);
    ph = (ph << 1) | (u64)hp_in;
    mh = (mh << 1) | (u64)hm_in;
    pv = mh | ~(xv | ph);
    mv = ph & xv;
}

// Bit-63 update that also hands back the pre-shift horizontal deltas (for a level-masked score, bpm_commons.h:60-61).
__device__ __forceinline__ void myers_step_hv(u64 eq, u64 &pv, u64 &mv, u32 hp_in, u32 hm_in, u32 &hp_out, u32 &hm_out,
                                              u64 &ph_raw, u64 &mh_raw)
{
    const u64 xv = eq | mv;
    const u64 eqh = eq | (u64)hm_in;
    const u64 xh = (((eqh & pv) + pv) ^ pv) | eqh;
    u64 ph = mv | ~(xh | pv);
    u64 mh = pv & xh;
    ph_raw = ph; mh_raw = mh;
    hp_out = (u32)(ph >> 63);
    hm_out = (u32)(mh >> 63);
    ph = (ph << 1) | (u64)hp_in;
    mh = (mh << 1) | (u64)hm_in;
    pv = mh | ~(xv | ph);
    mv = ph & xv;
}

// Same update, carry-out taken at bit `ob` (reference bpm_commons.h:49-68 with level_mask = 1<<ob).
__host__ __device__ __forceinline__ void myers_step_at(u64 eq, u64 &pv, u64 &mv, u32 hp_in, u32 hm_in, int ob,
                                              u32 &hp_out, u32 &hm_out)
{
    const u64 xv = eq | mv;
    const u64 eqh = eq | (u64)hm_in;
    const u64 xh = (((eqh & pv) + pv) ^ pv) | eqh;
    u64 ph = mv | ~(xh | pv);
    u64 mh = pv & xh;
    hp_out = (u32)(ph >> ob) & 1u;
    hm_out = (u32)(mh >> ob) & 1u;
    ph = (ph << 1) | (u64)hp_in;
    mh = (mh << 1) | (u64)hm_in;
    pv = mh | ~(xv | ph);
    mv = ph & xv;
}

// PEQ tables are stored [block][kPeqStride] (5 match masks + the mask of rows holding a character outside "ACGTN"
// = 48 bytes per 64-row block, 16-byte aligned):
// everything a thread needs for one block is three 16-byte loads from two DRAM sectors.
constexpr int kPeqStride = 6;

// Bytes [s, s+16) of the 32-byte concatenation a:b (a first), 0 <= s < 16: realigns a 16-byte window of codes that
// was fetched with two aligned 16-byte loads.
__device__ __forceinline__ uint4 realign16(const uint4 a, const uint4 b, int s)
{
    const bool w2 = (s & 8) != 0, w1 = (s & 4) != 0;
    const u32 t0 = w2 ? a.z : a.x, t1 = w2 ? a.w : a.y, t2 = w2 ? b.x : a.z, t3 = w2 ? b.y : a.w, t4 = w2 ? b.z : b.x,
              t5 = w2 ? b.w : b.y;
    const u32 v0 = w1 ? t1 : t0, v1 = w1 ? t2 : t1, v2 = w1 ? t3 : t2, v3 = w1 ? t4 : t3, v4 = w1 ? t5 : t4;
    const unsigned bs = (unsigned)(s & 3) * 8u;
    return make_uint4(__funnelshift_r(v0, v1, bs), __funnelshift_r(v1, v2, bs), __funnelshift_r(v2, v3, bs),
                      __funnelshift_r(v3, v4, bs));
}

__device__ __forceinline__ u64 funnel_r(u64 lo, u64 hi, unsigned sh)   // (hi:lo) >> sh, 0 <= sh < 64
{
    return sh ? ((lo >> sh) | (hi << (64 - sh))) : lo;
}

__host__ __device__ __forceinline__ int dec_digits(unsigned v)
{
    if (v < 100u) return v < 10u ? 1 : 2;               // the common case: short runs
    return v < 1000u ? 3 : v < 10000u ? 4 : v < 100000u ? 5 : v < 1000000u ? 6 : v < 10000000u ? 7
         : v < 100000000u ? 8 : v < 1000000000u ? 9 : 10;
}

#endif  // __CUDACC__

// op codes of the 2-bit packed alignment
constexpr int OP_M = 0, OP_X = 1, OP_I = 2, OP_D = 3;

}  // namespace qb
