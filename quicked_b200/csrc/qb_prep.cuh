// qb_prep.cuh — input preparation kernels: ASCII -> base codes, and PEQ (pattern match-mask) tables.
//
// Replaces the per-call CPU work of the reference's pattern compilers
// (reference bpm_banded.c:40-103 == bpm_windowed.c:41-122) and dna_encode (dna_text.c:41-46).
#pragma once
#include "qb_common.cuh"

namespace qb {

// Elementwise: codes[i] = enc_stored(raw[i]) (base code, bit 3 = not a plain ACGTN character).  16 bytes per thread,
// fully coalesced; the 256-entry table sits in shared memory (the arithmetic form costs ~25 integer ops per byte
// and made this kernel ALU-bound at 4x its HBM time).  HBM-bound: 2 B/char.
__global__ void __launch_bounds__(256) k_encode(const uint4 *__restrict__ raw, uint4 *__restrict__ codes, i64 n_vec)
{
    __shared__ unsigned char lut[256];
    lut[threadIdx.x] = (unsigned char)enc_stored(threadIdx.x);
    __syncthreads();
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
        const uint4 r = raw[i];
        u32 in[4] = {r.x, r.y, r.z, r.w}, out[4];
#pragma unroll
        for (int w = 0; w < 4; ++w)
            out[w] = (u32)lut[in[w] & 0xffu] | ((u32)lut[(in[w] >> 8) & 0xffu] << 8) | ((u32)lut[(in[w] >> 16) & 0xffu] << 16) |
                     ((u32)lut[in[w] >> 24] << 24);
        codes[i] = make_uint4(out[0], out[1], out[2], out[3]);
    }
}

// Descriptor of one table to build.
struct PeqJob {
    i64 src_off;   // first code byte of the (sub-)pattern, forward coordinates
    int m;         // (sub-)pattern length
    int rev;       // 1: table of the reversed (sub-)pattern
    i64 peq_off;   // destination, u64 index; layout [nbp][kPeqStride], nbp = ceil(m/64)+2
    i64 t_off = 0; // when flag >= 0: the pair's text, scanned for characters outside "ACGTN"
    int n = 0;
    int flag = -1; // index into the per-pair "odd characters" flags, -1: none wanted
};

// The forward-pattern job of every pair of a batch, derived on the device from its PairRec (nothing to upload).
__global__ void __launch_bounds__(256) k_make_peqjobs(const PairRec *__restrict__ pairs, int n, PeqJob *__restrict__ jobs)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const PairRec r = pairs[i];
    PeqJob j;
    j.src_off = r.p_off; j.m = r.m; j.rev = 0; j.peq_off = r.peq_off; j.t_off = r.t_off; j.n = r.n; j.flag = i;
    jobs[i] = j;
}

// One warp per table.  Each iteration covers one 64-row block: two coalesced 32-byte reads of codes, five ballots
// each.  Rows >= m inside the last block match every code (reference bpm_banded.c:77-86); the two extra blocks are 0.
// With job.flag >= 0 the warp also reports whether the pattern holds a character outside "ACGTN" or the text one outside
// "ACGT" (the WindowEd(S) kernel prices diagonal steps from the match masks only for pairs without any, and its compact
// variant keeps no match-mask row for code 4).
__global__ void __launch_bounds__(256) k_build_peq(const PeqJob *__restrict__ jobs, int n_jobs,
                                                   const unsigned char *__restrict__ codes, u64 *__restrict__ peq,
                                                   unsigned char *__restrict__ odd_flags)
{
    const int warp = (int)(((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (warp >= n_jobs) return;
    const PeqJob job = jobs[warp];
    if (job.m <= 0 || (job.flag >= 0 && job.n <= 0)) {        // a pair with an empty side has no table (and no space for one)
        if (job.flag >= 0 && lane == 0) odd_flags[job.flag] = 1;
        return;
    }
    const int nblk = (job.m + 63) >> 6, nbp = nblk + 2;
    u64 *dst = peq + job.peq_off;
    u32 any_odd = 0;
    for (int blk = 0; blk < nbp; ++blk) {
        u32 lo[kPeqStride], hi[kPeqStride];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int row = blk * 64 + half * 32 + lane;
            int code = -1;                       // -1: beyond the padded pattern (no match), 5: padding row (all match)
            bool odd = false;                    // row holds a character outside "ACGTN"
            if (row < job.m) {
                const unsigned sc = codes[job.src_off + (job.rev ? (job.m - 1 - row) : row)];
                code = (int)(sc & 7u); odd = (sc & kCodeOdd) != 0;
            } else if (blk < nblk) code = 5;
#pragma unroll
            for (int c = 0; c < kAlpha; ++c) {
                const u32 b = __ballot_sync(kFull, code == c || code == 5);
                if (half == 0) lo[c] = b; else hi[c] = b;
            }
            const u32 b = __ballot_sync(kFull, odd);
            if (half == 0) lo[kAlpha] = b; else hi[kAlpha] = b;
            any_odd |= b;
        }
        if (lane < kPeqStride) {
            u32 l = 0, h = 0;                    // lane 5: the mask of rows with an odd character
#pragma unroll
            for (int c = 0; c < kPeqStride; ++c) if (lane == c) { l = lo[c]; h = hi[c]; }
            dst[(i64)blk * kPeqStride + lane] = ((u64)h << 32) | l;
        }
    }
    if (job.flag >= 0) {
        const unsigned long long t0 = (unsigned long long)(codes + job.t_off), t1 = t0 + (unsigned long long)job.n, a0 = t0 & ~3ull;
        u32 acc = 0;
        for (unsigned long long a = a0 + 4ull * lane; a < t1; a += 128) {
            u32 msk = 0x0c0c0c0cu;                                   // kCodeOdd or code 4 (not A, C, G, T) in the bytes that belong to the text
            if (a < t0) msk <<= 8 * (unsigned)(t0 - a);
            if (a + 4 > t1) msk &= 0x0c0c0c0cu >> (8 * (unsigned)(a + 4 - t1));
            acc |= *reinterpret_cast<const u32 *>(a) & msk;
        }
        any_odd |= __ballot_sync(kFull, acc != 0);
        if (lane == 0) odd_flags[job.flag] = any_odd ? 1 : 0;
    }
}

// k_build_peq_pairs: the forward-pattern tables of a BIG batch, ONE PATTERN PER THREAD (the warp-per-pattern kernel
// above spends ~67 thread-instructions per pattern row on ballots and address arithmetic and is ALU-bound: 2.8 ms per
// 1 M patterns of 1 kbp).  A thread reads its pattern 8 code bytes at a time (aligned loads, realigned with a funnel
// shift), peels the three code bit planes and the odd-character plane off all 8 bytes with one multiply each
// (movemask: ((x >> k) & 0x01..01) * 0x0102040810204080 >> 56), and turns the four 64-row planes of a block into the
// six masks with a few 64-bit logic ops: ~6 instructions per row.  The text is scanned for odd characters on the way.
__device__ __forceinline__ u32 movemask8(u64 x, int k)     // bit k of each of the 8 bytes of x -> 8 bits
{
    return (u32)((((x >> k) & 0x0101010101010101ull) * 0x0102040810204080ull) >> 56);
}

// G = 1: one pattern per thread (reads up to a few kbp: enough patterns to fill the GPU).  G = 32: one pattern per warp,
// lane l takes blocks l, l + 32, ... and a 32nd of the text scan — 100 k patterns of 10 kbp are only 100 k threads
// otherwise (0.91 ms, 30 % issue slots; the blocks of a pattern are independent).
template <int G>
__global__ void __launch_bounds__(128) k_build_peq_pairs(const PairRec *__restrict__ pairs, int n, const unsigned char *__restrict__ codes,
                                                         u64 *__restrict__ peq, unsigned char *__restrict__ odd_flags)
{
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = gt / G, gl = gt % G;
    if (i >= n) return;
    const PairRec r = pairs[i];
    if (r.m <= 0 || r.n <= 0) { if (gl == 0) odd_flags[i] = 1; return; }       // no table (and no space for one)
    const int nblk = (r.m + 63) >> 6, nbp = nblk + 2;
    ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(peq + r.peq_off);
    const unsigned long long p0 = (unsigned long long)(codes + r.p_off);
    const u64 *src = reinterpret_cast<const u64 *>(p0 & ~7ull);
    const unsigned sh = 8u * (unsigned)(p0 & 7ull);
    u64 prev = (G == 1) ? __ldg(src) : 0ull;
    u64 any_odd = 0;
    for (int blk = gl; blk < nblk; blk += G) {
        if (G > 1) prev = __ldg(src + blk * 8);
        u64 b0 = 0, b1 = 0, b2 = 0, od = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const u64 next = __ldg(src + blk * 8 + q + 1);       // reads < 16 B past the pattern: inside the padded buffer
            const u64 w = sh ? ((prev >> sh) | (next << (64u - sh))) : prev;
            prev = next;
            b0 |= (u64)movemask8(w, 0) << (8 * q);
            b1 |= (u64)movemask8(w, 1) << (8 * q);
            b2 |= (u64)movemask8(w, 2) << (8 * q);
            od |= (u64)movemask8(w, 3) << (8 * q);
        }
        const int rows = r.m - blk * 64;                          // rows of the pattern in this block (>= 1)
        const u64 valid = rows >= 64 ? ~0ull : ((1ull << rows) - 1ull), pad = ~valid;   // rows >= m match every code
        const u64 lo = ~b2 & valid;
        const u64 e0 = (lo & ~b1 & ~b0) | pad, e1 = (lo & ~b1 & b0) | pad, e2 = (lo & b1 & ~b0) | pad, e3 = (lo & b1 & b0) | pad,
                  e4 = (b2 & valid) | pad;
        od &= valid;
        any_odd |= od;
        dst[blk * 3 + 0] = make_ulonglong2(e0, e1);
        dst[blk * 3 + 1] = make_ulonglong2(e2, e3);
        dst[blk * 3 + 2] = make_ulonglong2(e4, od);
    }
    if (gl == 0) {
#pragma unroll
        for (int k = 0; k < 6; ++k) dst[nblk * 3 + k] = make_ulonglong2(0, 0);   // the two blocks past the pattern
    }
    (void)nbp;
    // odd characters of the text: aligned 8-byte words, bytes outside the text masked off
    const unsigned long long t0 = (unsigned long long)(codes + r.t_off), t1 = t0 + (unsigned long long)r.n;
    for (unsigned long long a = (t0 & ~7ull) + 8ull * (unsigned)gl; a < t1; a += 8ull * G) {
        u64 msk = 0x0c0c0c0c0c0c0c0cull;                              // kCodeOdd, or code 4: a text character that is not A, C, G, T
        if (a < t0) msk <<= 8 * (unsigned)(t0 - a);
        if (a + 8 > t1) msk &= 0x0c0c0c0c0c0c0c0cull >> (8 * (unsigned)(a + 8 - t1));
        any_odd |= __ldg(reinterpret_cast<const u64 *>(a)) & msk;
    }
    if (G == 1) odd_flags[i] = any_odd ? 1 : 0;
    else {
        const unsigned any = __ballot_sync(kFull, any_odd != 0);   // G = 32: a pattern is a warp (blockDim is a multiple of 32)
        if (gl == 0) odd_flags[i] = any ? 1 : 0;
    }
}

}  // namespace qb
