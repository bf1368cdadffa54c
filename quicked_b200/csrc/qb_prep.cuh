// qb_prep.cuh — input preparation kernels: ASCII -> base codes, and PEQ (pattern match-mask) tables.
//
// Replaces the per-call CPU work of the reference's pattern compilers
// (reference bpm_banded.c:40-103 == bpm_windowed.c:41-122) and dna_encode (dna_text.c:41-46).
#pragma once
#include "qb_common.cuh"

namespace qb {

// Elementwise: codes[i] = enc(raw[i]).  16 bytes per thread, fully coalesced.  HBM-bound: 2 B/char.
__global__ void __launch_bounds__(256) k_encode(const uint4 *__restrict__ raw, uint4 *__restrict__ codes, i64 n_vec)
{
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
        const uint4 r = raw[i];
        u32 in[4] = {r.x, r.y, r.z, r.w}, out[4];
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            u32 o = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) o |= (u32)enc_base((in[w] >> (8 * b)) & 0xffu) << (8 * b);
            out[w] = o;
        }
        codes[i] = make_uint4(out[0], out[1], out[2], out[3]);
    }
}

// Descriptor of one table to build.
struct PeqJob {
    i64 src_off;   // first code byte of the (sub-)pattern, forward coordinates
    int m;         // (sub-)pattern length
    int rev;       // 1: table of the reversed (sub-)pattern
    i64 peq_off;   // destination, u64 index; layout [nbp][kPeqStride], nbp = ceil(m/64)+2
};

// One warp per table.  Each iteration covers one 64-row block: two coalesced 32-byte reads of codes, five ballots
// each.  Rows >= m inside the last block match every code (reference bpm_banded.c:77-86); the two extra blocks are 0.
__global__ void __launch_bounds__(256) k_build_peq(const PeqJob *__restrict__ jobs, int n_jobs,
                                                   const unsigned char *__restrict__ codes, u64 *__restrict__ peq)
{
    const int warp = (int)(((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (warp >= n_jobs) return;
    const PeqJob job = jobs[warp];
    const int nblk = (job.m + 63) >> 6, nbp = nblk + 2;
    u64 *dst = peq + job.peq_off;
    for (int blk = 0; blk < nbp; ++blk) {
        u32 lo[kAlpha], hi[kAlpha];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int row = blk * 64 + half * 32 + lane;
            int code = -1;                       // -1: beyond the padded pattern (no match), 5: padding row (all match)
            if (row < job.m) code = codes[job.src_off + (job.rev ? (job.m - 1 - row) : row)];
            else if (blk < nblk) code = 5;
#pragma unroll
            for (int c = 0; c < kAlpha; ++c) {
                const u32 b = __ballot_sync(kFull, code == c || code == 5);
                if (half == 0) lo[c] = b; else hi[c] = b;
            }
        }
        if (lane < kPeqStride) {
            u32 l = 0, h = 0;                    // lane 5: the pad word
#pragma unroll
            for (int c = 0; c < kAlpha; ++c) if (lane == c) { l = lo[c]; h = hi[c]; }
            dst[(i64)blk * kPeqStride + lane] = ((u64)h << 32) | l;
        }
    }
}

}  // namespace qb
