// qb_generate.cuh — the seeded generate_dataset twin ON THE DEVICE (reference edit model
// tools/generate_dataset/generate_dataset.c:108-199 mismatch / deletion / insertion at uniform positions, :204-245 long
// deletions, :366-410 driver).  Bit-identical to the host generator qb200_generate_pairs_ex (qb_capi.cu): the same
// per-pair splitmix64 stream, the same draws in the same order — so a batch can be born in HBM (2 GB of characters per
// 100 k pairs of 10 kbp never cross PCIe) and still be reproduced on the host for the parity checks.
//
// ONE CTA PER PAIR.  The text is counter-based (draw k of a splitmix64 stream is a pure function of the stream's start):
// every thread draws its own characters.  The edits are a chain — a position is drawn in the CURRENT pattern — so
// thread 0 replays them in order; a mismatch is a byte store, a deletion / insertion moves the pattern's tail, which the
// whole CTA does in shared memory, 512 words per step (funnel shifts, the way memmove would go through registers).
#pragma once
#include "qb_common.cuh"

namespace qb {

struct GenParams {
    u64 seed;
    i64 first_pair, stride;       // pair i of this call is pair first_pair + i of the job; its slot is raw + i * stride
    int length, num_errors, indels_num, indels_len;
};

__host__ __device__ __forceinline__ u64 gen_mix(u64 z)
{
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
constexpr u64 kGenGamma = 0x9e3779b97f4a7c15ull;
__device__ __forceinline__ u32 gen_below(u64 &s, u32 n) { s += kGenGamma; return (u32)(((gen_mix(s) >> 32) * (u64)n) >> 32); }

constexpr int kGenThreads = 128, kGenWordsPerThread = 4, kGenChunk = kGenThreads * kGenWordsPerThread;

// bytes [pos + D, len) of the word array move down to [pos, len - D) (memmove towards lower addresses), all threads
__device__ __forceinline__ void gen_shift_down(u32 *pw, int pos, int len, int D)
{
    const int w0 = pos >> 2, wl = (len - D - 1) >> 2;               // words that receive data
    if (len - D <= pos) return;
    const int q = D >> 2;
    const unsigned r8 = 8u * (unsigned)(D & 3);
    const u32 keep = (pos & 3) ? ((1u << (8 * (pos & 3))) - 1u) : 0u;  // bytes of word w0 below pos stay
    for (int base = w0; base <= wl; base += kGenChunk) {
        u32 v[kGenWordsPerThread];
#pragma unroll
        for (int k = 0; k < kGenWordsPerThread; ++k) {
            const int wi = base + (int)threadIdx.x * kGenWordsPerThread + k;
            v[k] = 0;
            if (wi <= wl) {
                const u32 a = pw[wi + q], b = pw[wi + q + 1];
                u32 x = r8 ? __funnelshift_r(a, b, r8) : a;
                if (wi == w0) x = (pw[wi] & keep) | (x & ~keep);
                v[k] = x;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kGenWordsPerThread; ++k) {
            const int wi = base + (int)threadIdx.x * kGenWordsPerThread + k;
            if (wi <= wl) pw[wi] = v[k];
        }
        __syncthreads();
    }
}
// bytes [pos, len) move up by one (memmove towards higher addresses); byte pos keeps its old value (the caller stores
// the inserted character there)
__device__ __forceinline__ void gen_shift_up1(u32 *pw, int pos, int len)
{
    if (len <= pos) return;
    const int w0 = pos >> 2, wl = len >> 2;                          // the last word that receives data holds byte `len`
    const u32 keep = (1u << (8 * (pos & 3))) - 1u;                   // bytes of word w0 below pos stay (0 when pos is word-aligned)
    for (int top = wl; top >= w0; top -= kGenChunk) {
        u32 v[kGenWordsPerThread];
#pragma unroll
        for (int k = 0; k < kGenWordsPerThread; ++k) {
            const int wi = top - ((int)threadIdx.x * kGenWordsPerThread + k);
            v[k] = 0;
            if (wi >= w0) {
                const u32 lo = wi > 0 ? pw[wi - 1] : 0u, hi = pw[wi];
                u32 x = __funnelshift_l(lo, hi, 8);
                if (wi == w0) x = (hi & keep) | (((hi & ~keep) << 8) & ~keep) | (hi & (0xffu << (8 * (pos & 3))));
                v[k] = x;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kGenWordsPerThread; ++k) {
            const int wi = top - ((int)threadIdx.x * kGenWordsPerThread + k);
            if (wi >= w0) pw[wi] = v[k];
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kGenThreads)
k_generate_pairs(GenParams g, int n_pairs, unsigned char *__restrict__ raw, int *__restrict__ pattern_len)
{
    extern __shared__ __align__(16) u32 gen_smem[];
    u32 *pw = gen_smem + 4;                                          // pattern bytes; gen_smem[0..3]: op, pos, len, more
    volatile int *ctl = reinterpret_cast<volatile int *>(gen_smem);
    unsigned char *pb = reinterpret_cast<unsigned char *>(pw);
    const int t = threadIdx.x;
    const int cap_words = (g.length + g.num_errors + 16) / 4;
    for (int i = blockIdx.x; i < n_pairs; i += gridDim.x) {
        u64 s = g.seed ^ 0x5851f42d4c957f2dull;                      // the pair's stream: hash (seed, index in the job)
        s = gen_mix(s + kGenGamma) ^ ((u64)(g.first_pair + i) * 0xd6e8feb86659fd93ull);
        s = gen_mix(s + kGenGamma);
        unsigned char *pat = raw + (i64)i * g.stride, *txt = pat + g.length + g.num_errors + 1;
        for (int w = t; w < cap_words; w += kGenThreads) pw[w] = 0;
        __syncthreads();
        for (int k = t; k < g.length; k += kGenThreads) {            // text: draw k is a function of s and k
            const u64 z = gen_mix(s + (u64)(k + 1) * kGenGamma);
            const unsigned char c = (unsigned char)"ACGT"[(u32)(((z >> 32) * 4ull) >> 32)];
            txt[k] = c; pb[k] = c;
        }
        s += (u64)g.length * kGenGamma;
        __syncthreads();
        int len = g.length;
        // ---- the edits: thread 0 draws, everybody moves ----
        int e = 0, d = 0, cnt = -1;                                   // edits done, long deletions done / to do (thread 0)
        for (;;) {
            if (t == 0) {
                int op = 0, pos = 0;                                  // op 0: finished, 1: delete one, 2: insert one, 3: long deletion
                while (e < g.num_errors) {
                    const u32 kind = gen_below(s, 3);
                    ++e;
                    if (kind == 0 && len > 0) {                       // mismatch (:108-129)
                        const u32 p = gen_below(s, (u32)len);
                        unsigned char c;
                        do { c = (unsigned char)"ACGT"[gen_below(s, 4)]; } while (c == pb[p]);
                        pb[p] = c;
                    } else if (kind == 1 && len > 1) {                // deletion (:131-149)
                        pos = (int)gen_below(s, (u32)len); op = 1;
                        break;
                    } else {                                          // insertion (:151-172)
                        pos = (int)gen_below(s, (u32)max(len, 1)); op = 2;
                        break;
                    }
                }
                if (!op && e >= g.num_errors && g.indels_num > 0 && g.indels_len > 0) {      // long deletions (:204-245)
                    if (cnt < 0) cnt = (int)gen_below(s, (u32)g.indels_num + 1);
                    while (d < cnt) {
                        const int p = (int)gen_below(s, (u32)max(len, 1));
                        ++d;
                        if (g.indels_len >= len) continue;
                        pos = p; op = 3;
                        break;
                    }
                }
                ctl[0] = op; ctl[1] = pos; ctl[2] = len;
            }
            __syncthreads();
            const int op = ctl[0], pos = ctl[1];
            len = ctl[2];
            __syncthreads();
            if (op == 0) break;
            if (op == 1) { gen_shift_down(pw, pos, len, 1); --len; }
            else if (op == 2) {
                gen_shift_up1(pw, pos, len);
                if (t == 0) pb[pos] = (unsigned char)"ACGT"[gen_below(s, 4)];
                ++len;
            } else {
                const int nl = len - g.indels_len;
                if (pos < nl) gen_shift_down(pw, pos, len, g.indels_len);
                len = nl;
            }
            __syncthreads();
        }
        for (int k = t; k < len; k += kGenThreads) pat[k] = pb[k];
        if (t == 0) { pat[len] = 0; txt[g.length] = 0; pattern_len[i] = len; }
        __syncthreads();
    }
}

}  // namespace qb
