// qb_capi.cu — the reference's six public symbols (include/quicked.h) on top of the batched GPU engine, plus the
// seeded dataset generator.  quicked_align() is a batch of one; it is correct but launch-latency bound (a 100 bp
// pair costs the CPU ~6 us, a kernel launch about as much) — throughput users call qb200_align_batch().
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/quicked_b200.h"

namespace {

// One engine context PER CALLING THREAD (reference threading model, SURVEY §8b: aligners are not thread-safe but
// different aligners run concurrently — align_benchmark keeps one per OpenMP thread, align_benchmark.c:246-284).
// A context is created on a thread's first quicked_align, handed back to a process-wide pool when the thread ends and
// reused by later threads; nothing is destroyed at process exit (the CUDA runtime may already be gone by then).
// QUICKED_B200_DEVICE = <n> (default 0) pins every thread to one GPU, "all" deals threads round-robin over the GPUs.
struct CtxPool {
    std::mutex mu;
    std::vector<qb200_ctx_t *> idle;
    int next_dev = 0;
    bool failed = false;
};
CtxPool &ctx_pool() { static CtxPool *p = new CtxPool; return *p; }

struct ThreadCtx {
    qb200_ctx_t *ctx = nullptr;
    ~ThreadCtx()
    {
        if (!ctx) return;
        CtxPool &p = ctx_pool();
        std::lock_guard<std::mutex> lock(p.mu);
        p.idle.push_back(ctx);
    }
};
thread_local ThreadCtx t_ctx;

qb200_ctx_t *thread_context()
{
    if (t_ctx.ctx) return t_ctx.ctx;
    CtxPool &p = ctx_pool();
    std::lock_guard<std::mutex> lock(p.mu);
    if (!p.idle.empty()) { t_ctx.ctx = p.idle.back(); p.idle.pop_back(); return t_ctx.ctx; }
    if (p.failed) return nullptr;
    int dev = 0;
    if (const char *e = getenv("QUICKED_B200_DEVICE")) {
        if (!strcmp(e, "all")) { const int n = qb200_device_count(); dev = n > 0 ? (p.next_dev++ % n) : 0; }
        else dev = atoi(e);
    }
    qb200_ctx_t *c = nullptr;
    if (qb200_create(&c, dev) != 0) { p.failed = true; return nullptr; }
    t_ctx.ctx = c;
    return c;
}

// counter_add semantics of the reference (profiler_counter.c:53-73): total, samples, min, max, running mean/var
void timer_add_sample(profiler_timer_t *t, uint64_t ns)
{
    if (!t) return;
    profiler_counter_t *c = &t->time_ns;
    c->total += ns;
    ++c->samples;
    if (c->samples == 1) {
        c->min = c->max = ns;
        c->m_oldM = c->m_newM = (double)ns;
        c->m_oldS = 0.0;
    } else {
        if (ns < c->min) c->min = ns;
        if (ns > c->max) c->max = ns;
        c->m_newM = c->m_oldM + ((double)ns - c->m_oldM) / (double)c->samples;
        c->m_newS = c->m_oldS + ((double)ns - c->m_oldM) * ((double)ns - c->m_newM);
        c->m_oldM = c->m_newM;
        c->m_oldS = c->m_newS;
    }
}

struct OwnedTimers { profiler_timer_t t[5]; };

}  // namespace

extern "C" {

bool quicked_check_error(quicked_status_t status) { return status < 0; }          // reference quicked.c:380

const char *quicked_status_msg(quicked_status_t status)                            // reference quicked.c:382-403
{
    switch (status) {
    case QUICKED_ERROR: return "ERROR: QuickEd has finished with unspecific error\n";
    case QUICKED_FAIL_NON_CONVERGENCE: return "ERROR: Hirschberg algorithm can not find a middle point of subsequence division!\n";
    case QUICKED_UNIMPLEMENTED: return "ERROR: The algorithm or parameter combination selected is not implemented\n";
    case QUICKED_UNKNOWN_ALGO: return "ERROR: Unknown algorithm selection\n";
    case QUICKED_EMPTY_SEQUENCE: return "ERROR: Tried to align an empty sequence\n";
    default: return "QuickEd finished without errors.\n";
    }
}

quicked_params_t quicked_default_params(void)                                      // reference quicked.c:308-321
{
    quicked_params_t p;
    memset(&p, 0, sizeof p);
    p.algo = QUICKED;
    p.bandwidth = 15;
    p.window_size = 9;
    p.overlap_size = 1;
    p.hew_threshold[0] = p.hew_threshold[1] = 40;
    p.hew_percentage[0] = p.hew_percentage[1] = 15;
    return p;
}

quicked_status_t quicked_new(quicked_aligner_t *aligner, quicked_params_t *params)  // reference quicked.c:323-352
{
    aligner->params = params;          // pointer, not a copy: callers mutate params after new (bindings/cpp/quicked.hpp:54-59)
    aligner->score = -1;
    aligner->cigar = NULL;
    aligner->mm_allocator = params->external_allocator;   // never dereferenced by this library
    if (!params->external_timer) {
        OwnedTimers *ot = (OwnedTimers *)calloc(1, sizeof(OwnedTimers));
        aligner->timer = &ot->t[0];
        aligner->timer_windowed_s = &ot->t[1];
        aligner->timer_windowed_l = &ot->t[2];
        aligner->timer_banded = &ot->t[3];
        aligner->timer_align = &ot->t[4];
    }
    return QUICKED_WIP;
}

quicked_status_t quicked_free(quicked_aligner_t *aligner)                          // reference quicked.c:354-378
{
    if (aligner->cigar) { free(aligner->cigar); aligner->cigar = NULL; }
    if (!aligner->params->external_timer && aligner->timer) {
        free(aligner->timer);          // the five owned timers are one allocation
        aligner->timer = aligner->timer_windowed_s = aligner->timer_windowed_l = aligner->timer_banded = aligner->timer_align = NULL;
    }
    return QUICKED_WIP;
}

quicked_status_t quicked_align(quicked_aligner_t *aligner, const char *pattern, const int pattern_len,
                               const char *text, const int text_len)               // reference quicked.c:405-437
{
    if (pattern_len == 0 || text_len == 0) return QUICKED_EMPTY_SEQUENCE;
    const quicked_params_t *prm = aligner->params;
    if (prm->algo != QUICKED && prm->algo != BANDED && prm->algo != WINDOWED && prm->algo != HIRSCHBERG)
        return QUICKED_UNKNOWN_ALGO;
    qb200_ctx_t *g_ctx = thread_context();
    if (!g_ctx) {
        fprintf(stderr, "quicked_b200: no CUDA device available; this library has no CPU fallback\n");
        return QUICKED_ERROR;
    }
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<char> seqs((size_t)pattern_len + (size_t)text_len + 1);
    memcpy(seqs.data(), pattern, (size_t)pattern_len);
    memcpy(seqs.data() + pattern_len, text, (size_t)text_len);
    int64_t poff = 0, toff = pattern_len;
    int32_t plen = pattern_len, tlen = text_len;
    qb200_batch_t b = {seqs.data(), (int64_t)pattern_len + text_len, 1, &poff, &plen, &toff, &tlen};
    int32_t score = -1, status = QUICKED_ERROR;
    int64_t off[2] = {0, 0};
    std::vector<char> cig((size_t)2 * ((size_t)pattern_len + text_len) + 16);     // quicked.c:48 bound
    qb200_results_t r = {&score, &status, cig.data(), (int64_t)cig.size(), off, 0};
    const int rc = qb200_align_batch(g_ctx, prm, &b, &r);
    if (rc != 0) {
        fprintf(stderr, "quicked_b200: %s (rc=%d)\n", qb200_last_error(g_ctx), rc);
        return QUICKED_ERROR;
    }
    aligner->score = score;
    if (!prm->only_score && status >= 0 && off[1] - off[0] > 1) {
        if (aligner->cigar) free(aligner->cigar);        // the reference leaks the previous string (quicked.c:46-51)
        aligner->cigar = (char *)malloc((size_t)(off[1] - off[0]));
        memcpy(aligner->cigar, cig.data() + off[0], (size_t)(off[1] - off[0]));
    }
    qb200_stats_t st;
    qb200_get_stats(g_ctx, &st);
    const uint64_t wall = (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
    timer_add_sample(aligner->timer, wall);
    if (prm->algo == QUICKED) {
        timer_add_sample(aligner->timer_windowed_s, (uint64_t)(st.ms_windowed_s * 1e6));
        if (st.pairs_stage2) timer_add_sample(aligner->timer_windowed_l, (uint64_t)(st.ms_windowed_l * 1e6));
        for (int64_t i = 0; i < st.banded_tries; ++i) timer_add_sample(aligner->timer_banded, (uint64_t)(st.ms_banded * 1e6 / (double)st.banded_tries));
        timer_add_sample(aligner->timer_align, (uint64_t)((st.ms_align_fill + st.ms_align_trace + st.ms_cigar) * 1e6));
    }
    return (quicked_status_t)status;
}

// ---- seeded generate_dataset twin (reference generate_dataset.c:52-63, :108-199, :366-410) ----
static inline uint64_t splitmix64(uint64_t &x)
{
    uint64_t z = (x += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
static inline uint32_t rand_below(uint64_t &s, uint32_t n) { return (uint32_t)(((splitmix64(s) >> 32) * (uint64_t)n) >> 32); }

// SAM-style CIGAR (reference cigar.c:193-240): the reference groups the one-character-per-op buffer after mapping
// X -> M (unless show_mismatches), but never maps the FIRST operation; matches print as '=' with show_mismatches.
int64_t qb200_cigar_to_sam(const char *cigar, int show_mismatches, char *out, int64_t capacity)
{
    if (!cigar || (!out && capacity > 0)) return 0;
    std::string res;
    char last_op = 0;
    long long last_len = 0;
    bool first = true;
    auto dump = [&]() {
        if (!last_len) return;
        char c = last_op;
        if (show_mismatches && c == 'M') c = '=';
        else if (c != 'M' && c != 'I' && c != 'D' && c != 'N' && c != '=' && c != 'X') c = '?';
        res += std::to_string(last_len);
        res += c;
    };
    auto feed = [&](char op, long long len) {            // len operations `op`, already mapped
        if (!len) return;
        if (op == last_op) { last_len += len; return; }
        dump();
        last_op = op; last_len = len;
    };
    for (const char *p = cigar; *p;) {
        long long len = 0;
        while (*p >= '0' && *p <= '9') { len = len * 10 + (*p - '0'); ++p; }
        if (!*p) break;
        const char op = *p++;
        if (len <= 0) continue;
        const char mapped = (!show_mismatches && op == 'X') ? 'M' : op;
        if (first) {                                      // operations[begin_offset] is used unmapped (cigar.c:209)
            first = false;
            last_op = op; last_len = 1;
            feed(mapped, len - 1);
        } else feed(mapped, len);
    }
    dump();
    const int64_t need = (int64_t)res.size() + 1;
    if (need > capacity) return -need;
    memcpy(out, res.c_str(), (size_t)need);
    return need - 1;
}

// Pairs [first_pair, first_pair + n_pairs) of the job `seed`: every pair has its own random stream (hash of seed and
// pair index), so any rank can generate any slice of the same job.  indels_num / indels_len: the reference's
// `--indels N,LEN` (generate_dataset.c:204-245): a uniform count in [0, N] of LEN-long deletions at uniform positions.
int64_t qb200_generate_pairs_ex(uint64_t seed, int64_t first_pair, int64_t n_pairs, int32_t length, double error, int32_t indels_num,
                                int32_t indels_len, char *seqs, int64_t *pattern_off, int32_t *pattern_len, int64_t *text_off,
                                int32_t *text_len)
{
    if (n_pairs < 0 || first_pair < 0 || length <= 0 || !seqs || indels_num < 0 || indels_len < 0) return QB200_ERR_ARG;
    const int num_errors = error >= 1.0 ? (int)error : (int)std::ceil((double)((float)length * (float)error));   // :370
    const int64_t stride = 2 * (int64_t)length + num_errors + 2;     // fixed slot per pair: pattern then text
    const char alphabet[4] = {'A', 'C', 'G', 'T'};
    unsigned nt = std::max(1u, std::thread::hardware_concurrency());
    if ((int64_t)nt > n_pairs) nt = (unsigned)std::max<int64_t>(1, n_pairs);
    std::vector<std::thread> th;
    for (unsigned w = 0; w < nt; ++w) {
        th.emplace_back([=]() {
            for (int64_t i = w; i < n_pairs; i += nt) {
                uint64_t s = seed ^ 0x5851f42d4c957f2dull;            // independent stream per pair: hash (seed, index in the job)
                s = splitmix64(s) ^ ((uint64_t)(first_pair + i) * 0xd6e8feb86659fd93ull);
                s = splitmix64(s);
                char *pat = seqs + i * stride, *txt = pat + length + num_errors + 1;
                for (int k = 0; k < length; ++k) txt[k] = alphabet[rand_below(s, 4)];
                memcpy(pat, txt, (size_t)length);
                int len = length;
                for (int e = 0; e < num_errors; ++e) {
                    const uint32_t kind = rand_below(s, 3);
                    if (kind == 0 && len > 0) {                        // mismatch (:108-129)
                        const uint32_t pos = rand_below(s, (uint32_t)len);
                        char c;
                        do { c = alphabet[rand_below(s, 4)]; } while (c == pat[pos]);
                        pat[pos] = c;
                    } else if (kind == 1 && len > 1) {                 // deletion (:131-149)
                        const uint32_t pos = rand_below(s, (uint32_t)len);
                        memmove(pat + pos, pat + pos + 1, (size_t)(len - 1 - (int)pos));
                        --len;
                    } else {                                           // insertion (:151-172)
                        const uint32_t pos = rand_below(s, (uint32_t)std::max(len, 1));
                        memmove(pat + pos + 1, pat + pos, (size_t)(len - (int)pos));
                        pat[pos] = alphabet[rand_below(s, 4)];
                        ++len;
                    }
                }
                if (indels_num > 0 && indels_len > 0) {                // large deletions (:204-245)
                    const uint32_t cnt = rand_below(s, (uint32_t)indels_num + 1);
                    for (uint32_t d = 0; d < cnt; ++d) {
                        const uint32_t pos = rand_below(s, (uint32_t)std::max(len, 1));
                        if (indels_len >= len) continue;
                        const int nl = len - indels_len;
                        if ((int)pos < nl) memmove(pat + pos, pat + pos + indels_len, (size_t)(nl - (int)pos));
                        len = nl;
                    }
                }
                pat[len] = 0;
                txt[length] = 0;
                pattern_off[i] = i * stride; pattern_len[i] = len;
                text_off[i] = i * stride + length + num_errors + 1; text_len[i] = length;
            }
        });
    }
    for (auto &t : th) t.join();
    return n_pairs * stride;
}

int64_t qb200_generate_pairs(uint64_t seed, int64_t n_pairs, int32_t length, double error, char *seqs,
                             int64_t *pattern_off, int32_t *pattern_len, int64_t *text_off, int32_t *text_len)
{
    return qb200_generate_pairs_ex(seed, 0, n_pairs, length, error, 0, 0, seqs, pattern_off, pattern_len, text_off, text_len);
}

}  // extern "C"
