// qb_align_benchmark — batch CLI on the GPU library with the reference tool's interface.
//
// Same flags, same `.seq` input and same output as the reference's align_benchmark for the algorithms on the path
// (reference tools/align_benchmark/align_benchmark_params.c:105-313, align_benchmark.c:73-99,
//  benchmark/benchmark_utils.c:151-170), so A/B runs and parity diffs are one command:
//
//   align_benchmark    -a quicked -i in.seq -o ref.out
//   qb_align_benchmark -a quicked -i in.seq -o gpu.out       &&  cmp ref.out gpu.out
//
// The reference's OpenMP batch loop (align_benchmark.c:232-306) is replaced by one qb200_align_batch() per batch.
// --check correct replays every CIGAR on its pair (cigar_check_alignment semantics, cigar.c:363-434).
//
// Streaming: a reader thread parses batch k+1 into page-locked buffers (`.seq` pairs or FASTA
// records, consecutive records = pattern, text) while the GPU aligns batch k and a writer thread formats batch k-1;
// --output-sam writes one SAM line per pair (query = text, reference = pattern, CIGAR through qb200_cigar_to_sam,
// the reference's cigar_sprint_SAM_CIGAR, cigar.c:504-529).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <getopt.h>
#include <strings.h>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "quicked_b200.h"

struct Options {
    std::string algo = "quicked", input, output, output_sam, input_format = "auto";
    bool output_full = false, force_scalar = false, check = false, check_score = false, only_score = false, sam_eqx = false;
    int bandwidth = 15, window_size = 9, overlap_size = 1, hew_threshold = 40, hew_percentage = 15;
    long batch_size = 1000000;
    int device = 0, verbose = 0;
};

static void usage()
{
    fprintf(stderr,
            "USE: ./qb_align_benchmark -a ALGORITHM -i PATH\n"
            "      Options::\n"
            "        [Algorithm]\n"
            "          --algorithm|a ALGORITHM   quicked | edit-banded | edit-windowed | edit-banded-hirschberg\n"
            "        [Input & Output]\n"
            "          --input|i PATH            '>pattern' / '<text' line pairs, or FASTA (records 2i, 2i+1 = pattern, text)\n"
            "          --input-format FMT        auto | seq | fasta\n"
            "          --output|o PATH           score<TAB>CIGAR per pair\n"
            "          --output-full PATH        lengths, score, sequences, CIGAR\n"
            "          --output-sam PATH         one SAM line per pair (query = text, reference = pattern); --sam-eqx: =/X ops\n"
            "        [Other Parameters]\n"
            "          --bandwidth INT  --window-size INT  --overlap-size INT\n"
            "          --hew-threshold INT  --hew-percentage INT  --force-scalar  --only-score\n"
            "        [Misc]\n"
            "          --check|c correct|score|alignment   replay every CIGAR on its pair; score / alignment: also compare the\n"
            "                                    score with the exact edit distance (the reference asks edlib; here a\n"
            "                                    multi-word bit-parallel checker in this tool: O(m*n/64) per pair)\n"
            "        [System]\n"
            "          --batch-size INT (pairs per GPU batch, default 1000000)   --device INT   --verbose|v\n");
}

static bool replay_ok(const char *cigar, const char *p, int m, const char *t, int n, int score)
{
    long i = 0, j = 0, cost = 0, num = 0;
    for (const char *c = cigar; *c; ++c) {
        if (*c >= '0' && *c <= '9') { num = num * 10 + (*c - '0'); continue; }
        for (long k = 0; k < num; ++k) {
            switch (*c) {
            case 'M': if (i >= m || j >= n || p[i] != t[j]) return false; ++i; ++j; break;
            case 'X': if (i >= m || j >= n || p[i] == t[j]) return false; ++i; ++j; ++cost; break;
            case 'D': if (i >= m) return false; ++i; ++cost; break;
            case 'I': if (j >= n) return false; ++j; ++cost; break;
            default: return false;
            }
        }
        num = 0;
    }
    return i == m && j == n && cost == score;
}

// Growable page-locked byte buffer (contents kept across growth).
struct PinnedBytes {
    char *p = nullptr;
    size_t cap = 0, n = 0;
    void reserve(size_t want)
    {
        if (want <= cap) return;
        size_t nc = std::max(want, cap + cap / 2 + (size_t)(1 << 20));
        char *q = (char *)qb200_host_alloc(nc);
        if (!q) { fprintf(stderr, "qb_align_benchmark: out of page-locked memory\n"); exit(3); }
        if (p) { memcpy(q, p, n); qb200_host_free(p); }
        p = q; cap = nc;
    }
    void append(const char *src, size_t len) { reserve(n + len + 1); memcpy(p + n, src, len); n += len; }
    ~PinnedBytes() { if (p) qb200_host_free(p); }
};

struct Batch {
    PinnedBytes seqs, cig;
    std::vector<int64_t> po, to, coff;
    std::vector<int32_t> pl, tl, score, status;
    std::vector<std::string> pname, tname;      // FASTA record names (empty for .seq input)
    int64_t first_index = 0;
    int state = 0;                              // 0 free, 1 parsed, 2 aligned
    bool last = false;
    void clear() { seqs.n = 0; po.clear(); to.clear(); pl.clear(); tl.clear(); pname.clear(); tname.clear(); last = false; }
    void add(const char *p, long m, const char *t, long n)
    {
        po.push_back((int64_t)seqs.n); pl.push_back((int32_t)std::max(m, 0L)); seqs.append(p, (size_t)std::max(m, 0L));
        to.push_back((int64_t)seqs.n); tl.push_back((int32_t)std::max(n, 0L)); seqs.append(t, (size_t)std::max(n, 0L));
    }
};

// Input: `.seq` ('>pattern' / '<text' lines, align_benchmark.c:73-99) or FASTA (multi-line records; consecutive
// records form a pair).  next() appends one pair to the batch; false at end of input.
struct PairReader {
    FILE *in; bool fasta = false;
    char *l1 = nullptr, *l2 = nullptr; size_t a1 = 0, a2 = 0;
    std::string pending_header, seq_a, seq_b, name_a, name_b;
    PairReader(FILE *f, const std::string &fmt) : in(f)
    {
        if (fmt == "fasta") fasta = true;
        else if (fmt == "auto") {                        // FASTA iff the second line does not start with '<' or '>'
            const long pos = ftell(in);
            ssize_t n1 = getline(&l1, &a1, in), n2 = getline(&l2, &a2, in);
            fasta = n1 > 0 && n2 > 0 && l1[0] == '>' && l2[0] != '<' && l2[0] != '>';
            fseek(in, pos, SEEK_SET);
        }
    }
    ~PairReader() { free(l1); free(l2); }
    static std::string name_of(const char *line) { const char *e = line; while (*e && *e != ' ' && *e != '\t' && *e != '\n' && *e != '\r') ++e; return std::string(line, e); }
    bool fasta_record(std::string &name, std::string &seq)
    {
        seq.clear();
        if (pending_header.empty()) {
            ssize_t n;
            while ((n = getline(&l1, &a1, in)) >= 0) if (l1[0] == '>') { pending_header = name_of(l1 + 1); if (pending_header.empty()) pending_header = "*"; break; }
            if (pending_header.empty()) return false;
        }
        name = pending_header; pending_header.clear();
        ssize_t n;
        while ((n = getline(&l1, &a1, in)) >= 0) {
            if (l1[0] == '>') { pending_header = name_of(l1 + 1); if (pending_header.empty()) pending_header = "*"; break; }
            while (n > 0 && (l1[n - 1] == '\n' || l1[n - 1] == '\r' || l1[n - 1] == ' ')) --n;
            seq.append(l1, (size_t)n);
        }
        return true;
    }
    bool next(Batch &b)
    {
        if (fasta) {
            if (!fasta_record(name_a, seq_a) || !fasta_record(name_b, seq_b)) return false;
            b.add(seq_a.data(), (long)seq_a.size(), seq_b.data(), (long)seq_b.size());
            b.pname.push_back(name_a); b.tname.push_back(name_b);
            return true;
        }
        ssize_t n1 = getline(&l1, &a1, in);
        if (n1 < 0) return false;
        ssize_t n2 = getline(&l2, &a2, in);
        if (n2 < 0) return false;
        while (n1 > 0 && (l1[n1 - 1] == '\n' || l1[n1 - 1] == '\r')) --n1;
        while (n2 > 0 && (l2[n2 - 1] == '\n' || l2[n2 - 1] == '\r')) --n2;
        const char *p = l1 + 1, *t = l2 + 1;
        long m = n1 - 1, n = n2 - 1;
        if (l1[0] == '<' && l2[0] == '>') { std::swap(p, t); std::swap(m, n); }   // generate_dataset.c:396-405 prints either order
        b.add(p, m, t, n);
        return true;
    }
};

// Exact global edit distance of (p, t): Myers' bit-vector recurrence over ceil(m/64) words per text column, unbanded —
// the job edlib does for the reference's `--check score` (benchmark_check.c:117-158).  Raw byte compare.
static long exact_edit_distance(const char *p, long m, const char *t, long n)
{
    if (m == 0) return n;
    if (n == 0) return m;
    const long W = (m + 63) / 64;
    std::vector<uint64_t> peq((size_t)W * 256, 0), pv((size_t)W, ~0ull), mv((size_t)W, 0);
    for (long i = 0; i < m; ++i) peq[(size_t)((unsigned char)p[i]) * W + i / 64] |= 1ull << (i & 63);
    long score = m;
    const int top = (int)((m - 1) & 63);
    for (long j = 0; j < n; ++j) {
        const uint64_t *eqc = &peq[(size_t)((unsigned char)t[j]) * W];
        uint64_t hp_in = 1, hm_in = 0;                   // global alignment: row 0 costs j
        for (long w = 0; w < W; ++w) {
            const uint64_t eq = eqc[w], xv = eq | mv[w], eqh = eq | hm_in;
            const uint64_t xh = (((eqh & pv[w]) + pv[w]) ^ pv[w]) | eqh;
            uint64_t ph = mv[w] | ~(xh | pv[w]), mh = pv[w] & xh;
            if (w == W - 1) score += (long)((ph >> top) & 1) - (long)((mh >> top) & 1);
            const uint64_t hp_out = ph >> 63, hm_out = mh >> 63;
            ph = (ph << 1) | hp_in; mh = (mh << 1) | hm_in;
            pv[w] = mh | ~(xv | ph); mv[w] = ph & xv;
            hp_in = hp_out; hm_in = hm_out;
        }
    }
    return score;
}

int main(int argc, char **argv)
{
    Options o;
    static struct option long_options[] = {
        {"algorithm", required_argument, 0, 'a'}, {"input", required_argument, 0, 'i'}, {"output", required_argument, 0, 'o'},
        {"output-full", required_argument, 0, 800}, {"output-sam", required_argument, 0, 801}, {"sam-eqx", no_argument, 0, 802},
        {"input-format", required_argument, 0, 803}, {"bandwidth", required_argument, 0, 2000}, {"window-size", required_argument, 0, 2001},
        {"overlap-size", required_argument, 0, 2002}, {"hew-threshold", required_argument, 0, 2003}, {"hew-percentage", required_argument, 0, 2004},
        {"force-scalar", no_argument, 0, 2005}, {"only-score", no_argument, 0, 2006}, {"check", required_argument, 0, 'c'},
        {"num-threads", required_argument, 0, 't'}, {"batch-size", required_argument, 0, 4000}, {"device", required_argument, 0, 4002},
        {"progress", required_argument, 0, 'P'}, {"verbose", no_argument, 0, 'v'}, {"help", no_argument, 0, 'h'}, {0, 0, 0, 0}};
    if (argc <= 1) { usage(); return 0; }
    int c, idx;
    while ((c = getopt_long(argc, argv, "a:i:o:P:c:vt:h", long_options, &idx)) != -1) {
        switch (c) {
        case 'a': o.algo = optarg; break;
        case 'i': o.input = optarg; break;
        case 'o': o.output = optarg; break;
        case 800: o.output = optarg; o.output_full = true; break;
        case 801: o.output_sam = optarg; break;
        case 802: o.sam_eqx = true; break;
        case 803: o.input_format = optarg; break;
        case 2000: o.bandwidth = atoi(optarg); break;
        case 2001: o.window_size = atoi(optarg); break;
        case 2002: o.overlap_size = atoi(optarg); break;
        case 2003: o.hew_threshold = atoi(optarg); break;
        case 2004: o.hew_percentage = atoi(optarg); break;
        case 2005: o.force_scalar = true; break;
        case 2006: o.only_score = true; break;
        case 'c':                                        // align_benchmark_params.c:205-227
            if (!strcasecmp(optarg, "correct")) o.check = true;
            else if (!strcasecmp(optarg, "score") || !strcasecmp(optarg, "alignment")) o.check = o.check_score = true;
            else if (!strcasecmp(optarg, "display")) {}
            else { fprintf(stderr, "Option '--check' must be in {'correct','score','alignment'}\n"); return 1; }
            break;
        case 't': case 'P': break;                       // accepted for command-line compatibility
        case 4000: o.batch_size = atol(optarg); break;
        case 4002: o.device = atoi(optarg); break;
        case 'v': o.verbose = 1; break;
        case 'h': usage(); return 1;
        default: fprintf(stderr, "Option not recognized \n"); return 1;
        }
    }
    quicked_params_t prm = quicked_default_params();
    if (o.algo == "quicked") prm.algo = QUICKED;
    else if (o.algo == "edit-banded") prm.algo = BANDED;
    else if (o.algo == "edit-windowed") prm.algo = WINDOWED;
    else if (o.algo == "edit-banded-hirschberg") prm.algo = HIRSCHBERG;
    else { fprintf(stderr, "Algorithm '%s' not recognized\n", o.algo.c_str()); return 1; }
    if (o.input_format != "auto" && o.input_format != "seq" && o.input_format != "fasta") { fprintf(stderr, "Input format '%s' not recognized\n", o.input_format.c_str()); return 1; }
    prm.bandwidth = (unsigned)o.bandwidth; prm.window_size = (unsigned)o.window_size; prm.overlap_size = (unsigned)o.overlap_size;
    prm.hew_threshold[0] = prm.hew_threshold[1] = (unsigned)o.hew_threshold;
    prm.hew_percentage[0] = prm.hew_percentage[1] = (unsigned)o.hew_percentage;
    prm.force_scalar = o.force_scalar; prm.only_score = o.only_score;

    FILE *in = fopen(o.input.c_str(), "r");
    if (!in) { fprintf(stderr, "Input file '%s' couldn't be opened\n", o.input.c_str()); return 1; }
    FILE *out = o.output.empty() ? nullptr : fopen(o.output.c_str(), "w");
    FILE *sam = o.output_sam.empty() ? nullptr : fopen(o.output_sam.c_str(), "w");
    qb200_ctx_t *gpu = nullptr;
    if (qb200_create(&gpu, o.device) != QB200_OK) { fprintf(stderr, "qb_align_benchmark: no CUDA device (there is no CPU fallback)\n"); return 2; }
    if (sam) fprintf(sam, "@HD\tVN:1.6\tSO:unknown\n@PG\tID:qb_align_benchmark\tPN:qb_align_benchmark\tDS:query=text reference=pattern algorithm=%s\n", o.algo.c_str());

    // ---- three stages over a ring of three batches: reader thread -> this thread (GPU) -> writer thread ----
    constexpr int kRing = 3;
    Batch ring[kRing];
    std::mutex mu;
    std::condition_variable cv;
    long total = 0, bad = 0, score_ok = 0, score_total = 0, score_diff = 0, cg_m = 0, cg_x = 0, cg_i = 0, cg_d = 0, bases = 0;
    qb200_stats_t st_acc;
    memset(&st_acc, 0, sizeof st_acc);
    double t_align = 0;
    int fatal = 0;
    const auto t_begin = std::chrono::steady_clock::now();

    std::thread reader([&] {
        PairReader pr(in, o.input_format);
        int64_t index = 0;
        for (int k = 0;; ++k) {
            Batch &b = ring[k % kRing];
            { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return fatal || b.state == 0; }); if (fatal) return; }
            b.clear();
            b.first_index = index;
            bool more = true;
            while ((long)b.po.size() < o.batch_size && (more = pr.next(b))) {}
            b.seqs.append("", 0); b.seqs.p[b.seqs.n++] = 0;
            index += (int64_t)b.po.size();
            b.last = !more;
            { std::lock_guard<std::mutex> lk(mu); b.state = 1; }
            cv.notify_all();
            if (b.last) return;
        }
    });
    std::thread writer([&] {
        std::string samcig;
        for (int k = 0;; ++k) {
            Batch &b = ring[k % kRing];
            { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return fatal || b.state == 2; }); if (fatal) return; }
            const int64_t np = (int64_t)b.po.size();
            for (int64_t i = 0; i < np; ++i) {
                const size_t q = (size_t)i;
                const char *cg = (prm.only_score || b.coff[q + 1] - b.coff[q] <= 1) ? "-" : b.cig.p + b.coff[q];
                const bool err = quicked_check_error((quicked_status_t)b.status[q]);
                const char *P = b.seqs.p + b.po[q], *T = b.seqs.p + b.to[q];
                if (o.check && !err && !prm.only_score && !replay_ok(cg, P, b.pl[q], T, b.tl[q], b.score[q])) ++bad;
                if (o.check && !err && !prm.only_score) {           // CIGAR breakdown (benchmark_check.c:64-74)
                    bases += b.pl[q];
                    for (const char *x = cg; *x;) {
                        long len = 0;
                        while (*x >= '0' && *x <= '9') { len = len * 10 + (*x - '0'); ++x; }
                        if (!*x) break;
                        const char op = *x++;
                        (op == 'M' ? cg_m : op == 'X' ? cg_x : op == 'I' ? cg_i : cg_d) += len;
                    }
                }
                if (o.check_score && !err) {
                    const long exact = exact_edit_distance(P, b.pl[q], T, b.tl[q]);
                    score_total += exact;
                    if (exact == b.score[q]) ++score_ok;
                    else {
                        score_diff += labs(exact - b.score[q]);
                        if (o.verbose) fprintf(stderr, "(#%lld)\t INACCURATE SCORE computed=%d\tcorrect=%ld\n", (long long)(b.first_index + i), b.score[q], exact);
                    }
                }
                if (out) {
                    if (o.output_full) {
                        fprintf(out, "%d\t%d\t", b.pl[q], b.tl[q]);
                        if (err) fprintf(out, "ERROR\t"); else fprintf(out, "%d\t", b.score[q]);
                        fwrite(P, 1, (size_t)b.pl[q], out); fputc('\t', out);
                        fwrite(T, 1, (size_t)b.tl[q], out);
                        fprintf(out, "\t%s\n", err ? (prm.only_score ? "-" : "ERROR") : cg);
                    } else if (err) fprintf(out, "ERROR\t%s\n", prm.only_score ? "-" : "ERROR");
                    else fprintf(out, "%d\t%s\n", b.score[q], cg);
                }
                if (sam) {
                    const long long id = (long long)(b.first_index + i);
                    if (b.tname.empty()) fprintf(sam, "text%lld", id); else fputs(b.tname[q].c_str(), sam);
                    if (err) fputs("\t4\t*\t0\t0\t*", sam);
                    else {
                        if (b.pname.empty()) fprintf(sam, "\t0\tpattern%lld\t1\t255\t", id); else fprintf(sam, "\t0\t%s\t1\t255\t", b.pname[q].c_str());
                        if (prm.only_score || cg[0] == '-') fputc('*', sam);
                        else {
                            samcig.resize(strlen(cg) + 16);
                            int64_t n = qb200_cigar_to_sam(cg, o.sam_eqx, &samcig[0], (int64_t)samcig.size());
                            if (n < 0) { samcig.resize((size_t)-n); n = qb200_cigar_to_sam(cg, o.sam_eqx, &samcig[0], (int64_t)samcig.size()); }
                            fwrite(samcig.data(), 1, (size_t)n, sam);
                        }
                    }
                    fputs("\t*\t0\t0\t", sam);
                    if (b.tl[q]) fwrite(T, 1, (size_t)b.tl[q], sam); else fputc('*', sam);
                    if (err) fputs("\t*\n", sam); else fprintf(sam, "\t*\tNM:i:%d\n", b.score[q]);
                }
            }
            total += np;
            const bool last = b.last;
            { std::lock_guard<std::mutex> lk(mu); b.state = 0; }
            cv.notify_all();
            if (last) return;
        }
    });
    for (int k = 0;; ++k) {
        Batch &b = ring[k % kRing];
        { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return b.state == 1; }); }
        const int64_t np = (int64_t)b.po.size();
        int rc = QB200_OK;
        if (np) {
            b.score.assign((size_t)np, -1); b.status.assign((size_t)np, -1); b.coff.assign((size_t)np + 1, 0);
            b.cig.reserve(std::max<size_t>(b.cig.cap, b.seqs.n / 2 + 1024));
            qb200_batch_t qb = {b.seqs.p, (int64_t)b.seqs.n, np, b.po.data(), b.pl.data(), b.to.data(), b.tl.data()};
            qb200_results_t r = {b.score.data(), b.status.data(), b.cig.p, (int64_t)b.cig.cap, b.coff.data(), 0};
            const auto t0 = std::chrono::steady_clock::now();
            rc = qb200_align_batch(gpu, &prm, &qb, &r);
            if (rc == QB200_ERR_CAPACITY) {
                b.cig.n = 0; b.cig.reserve((size_t)r.cigar_bytes + 1024);
                r.cigar = b.cig.p; r.cigar_capacity = (int64_t)b.cig.cap;
                rc = qb200_align_batch(gpu, &prm, &qb, &r);
            }
            t_align += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            qb200_stats_t st;
            if (qb200_get_stats(gpu, &st) == QB200_OK) {
                st_acc.ms_total += st.ms_total; st_acc.ms_prepare += st.ms_prepare; st_acc.ms_windowed_s += st.ms_windowed_s; st_acc.ms_windowed_l += st.ms_windowed_l;
                st_acc.ms_banded += st.ms_banded; st_acc.ms_align_fill += st.ms_align_fill; st_acc.ms_align_trace += st.ms_align_trace;
                st_acc.ms_cigar += st.ms_cigar; st_acc.ms_fused += st.ms_fused; st_acc.pairs_stage2 += st.pairs_stage2; st_acc.pairs_stage3 += st.pairs_stage3;
                st_acc.hirschberg_splits += st.hirschberg_splits; st_acc.kernel_launches += st.kernel_launches; st_acc.word_steps += st.word_steps;
                st_acc.h2d_bytes += st.h2d_bytes; st_acc.d2h_bytes += st.d2h_bytes;
            }
        }
        if (rc != QB200_OK) {
            fprintf(stderr, "qb_align_benchmark: %s (rc=%d)\n", qb200_last_error(gpu), rc);
            { std::lock_guard<std::mutex> lk(mu); fatal = 3; }
            cv.notify_all();
            break;
        }
        const bool last = b.last;
        { std::lock_guard<std::mutex> lk(mu); b.state = 2; }
        cv.notify_all();
        if (last) break;
    }
    reader.join(); writer.join();
    if (fatal) return fatal;
    const double t_total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
    fprintf(stderr, "...processed %ld reads (alignment = %2.3f seq/s)\n", total, t_align > 0 ? total / t_align : 0.0);
    fprintf(stderr, "[Benchmark]\n=> Total.reads            %ld\n=> Time.Benchmark      %9.2f s\n  => Time.Alignment    %9.2f s\n", total, t_total, t_align);
    if (o.verbose) {
        // the reference's QUICKED stage timers (align_benchmark.c:120-129) as the GPU stage times (CUDA events, summed over
        // the sub-batches of the pipeline: they overlap on the device, so their sum can exceed Time.Alignment)
        fprintf(stderr, "  => Time.Windowed Small %9.2f ms (WindowEd(S) bound)\n  => Time.Windowed Large %9.2f ms (%ld pairs)\n  => Time.Banded         %9.2f ms (%ld pairs)\n"
                        "  => Time.Align          %9.2f ms (fill %.2f + traceback %.2f + fused %.2f; %ld Hirschberg splits)\n  => Time.Prepare        %9.2f ms\n  => Time.CigarText      %9.2f ms\n"
                        "  => GPU.kernel_launches %ld   GPU.word_steps %ld   H2D %ld B   D2H %ld B\n",
                st_acc.ms_windowed_s, st_acc.ms_windowed_l, (long)st_acc.pairs_stage2, st_acc.ms_banded, (long)st_acc.pairs_stage3,
                st_acc.ms_align_fill + st_acc.ms_align_trace + st_acc.ms_fused, st_acc.ms_align_fill, st_acc.ms_align_trace, st_acc.ms_fused, (long)st_acc.hirschberg_splits,
                st_acc.ms_prepare, st_acc.ms_cigar, (long)st_acc.kernel_launches, (long)st_acc.word_steps, (long)st_acc.h2d_bytes, (long)st_acc.d2h_bytes);
    }
    if (o.check) {
        fprintf(stderr, "[Accuracy]\n => Alignments.Correct  %ld / %ld\n", total - bad, total);
        if (o.check_score) fprintf(stderr, " => Score.Correct       %ld / %ld\n   => Score.Total       %ld score uds.\n     => Score.Diff      %ld score uds.\n", score_ok, total, score_total, score_diff);
        if (!prm.only_score) fprintf(stderr, " => CIGAR.Breakdown\n   => CIGAR.Matches     %ld / %ld bases\n   => CIGAR.Mismatches  %ld\n   => CIGAR.Insertions  %ld\n   => CIGAR.Deletions   %ld\n", cg_m, bases, cg_x, cg_i, cg_d);
    }
    if (out) fclose(out);
    if (sam) fclose(sam);
    fclose(in);
    qb200_destroy(gpu);
    return bad ? 4 : 0;
}
