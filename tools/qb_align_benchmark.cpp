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
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <getopt.h>
#include <string>
#include <vector>

#include "quicked_b200.h"

struct Options {
    std::string algo = "quicked", input, output;
    bool output_full = false, force_scalar = false, check = false, only_score = false;
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
            "          --input|i PATH            '>pattern' / '<text' line pairs\n"
            "          --output|o PATH           score<TAB>CIGAR per pair\n"
            "          --output-full PATH        lengths, score, sequences, CIGAR\n"
            "        [Other Parameters]\n"
            "          --bandwidth INT  --window-size INT  --overlap-size INT\n"
            "          --hew-threshold INT  --hew-percentage INT  --force-scalar  --only-score\n"
            "        [Misc]\n"
            "          --check|c correct         replay every CIGAR on its pair\n"
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

int main(int argc, char **argv)
{
    Options o;
    static struct option long_options[] = {
        {"algorithm", required_argument, 0, 'a'}, {"input", required_argument, 0, 'i'}, {"output", required_argument, 0, 'o'},
        {"output-full", required_argument, 0, 800}, {"bandwidth", required_argument, 0, 2000}, {"window-size", required_argument, 0, 2001},
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
        case 2000: o.bandwidth = atoi(optarg); break;
        case 2001: o.window_size = atoi(optarg); break;
        case 2002: o.overlap_size = atoi(optarg); break;
        case 2003: o.hew_threshold = atoi(optarg); break;
        case 2004: o.hew_percentage = atoi(optarg); break;
        case 2005: o.force_scalar = true; break;
        case 2006: o.only_score = true; break;
        case 'c': o.check = true; break;
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
    prm.bandwidth = (unsigned)o.bandwidth; prm.window_size = (unsigned)o.window_size; prm.overlap_size = (unsigned)o.overlap_size;
    prm.hew_threshold[0] = prm.hew_threshold[1] = (unsigned)o.hew_threshold;
    prm.hew_percentage[0] = prm.hew_percentage[1] = (unsigned)o.hew_percentage;
    prm.force_scalar = o.force_scalar; prm.only_score = o.only_score;

    FILE *in = fopen(o.input.c_str(), "r");
    if (!in) { fprintf(stderr, "Input file '%s' couldn't be opened\n", o.input.c_str()); return 1; }
    FILE *out = o.output.empty() ? nullptr : fopen(o.output.c_str(), "w");
    qb200_ctx_t *gpu = nullptr;
    if (qb200_create(&gpu, o.device) != QB200_OK) { fprintf(stderr, "qb_align_benchmark: no CUDA device (there is no CPU fallback)\n"); return 2; }

    std::vector<char> seqs;
    std::vector<int64_t> po, to;
    std::vector<int32_t> pl, tl, score, status;
    std::vector<int64_t> coff;
    std::vector<char> cig;
    char *l1 = nullptr, *l2 = nullptr;
    size_t a1 = 0, a2 = 0;
    long total = 0, bad = 0;
    double t_align = 0;
    const auto t_begin = std::chrono::steady_clock::now();
    bool eof = false;
    while (!eof) {
        seqs.clear(); po.clear(); to.clear(); pl.clear(); tl.clear();
        while ((long)po.size() < o.batch_size) {          // align_benchmark.c:73-99: strip the '>' / '<' and the newline
            ssize_t n1 = getline(&l1, &a1, in);
            if (n1 < 0) { eof = true; break; }
            ssize_t n2 = getline(&l2, &a2, in);
            if (n2 < 0) { eof = true; break; }
            while (n1 > 0 && (l1[n1 - 1] == '\n' || l1[n1 - 1] == '\r')) --n1;
            while (n2 > 0 && (l2[n2 - 1] == '\n' || l2[n2 - 1] == '\r')) --n2;
            const char *p = l1 + 1, *t = l2 + 1;
            long m = n1 - 1, n = n2 - 1;
            if (l1[0] == '<' && l2[0] == '>') { std::swap(p, t); std::swap(m, n); }   // generate_dataset.c:396-405 prints either order
            po.push_back((int64_t)seqs.size()); pl.push_back((int32_t)std::max(m, 0L)); seqs.insert(seqs.end(), p, p + std::max(m, 0L));
            to.push_back((int64_t)seqs.size()); tl.push_back((int32_t)std::max(n, 0L)); seqs.insert(seqs.end(), t, t + std::max(n, 0L));
        }
        const int64_t np = (int64_t)po.size();
        if (!np) break;
        seqs.push_back(0);
        score.assign((size_t)np, -1); status.assign((size_t)np, -1); coff.assign((size_t)np + 1, 0);
        cig.resize(std::max<size_t>(cig.size(), seqs.size() / 2 + 1024));
        qb200_batch_t b = {seqs.data(), (int64_t)seqs.size(), np, po.data(), pl.data(), to.data(), tl.data()};
        qb200_results_t r = {score.data(), status.data(), cig.data(), (int64_t)cig.size(), coff.data(), 0};
        const auto t0 = std::chrono::steady_clock::now();
        int rc = qb200_align_batch(gpu, &prm, &b, &r);
        if (rc == QB200_ERR_CAPACITY) {
            cig.resize((size_t)r.cigar_bytes + 1024);
            r.cigar = cig.data(); r.cigar_capacity = (int64_t)cig.size();
            rc = qb200_align_batch(gpu, &prm, &b, &r);
        }
        t_align += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (rc != QB200_OK) { fprintf(stderr, "qb_align_benchmark: %s (rc=%d)\n", qb200_last_error(gpu), rc); return 3; }
        for (int64_t i = 0; i < np; ++i) {
            const char *cg = (prm.only_score || coff[(size_t)i + 1] - coff[(size_t)i] <= 1) ? "-" : cig.data() + coff[(size_t)i];
            const bool err = quicked_check_error((quicked_status_t)status[(size_t)i]);
            if (o.check && !err && !prm.only_score &&
                !replay_ok(cg, seqs.data() + po[(size_t)i], pl[(size_t)i], seqs.data() + to[(size_t)i], tl[(size_t)i], score[(size_t)i])) ++bad;
            if (!out) continue;
            if (o.output_full) {
                fprintf(out, "%d\t%d\t", pl[(size_t)i], tl[(size_t)i]);
                if (err) fprintf(out, "ERROR\t"); else fprintf(out, "%d\t", score[(size_t)i]);
                fwrite(seqs.data() + po[(size_t)i], 1, (size_t)pl[(size_t)i], out); fputc('\t', out);
                fwrite(seqs.data() + to[(size_t)i], 1, (size_t)tl[(size_t)i], out);
                fprintf(out, "\t%s\n", err ? (prm.only_score ? "-" : "ERROR") : cg);
            } else if (err) fprintf(out, "ERROR\t%s\n", prm.only_score ? "-" : "ERROR");
            else fprintf(out, "%d\t%s\n", score[(size_t)i], cg);
        }
        total += np;
    }
    const double t_total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
    fprintf(stderr, "...processed %ld reads (alignment = %2.3f seq/s)\n", total, t_align > 0 ? total / t_align : 0.0);
    fprintf(stderr, "[Benchmark]\n=> Total.reads            %ld\n=> Time.Benchmark      %9.2f s\n  => Time.Alignment    %9.2f s\n", total, t_total, t_align);
    if (o.check) fprintf(stderr, "[Accuracy]\n => Alignments.Correct  %ld / %ld\n", total - bad, total);
    if (out) fclose(out);
    fclose(in);
    free(l1); free(l2);
    qb200_destroy(gpu);
    return bad ? 4 : 0;
}
