// include/quicked.hpp — C++ wrapper with the reference binding's class surface (reference bindings/cpp/quicked.hpp:46-73)
// on top of the drop-in C API, plus a batched alignMany() on the additive GPU entry point.  Header-only.
#ifndef QUICKED_HPP
#define QUICKED_HPP

#include <cstdint>
#include <exception>
#include <string>
#include <utility>
#include <vector>

#include "quicked_b200.h"

namespace quicked {

class QuickedException : public std::exception {
    quicked_status_t status;

  public:
    explicit QuickedException(quicked_status_t s) : status(s) {}
    const char *what() const noexcept override { return quicked_status_msg(status); }
};

class QuickedAligner {
  public:
    QuickedAligner()
    {
        params = quicked_default_params();
        const quicked_status_t st = quicked_new(&aligner, &params);
        if (quicked_check_error(st)) throw QuickedException(st);
    }
    ~QuickedAligner()
    {
        quicked_free(&aligner);
        if (gpu) qb200_destroy(gpu);
    }
    QuickedAligner(const QuickedAligner &) = delete;
    QuickedAligner &operator=(const QuickedAligner &) = delete;

    void align(std::string *pattern, std::string *text)
    {
        const quicked_status_t st = quicked_align(&aligner, pattern->c_str(), (int)pattern->length(), text->c_str(), (int)text->length());
        if (quicked_check_error(st)) throw QuickedException(st);
    }

    struct Result { int status; int score; std::string cigar; };
    // Batched: one GPU pass over all pairs (same parameters as the single-pair calls).
    std::vector<Result> alignMany(const std::vector<std::pair<std::string, std::string>> &pairs, int device = 0)
    {
        if (!gpu && qb200_create(&gpu, device) != QB200_OK) throw QuickedException(QUICKED_ERROR);
        std::string seqs;
        std::vector<int64_t> po, to;
        std::vector<int32_t> pl, tl;
        for (const auto &pr : pairs) {
            po.push_back((int64_t)seqs.size()); pl.push_back((int32_t)pr.first.size()); seqs += pr.first;
            to.push_back((int64_t)seqs.size()); tl.push_back((int32_t)pr.second.size()); seqs += pr.second;
        }
        const int64_t n = (int64_t)pairs.size();
        std::vector<int32_t> score((size_t)n), status((size_t)n);
        std::vector<int64_t> off((size_t)n + 1);
        std::vector<char> cig(seqs.size() + 16 * (size_t)n + 64);
        qb200_batch_t b = {seqs.data(), (int64_t)seqs.size(), n, po.data(), pl.data(), to.data(), tl.data()};
        qb200_results_t r = {score.data(), status.data(), cig.data(), (int64_t)cig.size(), off.data(), 0};
        int rc = qb200_align_batch(gpu, &params, &b, &r);
        if (rc == QB200_ERR_CAPACITY) {
            cig.resize((size_t)r.cigar_bytes + 64);
            r.cigar = cig.data(); r.cigar_capacity = (int64_t)cig.size();
            rc = qb200_align_batch(gpu, &params, &b, &r);
        }
        if (rc != QB200_OK) throw QuickedException(QUICKED_ERROR);
        std::vector<Result> out((size_t)n);
        for (int64_t i = 0; i < n; ++i)
            out[(size_t)i] = {status[(size_t)i], score[(size_t)i],
                              (off[(size_t)i + 1] - off[(size_t)i] > 1) ? std::string(cig.data() + off[(size_t)i]) : std::string()};
        return out;
    }

    void setAlgorithm(quicked_algo_t algo) { params.algo = algo; }
    void setOnlyScore(bool v) { params.only_score = v; }
    void setBandwidth(unsigned int v) { params.bandwidth = v; }
    void setWindowSize(unsigned int v) { params.window_size = v; }
    void setOverlapSize(unsigned int v) { params.overlap_size = v; }
    void setForceScalar(bool v) { params.force_scalar = v; }
    void setHEWThreshold(unsigned int v) { params.hew_threshold[0] = params.hew_threshold[1] = v; }
    void setHEWPercentage(unsigned int v) { params.hew_percentage[0] = params.hew_percentage[1] = v; }

    int getScore() { return aligner.score; }
    std::string getCigar() { return std::string(aligner.cigar ? aligner.cigar : "NULL"); }

  private:
    quicked_aligner_t aligner;
    quicked_params_t params;
    qb200_ctx_t *gpu = nullptr;
};

}  // namespace quicked
#endif  // QUICKED_HPP
