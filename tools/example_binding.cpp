// Mirrors the reference's examples/bindings/basic.cpp on the B200 library (row f3): single pair + a small batch.
#include <iostream>
#include "quicked.hpp"

int main()
{
    std::string pattern = "ACGT", text = "ACTT";
    try {
        quicked::QuickedAligner aligner;
        aligner.align(&pattern, &text);
        std::cout << "Score: " << aligner.getScore() << "\nCigar: " << aligner.getCigar() << "\n";
        aligner.setAlgorithm(BANDED);
        aligner.setBandwidth(50);
        auto res = aligner.alignMany({{"GATTACA", "GATCACA"}, {"ACGTACGT", "ACGTCGT"}});
        for (const auto &r : res) std::cout << r.score << "\t" << r.cigar << "\n";
    } catch (quicked::QuickedException &e) {
        std::cerr << e.what() << std::endl;
        return 1;
    }
    return 0;
}
