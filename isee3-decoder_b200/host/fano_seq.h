// fano_seq.h -- host-side sequential (Fano) decoder and its metric table for the K=24 r=1/2 code: the cheap first try of
// the reference's frame decoder before the Viterbi decoder is asked (decode.c:184-204).
//
// This is CPU code by nature (one path explored at a time, data-dependent back-tracking); it exists so that decode_block
// can run the reference's Fano-first policy and hand only the frames Fano gives up on to the GPU, in one batch.
// Results are the reference's, value for value: decoded bytes, number of decoded bits, final path metric and cycle count
// of fano() (fano.c:38-205) and the integer tables of gen_met() (metrics.c:24-89) -- checked in the CPU test tier against
// the unmodified reference (tests/test_host_logic.py).
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace v224host {

// Log-likelihood metric table for an 8-bit quantised BPSK/AWGN channel (metrics.c:24-89): table[sent][received].
// Bin s collects (s-128-0.5, s-128+0.5], bins 0 and 255 take the tails; metric = (1 + log2 P(s|sent) - log2(P(s|0)+P(s|1))
// - bias) * scale, rounded to nearest-even; both entries are -bias*scale where the two probabilities are equal, and
// -33*scale where the sent symbol's probability underflowed to zero.
inline void fano_metric_table(int table[2][256], double signal, double noise, double bias, double scale)
{
    const double inv = 1. / noise;
    auto cdf = [](double x) { return 0.5 + 0.5 * std::erf(x / M_SQRT2); };           // metrics.c:17-19
    double below0 = 0.0, below1 = 0.0;
    for (int s = 0; s < 256; s++) {
        const double upto0 = s != 255 ? cdf((s - 128 + 0.5 + signal) * inv) : 1.0;    // metrics.c:57-58
        const double upto1 = s != 255 ? cdf((s - 128 + 0.5 - signal) * inv) : 1.0;
        const double p0 = upto0 - below0, p1 = upto1 - below1;
        below0 = upto0;
        below1 = upto1;
        double m0, m1;
        if (p0 == p1) {
            m0 = m1 = -bias;                                                          // metrics.c:67-71
        } else {
            m0 = p0 == 0 ? -33.0 : 1 + std::log2(p0) - std::log2(p1 + p0) - bias;     // metrics.c:77-78
            m1 = p1 == 0 ? -33.0 : 1 + std::log2(p1) - std::log2(p1 + p0) - bias;
        }
        table[0][s] = (int)std::lrint(m0 * scale);
        table[1][s] = (int)std::lrint(m1 * scale);
    }
}

struct FanoOutcome {
    int bits;                 // decoded bits (== nbits on success), fano.c:203
    unsigned long metric;     // cumulative metric of the node the search stopped at, fano.c:190
    unsigned long cycles;     // loop count (limit + 1 when the search timed out), fano.c:191
};

class FanoDecoder {
public:
    static constexpr int K = 24;
    static constexpr unsigned long long POLY1 = 073665667ull, POLY2 = 073665665ull;     // code.h:59-60
    static constexpr int G1FLIP = 0, G2FLIP = 1;                                        // code.h:62-63

    // One frame: `symbols` = 2*nbits soft symbols, `data` receives bits/8 bytes (MSB first).  The last K-1 bits of the
    // frame are forced to `tail`; the encoder starts in `start`.  cycles_per_bit * nbits bounds the search.
    FanoOutcome decode(unsigned char *data, const unsigned char *symbols, unsigned nbits, const int table[2][256], int delta,
                       unsigned long cycles_per_bit, unsigned long long start, unsigned long long tail)
    {
        const long n = (long)nbits, first_forced = n - (K - 1);
        pair_metric_.resize(4 * (size_t)n);
        gamma_.resize(n); reg_.resize(n); best_.resize(n); other_.resize(n); second_.assign(n, 0);
        // the four pair metrics of every position: index = 2 * (POLY1 symbol sent) + (POLY2 symbol sent), fano.c:75-85
        for (long k = 0; k < n; k++) {
            const unsigned char a = symbols[2 * k], b = symbols[2 * k + 1];
            int *m = &pair_metric_[4 * (size_t)k];
            m[0] = table[0][a] + table[0][b];
            m[1] = table[0][a] + table[1][b];
            m[2] = table[1][a] + table[0][b];
            m[3] = table[1][a] + table[1][b];
        }
        // arrive at position k with register `r` (new bit still 0): rank the two branches, or take the forced tail bit
        auto arrive = [&](long k, unsigned long long r, bool may_force) {
            const int zero_syms = pair_of(r);
            const int *m = &pair_metric_[4 * (size_t)k];
            if (may_force && k >= first_forced) {                                       // fano.c:139-145
                const int bit = (int)((tail >> (n - k - 1)) & 1);
                r += (unsigned long long)bit;
                best_[k] = m[(bit ? 3 : 0) ^ zero_syms];
            } else {                                                                    // fano.c:94-106,146-160
                const int m0 = m[zero_syms], m1 = m[3 ^ zero_syms];                     // both polynomials are odd: the 1-branch sends the complement
                if (m0 > m1) { best_[k] = m0; other_[k] = m1; }
                else         { best_[k] = m1; other_[k] = m0; r |= 1; }
            }
            reg_[k] = r;
            second_[k] = 0;
        };
        long k = 0;
        long threshold = 0;
        gamma_[0] = 0;
        arrive(0, start << 1, false);
        const unsigned long limit = cycles_per_bit * nbits;                              // fano.c:108
        unsigned long cycle = 1;
        for (; cycle <= limit; cycle++) {
            const long ahead = gamma_[k] + (second_[k] ? other_[k] : best_[k]);
            if (ahead >= threshold) {
                // forward; on the first visit of this node raise the threshold as far as the new metric allows (fano.c:123-131)
                if (gamma_[k] < threshold + delta)
                    while (ahead >= threshold + delta) threshold += delta;
                if (k + 1 == n) break;                                                   // end of frame reached (fano.c:133-136)
                k++;
                gamma_[k] = ahead;
                arrive(k, reg_[k - 1] << 1, true);
                continue;
            }
            // blocked: walk back to the nearest node whose other branch is still untried; if the way back is itself
            // below the threshold, lower the threshold and retry this node's best branch (fano.c:165-187)
            for (;;) {
                if (k == 0 || gamma_[k - 1] < threshold) {
                    threshold -= delta;
                    if (second_[k]) { second_[k] = 0; reg_[k] ^= 1; }
                    break;
                }
                k--;
                if (k < first_forced && !second_[k]) { second_[k] = 1; reg_[k] ^= 1; break; }
            }
        }
        FanoOutcome out;
        out.metric = (unsigned long)gamma_[k];
        out.cycles = cycle;
        out.bits = (int)(k + 1);
        for (long j = 0; j < out.bits / 8; j++) data[j] = (unsigned char)reg_[8 * j + 7];   // fano.c:196-201
        return out;
    }

private:
    // symbol pair an encoder register sends: POLY1 symbol in bit 1, POLY2 symbol in bit 0 (fano.c:29-36)
    static int pair_of(unsigned long long r)
    {
        return ((__builtin_parityll(r & POLY1) << 1) ^ G1FLIP) | (__builtin_parityll(r & POLY2) ^ G2FLIP);
    }
    std::vector<int> pair_metric_, best_, other_;
    std::vector<long> gamma_;
    std::vector<unsigned long long> reg_;
    std::vector<unsigned char> second_;
};

} // namespace v224host
