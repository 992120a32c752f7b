// v224_pairing.cpp -- v224x_pair_symbols(): the host half of the reference's streaming driver as a library call.
// Plain C++ (compiled by g++, no CUDA): the sync-correlator phase flip of vdecode.c:107-140 stays on the host, as
// BASELINE's north_star asks; this entry runs it over a whole buffer so that block callers (bench.py, Python, the
// multi-GPU decode) pay for it once per stream and inside their timed region.
#include <thread>
#include <cstdlib>
#include "../host/pairing.h"
#include "../../include/viterbi224_b200.h"

extern "C" long long v224x_pair_symbols(const unsigned char *soft, long long nsyms, int start_phase, int dontflip, int delay,
                                        unsigned char *pairs_out, unsigned char *cmp_out, long long *flip_at, int flip_cap, int *nflips)
{
    if (nflips) *nflips = 0;
    if (!soft || !pairs_out || nsyms < 0) return -1;
    v224host::SymbolPairer pr(start_phase, dontflip != 0, delay);
    std::vector<unsigned long long> flips;
    // The 34-tap correlation is nearly all of the work and does not depend on the flip decisions: for long buffers it is
    // computed up front by a few host threads, slice by slice; the sequential pass below then only compares peaks and
    // copies symbols (and recomputes the 33 positions behind every dropped symbol itself).
    static thread_local std::vector<int16_t> pre;        // kept between calls: block callers come back with the same size
    pre.clear();
    unsigned nt = std::thread::hardware_concurrency() / 4;      // several ranks of one box may be pairing at the same time
    if (const char *e = getenv("V224_PAIR_THREADS")) nt = (unsigned)atoi(e);
    nt = nt < 1 ? 1 : (nt > 4 ? 4 : nt);
    if (!dontflip && nsyms >= (1ll << 22) && nt > 1) {
        pre.resize((size_t)nsyms);
        std::vector<std::thread> th;
        for (unsigned t = 1; t < nt; t++)
            th.emplace_back(v224host::SymbolPairer::correlate_block, soft, (size_t)(nsyms * t / nt), (size_t)(nsyms * (t + 1) / nt), pre.data());
        v224host::SymbolPairer::correlate_block(soft, 0, (size_t)(nsyms / nt), pre.data());
        for (auto &t : th) t.join();
    }
    // in slices: the run buffers of the correlator stay cache-sized whatever the caller hands over
    long long done = 0, npairs = 0;
    while (done < nsyms) {
        const long long n = nsyms - done < (1ll << 20) ? nsyms - done : (1ll << 20);
        npairs += (long long)pr.feed(soft + done, (size_t)n, pairs_out + 2 * npairs, cmp_out ? cmp_out + 2 * npairs : nullptr, &flips,
                                     pre.empty() ? nullptr : pre.data() + done);
        done += n;
    }
    if (nflips) *nflips = (int)flips.size();
    for (size_t i = 0; i < flips.size() && (int)i < flip_cap && flip_at; i++) flip_at[i] = (long long)flips[i];
    return npairs;
}
