// v224_pairing.cpp -- v224x_pair_symbols(): the host half of the reference's streaming driver as a library call.
// Plain C++ (compiled by g++, no CUDA): the sync-correlator phase flip of vdecode.c:107-140 stays on the host, as
// BASELINE's north_star asks; this entry runs it over a whole buffer so that block callers (bench.py, Python, the
// multi-GPU decode) pay for it once per stream and inside their timed region.
#include "../host/pairing.h"
#include "../../include/viterbi224_b200.h"

extern "C" long long v224x_pair_symbols(const unsigned char *soft, long long nsyms, int start_phase, int dontflip, int delay,
                                        unsigned char *pairs_out, unsigned char *cmp_out, long long *flip_at, int flip_cap, int *nflips)
{
    if (nflips) *nflips = 0;
    if (!soft || !pairs_out || nsyms < 0) return -1;
    v224host::SymbolPairer pr(start_phase, dontflip != 0, delay);
    std::vector<unsigned long long> flips;
    // in slices: the run buffers of the correlator stay cache-sized whatever the caller hands over
    long long done = 0, npairs = 0;
    while (done < nsyms) {
        const long long n = nsyms - done < (1ll << 20) ? nsyms - done : (1ll << 20);
        npairs += (long long)pr.feed(soft + done, (size_t)n, pairs_out + 2 * npairs, cmp_out ? cmp_out + 2 * npairs : nullptr, &flips);
        done += n;
    }
    if (nflips) *nflips = (int)flips.size();
    for (size_t i = 0; i < flips.size() && (int)i < flip_cap && flip_at; i++) flip_at[i] = (long long)flips[i];
    return npairs;
}
