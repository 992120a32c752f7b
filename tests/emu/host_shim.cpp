// Test helper: C entry points around the host-side headers of the block drivers (isee3-decoder_b200/host/*.h), so that the
// CPU test tier can call them through ctypes and compare with the unmodified reference (oracle/_ref/libv224_reffano.so).
#include "../../isee3-decoder_b200/host/fano_seq.h"
#include "../../isee3-decoder_b200/host/hostfmt.h"
#include <cstring>

extern "C" void shim_fano_metric_table(int *table, double signal, double noise, double bias, double scale)
{
    v224host::fano_metric_table(reinterpret_cast<int(*)[256]>(table), signal, noise, bias, scale);
}

extern "C" int shim_fano(unsigned long *metric, unsigned long *cycles, unsigned char *data, const unsigned char *symbols, unsigned nbits,
                         const int *table, int delta, unsigned long maxcycles, unsigned long long start, unsigned long long tail)
{
    static v224host::FanoDecoder dec;
    const v224host::FanoOutcome o = dec.decode(data, symbols, nbits, reinterpret_cast<const int(*)[256]>(table), delta, maxcycles, start, tail);
    *metric = o.metric;
    *cycles = o.cycles;
    return o.bits;
}

extern "C" void shim_format_hms(double t, char *out, int n)
{
    strncpy(out, v224host::format_hms(t).c_str(), (size_t)n - 1);
    out[n - 1] = 0;
}
