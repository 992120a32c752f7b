// hostfmt.h -- output formats the reference's host programs share, restated for the block-mode drivers.
#pragma once
#include <cstdio>
#include <string>

namespace v224host {

// A time in seconds as [d:][hh:]mm:ss.sss -- the string timeformat.c:27-66 (format_hms) builds: days only when
// non-zero, hours when days or hours are non-zero, seconds with a leading zero below ten and three decimals.
inline std::string format_hms(double t)
{
    const int days = (int)(t / 86400.);
    t -= days * 86400;
    const int hours = (int)(t / 3600.);
    t -= hours * 3600;
    const int minutes = (int)(t / 60.);
    t -= minutes * 60;
    char buf[64];
    std::string out;
    if (days > 0) { snprintf(buf, sizeof buf, "%d:", days); out += buf; }
    if (days > 0 || hours > 0) { snprintf(buf, sizeof buf, "%02d:", hours); out += buf; }
    snprintf(buf, sizeof buf, "%02d:", minutes);
    out += buf;
    if (t < 10.0) out += "0";
    snprintf(buf, sizeof buf, "%.3lf", t);
    out += buf;
    return out;
}

// One 1024-bit minor frame as hex: 16 bytes per line, single spaces, the layout both decode.c:254-260 and
// framer.c:77-87 print.
inline void print_frame_hex(FILE *f, const unsigned char *data, int nbytes)
{
    for (int i = 0; i < nbytes; i++) {
        fprintf(f, "%02x", data[i]);
        fputc((i % 16) == 15 ? '\n' : ' ', f);
    }
}

} // namespace v224host
