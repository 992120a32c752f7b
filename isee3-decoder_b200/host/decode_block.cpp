// decode_block -- framed (minor-frame) decoder above libviterbi224_b200 (host program, no CUDA in this file).
//
// Same input, same standard output as the reference's frame decoder in its Viterbi-only mode, `decode -V`
// (decode.c:44-289): symdemod-format soft symbols on stdin; for every 1024-bit minor frame a header line and a hex dump.
// Per frame the reference runs   init(SYNCWORD & 0xffffff) / update(1024) / chainback(1024, SYNCWORD & 0xffffff)
// (decode.c:220-222) on the 2048 symbols behind the frame-sync position, checks that the decoded frame ends in the
// 40-bit sync word (decode.c:241-249) and, only when it does NOT ("no lock"), searches the next frame's position with
// the 34-tap sync correlator (decode.c:162-181); after a good frame the next one is taken to start exactly 2048
// symbols later.
//
// That makes a run of good frames a batch of independent, equally spaced frames -- the data-parallel axis the GPU
// wants.  This program decodes such a run speculatively with ONE v224x_decode_frames call (frames side by side in a
// lockstep launch), then replays the reference's lock logic over the results in order: everything up to and including
// the first frame that fails the sync-word check is what the reference would have printed; the frames behind it were
// speculation and are decoded again from wherever the correlator puts the next frame.  After a bad frame the batch
// restarts at one frame and doubles while frames keep locking.  The output is byte-for-byte the reference's.
//
// The Fano decoder and the Fano-first policy (decode.c:184-204) are not part of this path: only -V behaviour exists.
//   -n  do not print bad frames      -r symrate  for the time stamp (default 1024)      -B frames  largest batch (64)
//   -L n  frames side by side in one launch (1..4, default 4)
//   -S  sync search only: print the frame positions found with lock never asserted (no GPU; CPU test tier)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <clocale>
#include <vector>
#include <unistd.h>
#include "../../include/viterbi224.h"
#include "../../include/viterbi224_b200.h"
#include "hostfmt.h"

namespace {

constexpr int FRAMEBITS = 1024;                       // decode.c:21
constexpr int FRAMESYMBOLS = 2 * FRAMEBITS;           // decode.c:22
constexpr int SYNCBITS = 34;                          // decode.c:23
constexpr unsigned long long SYNCWORD = 0x12fc819fbeull;   // decode.c:24
constexpr unsigned long long POLY1 = 073665667ull, POLY2 = 073665665ull;   // code.h:59-60
constexpr int G1FLIP = 0, G2FLIP = 1;                 // code.h:62-63

// The 34 correlator taps (decode.c:37-40 lists them as constants): the last 34 symbols of the encoded sync word.
void sync_taps(int taps[SYNCBITS])
{
    int sym[80];
    unsigned long long reg = 0;
    for (int i = 39; i >= 0; i--) {
        reg = (reg << 1) | ((SYNCWORD >> i) & 1);
        sym[2 * (39 - i)] = G1FLIP ^ __builtin_parityll(reg & POLY1);
        sym[2 * (39 - i) + 1] = G2FLIP ^ __builtin_parityll(reg & POLY2);
    }
    for (int k = 0; k < SYNCBITS; k++) taps[k] = sym[80 - SYNCBITS + k];
}

// stdin as a window over the symbol stream, addressed by absolute symbol index: have(n) blocks until symbols [.., n)
// are there (false at end of input), drop_before(a) forgets what lies in front of a.
struct Input {
    std::vector<unsigned char> buf;
    unsigned long long origin = 0;      // absolute index of buf[0]
    bool eof = false;
    bool have(unsigned long long n)
    {
        while (origin + buf.size() < n && !eof) {
            const size_t old = buf.size(), need = (size_t)(n - origin) - old, want = need > (1u << 16) ? need : (1u << 16);
            buf.resize(old + want);
            const size_t got = fread(buf.data() + old, 1, want, stdin);
            buf.resize(old + got);
            if (got == 0) eof = true;
        }
        return origin + buf.size() >= n;
    }
    const unsigned char *at(unsigned long long a) const { return buf.data() + (size_t)(a - origin); }
    void drop_before(unsigned long long a)
    {
        if (a - origin < (1u << 24)) return;
        buf.erase(buf.begin(), buf.begin() + (size_t)(a - origin));
        origin = a;
    }
};

// decode.c:162-181: the first position of the largest correlation over one frame of positions
int sync_search(const unsigned char *s, const int taps[SYNCBITS])
{
    int best = -1000000, where = -1;
    for (int i = 0; i < FRAMESYMBOLS; i++) {
        int sum = 0;
        for (int k = 0; k < SYNCBITS; k++) {
            const int v = (int)s[i + k] - 128;
            sum += taps[k] ? v : -v;
        }
        if (sum > best) { best = sum; where = i; }
    }
    return where;
}

} // namespace

int main(int argc, char *argv[])
{
    int no_bad = 0, max_batch = 64, nlock = 4, sync_only = 0, viterbi_only = 0;
    double symrate = 1024;
    const char *lang = getenv("LANG");
    setlocale(LC_ALL, lang ? lang : "en_US.utf8");                       // decode.c:62-65
    int opt;
    while ((opt = getopt(argc, argv, "nFVvr:s:m:d:pB:L:S")) != -1) {
        switch (opt) {
        case 'n': no_bad = 1; break;
        case 'V': viterbi_only = 1; break;
        case 'F': fprintf(stderr, "%s: the Fano decoder is not part of this program (Viterbi path only)\n", argv[0]); return 1;
        case 'r': symrate = atof(optarg); break;
        case 'B': max_batch = atoi(optarg); break;
        case 'L': nlock = atoi(optarg); break;
        case 'S': sync_only = 1; break;
        default: break;                                                   // -v -p -s -m -d: accepted, no effect without Fano
        }
    }
    if (!viterbi_only && !sync_only) fprintf(stderr, "%s: Fano-first decoding is not part of this program; running as -V\n", argv[0]);
    if (max_batch < 1) max_batch = 1;
    int taps[SYNCBITS];
    sync_taps(taps);
    Input in;

    if (sync_only) {
        // the correlator alone, as if no frame ever locked: one position per frame, each searched from the previous one
        unsigned long long base = 0;
        while (in.have(base + FRAMESYMBOLS + SYNCBITS)) {
            const int ss = sync_search(in.at(base), taps);
            if (!in.have(base + ss + FRAMESYMBOLS + SYNCBITS)) break;
            printf("%llu\n", base + ss + SYNCBITS);
            base += ss + FRAMESYMBOLS;
            in.drop_before(base);
        }
        return 0;
    }

    printf("%s: Fano %s; Viterbi %s\n", argv[0], "disabled", "enabled");  // decode.c:110-112
    if (no_bad) printf("%s: Not displaying bad frames\n", argv[0]);       // decode.c:114-115
    void *vd = create_viterbi224(FRAMEBITS);
    if (!vd) {
        printf("%s: cannot set up the Viterbi decoder: %s\n", argv[0], v224x_last_error());
        return 2;                                                         // decode.c:141-146
    }
    std::vector<unsigned char> data((size_t)max_batch * (FRAMEBITS / 8));
    std::vector<unsigned int> states(max_batch, (unsigned int)(SYNCWORD & 0xffffff));
    unsigned long long frames = 1;
    unsigned long long base = 0;        // absolute index of the reference's symbols[0] (its total_symbols)
    int lock = 0, batch = 1;
    unsigned long long launches = 0, wasted = 0;

    for (;;) {
        if (!in.have(base + FRAMESYMBOLS + SYNCBITS)) break;             // decode.c:152-161
        int sync_start = 0;
        if (!lock) {
            sync_start = sync_search(in.at(base), taps);         // decode.c:162-181
            if (!in.have(base + sync_start + FRAMESYMBOLS + SYNCBITS)) break;   // decode.c:183-192
        }
        // a run of frames, 2048 symbols apart: as many as are complete, at most `batch`
        const unsigned long long first = base + sync_start + SYNCBITS;
        int nb = 1;
        while (nb < batch && in.have(first + (unsigned long long)(nb + 1) * FRAMESYMBOLS)) nb++;
        if (v224x_decode_frames(vd, in.at(first), nb, FRAMEBITS, states.data(), states.data(), data.data(), nlock) < 0) {
            fprintf(stderr, "%s: decode failed: %s\n", argv[0], v224x_last_error());
            return 1;
        }
        launches++;
        int used = 0;
        for (int f = 0; f < nb; f++) {
            const unsigned char *d = data.data() + (size_t)f * (FRAMEBITS / 8);
            unsigned long long lastword = 0;
            for (int i = 123; i < 128; i++) lastword = (lastword << 8) | d[i];
            lock = lastword == SYNCWORD;                                  // decode.c:241-249
            if (lock || !no_bad) {                                        // decode.c:251-267
                const unsigned long long start_symbol = base + sync_start + SYNCBITS;
                printf("Frame %'llu at symbol %'llu (%s) with %s %s\n", frames, start_symbol,
                       v224host::format_hms(start_symbol / symrate).c_str(), "Viterbi", !lock ? "(bad)" : "");
                v224host::print_frame_hex(stdout, d, FRAMEBITS / 8);
                putchar('\n');
                fflush(stdout);
            }
            frames++;
            base += sync_start + FRAMESYMBOLS;                            // decode.c:270-282
            sync_start = 0;
            used++;
            if (!lock) break;                                             // what follows was decoded on a guess that no longer holds
        }
        wasted += (unsigned long long)(nb - used);
        batch = lock ? (2 * batch < max_batch ? 2 * batch : max_batch) : 1;
        in.drop_before(base);           // symbols in front of `base` are never looked at again
    }
    delete_viterbi224(vd);
    if (getenv("V224_HOST_STATS")) fprintf(stderr, "%s: %llu frames, %llu launches, %llu speculative frames discarded\n", argv[0], frames - 1, launches, wasted);
    return 0;
}
