// decode_block -- framed (minor-frame) decoder above libviterbi224_b200 (host program, no CUDA in this file).
//
// Same command line, same input, same standard output as the reference's frame decoder `decode` (decode.c:44-289) in all
// three of its modes -- Fano first with Viterbi fallback (default), -V Viterbi only, -F Fano only: symdemod-format soft
// symbols on stdin; for every 1024-bit minor frame a header line and a hex dump.
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
// Fano-first mode (decode.c:184-231): every frame of the run gets the sequential decoder on the host (fano_seq.h, the
// reference's fano() value for value); the reference asks the Viterbi decoder only for frames Fano gave up on, and only
// when the previous frame locked (or -p).  Here the frames of a run that Fano gave up on are gathered and decoded by the
// GPU in ONE batch, and the same replay of the lock logic picks, per frame, what the reference would have printed.
// -F runs without any GPU.
//   -n -F -V -v -p -r symrate -s fano_scale -m fano_maxcycles -d fano_delta   as in the reference (decode.c:78-108)
//   -B frames  largest batch (64)
//   -L n  frames side by side in one launch (1..4, default 4)
//   -S  sync search only: print the frame positions found with lock never asserted (no GPU; CPU test tier)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <clocale>
#include <vector>
#include <chrono>
#include <unistd.h>
#include <poll.h>
#include <cerrno>
#include "../../include/viterbi224.h"
#include "../../include/viterbi224_b200.h"
#include "hostfmt.h"
#include "fano_seq.h"

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
// are there (false at end of input), have_now(n) only takes what is already waiting in the pipe (a live symdemod delivers
// about a thousand symbols per second: a speculative batch must not sit and wait for frames that have not been received
// yet), drop_before(a) forgets what lies in front of a.  read(2) returns what is there; fread would hold out for full buffers.
struct Input {
    std::vector<unsigned char> buf;
    unsigned long long origin = 0;      // absolute index of buf[0]
    bool eof = false;
    bool fill(unsigned long long n, bool wait)
    {
        while (origin + buf.size() < n && !eof) {
            if (!wait) {
                struct pollfd pf = {0, POLLIN, 0};
                if (poll(&pf, 1, 0) <= 0) break;                 // nothing waiting right now
            }
            const size_t old = buf.size(), need = (size_t)(n - origin) - old, want = need > (1u << 16) ? need : (1u << 16);
            buf.resize(old + want);
            const ssize_t got = read(0, buf.data() + old, want);
            buf.resize(old + (got > 0 ? (size_t)got : 0));
            if (got < 0 && errno == EINTR) continue;
            if (got <= 0) eof = true;
        }
        return origin + buf.size() >= n;
    }
    bool have(unsigned long long n) { return fill(n, true); }
    bool have_now(unsigned long long n) { return fill(n, false); }
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
    int no_bad = 0, max_batch = 64, nlock = 4, sync_only = 0;
    int viterbi_enabled = 1, fano_enabled = 1, persistent = 0;           // decode.c:67-76
    double symrate = 1024, fano_scale = 8;
    unsigned long fano_maxcycles = 100;
    int fano_delta = (int)(4 * fano_scale);                               // fixed before the options are read (decode.c:73)
    const char *lang = getenv("LANG");
    setlocale(LC_ALL, lang ? lang : "en_US.utf8");                       // decode.c:62-65
    int opt;
    while ((opt = getopt(argc, argv, "nFVvr:s:m:d:pB:L:S")) != -1) {
        switch (opt) {
        case 'p': persistent = 1; break;
        case 'n': no_bad = 1; break;
        case 'F': viterbi_enabled = 0; break;
        case 'V': fano_enabled = 0; break;
        case 'v': break;
        case 'r': symrate = atof(optarg); break;
        case 's': fano_scale = atof(optarg); break;
        case 'm': fano_maxcycles = (unsigned long)atol(optarg); break;
        case 'd': fano_delta = atoi(optarg); break;
        case 'B': max_batch = atoi(optarg); break;
        case 'L': nlock = atoi(optarg); break;
        case 'S': sync_only = 1; break;
        default:
            printf("usage: %s [-F] [-V] [-v] [-r symrate] [-s fano_scale] [-m fano_maxcycles] [-d fano_delta]\n", argv[0]);
            break;
        }
    }
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

    printf("%s: Fano %s; Viterbi %s\n", argv[0], fano_enabled ? "enabled" : "disabled", viterbi_enabled ? "enabled" : "disabled");   // decode.c:110-112
    if (no_bad) printf("%s: Not displaying bad frames\n", argv[0]);      // decode.c:114-115
    if (!fano_enabled && !viterbi_enabled) {
        printf("%s: Specify only one of -F or -V\n", argv[0]);           // decode.c:117-120
        return 1;
    }
    using clk = std::chrono::steady_clock;
    auto secs = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    const clk::time_point t_start = clk::now();
    double t_create = 0, t_gpu = 0, t_fano = 0;
    int mettab[2][256];
    v224host::FanoDecoder fano;
    void *vd = nullptr;
    if (fano_enabled) {
        // symdemod scales to a total amplitude of 100; signal and noise amplitude assumed at the decoder's threshold,
        // Es/N0 = 0 dB (decode.c:121-137)
        const double total_amp = 100., est_esn0 = 1.0;
        const double noise_amp = total_amp / sqrt(1 + 2 * est_esn0), sig_amp = noise_amp * sqrt(2 * est_esn0);
        printf("%s: Fano decoder params: delta %'d; scale %'.1lf; maxcycles %'lu; signal %.1lf; noise %.1lf\n", argv[0], fano_delta, fano_scale,
               fano_maxcycles, sig_amp, noise_amp);
        v224host::fano_metric_table(mettab, sig_amp, noise_amp, 0.5, fano_scale);
    }
    if (viterbi_enabled && !fano_enabled) {
        // Viterbi only: the decoder is needed from the first frame on (decode.c:138-147); with Fano first it is created
        // when the first frame falls through to it
        const clk::time_point t0 = clk::now();
        vd = create_viterbi224(FRAMEBITS);
        t_create += secs(t0, clk::now());
        if (!vd) {
            printf("%s: cannot set up the Viterbi decoder: %s\n", argv[0], v224x_last_error());
            return 2;
        }
    }
    const unsigned int syncstate = (unsigned int)(SYNCWORD & 0xffffff);   // known start and end state of every frame (decode.c:220-222)
    const int FB = FRAMEBITS / 8;
    std::vector<unsigned char> fano_data((size_t)max_batch * FB), vit_data((size_t)max_batch * FB), vit_syms;
    std::vector<unsigned int> states(max_batch, syncstate);
    std::vector<int> fano_bits(max_batch), vit_slot(max_batch);
    unsigned long long frames = 1;
    unsigned long long base = 0;        // absolute index of the reference's symbols[0] (its total_symbols)
    int lock = 0, batch = 1;
    unsigned long long launches = 0, wasted = 0, n_fano_ok = 0, n_viterbi = 0;

    for (;;) {
        if (!in.have(base + FRAMESYMBOLS + SYNCBITS)) break;             // decode.c:152-161
        int sync_start = 0;
        if (!lock) {
            sync_start = sync_search(in.at(base), taps);                  // decode.c:162-181
            if (!in.have(base + sync_start + FRAMESYMBOLS + SYNCBITS)) break;   // decode.c:183-192
        }
        // a run of frames, 2048 symbols apart: as many as are complete, at most `batch`
        const unsigned long long first = base + sync_start + SYNCBITS;
        int nb = 1;
        while (nb < batch && in.have_now(first + (unsigned long long)(nb + 1) * FRAMESYMBOLS)) nb++;

        // 1. Fano on the host for every frame of the run (decode.c:196-204; the cycle limit it passes is the constant 100)
        if (fano_enabled) {
            const clk::time_point t0 = clk::now();
            memset(fano_data.data(), 0, (size_t)nb * FB);
            for (int f = 0; f < nb; f++)
                fano_bits[f] = fano.decode(&fano_data[(size_t)f * FB], in.at(first + (unsigned long long)f * FRAMESYMBOLS), FRAMEBITS, mettab,
                                           fano_delta, 100, syncstate, syncstate).bits;
            t_fano += secs(t0, clk::now());
        }
        // 2. the frames the reference would hand to the Viterbi decoder if every earlier frame of the run locks
        //    (decode.c:205-231): all of them without Fano; with Fano those it gave up on, frame 0 only if the previous
        //    frame locked or -p.  One batch on the GPU, gathered into consecutive slots.
        int nv = 0;
        for (int f = 0; f < nb; f++) {
            vit_slot[f] = -1;
            const bool prev_lock = f ? true : lock != 0;
            if (viterbi_enabled && (!fano_enabled || ((persistent || prev_lock) && fano_bits[f] != FRAMEBITS))) vit_slot[f] = nv++;
        }
        if (nv) {
            const clk::time_point t0 = clk::now();
            if (!vd) { vd = create_viterbi224(FRAMEBITS); t_create += secs(t0, clk::now()); }
            if (!vd) {
                // (the reference prints a notice and goes on with Fano's result, decode.c:212-215)
                printf("%s: cannot set up the Viterbi decoder: %s\n", argv[0], v224x_last_error());
                for (int f = 0; f < nb; f++) vit_slot[f] = -1;
                nv = 0;
            }
        }
        if (nv) {
            const unsigned char *src = in.at(first);
            if (nv != nb) {
                vit_syms.resize((size_t)nv * FRAMESYMBOLS);
                for (int f = 0; f < nb; f++)
                    if (vit_slot[f] >= 0) memcpy(&vit_syms[(size_t)vit_slot[f] * FRAMESYMBOLS], in.at(first + (unsigned long long)f * FRAMESYMBOLS), FRAMESYMBOLS);
                src = vit_syms.data();
            }
            const clk::time_point t0 = clk::now();
            if (v224x_decode_frames(vd, src, nv, FRAMEBITS, states.data(), states.data(), vit_data.data(), nlock) < 0) {
                fprintf(stderr, "%s: decode failed: %s\n", argv[0], v224x_last_error());
                return 1;
            }
            t_gpu += secs(t0, clk::now());
            launches++;
        }
        // 3. replay the reference's per-frame logic in order
        int used = 0;
        for (int f = 0; f < nb; f++) {
            const char *decoder = "None";
            const unsigned char *d = &fano_data[(size_t)f * FB];
            int result = 0;
            if (fano_enabled) { decoder = "Fano"; result = fano_bits[f]; }
            if (vit_slot[f] >= 0) { decoder = "Viterbi"; result = FRAMEBITS; d = &vit_data[(size_t)vit_slot[f] * FB]; n_viterbi++; }
            else if (result == FRAMEBITS) n_fano_ok++;
            lock = 0;
            if (result == FRAMEBITS) {                                    // decode.c:237-249
                unsigned long long lastword = 0;
                for (int i = 123; i < 128; i++) lastword = (lastword << 8) | d[i];
                lock = lastword == SYNCWORD;
            }
            if (lock || !no_bad) {                                        // decode.c:251-267
                const unsigned long long start_symbol = base + sync_start + SYNCBITS;
                printf("Frame %'llu at symbol %'llu (%s) with %s %s\n", frames, start_symbol,
                       v224host::format_hms(start_symbol / symrate).c_str(), decoder, !lock ? "(bad)" : "");
                v224host::print_frame_hex(stdout, d, FB);
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
    if (vd) delete_viterbi224(vd);
    if (getenv("V224_HOST_STATS"))
        fprintf(stderr, "%s: %llu frames (%llu by Fano, %llu by Viterbi), %llu launches, %llu speculative frames discarded; "
                "seconds: total %.2f, create %.2f, decode_frames %.2f, fano %.2f\n", argv[0], frames - 1,
                n_fano_ok, n_viterbi, launches, wasted, secs(t_start, clk::now()), t_create, t_gpu, t_fano);
    return 0;
}
