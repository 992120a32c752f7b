// vdecode_block -- block-mode streaming driver for libviterbi224_b200 (host program, no CUDA in this file).
//
// Same command line and the same standard output as the reference's streaming decoder `vdecode`
// (vdecode.c:38-189: soft symbols on stdin, one ASCII '0'/'1' per decoded bit on stdout, status lines on stderr),
// so it sits in the reference's shell pipeline unchanged:      symdemod | vdecode_block -d 200 | framer
//
// The reference asks the decoder for one bit per symbol pair (update(1) + decodebit(delay, 0), vdecode.c:145-152).
// Which symbols form a pair is decided by its sync correlator (vdecode.c:107-140) from RECEIVED SYMBOLS ONLY, never
// from decoder output, so this program runs that decision ahead of the decoder: it turns the input into the exact
// sequence of pairs vdecode would hand to update_viterbi224_blk -- including the stale-symbol pair that follows every
// phase flip -- collects them in blocks, and decodes a block with ONE call (v224x_stream_decode_seg, whose output is
// what the per-bit loop returns).  Everything vdecode derives from decoder output (start-up suppression, the
// re-encode symbol-error tally, the status lines) is replayed per pair afterwards from values recorded on the way in.
//
// Extra options: -B pairs  block size (default 262144; latency = one block), -S n  decoders in lockstep (default 4),
// -P  pairs only: write the symbol pairs that would go to the decoder (2 bytes each) to stdout and exit -- no GPU needed;
// the CPU test tier checks the pairing / phase-flip logic through it.
// -f  frames instead of bits: standard output is what `vdecode | framer` prints (framer.c:61-95: a 1024-bit shift
// register over the decoded bits; whenever its last 40 bits are the sync word, a header line and the hex dump of the
// frame), -r bitrate  for the frame time stamp (framer's option, default 512).  Saves the one-character-per-bit pipe.
// -b  with -f: standard input already is decoded bits ('0'/'1' characters): only the framing runs -- a `framer`; no GPU.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <clocale>
#include <vector>
#include <unistd.h>
#include <cerrno>
#include "../../include/viterbi224.h"
#include "../../include/viterbi224_b200.h"
#include "hostfmt.h"

namespace {

// code constants of the reference's active code block (code.h:54-63)
constexpr int K = 24;
constexpr unsigned long long POLY1 = 073665667ull, POLY2 = 073665665ull;
constexpr int G1FLIP = 0, G2FLIP = 1;
constexpr int FRAME_SYMBOLS = 2048;       // symbols per minor frame (vdecode.c:14-15)
constexpr int NTAPS = 34;                 // usable encoded sync symbols (vdecode.c:16)
constexpr int RING = 4096;                // vdecode's symbol history (vdecode.c:20); its size shows in the tally, so it is kept
constexpr unsigned long long SYNCWORD = 0x12fc819fbeull;   // decode.c:24; the correlator taps are its encoding

inline int parity64(unsigned long long x) { return __builtin_parityll(x); }

// read(2) hands over what has arrived (fread would hold out for a full buffer -- minutes on a live 1 ksymbol/s stream,
// on top of the one block of latency -B asks for); 0 at end of input
inline ssize_t read_some(unsigned char *dst, size_t cap)
{
    for (;;) {
        const ssize_t got = read(0, dst, cap);
        if (got < 0 && errno == EINTR) continue;
        return got;
    }
}

// The 34 encoded sync symbols (vdecode.c:27-30 lists them as constants; they are the tail of encode(SYNCWORD)).
void sync_taps(int taps[NTAPS])
{
    int sym[80];
    unsigned long long reg = 0;
    for (int i = 39; i >= 0; i--) {
        reg = (reg << 1) | ((SYNCWORD >> i) & 1);
        sym[2 * (39 - i)] = G1FLIP ^ parity64(reg & POLY1);
        sym[2 * (39 - i) + 1] = G2FLIP ^ parity64(reg & POLY2);
    }
    for (int k = 0; k < NTAPS; k++) taps[k] = sym[80 - NTAPS + k];
}

struct PairRec {
    unsigned char s0, s1;     // the pair handed to the decoder
    unsigned char c1, c2;     // hard-sliced history symbols vdecode compares the re-encoded pair with (vdecode.c:176-177)
};

} // namespace

int main(int argc, char *argv[])
{
    int delay = 200, interval = 1024, quiet = 0, dontflip = 0, phase = 0, nseg = 4, pairs_only = 0, framing = 0, bitrate = 512, bits_in = 0;
    long block = 262144;
    const char *lang = getenv("LANG");
    setlocale(LC_ALL, lang ? lang : "en_US.utf8");                       // vdecode.c:59-62 (thousands separators in the status line)
    int opt;
    while ((opt = getopt(argc, argv, "d:pi:qFB:S:Pfr:b")) != -1) {
        switch (opt) {
        case 'F': dontflip = 1; break;
        case 'q': quiet = 1; break;
        case 'p': phase = 1; break;
        case 'i': interval = atoi(optarg); break;
        case 'd': delay = atoi(optarg); break;
        case 'B': block = atol(optarg); break;
        case 'S': nseg = atoi(optarg); break;
        case 'P': pairs_only = 1; break;
        case 'f': framing = 1; break;
        case 'b': bits_in = 1; break;
        case 'r': bitrate = atoi(optarg); break;
        default: break;
        }
    }
    if (delay < 24) {                                                     // vdecode.c:86-91
        fprintf(stderr, "%s: decoder delay too small, using 200\n", argv[0]);
        delay = 200;
    } else if (delay > 1024) {
        fprintf(stderr, "%s: Warning; excessive decode delay; 1MB/bit needed\n", argv[0]);
    }
    if (block < 1024) block = 1024;
    const int ring_rows = delay + 8192;                                   // the library works through a block in chunks of (rows - delay)
    void *vd = nullptr;
    if (!pairs_only && !(framing && bits_in)) {
        vd = create_viterbi224(ring_rows);
        if (!vd) { fprintf(stderr, "%s: create_viterbi224 failed: %s\n", argv[0], v224x_last_error()); return 1; }
        init_viterbi224(vd, 0);                                           // vdecode.c:96
    }

    int taps[NTAPS];
    sync_taps(taps);
    unsigned char hist[RING];
    for (int i = 0; i < RING; i += 2) { hist[i] = G1FLIP ? 255 : 0; hist[i + 1] = G2FLIP ? 255 : 0; }    // vdecode.c:55-58
    int slot = phase;                    // ring slot of the next input symbol; its low bit is the decoder's symbol phase
    unsigned char even_sym = 0;          // the last symbol that landed on an even slot (first half of the next pair)
    int frame_count = 0, peak_in = -1000000, peak_out = -1000000;
    const int back = 2 * (delay + K - 2);

    std::vector<PairRec> pairs;
    std::vector<size_t> flip_at;         // a phase flip happened before the pair with this index (for the notice's place on stderr)
    std::vector<unsigned char> syms, bits;
    pairs.reserve(block); syms.reserve(2 * block); bits.resize(block);
    std::vector<unsigned char> inbuf(1 << 20);
    std::vector<char> outbuf;

    // per-pair state of the output side (vdecode.c:147-184)
    int startup = delay;
    unsigned long long re_encoder = 0, symerrs = 0, nbits = 0;
    // -f: framer.c's 1024-bit shift register (newest bit at the end), frame and bit counters
    unsigned char shreg[128] = {0};
    unsigned long long fr_frames = 1, fr_bits = 0, last40 = 0;
    int shpos = 0;                       // the register is kept as a ring of bits: bit index of the oldest bit
    auto frame_bit = [&](int bit) {
        // overwrite the oldest bit with the newest (ring instead of shifting 1024 bits along, framer.c:66-72)
        unsigned char &b = shreg[shpos >> 3];
        const unsigned char m = (unsigned char)(0x80u >> (shpos & 7));
        b = bit ? (b | m) : (b & ~m);
        shpos = (shpos + 1) & 1023;
        last40 = ((last40 << 1) | (unsigned long long)bit) & 0xffffffffffull;
        if (last40 == SYNCWORD) {                                        // framer.c:74
            printf("Frame %'llu at bit %'llu (%s)\n", fr_frames, fr_bits, v224host::format_hms((double)fr_bits / bitrate).c_str());
            unsigned char fr[128];
            for (int i = 0; i < 128; i++) {
                // 8 bits starting at ring bit shpos + 8 i
                const int p = (shpos + 8 * i) & 1023, sh = p & 7;
                const unsigned hi = shreg[p >> 3], lo = shreg[((p >> 3) + 1) & 127];
                fr[i] = (unsigned char)(((hi << 8 | lo) >> (8 - sh)) & 0xff);
            }
            v224host::print_frame_hex(stdout, fr, 128);
            fr_frames++;
            putchar('\n');
            fflush(stdout);
        }
        fr_bits++;
    };

    auto flush_block = [&]() -> int {
        const int n = (int)pairs.size();
        if (n == 0) {
            for (size_t f = 0; f < flip_at.size(); f++) fprintf(stderr, "%s: flipping phase\n", argv[0]);
            flip_at.clear();
            return 0;
        }
        syms.resize(2 * (size_t)n);
        for (int i = 0; i < n; i++) { syms[2 * i] = pairs[i].s0; syms[2 * i + 1] = pairs[i].s1; }
        if (pairs_only) {
            fwrite(syms.data(), 1, syms.size(), stdout);
            for (size_t f = 0; f < flip_at.size(); f++) fprintf(stderr, "%s: flipping phase\n", argv[0]);
            pairs.clear();
            flip_at.clear();
            return 0;
        }
        bits.resize(n);
        if (v224x_stream_decode_seg(vd, syms.data(), n, delay, bits.data(), nseg, -1, nullptr) < 0) {
            fprintf(stderr, "%s: decode failed: %s\n", argv[0], v224x_last_error());
            return -1;
        }
        outbuf.clear();
        size_t nf = 0;
        for (int i = 0; i < n; i++) {
            while (nf < flip_at.size() && flip_at[nf] == (size_t)i) { fprintf(stderr, "%s: flipping phase\n", argv[0]); nf++; }
            if (startup == 0) {
                const int bit = bits[i];
                if (framing) frame_bit(bit);
                else outbuf.push_back(bit ? '1' : '0');
                re_encoder = (re_encoder << 1) | (unsigned long long)bit;
            } else {
                startup--;
            }
            const int e1 = G1FLIP ^ parity64(re_encoder & POLY1), e2 = G2FLIP ^ parity64(re_encoder & POLY2);
            if (startup == 0) symerrs += (unsigned long long)((e1 ^ pairs[i].c1) + (e2 ^ pairs[i].c2));
            if (!quiet && interval != 0 && (++nbits % (unsigned long long)interval) == 0) {
                fprintf(stderr, "%s: bits %'llu; symerrs %'llu/%'d %'.3lg%%\n", argv[0], nbits, symerrs, 2 * interval,
                        100. * symerrs / (2. * interval));
                symerrs = 0;
            }
        }
        for (; nf < flip_at.size(); nf++) fprintf(stderr, "%s: flipping phase\n", argv[0]);
        if (!outbuf.empty()) fwrite(outbuf.data(), 1, outbuf.size(), stdout);
        fflush(stdout);
        pairs.clear();
        flip_at.clear();
        return 0;
    };

    if (framing && bits_in) {
        for (;;) {
            const ssize_t got = read_some(inbuf.data(), inbuf.size());
            if (got <= 0) break;
            for (ssize_t p = 0; p < got; p++) frame_bit(inbuf[p] == '1');      // framer.c:65: anything but '1' counts as 0
        }
        return 0;
    }
    for (;;) {
        const ssize_t got = read_some(inbuf.data(), inbuf.size());
        if (got <= 0) break;
        for (ssize_t p = 0; p < got; p++) {
            const unsigned char c = inbuf[p];
            hist[slot] = c;
            if ((slot & 1) == 0) even_sym = c;
            if (!dontflip) {
                // correlate the newest 34 symbols with the encoded sync pattern
                int sum = 0;
                for (int k = 0; k < NTAPS; k++) {
                    const int v = (int)hist[(RING + slot + k - (NTAPS - 1)) % RING] - 128;
                    sum += taps[k] ? v : -v;
                }
                if ((slot & 1) == 0) {
                    if (sum > peak_out) peak_out = sum;
                } else {
                    if (sum > peak_in) peak_in = sum;
                    if (++frame_count >= FRAME_SYMBOLS) {
                        // once per frame: did the other symbol phase see the stronger sync?
                        frame_count = 0;
                        if (peak_out > peak_in) {
                            if (!quiet) flip_at.push_back(pairs.size());     // the notice is printed where vdecode prints it
                            slot += (slot & 1) ? -1 : 1;          // this symbol is not decoded; the next one reuses its slot
                        }
                        peak_in = peak_out = -1000000;
                    }
                }
            }
            if (slot & 1) {
                PairRec r;
                r.s0 = even_sym; r.s1 = c;
                // (for delays beyond 2035 the reference's index goes negative -- undefined there; wrapped here)
                r.c1 = hist[(((slot - back - 1) % RING) + RING) % RING] > 128;
                r.c2 = hist[(((slot - back) % RING) + RING) % RING] > 128;
                pairs.push_back(r);
                if ((long)pairs.size() >= block && flush_block()) return 1;
            }
            slot = (slot + 1) % RING;
        }
    }
    if (flush_block()) return 1;
    fflush(stdout);
    delete_viterbi224(vd);
    return 0;
}
