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
// Extra options: -B pairs  block size (default 262144 per GPU; latency = one block), -S n  decoders in lockstep per GPU (default 4),
// -G n  GPUs: every block is cut into n time segments decoded side by side, one per GPU (v224x_multi_stream_decode: every
// GPU-to-GPU hand-over verified on the device, a range whose decoder had not converged is decoded again -- the output
// is the one-GPU output), -D a,b,..  the CUDA device ordinals to use with -G (default 0 .. n-1; an ordinal may repeat: its
// ranges then share that GPU), -v  a summary of the hand-over checks on stderr at the end,
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
#include "pairing.h"

namespace {

using v224host::SymbolPairer;
constexpr int K = SymbolPairer::K;
constexpr unsigned long long POLY1 = SymbolPairer::POLY1, POLY2 = SymbolPairer::POLY2, SYNCWORD = SymbolPairer::SYNCWORD;
constexpr int G1FLIP = SymbolPairer::G1FLIP, G2FLIP = SymbolPairer::G2FLIP;

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

} // namespace

int main(int argc, char *argv[])
{
    int delay = 200, interval = 1024, quiet = 0, dontflip = 0, phase = 0, nseg = 4, pairs_only = 0, framing = 0, bitrate = 512, bits_in = 0;
    int ngpu = 1, verbose = 0;
    const char *devlist = nullptr;
    long block = 0;
    const char *lang = getenv("LANG");
    setlocale(LC_ALL, lang ? lang : "en_US.utf8");                       // vdecode.c:59-62 (thousands separators in the status line)
    int opt;
    while ((opt = getopt(argc, argv, "d:pi:qFB:S:Pfr:bG:vD:")) != -1) {
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
        case 'G': ngpu = atoi(optarg); break;
        case 'v': verbose = 1; break;
        case 'D': devlist = optarg; break;
        default: break;
        }
    }
    if (delay < 24) {                                                     // vdecode.c:86-91
        fprintf(stderr, "%s: decoder delay too small, using 200\n", argv[0]);
        delay = 200;
    } else if (delay > 1024) {
        fprintf(stderr, "%s: Warning; excessive decode delay; 1MB/bit needed\n", argv[0]);
    }
    if (ngpu < 1) ngpu = 1;
    if (block <= 0) block = 262144l * ngpu;
    if (block < 1024) block = 1024;
    const int ring_rows = delay + 8192;                                   // the library works through a block in chunks of (rows - delay)
    void *vd = nullptr;
    v224x_multi *vm = nullptr;
    if (!pairs_only && !(framing && bits_in)) {
        if (ngpu > 1) {
            std::vector<int> devs;
            for (const char *q = devlist; q && *q;) {
                devs.push_back(atoi(q));
                q = strchr(q, ',');
                if (q) q++;
            }
            if (!devs.empty() && (int)devs.size() != ngpu) { fprintf(stderr, "%s: -D lists %zu devices, -G says %d\n", argv[0], devs.size(), ngpu); return 1; }
            vm = v224x_multi_create(devs.empty() ? nullptr : devs.data(), ngpu, ring_rows);
            if (!vm) { fprintf(stderr, "%s: v224x_multi_create failed: %s\n", argv[0], v224x_last_error()); return 1; }
            v224x_multi_init(vm, 0);
        } else {
            vd = create_viterbi224(ring_rows);
            if (!vd) { fprintf(stderr, "%s: create_viterbi224 failed: %s\n", argv[0], v224x_last_error()); return 1; }
            init_viterbi224(vd, 0);                                       // vdecode.c:96
        }
    }

    // vdecode's pairing and phase-flip logic (vdecode.c:101-140), run ahead of the decoder
    SymbolPairer pairer(phase, dontflip != 0, delay);
    std::vector<unsigned long long> flips_abs;      // index (since the start of the stream) of the first pair after each phase flip
    unsigned long long pairs_before = 0;            // pairs of earlier blocks
    std::vector<size_t> flip_at;         // a phase flip happened before the pair with this index (for the notice's place on stderr)
    std::vector<unsigned char> syms(2 * ((size_t)block + 2)), cmps(2 * ((size_t)block + 2)), bits(block + 2);
    size_t npairs = 0;                   // pairs collected for the running block
    std::vector<unsigned char> inbuf(1 << 20);
    std::vector<char> outbuf;
    long long tot_verified = 0, tot_redone = 0, tot_inner_verified = 0, tot_inner_redone = 0, tot_blocks = 0;
    int worst_spread = 0;

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
        const int n = (int)npairs;
        flip_at.clear();
        for (unsigned long long f : flips_abs) flip_at.push_back((size_t)(f - pairs_before));
        flips_abs.clear();
        pairs_before += npairs;
        npairs = 0;
        if (n == 0) {
            for (size_t f = 0; f < flip_at.size() && !quiet; f++) fprintf(stderr, "%s: flipping phase\n", argv[0]);
            return 0;
        }
        if (pairs_only) {
            fwrite(syms.data(), 1, 2 * (size_t)n, stdout);
            for (size_t f = 0; f < flip_at.size() && !quiet; f++) fprintf(stderr, "%s: flipping phase\n", argv[0]);
            return 0;
        }
        if (vm) {
            v224x_multi_report rep;
            if (v224x_multi_stream_decode(vm, syms.data(), n, delay, bits.data(), nseg, -1, &rep) < 0) {
                fprintf(stderr, "%s: decode failed: %s\n", argv[0], v224x_last_error());
                return -1;
            }
            tot_verified += rep.handovers_verified; tot_redone += rep.ranges_redone;
            tot_inner_verified += rep.inner_verified; tot_inner_redone += rep.inner_redone;
            if (rep.worst_spread > worst_spread) worst_spread = rep.worst_spread;
        } else {
            v224x_seg_report rep;
            if (v224x_stream_decode_seg(vd, syms.data(), n, delay, bits.data(), nseg, -1, &rep) < 0) {
                fprintf(stderr, "%s: decode failed: %s\n", argv[0], v224x_last_error());
                return -1;
            }
            tot_inner_verified += rep.verified; tot_inner_redone += rep.redone;
            if (rep.worst_spread > worst_spread) worst_spread = rep.worst_spread;
        }
        tot_blocks++;
        outbuf.clear();
        size_t nf = 0;
        for (int i = 0; i < n; i++) {
            while (nf < flip_at.size() && flip_at[nf] == (size_t)i) { if (!quiet) fprintf(stderr, "%s: flipping phase\n", argv[0]); nf++; }
            if (startup == 0) {
                const int bit = bits[i];
                if (framing) frame_bit(bit);
                else outbuf.push_back(bit ? '1' : '0');
                re_encoder = (re_encoder << 1) | (unsigned long long)bit;
            } else {
                startup--;
            }
            const int e1 = G1FLIP ^ parity64(re_encoder & POLY1), e2 = G2FLIP ^ parity64(re_encoder & POLY2);
            if (startup == 0) symerrs += (unsigned long long)((e1 ^ cmps[2 * i]) + (e2 ^ cmps[2 * i + 1]));
            if (!quiet && interval != 0 && (++nbits % (unsigned long long)interval) == 0) {
                fprintf(stderr, "%s: bits %'llu; symerrs %'llu/%'d %'.3lg%%\n", argv[0], nbits, symerrs, 2 * interval,
                        100. * symerrs / (2. * interval));
                symerrs = 0;
            }
        }
        for (; nf < flip_at.size(); nf++) if (!quiet) fprintf(stderr, "%s: flipping phase\n", argv[0]);
        if (!outbuf.empty()) fwrite(outbuf.data(), 1, outbuf.size(), stdout);
        fflush(stdout);
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
        for (ssize_t p = 0; p < got;) {
            // n symbols give at most n / 2 + 1 pairs: never more than the block has room for
            const size_t room = (size_t)block - npairs;
            const size_t take = std::min((size_t)(got - p), room > 1 ? 2 * (room - 1) : (size_t)1);
            npairs += pairer.feed(inbuf.data() + p, take, syms.data() + 2 * npairs, cmps.data() + 2 * npairs, &flips_abs);
            p += (ssize_t)take;
            if ((long)npairs >= block && flush_block()) return 1;
        }
    }
    if (flush_block()) return 1;
    fflush(stdout);
    if (verbose)
        fprintf(stderr, "%s: %lld blocks on %d GPU(s); GPU-to-GPU hand-overs verified %lld, ranges redone %lld; lockstep hand-overs verified %lld, "
                "segments redone %lld; worst snapshot spread %d; residual differences vs the sequential decode: 0 (by construction)\n",
                argv[0], tot_blocks, ngpu, tot_verified, tot_redone, tot_inner_verified, tot_inner_redone, worst_spread);
    if (vm) v224x_multi_delete(vm);
    delete_viterbi224(vd);
    return 0;
}
