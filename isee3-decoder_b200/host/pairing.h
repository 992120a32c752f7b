// pairing.h -- the host side of the reference's streaming driver: which received symbols form the pairs that go to
// update_viterbi224_blk (vdecode.c:101-140,186).  Plain C++, no CUDA.  Shared by the block driver
// (host/vdecode_block.cpp) and by the library's v224x_pair_symbols() (csrc/v224_pairing.cpp).
//
// vdecode.c keeps the last 4096 symbols in a ring (`oldsymbols`), correlates the newest 34 with the encoded sync word
// after EVERY symbol, remembers the strongest correlation seen on even and on odd ring slots, and once per 2048
// odd-slot symbols compares the two: if the even ("out of phase") slots saw the stronger peak, the ring position steps
// back by one (vdecode.c:126-133).  The symbol that triggered the decision is then not decoded, the next input symbol
// overwrites its slot and is decoded together with the stale even-slot symbol -- one garbage pair -- and from there on
// the pairing is shifted by one symbol.  The decision uses received symbols only, never decoder output, so the whole
// pair sequence can be produced ahead of the decoder.
//
// The correlation is the expensive part (34 taps per symbol).  Between two decisions no flip can happen, so it is
// computed for a whole run of symbols at once over a linear copy of the history (tap-major loops the compiler
// vectorises); everything else is a few operations per symbol.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>

namespace v224host {

class SymbolPairer {
public:
    // code constants of the reference's active code block (code.h:54-63)
    static constexpr int K = 24;
    static constexpr unsigned long long POLY1 = 073665667ull, POLY2 = 073665665ull;
    static constexpr int G1FLIP = 0, G2FLIP = 1;
    static constexpr int FRAME_SYMBOLS = 2048;        // odd-slot symbols between two phase decisions (vdecode.c:14-15,120)
    static constexpr int NTAPS = 34;                  // vdecode.c:16
    static constexpr int RING = 4096;                 // vdecode.c:20
    static constexpr unsigned long long SYNCWORD = 0x12fc819fbeull;   // decode.c:24

    // start_phase: vdecode -p (vdecode.c:77); dontflip: vdecode -F (vdecode.c:71); delay: decode delay, only for the
    // positions of the re-encode comparison symbols (vdecode.c:176-177)
    SymbolPairer(int start_phase, bool dontflip, int delay) : slot_(start_phase & 1), dontflip_(dontflip), back_(2 * (delay + K - 2))
    {
        for (int i = 0; i < RING; i += 2) { ring_[i] = G1FLIP ? 255 : 0; ring_[i + 1] = G2FLIP ? 255 : 0; }   // vdecode.c:55-58
        sync_taps(taps_);
        // linear history in front of the first symbol: the preset ring slots slot-33 .. slot-1
        for (int k = 0; k < NTAPS - 1; k++) tail_[k] = (int16_t)((int)ring_[(RING + slot_ - (NTAPS - 1) + k) % RING] - 128);
    }

    // The 34 encoded sync symbols (vdecode.c:27-30 lists them as constants; they are the tail of encode(SYNCWORD)).
    static void sync_taps(int taps[NTAPS])
    {
        int sym[80];
        unsigned long long reg = 0;
        for (int i = 39; i >= 0; i--) {
            reg = (reg << 1) | ((SYNCWORD >> i) & 1);
            sym[2 * (39 - i)] = G1FLIP ^ __builtin_parityll(reg & POLY1);
            sym[2 * (39 - i) + 1] = G2FLIP ^ __builtin_parityll(reg & POLY2);
        }
        for (int k = 0; k < NTAPS; k++) taps[k] = sym[80 - NTAPS + k];
    }

    // Consume n received symbols.  For every pair handed to the decoder: 2 bytes appended at syms_out (the pair) and, if
    // cmp_out is not NULL, 2 bytes at cmp_out (the hard-sliced history symbols vdecode compares the re-encoded pair
    // with, vdecode.c:176-177).  Both need room for n/2 + 1 pairs.  Every phase flip appends the index of the next pair
    // (counted since construction) to *flips, if given.  Returns the number of pairs produced by this call.
    // pre (optional): pre[i] = correlation of in[i-33 .. i] taken over the input as it stands (correlate_block(), which
    // callers may run on slices of the buffer in parallel); it is used wherever the 34 newest history symbols ARE the 34
    // newest input symbols, i.e. everywhere except right behind the start of the stream and behind a dropped symbol.
    size_t feed(const unsigned char *in, size_t n, unsigned char *syms_out, unsigned char *cmp_out, std::vector<unsigned long long> *flips,
                const int16_t *pre = nullptr)
    {
        size_t npairs = 0;
        while (n) {
            size_t count = n;
            if (!dontflip_) {
                // no decision can fall inside the run: it ends at the odd-slot symbol that completes the frame count
                const size_t odd_needed = (size_t)(FRAME_SYMBOLS - frame_count_);
                count = std::min(n, (slot_ & 1) ? 2 * odd_needed - 1 : 2 * odd_needed);
                correlate_run(in, count, pre);
                if (pre) pre += count;
            }
            if (cmp_out) npairs += emit_run_with_history(in, count, syms_out + 2 * npairs, cmp_out + 2 * npairs, flips, npairs);
            else         npairs += emit_run(in, count, syms_out + 2 * npairs, flips, npairs);
            if (!dontflip_) carry_tail(in, count);
            in += count;
            n -= count;
        }
        pairs_total_ += npairs;
        return npairs;
    }

    // pre[i] for i in [from, to): the correlation of in[i-33 .. i] (positions before the buffer count as erasures; feed()
    // never uses those).  Independent of all state: callers may run it on slices of a buffer in parallel.
    static void correlate_block(const unsigned char *in, size_t from, size_t to, int16_t *pre)
    {
        int taps[NTAPS];
        sync_taps(taps);
        constexpr size_t CH = 1 << 15;
        std::vector<int16_t> lin(CH + NTAPS - 1 + 16);
        for (size_t a = from; a < to; a += CH) {
            const size_t b = std::min(to, a + CH);
            for (size_t k = 0; k < NTAPS - 1; k++) lin[k] = a + k >= NTAPS - 1 ? (int16_t)((int)in[a + k - (NTAPS - 1)] - 128) : (int16_t)0;
            widen(in + a, b - a, lin.data() + NTAPS - 1);
            correlate(lin.data(), b - a, taps, pre + a);
        }
    }

    int phase() const { return slot_ & 1; }
    unsigned long long pairs_total() const { return pairs_total_; }

private:
    // sums[i] = correlation of the 34 newest symbols after in[i] arrived (vdecode.c:111-117); peaks per slot parity
    void correlate_run(const unsigned char *in, size_t count, const int16_t *pre)
    {
        // the first `own` symbols of the run still see history that is not plain input (preset ring, a dropped symbol)
        const size_t own = pre ? std::min(count, stale_) : count;
        lin_.resize(own + NTAPS - 1 + 16);
        std::memcpy(lin_.data(), tail_, (NTAPS - 1) * sizeof(int16_t));
        widen(in, own, lin_.data() + NTAPS - 1);
        acc_.resize(own + 16);
        correlate(lin_.data(), own, taps_, acc_.data());
        stale_ -= std::min(stale_, count);
        int pe = -1000000, po = -1000000;              // peaks on in[even i] / in[odd i]
        peaks(acc_.data(), 0, own, pe, po);
        if (pre) peaks(pre, own, count, pe, po);
        // in[i] lands on slot slot_ + i: even slots feed the out-of-phase peak, odd slots the in-phase one
        if (slot_ & 1) std::swap(pe, po);              // in[0] sits on an odd slot
        peak_out_ = std::max(peak_out_, pe);
        peak_in_ = std::max(peak_in_, po);
    }
    static void peaks(const int16_t *a, size_t from, size_t to, int &pe, int &po)
    {
        int16_t me = -32768, mo = -32768;
        size_t i = from;
        if (i < to && (i & 1)) { mo = std::max(mo, a[i]); i++; }
        for (; i + 1 < to; i += 2) { me = std::max(me, a[i]); mo = std::max(mo, a[i + 1]); }
        if (i < to) me = std::max(me, a[i]);
        if (me > -32768) pe = std::max(pe, (int)me);
        if (mo > -32768) po = std::max(po, (int)mo);
    }
    static void widen(const unsigned char *in, size_t count, int16_t *out)
    {
        for (size_t i = 0; i < count; i++) out[i] = (int16_t)((int)in[i] - 128);
    }
    // tap-major: every inner loop is a plain vector add / subtract over the run (|sum| <= 34 * 128 fits 16 bits)
#if defined(__GNUC__) && defined(__x86_64__) && !defined(__CUDACC__)
    __attribute__((target_clones("avx2", "default")))
#endif
    static void correlate(const int16_t *lin, size_t count, const int *taps, int16_t *acc)
    {
        for (size_t i = 0; i < count; i++) acc[i] = 0;
        for (int k = 0; k < NTAPS; k++) {
            const int16_t *x = lin + k;
            if (taps[k]) for (size_t i = 0; i < count; i++) acc[i] = (int16_t)(acc[i] + x[i]);
            else         for (size_t i = 0; i < count; i++) acc[i] = (int16_t)(acc[i] - x[i]);
        }
    }
    // The pairs of one run, symbol by symbol as vdecode.c:107-186 goes through them, with the 4096-symbol ring kept up
    // to date: the re-encode comparison reads it (and shows its size: a delay beyond the ring reads newer symbols).
    size_t emit_run_with_history(const unsigned char *in, size_t count, unsigned char *syms_out, unsigned char *cmp_out,
                                 std::vector<unsigned long long> *flips, size_t before)
    {
        size_t npairs = 0;
        for (size_t i = 0; i < count; i++) {
            const unsigned char c = in[i];
            ring_[slot_] = c;
            bool decoded = (slot_ & 1) != 0;
            if (!decoded) {
                even_sym_ = c;
            } else if (!dontflip_ && ++frame_count_ >= FRAME_SYMBOLS) {
                // once per frame: did the other symbol phase see the stronger sync?  (vdecode.c:120-136)
                frame_count_ = 0;
                if (peak_out_ > peak_in_) {
                    if (flips) flips->push_back(pairs_total_ + before + npairs);
                    decoded = false;           // this symbol is not decoded; the next one reuses its slot
                    slot_ -= 1;
                    dropped_last_ = true;
                }
                peak_in_ = peak_out_ = -1000000;
            }
            if (decoded) {
                syms_out[2 * npairs] = even_sym_;
                syms_out[2 * npairs + 1] = c;
                // (for delays beyond 2035 the reference's index goes negative -- undefined there; wrapped here)
                cmp_out[2 * npairs] = ring_[(slot_ - back_ - 1) & (RING - 1)] > 128;
                cmp_out[2 * npairs + 1] = ring_[(slot_ - back_) & (RING - 1)] > 128;
                npairs++;
            }
            slot_ = (slot_ + 1) & (RING - 1);
        }
        return npairs;
    }
    // The same without the ring (nobody asked for the comparison symbols): the pair stream is the received stream with
    // the symbols dropped by phase flips taken out, so a run is one copy.
    size_t emit_run(const unsigned char *in, size_t count, unsigned char *syms_out, std::vector<unsigned long long> *flips, size_t before)
    {
        size_t npairs = 0, i = 0;
        if ((slot_ & 1) && count) {                    // the run's first symbol completes the pair of the held even symbol
            syms_out[0] = even_sym_;
            syms_out[1] = in[0];
            npairs = 1;
            i = 1;
        }
        const size_t whole = (count - i) / 2;
        std::memcpy(syms_out + 2 * npairs, in + i, 2 * whole);
        npairs += whole;
        i += 2 * whole;
        if (whole) even_sym_ = in[i - 2];
        if (i < count) even_sym_ = in[i];              // a trailing even-slot symbol waits for its partner
        const size_t odd_in_run = (count + (slot_ & 1)) / 2;
        slot_ = (int)((slot_ + count) & (RING - 1));
        if (!dontflip_ && (frame_count_ += (int)odd_in_run) >= FRAME_SYMBOLS) {
            // the run ended on the odd-slot symbol that completes the frame (vdecode.c:120-136)
            frame_count_ = 0;
            if (peak_out_ > peak_in_) {
                npairs -= 1;                           // that symbol is not decoded; the next one reuses its slot
                if (flips) flips->push_back(pairs_total_ + before + npairs);
                slot_ = (slot_ - 1) & (RING - 1);
                dropped_last_ = true;
            }
            peak_in_ = peak_out_ = -1000000;
        }
        return npairs;
    }
    // the 33 newest history symbols in front of the next run; a symbol dropped by a phase flip is not history
    void carry_tail(const unsigned char *in, size_t count)
    {
        const size_t kept = count - (dropped_last_ ? 1 : 0);
        if (dropped_last_) stale_ = NTAPS - 1;         // the window of the next 33 symbols skips the dropped one
        dropped_last_ = false;
        constexpr size_t H = NTAPS - 1;
        if (kept >= H) {
            for (size_t k = 0; k < H; k++) tail_[k] = (int16_t)((int)in[kept - H + k] - 128);
        } else {
            std::memmove(tail_, tail_ + kept, (H - kept) * sizeof(int16_t));
            for (size_t k = 0; k < kept; k++) tail_[H - kept + k] = (int16_t)((int)in[k] - 128);
        }
    }

    unsigned char ring_[RING];
    int taps_[NTAPS];
    int16_t tail_[NTAPS - 1];
    int slot_;
    bool dontflip_;
    int back_;
    unsigned char even_sym_ = 0;       // the last symbol that landed on an even slot (first half of the next pair)
    int frame_count_ = 0, peak_in_ = -1000000, peak_out_ = -1000000;
    bool dropped_last_ = false;
    size_t stale_ = NTAPS - 1;         // upcoming symbols whose 34-symbol window still holds non-input history
    unsigned long long pairs_total_ = 0;
    std::vector<int16_t> lin_, acc_;
};

} // namespace v224host
