// v224_common.cuh -- shared definitions of the B200 viterbi224 runtime and kernels.
//
// Code parameters follow the reference's active code block (code.h:54-63, MCQLI-24).
#pragma once
#include <cstdint>
#include <cstddef>
#include <cuda_runtime.h>

// The fused-pass translation unit is compiled twice: tiles of 64 columns (namespace v224, the lockstep multi-decoder shape)
// and tiles of 32 columns (-DV224_TILE_COLS_LOG2=5 -DV224_NS=v224t32: twice as many tiles of half the latency and a quarter
// instead of half of the previous pass as each tile's dependency -- the shape a decoder running alone wants).  Everything
// that depends on the tile width lives in that unit's namespace; the two decision-row layouts are told apart by the row tag.
#ifndef V224_NS
#define V224_NS v224
#endif
#ifndef V224_TILE_COLS_LOG2
#define V224_TILE_COLS_LOG2 6
#endif

namespace V224_NS {

constexpr int      K        = 24;                  // code.h:61
constexpr uint32_t POLY1    = 073665667u;          // code.h:59
constexpr uint32_t POLY2    = 073665665u;          // code.h:60
constexpr int      G1FLIP   = 0;                   // code.h:62
constexpr int      G2FLIP   = 1;                   // code.h:63
constexpr uint32_t NSTATES  = 1u << (K - 1);       // 2^23 path metrics
constexpr uint32_t NBFLY    = 1u << (K - 2);       // 2^22 butterflies per stage
constexpr uint32_t STATEMASK = NSTATES - 1;
constexpr uint32_t ROWWORDS = NSTATES / 32;        // 2^18 words = 1 MiB per decision row
constexpr size_t   ROWBYTES = ROWWORDS * 4;
constexpr size_t   METRICBYTES = (size_t)NSTATES * 2;   // 16 MiB

constexpr int RENORM_TRIGGER = 25000;              // viterbi224_sse2.c:351
constexpr int INIT_BIAS      = 5000;               // viterbi224_sse2.c:45 (SHRT_MIN+5000)
// The packed 16-bit fast arithmetic compares metrics through a +0x8000 bias and needs
// max - min + 510*8 to stay far below 2^15.  A legal trellis state never exceeds
// 5000 + 23*510 = 16730 (any state is reachable from the best one in 23 stages); larger spreads
// (only loadable through v224x_set_state) take the exact integer kernel instead.
constexpr long long MAX_FAST_SPREAD = 24000;

// Fused pass geometry: FK stages per HBM pass, done as two register rounds of FR stages.
constexpr int FR = 4;
constexpr int FK = 2 * FR;                         // 8 trellis stages per pass
// A tile is all 256 rows (m) x 64 consecutive columns (j) of the [256][2^15] view of the metric array (32 KiB).
// A thread holds 16 rows x NQ packed registers (2*NQ columns).  NQ = 2: 256 threads per tile at <= 64 registers,
// 4 tiles = 32 warps per SM (default).  NQ = 4: 128 threads per tile at 128 registers, 16 warps per SM -- the
// first-generation shape, kept as a build option for A/B runs (profiles/).
#ifndef V224_NQ
#define V224_NQ 2
#endif
constexpr int NQ = V224_NQ;                        // packed 2x16-bit registers per row per thread
static_assert(NQ == 1 || NQ == 2 || NQ == 4, "columns per thread");
constexpr int COLW_LOG2 = NQ == 1 ? 1 : NQ == 2 ? 2 : 3;   // log2(columns per thread)
constexpr int COLW = 1 << COLW_LOG2;
constexpr int FUSED_COLS_LOG2 = V224_TILE_COLS_LOG2;
static_assert(FUSED_COLS_LOG2 == 6 || FUSED_COLS_LOG2 == 5, "tile width");
constexpr int FUSED_TILE_COLS = 1 << FUSED_COLS_LOG2;  // columns (j) per tile
constexpr int FUSED_COLGROUPS = FUSED_TILE_COLS / COLW; // column groups (one per thread column) per tile: 16 / 8
constexpr int FUSED_THREADS   = 16 * FUSED_COLGROUPS;  // 16 row groups x column groups: 256 / 128
// CTA = FUSED_THREADS compute threads + two protocol warps (v224_acs_persist.cu) at 64 registers: 3 CTAs of 320 threads per
// SM with 64-column tiles, up to 5 CTAs of 192 threads with 32-column tiles; the round-1 -> round-2 exchange buffer is
// double-buffered (it doubles as the landing zone of the next tile's tensor copy): 3 x 67 KiB / 5 x 35 KiB of shared memory.
#ifndef V224_CTAS_PER_SM
#define V224_CTAS_PER_SM (V224_TILE_COLS_LOG2 == 6 ? 3 : V224_NQ == 1 ? 3 : 5)
#endif
constexpr int FUSED_CTAS_PER_SM = V224_CTAS_PER_SM;
#ifndef V224_XCHG_BUFS
#define V224_XCHG_BUFS ((V224_CTAS_PER_SM <= 3 || V224_TILE_COLS_LOG2 == 5) ? 2 : 1)
#endif
constexpr int XCHG_BUFS = V224_XCHG_BUFS;
// The producer warp prefetches a tile's input into its exchange buffer with one tensor copy (TMA engine)
// while the compute warps still work on the previous tile; needs the double buffer.
#ifndef V224_BULK_LOAD
#define V224_BULK_LOAD (V224_XCHG_BUFS == 2)
#endif
constexpr bool BULK_LOAD = V224_BULK_LOAD;
static_assert(!BULK_LOAD || XCHG_BUFS == 2, "bulk prefetch needs the double exchange buffer");
constexpr int FUSED_TILES     = (1 << (23 - FK)) / FUSED_TILE_COLS;   // 512 tiles per pass
constexpr int TILE_CLASSES    = 128 / FUSED_TILE_COLS; // tile t of pass n+1 reads the 256 tiles == (t >> 8) mod TILE_CLASSES of pass n

// Path-metric buffers rotate A -> B -> C -> A.  Three (not two) so that, in the persistent kernel,
// pass n+1 may already be writing while pass n is still being validated: pass n's input stays
// intact until pass n is resolved, which makes an invalidated pass restartable.
constexpr int NBUF = 3;
constexpr int PSLOTS = 4;                          // in-flight pass bookkeeping slots (ring)

// Decision-row formats (row_fmt[] tags).  0 = canonical (bit index = new-state number, the
// reference's layout, viterbi224_sse2.c:141,324); 1..8 = written by stage t of a fused pass with 64-column tiles,
// 9..16 = by stage t - 8 of a fused pass with 32-column tiles (two registers per row and thread), 17..24 = by stage t - 16 of
// the 32-column build with ONE register per row and thread, each in that kernel's thread-major layout
// (see fused_bit_address()).
constexpr uint8_t ROWFMT_CANON = 0;
constexpr uint8_t ROWFMT_FUSED_BASE = FUSED_COLS_LOG2 == 6 ? 0 : NQ == 2 ? 8 : 16;    // tag of stage t = base + t

// Device-resident control block.  One per decoder handle.  The last CTA of every pass
// ("resolver") folds the pass's statistics into it; the host only reads it back at the end of
// an ABI call.
//
// Metric representation: HBM holds P (uint16, unsigned).  The reference's int16 metric is
// R = P + O for the 64-bit offset O below.  Decisions depend on metric differences only, so
// any O is exact as long as neither side saturates; the resolver tracks the reference's
// renormalisation test (viterbi224_sse2.c:351-377) on R virtually and keeps P small.
// Per-pass statistics.  Same-address atomics serialise in L2 at a few ns each, and a pass issues thousands of
// them, so minima / maxima are spread over STAT_BUCKETS buckets, each in its own 32-byte sector.
constexpr int STAT_BUCKETS = 16;
struct PassStats {
    unsigned s0[FK + 1];                        // P of state 0 after stage t
    unsigned minP[FK + 1][STAT_BUCKETS][8];     // min of P after stage t ([FK] always, the others in careful passes); [..][0] used
    unsigned maxP[STAT_BUCKETS][8];             // max of P after the last stage; [..][0] used
};
__host__ __device__ inline void stats_reset(PassStats &s)
{
    for (int t = 0; t <= FK; t++) {
        s.s0[t] = 0;
        for (int b = 0; b < STAT_BUCKETS; b++) s.minP[t][b][0] = 0xffffffffu;
    }
    for (int b = 0; b < STAT_BUCKETS; b++) s.maxP[b][0] = 0;
}
__host__ __device__ inline unsigned stats_min(const PassStats &s, int t)
{
    unsigned m = 0xffffffffu;
    for (int b = 0; b < STAT_BUCKETS; b++) { const unsigned v = *(volatile const unsigned *)&s.minP[t][b][0]; m = v < m ? v : m; }
    return m;
}
__host__ __device__ inline unsigned stats_max(const PassStats &s)
{
    unsigned m = 0;
    for (int b = 0; b < STAT_BUCKETS; b++) { const unsigned v = *(volatile const unsigned *)&s.maxP[b][0]; m = v > m ? v : m; }
    return m;
}

// Bookkeeping of one in-flight pass of the persistent kernel.  Two 64-bit words carry everything a tile's protocol
// warp has to poll, so that one round of parallel loads decides "may this tile start":
//   done_word = [63:24] tiles done per class, 10 bits each (class = tile index mod TILE_CLASSES) | [23:11] pass number mod 2^13
//               | [10:0] tiles done      (the pass tag is written when the slot is reset for that pass)
//   pass_word = (pass number + 1) << 32 | careful << 31 | discard << 30 | measure << 29 | sub     published by the resolver of pass n-2 (or the launch)
//               measure: the pass's tiles reduce the minimum and maximum of its output.  They are only needed to keep bounds tight (the
//               subtraction that keeps P small, the saturation / spread watch), so most passes skip them and the resolver carries
//               rigorous bounds instead: a metric never decreases, and the largest one grows by at most 255 per stage (a survivor is
//               the smaller of two candidates whose branch metrics add up to 510)
//               discard: the pass cannot be invalidated (the launch / resolver has proved that the reference cannot saturate before it
//               ends and that its metric spread stays in range), so its tiles may drop their input lines from the L2
struct PassSlot {
    PassStats st;
    unsigned long long done_word;
    unsigned long long pass_word;
};
__host__ __device__ inline unsigned long long make_pass_word(int n, int careful, int sub, int discard = 0, int measure = 1)
{
    return ((unsigned long long)(unsigned)(n + 1) << 32) | ((unsigned long long)(careful ? 1u : 0u) << 31) |
           ((unsigned long long)(discard ? 1u : 0u) << 30) | ((unsigned long long)(measure ? 1u : 0u) << 29) | ((unsigned)sub & 0x1fffffffu);
}
__host__ __device__ inline int pass_word_sub(unsigned long long w) { return (int)(w & 0x1fffffffu); }
__host__ __device__ inline bool pass_word_measure(unsigned long long w) { return ((w >> 29) & 1u) != 0; }
constexpr int MEASURE_EVERY = 4;                   // a pass in this many reduces min / max even when nothing else asks for it
constexpr long long MAX_GROWTH_PER_STAGE = 255;    // of the largest path metric (see pass_word)
__host__ __device__ inline bool pass_word_careful(unsigned long long w) { return ((w >> 31) & 1u) != 0; }
__host__ __device__ inline bool pass_word_discard(unsigned long long w) { return ((w >> 30) & 1u) != 0; }
// A pass may drop its input lines only if nothing can invalidate it: the reference cannot saturate before the pass ends (the
// largest metric grows by at most 510 per stage; non-careful passes are not even started otherwise) and the spread, which also
// grows by at most 510 per stage, stays below the fast path's limit (`stages_ahead` = stages between the statistics maxR and
// spread come from and the end of the pass).
__host__ __device__ inline int discard_ok(long long maxR, long long spread, int stages_ahead)
{
    return maxR + 510ll * stages_ahead <= 32767 && spread + 510ll * stages_ahead <= MAX_FAST_SPREAD;
}
__host__ __device__ inline unsigned done_class_count(unsigned long long w, unsigned cls) { return (unsigned)(w >> (24 + 10 * cls)) & 0x3ffu; }
__host__ __device__ inline unsigned long long done_increment(unsigned cls) { return 1ull | (1ull << (24 + 10 * cls)); }
__host__ __device__ inline unsigned long long done_word_fresh(int pass) { return (unsigned long long)((unsigned)pass & 0x1fffu) << 11; }
__host__ __device__ inline unsigned done_word_pass(unsigned long long w) { return (unsigned)(w >> 11) & 0x1fffu; }
__host__ __device__ inline unsigned done_total(unsigned long long w) { return (unsigned)w & 0x7ffu; }
static_assert(TILE_CLASSES <= 4 && FUSED_TILES <= 1024, "done_word layout");
struct PersistCtl {
    unsigned next_item;     // dynamic work queue head: item = pass * 512 + order index
    unsigned resolved_upto; // number of passes resolved (in order)
    int stop_pass;          // passes >= stop_pass must not run (saturation watch / invalidated pass)
    int npasses;
    int force_careful;
    int no_discard;         // passes never drop their consumed input lines from the L2 (option)
    int measure_all;        // every pass reduces min / max (option "measure_all")
    unsigned lbP, ubP;      // bounds of the stored metrics (P) of the last resolved pass's output: min >= lbP, max <= ubP
    const uint32_t *passtab; // the launch's per-pass tables (the resolver reads the growth bound of state 0 from them)
    long long Ostore;       // R = P_stored + Ostore for the output of the last resolved pass
    long long maxR_prev;    // largest reference metric at the output of the last resolved pass
    PassSlot slot[PSLOTS];
};

struct Ctl {
    long long O;            // R = P + O
    long long renormals;    // the reference's running `renormals` (viterbi224_sse2.c:33,367)
    long long T;            // trellis stages since init (ring position = T % len)
    int  renorm_count;      // renormalisations since the last init (update returns the difference over the call)
    int  pad0;
    int  sub;               // amount the next pass subtracts from every P while loading
    int  cur;               // which metric buffer is the "old" one
    long long R0;           // reference metric of state 0 after the last stage (renorm trigger watch)
    long long maxR;         // largest reference metric after the last stage (saturation watch)
    long long spread;       // max - min after the last stage (packed-arithmetic range watch)
    int  error;             // sticky: internal invariant violated
    unsigned ticket;        // CTA completion counter of the running pass
    // counters for tests / bench
    unsigned n_fused, n_single, n_careful, n_sat, n_invalidated;
    // ---- everything above is what the host mirrors after each call (CTL_HOST_BYTES) ----
    PassStats st;           // statistics of the running per-pass / single-stage kernel, reset by its resolver
    PersistCtl pc;
    unsigned spec_first;    // per-bit streaming: (stage counter after the stage) << 1 | first decision of the speculative decodebit walk
    unsigned pad1;
};
constexpr size_t CTL_HOST_BYTES = offsetof(Ctl, st);

// ---- thread <-> (row group, column group) maps of the two register rounds --------------------
// Round 1 (stages 1-4): thread = (ml, g), a warp = 32/CG row groups x CG column groups: every 64/128-bit load
// instruction reads whole 128-byte lines.  Round 2 (stages 5-8): thread = (mh, g), a warp = 4 row groups x 8
// column groups, so that the 32-byte metric stores of four lanes make one 128-byte line.
__host__ __device__ inline void round1_map(uint32_t tid, uint32_t &thr, uint32_t &g)
{
    thr = tid / FUSED_COLGROUPS;
    g = tid % FUSED_COLGROUPS;
}
__host__ __device__ inline void round2_map(uint32_t tid, uint32_t &thr, uint32_t &g)
{
    const uint32_t w = tid >> 5, lane = tid & 31;
    thr = (w / (FUSED_COLGROUPS / 8)) * 4 + (lane >> 3);
    g = (w % (FUSED_COLGROUPS / 8)) * 8 + (lane & 7);
}
__host__ __device__ inline uint32_t round2_tid_cg(uint32_t thr, uint32_t g, uint32_t colgroups)
{
    const uint32_t w = (thr >> 2) * (colgroups / 8) + (g >> 3), lane = ((thr & 3) << 3) | (g & 7);
    return (w << 5) | lane;
}
__host__ __device__ inline uint32_t round2_tid(uint32_t thr, uint32_t g) { return round2_tid_cg(thr, g, FUSED_COLGROUPS); }

// shared-memory exchange element (row m, column group g) between the two register rounds, in units of NQ words.  A
// half-warp of round 2 reads rows 16 apart (same banks): 64-column tiles swap the 64-byte halves of the rows with odd
// mh bit 1 position, 32-column tiles (64-byte rows) swap neighbouring rows of odd mh -- either way the rows a warp reads
// together land on disjoint banks, and a round-1 thread still writes exactly the rows its own warp read (in place).
__host__ __device__ inline uint32_t xchg_index(uint32_t m, uint32_t g)
{
    if (NQ == 4) return m * FUSED_COLGROUPS + g;
    if (NQ == 1) {
        // 64-byte rows of 16 one-word elements: a round-2 warp reads 8 elements (32 bytes) of four rows 16 apart -- odd mh swap
        // neighbouring rows, mh bit 1 swaps the 32-byte halves: the four pieces land on the four quarters of the banks
        const uint32_t mh = m >> 4;
        return (m ^ (mh & 1u)) * FUSED_COLGROUPS + (g ^ ((mh & 2u) << 2));
    }
    if (FUSED_COLGROUPS == 16) return m * FUSED_COLGROUPS + (g ^ ((m >> 1) & 8u));
    return (m ^ ((m >> 4) & 1u)) * FUSED_COLGROUPS + g;
}

// Decision-row formats: 0 = canonical; otherwise written by stage t of a fused pass (tag = t for 64-column tiles, 8 + t
// for 32-column tiles).  Fused rows hold the COMPLEMENT of the decision bit (the sign bit the butterfly produces is the
// inverted decision; flipping it once per traceback read is cheaper than once per state update).
constexpr uint32_t FUSED_ROWS_COMPLEMENTED = 1u;
// Where a fused-format decision bit lives: state s after stage t -> bit index inside the 2^23-bit row.
// Slot fields (23 bits): mh[4] | ml[4] | j[15], j = tile | g | q | h.  The thread (tile, tid) owns the NQ
// consecutive words starting at word (tile * threads + tid) * NQ; inside them
//   word = side * NQ/2 + q/2, byte = (q & 1) * 2 + h, bit = pair index
// where `inner` is the 4-bit row index a thread holds in that round (mh in round 1, ml in round 2), the stage's
// butterfly pairs the two rows that differ in bit sb of `inner`, side = that bit, pair index = the other three.
__host__ __device__ inline uint32_t fused_bit_address(int fmt, uint32_t s)
{
    // geometry of the build that wrote the row: 1..8 64-column tiles, 9..16 32-column tiles, 17..24 32-column tiles with one
    // register per row and thread (all other builds: NQ registers, NQ = this build's)
    const int cols_log2 = fmt > 8 ? 5 : 6;
    const int nq = fmt > 16 ? 1 : (NQ == 1 ? 2 : NQ);
    const int t = fmt > 16 ? fmt - 16 : fmt > 8 ? fmt - 8 : fmt;
    const int colw_log2 = nq == 1 ? 1 : nq == 2 ? 2 : 3;
    const uint32_t colgroups = (1u << cols_log2) >> colw_log2, threads = 16 * colgroups;
    uint32_t p = ((s >> t) | (s << (23 - t))) & STATEMASK;      // the slot its survivor sits in during the pass
    const uint32_t mh = (p >> 19) & 15, ml = (p >> 15) & 15, j = p & 32767;
    const uint32_t tile = j >> cols_log2, g = (j & ((1u << cols_log2) - 1)) >> colw_log2;
    const uint32_t q = (j & ((1u << colw_log2) - 1)) >> 1, h = j & 1;
    const bool r1 = t <= FR;
    const uint32_t inner = r1 ? mh : ml;
    const uint32_t tid = r1 ? ml * colgroups + g : round2_tid_cg(mh, g, colgroups);
    const int sb = FR - 1 - ((t - 1) % FR);
    const uint32_t side = (inner >> sb) & 1;
    const uint32_t pidx = ((inner >> (sb + 1)) << sb) | (inner & ((1u << sb) - 1));
    if (nq == 1)     // one word per thread and stage: byte = side * 2 + half
        return (tile * threads + tid) * 32 + ((side << 1) | h) * 8 + pidx;
    const uint32_t word = (tile * threads + tid) * nq + side * (nq / 2) + (q >> 1);
    return word * 32 + (((q & 1) << 1) | h) * 8 + pidx;
}

} // namespace V224_NS
