// v224_common.cuh -- shared definitions of the B200 viterbi224 runtime and kernels.
//
// Code parameters follow the reference's active code block (code.h:54-63, MCQLI-24).
#pragma once
#include <cstdint>
#include <cstddef>
#include <cuda_runtime.h>

namespace v224 {

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
// Tile width.  64 columns (512 tiles of 128 threads, 4 CTAs/SM) measured 17.3 us per pass on B200,
// 32 columns (1024 tiles of 64 threads, 8 CTAs/SM: better balanced, 7 vs 6 tiles per SM) 19.2 us:
// the finer tiles lose more to per-tile overhead than they gain in balance (profiles/).
#ifndef V224_TILE_COLS_LOG2
#define V224_TILE_COLS_LOG2 6
#endif
constexpr int FUSED_COLS_LOG2 = V224_TILE_COLS_LOG2;
constexpr int FUSED_TILE_COLS = 1 << FUSED_COLS_LOG2;  // columns (j) per tile
constexpr int FUSED_COLGROUPS = FUSED_TILE_COLS / 8;   // 8 columns (4 packed registers) per thread
constexpr int FUSED_THREADS   = 16 * FUSED_COLGROUPS;  // 16 row groups x column groups
constexpr int FUSED_CTAS_PER_SM = 512 / FUSED_THREADS; // 16 warps per SM at 128 registers
constexpr int FUSED_TILES     = (1 << (23 - FK)) / FUSED_TILE_COLS;   // tiles per pass
constexpr int TILE_CLASSES    = 128 / FUSED_TILE_COLS; // tile t of pass n+1 reads the tiles == (t >> 8) mod TILE_CLASSES of pass n
static_assert(FUSED_TILE_COLS == 32 || FUSED_TILE_COLS == 64, "tile width");

// Path-metric buffers rotate A -> B -> C -> A.  Three (not two) so that, in the persistent kernel,
// pass n+1 may already be writing while pass n is still being validated: pass n's input stays
// intact until pass n is resolved, which makes an invalidated pass restartable.
constexpr int NBUF = 3;
constexpr int PSLOTS = 4;                          // in-flight pass bookkeeping slots (ring)

// Decision-row formats (row_fmt[] tags).  0 = canonical (bit index = new-state number, the
// reference's layout, viterbi224_sse2.c:141,324); other values = written by a stage of a fused
// pass in that kernel's thread-major layout (see fused_bit_address()).
constexpr uint8_t ROWFMT_CANON = 0;

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

// Bookkeeping of one in-flight pass of the persistent kernel.
struct PassSlot {
    PassStats st;
    unsigned done[TILE_CLASSES];   // finished tiles by (tile index mod TILE_CLASSES): what the next pass waits on
    unsigned done_total;
    int sub;                // what this pass subtracts from every P while loading
    int careful;            // this pass records per-stage minima
};
struct PersistCtl {
    unsigned next_item;     // dynamic work queue head: item = pass * 512 + order index
    unsigned next_rank;     // balanced mode: SM ranks handed out in arrival order
    int sm_rank[256];       // balanced mode: rank of SM %smid (-1 until its first CTA arrives)
    unsigned sm_slots[256]; // balanced mode: CTAs that arrived on SM %smid
    unsigned resolved_upto; // number of passes resolved (in order)
    int stop_pass;          // passes >= stop_pass must not run (saturation watch / invalidated pass)
    int npasses;
    int force_careful;
    int pad;
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
};
constexpr size_t CTL_HOST_BYTES = offsetof(Ctl, st);

// ---- tile partitions of the 4096 column groups (8 columns each) -------------------------------
// UNIFORM: tile t = groups [t*CG, (t+1)*CG), CG = FUSED_COLGROUPS.
// BALANCED (B200, 148 SMs x 4 CTAs): every SM gets 14 or 13 warps' worth of work instead of 16 or 12:
//   SM rank r < 124 owns 28 consecutive groups as tiles of 8,8,6,6; rank r >= 124 owns 26 as 8,6,6,6
//   (124*28 + 24*26 = 4096).  A tile of 6 groups uses 3 warps of its CTA, one of 8 groups all 4.
constexpr int BAL_SMS = 148, BAL_CTAS_PER_SM = 4, BAL_BIG_SMS = 124, BAL_TILES = BAL_SMS * BAL_CTAS_PER_SM;
__host__ __device__ inline void balanced_tile(uint32_t rank, uint32_t slot, uint32_t &g0, uint32_t &ncg)
{
    const bool big = rank < BAL_BIG_SMS;
    const uint32_t base = big ? rank * 28u : BAL_BIG_SMS * 28u + (rank - BAL_BIG_SMS) * 26u;
    if (big) { ncg = slot < 2 ? 8u : 6u; g0 = base + (slot < 2 ? slot * 8u : 16u + (slot - 2) * 6u); }
    else     { ncg = slot < 1 ? 8u : 6u; g0 = base + (slot < 1 ? 0u : 8u + (slot - 1) * 6u); }
}
// the balanced tile that contains column group G
__host__ __device__ inline void balanced_tile_of_group(uint32_t G, uint32_t &g0, uint32_t &ncg)
{
    uint32_t rank, r;
    if (G < BAL_BIG_SMS * 28u) { rank = G / 28u; r = G % 28u; balanced_tile(rank, r < 8 ? 0 : r < 16 ? 1 : r < 22 ? 2 : 3, g0, ncg); }
    else { const uint32_t Gp = G - BAL_BIG_SMS * 28u; rank = BAL_BIG_SMS + Gp / 26u; r = Gp % 26u; balanced_tile(rank, r < 8 ? 0 : r < 14 ? 1 : r < 20 ? 2 : 3, g0, ncg); }
}

// Decision-row formats: 0 = canonical; t (1..8) = stage t of a pass over UNIFORM tiles; 8 + t = over BALANCED tiles.
constexpr uint8_t ROWFMT_BALANCED = 8;

// where a fused-format decision bit lives: fmt as above, state s after that stage.
// Returns the bit index inside the 2^23-bit row.  Slot fields (23 bits): mh[4] | ml[4] | G[12] | q[2] | h[1];
// a thread (row group thr, column group G) of tile (g0, ncg) owns the 16-byte chunk g0*16 + thr*ncg + (G - g0).
__host__ __device__ inline uint32_t fused_bit_address(int fmt, uint32_t s)
{
    const int t = fmt > ROWFMT_BALANCED ? fmt - ROWFMT_BALANCED : fmt;
    // slot p = state rotated right by t (the slot its survivor sits in during the pass)
    uint32_t p = ((s >> t) | (s << (23 - t))) & STATEMASK;
    uint32_t mh = (p >> 19) & 15, ml = (p >> 15) & 15, G = (p >> 3) & 4095;
    uint32_t q = (p >> 1) & 3, h = p & 1;
    uint32_t g0, ncg;
    if (fmt > ROWFMT_BALANCED) balanced_tile_of_group(G, g0, ncg);
    else { ncg = FUSED_COLGROUPS; g0 = G - G % FUSED_COLGROUPS; }
    uint32_t thr   = (t <= FR) ? ml : mh;       // thread row-group in this round
    uint32_t inner = (t <= FR) ? mh : ml;       // register row index in this round
    uint32_t chunk = g0 * 16 + thr * ncg + (G - g0);              // 16-byte chunk per thread
    uint32_t w     = inner >> 2;                                  // word in chunk
    uint32_t i     = ((inner & 3) << 1) | (q >> 1);               // bit in byte
    uint32_t byte  = ((q & 1) << 1) | h;                          // byte in word
    return chunk * 128 + w * 32 + byte * 8 + i;
}

} // namespace v224
