// fused_emu.cu -- TEST INFRASTRUCTURE: runs the arithmetic core of the fused 8-stage CUDA pass
// (isee3-decoder_b200/csrc/v224_fused_core.cuh) on the host, thread by thread, so that the
// index algebra (slot rotation, per-stage label masks, decision-bit layout) can be checked
// against the CPU oracle in the no-GPU test tier.  It mirrors k_acs_fused's data movement
// exactly; it is never linked into the product library.
#define V224_HOST_EMU 1
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../isee3-decoder_b200/csrc/v224_fused_core.cuh"

using namespace v224;

template <int T>
static void emu_stage(uint32_t (&A)[16][4], uint32_t pbase, const uint32_t *optab, uint32_t *rows, uint32_t chunk,
                      uint32_t *s0, uint32_t *minP, bool first)
{
    uint32_t dw[4];
    acs_stage<T>(A, pbase, optab, dw);
    uint32_t *row = rows + (size_t)(T - 1) * ROWWORDS;
    for (int w = 0; w < 4; w++) row[chunk * 4 + w] = dw[w];
    if (first) s0[T] = A[0][0] & 0xffffu;
    uint32_t mn = tile_min(A);
    if (mn < minP[T]) minP[T] = mn;
}

struct EmuTile { uint32_t g0, ncg; };

extern "C" {

// One fused pass: oldP/newP 2^23 uint16, rows = 8 decision rows in fused layout, syms = 16 bytes.
// balanced = 0: uniform tiles of FUSED_COLGROUPS column groups; 1: the 592-tile balanced partition.
// stats: s0[1..8], minP[1..8], maxP_end written to out_stats[0..8], [9..17], [18].
void emu_fused_pass(const uint16_t *oldP, uint16_t *newP, uint32_t *rows, const uint8_t *syms, int sub, uint32_t *out_stats, int balanced)
{
    alignas(16) static uint32_t optab[OPTAB_WORDS];
    for (int e = 0; e < OPTAB_WORDS; e++) optab[e] = optab_entry(e, syms);
    uint32_t s0[FK + 1] = {0}, minP[FK + 1];
    for (int t = 0; t <= FK; t++) minP[t] = 0xffffffffu;
    uint32_t maxP = 0;
    const uint32_t sub2 = (uint32_t)sub * 0x10001u;
    std::vector<EmuTile> tiles;
    if (balanced) {
        for (uint32_t rank = 0; rank < BAL_SMS; rank++)
            for (uint32_t slot = 0; slot < BAL_CTAS_PER_SM; slot++) { EmuTile t; balanced_tile(rank, slot, t.g0, t.ncg); tiles.push_back(t); }
    } else {
        for (uint32_t t = 0; t < FUSED_TILES; t++) tiles.push_back({t * FUSED_COLGROUPS, (uint32_t)FUSED_COLGROUPS});
    }
    std::vector<uint32_t> tile(256 * 8 * 4);
    for (const EmuTile &tl : tiles) {
        const uint32_t g0 = tl.g0, ncg = tl.ncg, nthreads = 16 * ncg;
        // round 1
        for (uint32_t tid = 0; tid < nthreads; tid++) {
            const uint32_t thr = tid / ncg, g = tid % ncg, G = g0 + g, chunk = g0 * 16 + tid;
            uint32_t A[16][4];
            for (int mh = 0; mh < 16; mh++) {
                const uint32_t *src = reinterpret_cast<const uint32_t *>(oldP + ((size_t)(mh * 16 + thr) * 32768 + G * 8));
                for (int q = 0; q < 4; q++) A[mh][q] = src[q] - sub2;
            }
            const uint32_t pbase = (thr << 15) | (G << 3);
            const bool first = G == 0 && thr == 0;
            emu_stage<1>(A, pbase, optab, rows, chunk, s0, minP, first);
            emu_stage<2>(A, pbase, optab, rows, chunk, s0, minP, first);
            emu_stage<3>(A, pbase, optab, rows, chunk, s0, minP, first);
            emu_stage<4>(A, pbase, optab, rows, chunk, s0, minP, first);
            for (int mh = 0; mh < 16; mh++)
                for (int q = 0; q < 4; q++) tile[((mh * 16 + thr) * ncg + g) * 4 + q] = A[mh][q];
        }
        // round 2
        for (uint32_t tid = 0; tid < nthreads; tid++) {
            const uint32_t thr = tid / ncg, g = tid % ncg, G = g0 + g, chunk = g0 * 16 + tid;
            uint32_t A[16][4];
            for (int ml = 0; ml < 16; ml++)
                for (int q = 0; q < 4; q++) A[ml][q] = tile[((thr * 16 + ml) * ncg + g) * 4 + q];
            const uint32_t pbase = (thr << 19) | (G << 3);
            const bool first = G == 0 && thr == 0;
            emu_stage<5>(A, pbase, optab, rows, chunk, s0, minP, first);
            emu_stage<6>(A, pbase, optab, rows, chunk, s0, minP, first);
            emu_stage<7>(A, pbase, optab, rows, chunk, s0, minP, first);
            emu_stage<8>(A, pbase, optab, rows, chunk, s0, minP, first);
            uint32_t mx = tile_max(A);
            if (mx > maxP) maxP = mx;
            for (int q = 0; q < 4; q++)
                for (int h = 0; h < 2; h++)
                    for (int ml = 0; ml < 16; ml++)
                        newP[((size_t)(G * 8 + q * 2 + h) << 8) + thr * 16 + ml] = (uint16_t)(A[ml][q] >> (16 * h));
        }
    }
    for (int t = 0; t <= FK; t++) { out_stats[t] = s0[t]; out_stats[FK + 1 + t] = minP[t]; }
    out_stats[2 * (FK + 1)] = maxP;
}

// fused-layout row written by stage t -> canonical (reference) layout
void emu_canon_row(int fmt, const uint32_t *fused_row, uint32_t *canon_row)
{
    memset(canon_row, 0, ROWBYTES);
    for (uint32_t s = 0; s < NSTATES; s++) {
        const uint32_t a = fused_bit_address(fmt, s);
        canon_row[s >> 5] |= ((fused_row[a >> 5] >> (a & 31)) & 1u) << (s & 31);
    }
}
}
