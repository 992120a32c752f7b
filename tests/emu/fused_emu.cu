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

using namespace V224_NS;

template <int T>
static void emu_stage(uint32_t (&A)[16][NQ], uint32_t labels, const uint32_t *optab, uint32_t *rows, uint32_t chunk,
                      uint32_t *s0, uint32_t *minP, bool first)
{
    uint32_t dw[NQ];
    acs_stage<T>(A, labels, optab, dw);
    uint32_t *row = rows + (size_t)(T - 1) * ROWWORDS;
    for (int w = 0; w < NQ; w++) row[chunk * NQ + w] = dw[w];
    if (first) s0[T] = A[0][0] & 0xffffu;
    uint32_t mn = tile_min(A);
    if (mn < minP[T]) minP[T] = mn;
}

// (the shared-memory exchange index xchg_index() is the kernel's own, from v224_common.cuh)

extern "C" {

int emu_nq(void) { return NQ; }
int emu_tile_cols(void) { return FUSED_TILE_COLS; }
int emu_rowfmt_base(void) { return ROWFMT_FUSED_BASE; }

// One fused pass: oldP/newP 2^23 uint16, rows = 8 decision rows in fused layout, syms = 16 bytes.
// stats: s0[1..8], minP[1..8], maxP_end written to out_stats[0..8], [9..17], [18].
void emu_fused_pass(const uint16_t *oldP, uint16_t *newP, uint32_t *rows, const uint8_t *syms, int sub, uint32_t *out_stats)
{
    alignas(16) static uint32_t optab[OPTAB_WORDS];
    for (int e = 0; e < OPTAB_WORDS; e++) optab[e] = optab_entry(e, syms);
    uint32_t s0[FK + 1] = {0}, minP[FK + 1];
    for (int t = 0; t <= FK; t++) minP[t] = 0xffffffffu;
    uint32_t maxP = 0;
    const uint32_t sub2 = (uint32_t)sub * 0x10001u;
    std::vector<uint32_t> tile(256 * FUSED_TILE_COLS / 2);
    for (uint32_t tau = 0; tau < (uint32_t)FUSED_TILES; tau++) {
        // round 1
        for (uint32_t tid = 0; tid < (uint32_t)FUSED_THREADS; tid++) {
            uint32_t thr, g;
            round1_map(tid, thr, g);
            const uint32_t G = tau * FUSED_COLGROUPS + g, chunk = tau * FUSED_THREADS + tid;
            uint32_t A[16][NQ];
            for (int mh = 0; mh < 16; mh++) {
                const uint32_t *src = reinterpret_cast<const uint32_t *>(oldP + ((size_t)(mh * 16 + thr) * 32768 + G * COLW));
                for (int q = 0; q < NQ; q++) A[mh][q] = src[q] - sub2;
            }
            // the kernel's split of the branch labels: thread part (fixed per thread) ^ tile part (per tile, from the producer warp)
            uint32_t t2, g2;
            round2_map(tid, t2, g2);
            const uint32_t labels = packed_thread_labels((thr << 15) | (g << COLW_LOG2), (t2 << 19) | (g2 << COLW_LOG2)) ^
                                    packed_tile_labels(tau << FUSED_COLS_LOG2);
            const bool first = tau == 0 && tid == 0;
            emu_stage<1>(A, labels, optab, rows, chunk, s0, minP, first);
            emu_stage<2>(A, labels, optab, rows, chunk, s0, minP, first);
            emu_stage<3>(A, labels, optab, rows, chunk, s0, minP, first);
            emu_stage<4>(A, labels, optab, rows, chunk, s0, minP, first);
            for (int mh = 0; mh < 16; mh++)
                for (int q = 0; q < NQ; q++) tile[xchg_index(mh * 16 + thr, g) * NQ + q] = A[mh][q];
        }
        // round 2
        for (uint32_t tid = 0; tid < (uint32_t)FUSED_THREADS; tid++) {
            uint32_t thr, g;
            round2_map(tid, thr, g);
            const uint32_t G = tau * FUSED_COLGROUPS + g, chunk = tau * FUSED_THREADS + tid;
            uint32_t A[16][NQ];
            for (int ml = 0; ml < 16; ml++)
                for (int q = 0; q < NQ; q++) A[ml][q] = tile[xchg_index(thr * 16 + ml, g) * NQ + q];
            uint32_t t1, g1;
            round1_map(tid, t1, g1);
            const uint32_t labels = packed_thread_labels((t1 << 15) | (g1 << COLW_LOG2), (thr << 19) | (g << COLW_LOG2)) ^
                                    packed_tile_labels(tau << FUSED_COLS_LOG2);
            const bool first = tau == 0 && tid == 0;
            emu_stage<5>(A, labels, optab, rows, chunk, s0, minP, first);
            emu_stage<6>(A, labels, optab, rows, chunk, s0, minP, first);
            emu_stage<7>(A, labels, optab, rows, chunk, s0, minP, first);
            emu_stage<8>(A, labels, optab, rows, chunk, s0, minP, first);
            uint32_t mx = tile_max(A);
            if (mx > maxP) maxP = mx;
            for (int q = 0; q < NQ; q++)
                for (int h = 0; h < 2; h++)
                    for (int ml = 0; ml < 16; ml++)
                        newP[((size_t)(G * COLW + q * 2 + h) << 8) + thr * 16 + ml] = (uint16_t)(A[ml][q] >> (16 * h));
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
        canon_row[s >> 5] |= (((fused_row[a >> 5] >> (a & 31)) & 1u) ^ FUSED_ROWS_COMPLEMENTED) << (s & 31);
    }
}
}
