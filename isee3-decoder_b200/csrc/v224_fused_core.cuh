// v224_fused_core.cuh -- arithmetic core of the fused 8-stage ACS pass.
//
// Index algebra (not in the reference; it follows from the butterfly's rotate-left index map,
// viterbi224_sse2.c:296-299,326-327: old {b, b+2^22} -> new {2b, 2b+1}):
//
//   Keep every value in a fixed "slot" p (23 bits) for the whole pass.  If slot p holds state
//   rotl^t(p) after t stages, stage t+1 pairs the slots that differ in slot bit 22-t, and the
//   survivor for input bit 0 / 1 stays in the slot whose bit 22-t is 0 / 1.  An 8-stage pass
//   therefore only ever combines slots that differ in slot bits 22..15 ("m", 8 bits); slot bits
//   14..0 ("j") never mix.  A tile is all 256 m x 64 consecutive j.  Stages 1-4 butterfly
//   over m's high nibble mh (a thread holds the 16 mh rows of one ml), stages 5-8 over the low
//   nibble ml (a thread holds the 16 ml rows of one mh); the two rounds exchange through shared
//   memory once.  After 8 stages slot (m, j) holds state (j << 8) | m.
//
//   Two columns j, j+1 share one 32-bit register (packed 16-bit metrics); both halves run the
//   same butterfly with their own branch metric.  A thread holds NQ such registers per row.
//
// Branch metrics: for the butterfly whose bit-0 member sits in slot p (stage bit cleared) the
// expected symbols are parity(reg24 & POLYn) with reg24 = rotl^(t-1)(p) << 1 (viterbi224_sse2.c
// :74-77 evaluates the same parity into Branchtab224).  Parity is linear, so it splits into a
// per-thread part and a compile-time per-register part; per stage a thread needs only four
// packed (x) and four packed (0x8000 - delta) operands.
//
// This header is also compiled for the host by tests/emu (V224_HOST_EMU) to check the index
// algebra against the CPU oracle without a GPU; the product only uses the device path.
#pragma once
#include "v224_common.cuh"

#if defined(V224_HOST_EMU) && !defined(__CUDA_ARCH__)
#define V224_HD
namespace V224_NS {
static inline uint32_t f_popc(uint32_t x) { return (uint32_t)__builtin_popcount(x); }
static inline uint32_t f_addmin_u16x2(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t lo = ((a & 0xffff) + (b & 0xffff)) & 0xffff, hi = ((a >> 16) + (b >> 16)) & 0xffff;
    uint32_t clo = c & 0xffff, chi = c >> 16;
    return (lo < clo ? lo : clo) | ((hi < chi ? hi : chi) << 16);
}
static inline uint32_t f_prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint64_t pool = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        uint32_t n = (sel >> (4 * i)) & 0xf;
        uint32_t byte = (uint32_t)(pool >> (8 * (n & 7))) & 0xff;
        if (n & 8) byte = (byte & 0x80) ? 0xff : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
static inline uint32_t f_add_u16x2(uint32_t a, uint32_t b)
{
    return (((a & 0xffff) + (b & 0xffff)) & 0xffff) | (((a >> 16) + (b >> 16)) << 16);
}
static inline uint32_t f_sub_add(uint32_t c, uint32_t a, uint32_t k) { return c - a + k; }
static inline uint32_t f_minu2(uint32_t a, uint32_t b)
{
    uint32_t lo = (a & 0xffff) < (b & 0xffff) ? (a & 0xffff) : (b & 0xffff);
    uint32_t hi = (a >> 16) < (b >> 16) ? (a >> 16) : (b >> 16);
    return lo | (hi << 16);
}
static inline uint32_t f_maxu2(uint32_t a, uint32_t b)
{
    uint32_t lo = (a & 0xffff) > (b & 0xffff) ? (a & 0xffff) : (b & 0xffff);
    uint32_t hi = (a >> 16) > (b >> 16) ? (a >> 16) : (b >> 16);
    return lo | (hi << 16);
}
}
#else
#define V224_HD __device__ __forceinline__
namespace V224_NS {
V224_HD uint32_t f_popc(uint32_t x) { return (uint32_t)__popc(x); }
V224_HD uint32_t f_addmin_u16x2(uint32_t a, uint32_t b, uint32_t c) { return __viaddmin_u16x2(a, b, c); }   // VIADDMNMX.U16x2
// PRMT in its default mode: bit 3 of a selector nibble replicates the selected byte's sign bit.
// (__byte_perm() masks the selector to 3 bits per nibble and would drop that.)
V224_HD uint32_t f_prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}
V224_HD uint32_t f_add_u16x2(uint32_t a, uint32_t b) { return __vadd2(a, b); }                                // VIADD.16x2 (ALU pipe)
V224_HD uint32_t f_sub_add(uint32_t c, uint32_t a, uint32_t k)                                                // c - a + k, kept together
{
    uint32_t r;
    asm("{\n\t.reg .u32 t;\n\tsub.u32 t, %1, %2;\n\tadd.u32 %0, t, %3;\n\t}" : "=r"(r) : "r"(c), "r"(a), "r"(k));
    return r;
}
V224_HD uint32_t f_minu2(uint32_t a, uint32_t b) { return __vminu2(a, b); }                                   // VIMNMX.U16x2
V224_HD uint32_t f_maxu2(uint32_t a, uint32_t b) { return __vmaxu2(a, b); }
}
#endif

// Pipe balance of the butterfly (profiles/r02_ab_pipe_balance.txt).  The SM has two integer-capable pipes, both 16 lanes wide: ALU
// (VIADDMNMX, PRMT, LOP3, IADD3, VIADD.16x2) and FMA-heavy (IMAD.IADD -- what ptxas makes of a plain two-input add when the ALU
// pipe already carries the rest).  Written with two-input adds only, a pair of packed butterflies is 5 FMA-heavy + 4 ALU
// instructions, and ncu shows FMA-heavy as the busier pipe (63 % against 58 %).  Bit pidx of V224_IADD3_MASK (and bit q of
// V224_IADD3_QMASK) turns a pair's three adds  u = c - a, u + K0, u + K1  into two three-input adds  c - a + K0, c - a + K1
// (IADD3, ALU pipe): one instruction fewer, 3 off the FMA-heavy pipe, 2 onto the ALU pipe.  One pair index in eight (2 of a stage's
// 16 pairs) balances the pipes and is the measured optimum (8.35 -> 8.14 us per pass with 4 decoders in lockstep, 13.62 -> 13.40
// for one alone); more of them overload the ALU pipe.  V224_ALU_ADD_MASK moves c + x of the selected pairs onto the ALU pipe as a
// packed VIADD.16x2 instead (same value: no half overflows) -- measured, smaller gain, off.
#ifndef V224_ALU_ADD_MASK
#define V224_ALU_ADD_MASK 0
#endif
#ifndef V224_IADD3_MASK
#define V224_IADD3_MASK 0x01
#endif
#ifndef V224_IADD3_QMASK
#define V224_IADD3_QMASK 0xff
#endif

namespace V224_NS {

// Slot-bit mask of POLY at stage t (1-based): slot bit s feeds register bit ((s+t-1) mod 23)+1.
__host__ __device__ constexpr uint32_t slotmask(uint32_t poly, int t)
{
    uint32_t m = 0;
    for (int s = 0; s < 23; s++)
        if ((poly >> (((s + t - 1) % 23) + 1)) & 1u) m |= 1u << s;
    return m;
}

// label = e1 | e2 << 1 of the linear (un-flipped) part
template <int T> V224_HD uint32_t slot_label(uint32_t p)
{
    constexpr uint32_t m1 = slotmask(POLY1, T), m2 = slotmask(POLY2, T);
    // POLY1 ^ POLY2 is a single register bit (code.h:59-60), so the two masks differ in at most one slot bit:
    // one population count serves both parities
    constexpr uint32_t dm = m1 ^ m2;
    static_assert((dm & (dm - 1)) == 0, "the polynomials differ in one bit");
    const uint32_t e1 = f_popc(p & m1) & 1u;
    const uint32_t e2 = dm ? (e1 ^ ((p & dm) ? 1u : 0u)) : e1;
    return e1 | (e2 << 1);
}
constexpr uint32_t FLIP_LABEL = (uint32_t)G1FLIP | ((uint32_t)G2FLIP << 1);

// The label is linear in the slot bits, so a thread's label at stage t is
//   label_t(thread bits) ^ label_t(tile bits) ^ FLIP_LABEL
// with disjoint thread / tile bit fields.  Both parts are packed two bits per stage (stage t at bits 2t-2, 2t-1):
// the thread part once per kernel, the tile part (with the flip folded in) once per tile -- a stage then needs one
// shift and one mask instead of the parities.
template <int T> V224_HD uint32_t packed_label_stage(uint32_t p) { return slot_label<T>(p) << (2 * (T - 1)); }
// stages 1..FR see round-1 thread bits p1, stages FR+1..FK round-2 thread bits p2
V224_HD uint32_t packed_thread_labels(uint32_t p1, uint32_t p2)
{
    return packed_label_stage<1>(p1) | packed_label_stage<2>(p1) | packed_label_stage<3>(p1) | packed_label_stage<4>(p1) |
           packed_label_stage<5>(p2) | packed_label_stage<6>(p2) | packed_label_stage<7>(p2) | packed_label_stage<8>(p2);
}
V224_HD uint32_t packed_tile_labels(uint32_t ptile)
{
    uint32_t flip = 0;
    for (int t = 0; t < FK; t++) flip |= FLIP_LABEL << (2 * t);
    return (packed_label_stage<1>(ptile) | packed_label_stage<2>(ptile) | packed_label_stage<3>(ptile) | packed_label_stage<4>(ptile) |
            packed_label_stage<5>(ptile) | packed_label_stage<6>(ptile) | packed_label_stage<7>(ptile) | packed_label_stage<8>(ptile)) ^ flip;
}
static_assert(FK == 8 && FR == 4, "packed labels are written out for 8 stages in two rounds");

// Operand table in shared memory: optab[(t-1)*32 + beta*8 + {0..3: X[beta^i], 4..7: K[beta^i]}].
// The per-pass table in global memory (PassTab) carries it plus the ring rows of the pass's eight stages.
constexpr int OPTAB_WORDS = FK * 32;
// Word OPTAB_WORDS + FK: the most the metric of state 0 can grow over the pass -- the cost of staying in state 0 for all its stages,
// sum of s0 + (255 - s1) (the all-zero transition expects symbols (G1FLIP, G2FLIP) = (0, 1), viterbi224_sse2.c:75-76,292): the
// survivor of state 0 is the minimum over two candidates one of which is that transition.  The resolver predicts from it which
// passes can reach the renormalisation trigger (and have to record per-stage minima).
constexpr int PASSTAB_CZ = OPTAB_WORDS + FK;
constexpr int PASSTAB_WORDS = OPTAB_WORDS + FK + 4;  // 268 words = 67 x 16 bytes

// Entry `e` (0 .. FK*32) of the operand table for the pass whose symbols are sym[2*(t-1)], sym[2*(t-1)+1].
template <typename SymPtr>
V224_HD uint32_t optab_entry(int e, SymPtr sym)
{
    const int t = e / 32 + 1, beta = (e / 8) & 3, isK = (e / 4) & 1, i = e & 3;
    const int s0 = sym[2 * (t - 1)], s1 = sym[2 * (t - 1) + 1];
    // label of slot bit 0 (the upper half's extra label) at stage t
    uint32_t m1 = slotmask(POLY1, t), m2 = slotmask(POLY2, t);
    const uint32_t hc = (m1 & 1u) | ((m2 & 1u) << 1);
    const uint32_t l_lo = (uint32_t)(beta ^ i), l_hi = l_lo ^ hc;
    // (BT0 ^ s0) + (BT1 ^ s1), viterbi224_sse2.c:292
    const int x_lo = ((l_lo & 1) ? 255 - s0 : s0) + ((l_lo & 2) ? 255 - s1 : s1);
    const int x_hi = ((l_hi & 1) ? 255 - s0 : s0) + ((l_hi & 2) ? 255 - s1 : s1);
    if (!isK) return (uint32_t)x_lo | ((uint32_t)x_hi << 16);
    const int d_lo = 2 * x_lo - 510, d_hi = 2 * x_hi - 510;     // x - (510 - x), :293
    return (uint32_t)(0x8000 - d_lo) | ((uint32_t)(0x8000 - d_hi) << 16);
}

// One trellis stage over a thread's 16 rows x NQ packed registers.
//   A[inner][q]  : packed P metrics, halves = columns 2q, 2q+1 of the thread's 2*NQ columns
//   labels       : packed_thread_labels(..) ^ packed_tile_labels(..): the thread's label (flip included) of stage t at bits 2t-2, 2t-1
//   optab        : shared operand table
//   dw[NQ]       : returns the 32*NQ decision bits of this thread/stage in fused layout (fused_bit_address())
// the thread's four packed X and four packed K of stage T (two 128-bit shared loads)
template <int T>
V224_HD void stage_operands(uint32_t labels, const uint32_t *optab, uint32_t (&Xv)[4], uint32_t (&Kv)[4])
{
    // operand row of this thread's label: (T-1)*32 + beta*8 words
    const uint32_t boff = (T == 1 ? (labels << 3) : T <= 2 ? (labels << 1) : (labels >> (2 * (T - 1) - 3))) & 24u;
    const uint32_t *tab = optab + (T - 1) * 32 + boff;
#pragma unroll
    for (int i = 0; i < 4; i++) { Xv[i] = tab[i]; Kv[i] = tab[4 + i]; }
}

// PF_AT >= 0: the operands of stage T + 1 are loaded into Xn / Kn right before pair index PF_AT (labels / optab as for stage_operands)
template <int T, int PF_AT = -1>
V224_HD void acs_stage_body(uint32_t (&A)[16][NQ], const uint32_t (&Xv)[4], const uint32_t (&Kv)[4], uint32_t (&dw)[NQ],
                            uint32_t labels = 0, const uint32_t *optab = nullptr, uint32_t *Xn = nullptr, uint32_t *Kn = nullptr);

template <int T>
V224_HD void acs_stage(uint32_t (&A)[16][NQ], uint32_t labels, const uint32_t *optab, uint32_t (&dw)[NQ])
{
    uint32_t Xv[4], Kv[4];
    stage_operands<T>(labels, optab, Xv, Kv);
    acs_stage_body<T>(A, Xv, Kv, dw);
}

template <int T, int PF_AT>
V224_HD void acs_stage_body(uint32_t (&A)[16][NQ], const uint32_t (&Xv)[4], const uint32_t (&Kv)[4], uint32_t (&dw)[NQ],
                            uint32_t labels, const uint32_t *optab, uint32_t *Xn, uint32_t *Kn)
{
    constexpr int sb = FR - 1 - ((T - 1) % FR);          // stage bit inside `inner`
    constexpr int ishift = (T <= FR) ? 19 : 15;           // slot position of `inner`
#pragma unroll
    for (int w = 0; w < NQ; w++) dw[w] = 0;
#pragma unroll
    for (int pidx = 0; pidx < 8; pidx++) {
        if (PF_AT == pidx && T < FK) {
            constexpr int TN = T < FK ? T + 1 : T;
            const uint32_t boff = (TN <= 2 ? (labels << 1) : (labels >> (2 * (TN - 1) - 3))) & 24u;
            const uint32_t *tabn = optab + (TN - 1) * 32 + boff;
#pragma unroll
            for (int i = 0; i < 4; i++) { Xn[i] = tabn[i]; Kn[i] = tabn[4 + i]; }
        }
        const int ia = ((pidx >> sb) << (sb + 1)) | (pidx & ((1 << sb) - 1));
        const int ic = ia | (1 << sb);
        uint32_t D0[NQ], D1[NQ];
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            const uint32_t off = ((uint32_t)ia << ishift) | ((uint32_t)q << 1);
            const uint32_t c = slot_label<T>(off);
            const uint32_t X = Xv[c], Y = Xv[c ^ 3], K0 = Kv[c], K1 = Kv[c ^ 3];
            const uint32_t a = A[ia][q], cc = A[ic][q];
            // m0 = a+x, m1 = c+y, m2 = a+y, m3 = c+x          (viterbi224_sse2.c:296-299)
            const uint32_t t0 = cc + Y, t1 = ((V224_ALU_ADD_MASK >> pidx) & 1) ? f_add_u16x2(cc, X) : cc + X;
            // decision0 = m0 > m1  <=>  (c - a) - delta < 0  <=> bit15 of D0 clear   (:316)
            // decision1 = m2 > m3  <=>  (c - a) + delta < 0  <=> bit15 of D1 clear   (:317)
            if (((V224_IADD3_MASK >> pidx) & 1) && ((V224_IADD3_QMASK >> q) & 1)) {
                // one three-input add per decision word (IADD3, ALU pipe) for the pairs that balance the two integer pipes
                D0[q] = f_sub_add(cc, a, K0);
                D1[q] = f_sub_add(cc, a, K1);
            } else {
                const uint32_t u = cc - a;
                D0[q] = u + K0;
                D1[q] = u + K1;
            }
            A[ia][q] = f_addmin_u16x2(a, X, t0);          // min(m0, m1) -> state 2b    (:319)
            A[ic][q] = f_addmin_u16x2(a, Y, t1);          // min(m2, m3) -> state 2b+1  (:320)
        }
        // gather the sign bits: word = side * NQ/2 + q/2, byte = (q&1)*2 + half, bit = pair index
        // (one register per row: a single word, byte = side * 2 + half)
        if (NQ == 1) {
            const uint32_t s01 = f_prmt(D0[0], D1[0], 0xfdb9);
            dw[0] = (s01 & (0x01010101u << pidx)) | dw[0];
        }
#pragma unroll
        for (int qq = 0; qq < NQ / 2; qq++) {
            const uint32_t s0 = f_prmt(D0[2 * qq], D0[2 * qq + 1], 0xfdb9);
            const uint32_t s1 = f_prmt(D1[2 * qq], D1[2 * qq + 1], 0xfdb9);
            // bit 15 of D is the INVERTED decision; it is stored as it is (fused rows hold complemented bits, the
            // traceback kernels flip them back: FUSED_ROWS_COMPLEMENTED) -- one LOP3 per 4 decisions, no final XOR.
            // (Measured and dropped, profiles/r01_ab_imad_gather.txt, r01_ab_tree_gather.txt: merging with multiply-adds on
            // the FMA pipe, and merging the eight 0x00/0xff words pairwise with bit-selects -- 7 instead of 8 LOP3 -- are
            // both exact and both 1-2 % slower.)
            dw[qq] = (s0 & (0x01010101u << pidx)) | dw[qq];
            dw[NQ / 2 + qq] = (s1 & (0x01010101u << pidx)) | dw[NQ / 2 + qq];
        }
    }
}

// packed min / max over a thread's 64 registers
V224_HD uint32_t tile_min(const uint32_t (&A)[16][NQ])
{
    uint32_t m = A[0][0];
#pragma unroll
    for (int i = 0; i < 16; i++)
#pragma unroll
        for (int q = 0; q < NQ; q++) m = f_minu2(m, A[i][q]);
    uint32_t lo = m & 0xffff, hi = m >> 16;
    return lo < hi ? lo : hi;
}
V224_HD uint32_t tile_max(const uint32_t (&A)[16][NQ])
{
    uint32_t m = A[0][0];
#pragma unroll
    for (int i = 0; i < 16; i++)
#pragma unroll
        for (int q = 0; q < NQ; q++) m = f_maxu2(m, A[i][q]);
    uint32_t lo = m & 0xffff, hi = m >> 16;
    return lo > hi ? lo : hi;
}

} // namespace v224
