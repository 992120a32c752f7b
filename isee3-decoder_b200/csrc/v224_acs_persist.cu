// v224_acs_persist.cu -- the fused ACS pass of the B200 viterbi224 decoder (sm_100a).
//
//   k_acs_persist    8 trellis stages per HBM pass (viterbi224_sse2.c:264-328, x8); one launch runs many passes of
//                    1..MAX_CTX independent decoders as a dataflow over a dynamic tile queue
//   k_build_passtab  per-pass operand tables + ring rows
//   k_persist_begin  launch prologue on the control block
//
// CTA = FUSED_THREADS / 32 compute warps + 2 protocol warps (warp specialisation: producer and retirer), 64 registers per
// thread.  The producer runs the dataflow one tile ahead of the compute warps: it claims the next (pass, decoder, tile)
// item, polls the pass parameters and the previous pass's completion counters (L2 round trips), issues the tile's
// tensor copy and the pass table's bulk copy (TMA) and hands the tile over through a shared-memory mbarrier.  The
// retirer waits until the compute warps have issued a tile's stores, publishes the tile (release) and, for the last
// tile of a pass, runs the resolver.  The compute warps therefore never wait for an L2 round trip of the protocol,
// only for data.
//
// Built twice (isee3-decoder_b200/build.py): tiles of 64 columns -- 256 compute threads, 3 CTAs per SM, the shape of
// the lockstep multi-decoder launches -- and tiles of 32 columns (-DV224_TILE_COLS_LOG2=5, namespace v224t32) -- 128
// compute threads, up to 5 CTAs per SM, for a decoder running alone.
//
// This file is compiled with -Xptxas -O1: at the default level ptxas hoists the decision-bit gather of a whole stage
// behind the butterflies and spills (~500 bytes per thread at 64 registers); in source order the tile body fits in both
// builds (csrc/ptxas.log).
#include "v224_common.cuh"
#include "v224_fused_core.cuh"
#include "v224_kernels.h"
#include <mutex>
#include <cuda.h>          // CUtensorMap and the cuTensorMapEncodeTiled prototype only; the entry point is fetched at run time

namespace V224_NS {

#ifdef V224_TRACE
// per (pass < 64, decoder < 4, tile < 512): 8 event timestamps (globaltimer ns) + the SM the tile ran on
__device__ unsigned long long g_trace[64 * 4 * 512 * 8];
__device__ unsigned g_smid[64 * 4 * 512];
__device__ __forceinline__ unsigned long long gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define TRACE(key, tau, ev) do { if ((threadIdx.x & 31) == 0 && (key) >= 0 && (key) < 256) { g_trace[(((size_t)(key)) * 512 + (tau)) * 8 + (ev)] = gtime(); if ((ev) == 1) { unsigned sm_; asm volatile("mov.u32 %0, %smid;" : "=r"(sm_)); g_smid[(size_t)(key) * 512 + (tau)] = sm_ | (blockIdx.x << 16); } } } while (0)
#define TKEY(n, s) ((n) < 64 ? (n) * 4 + (int)(s) : -1)
#else
#define TRACE(key, tau, ev) do { } while (0)
#define TKEY(n, s) 0
#endif

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_cs_v4(void *p, uint4 v)   // streaming store: decision rows are write-once
{
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_cs_v2(void *p, uint32_t x, uint32_t y)
{
    asm volatile("st.global.cs.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void st_cs_b32(void *p, uint32_t x)
{
    asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(x) : "memory");
}
__device__ __forceinline__ void st_v8(void *p, const uint32_t (&v)[8])   // 256-bit store (sm_100+)
{
    asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
                 "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_relaxed(const void *p)
{
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed64(const void *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_acq_rel()
{
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
}
__device__ __forceinline__ void st_release(unsigned *p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long atom_acq_rel_add64(unsigned long long *p, unsigned long long v)
{
    unsigned long long old;
    asm volatile("atom.acq_rel.gpu.global.add.u64 %0, [%1], %2;" : "=l"(old) : "l"(p), "l"(v) : "memory");
    return old;
}
// shared-memory mbarriers (CTA scope)
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t *bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    // the third operand is the suspend-time hint: the warp sleeps in hardware until the phase completes (or that long)
    // instead of spinning through the issue slots the compute warps need
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 ::"r"(smem_addr(bar)), "r"(parity), "r"(0x989680u) : "memory");
}
// The protocol warps' waits last about as long as a tile's compute time, and the try_wait loop above comes round every
// ~35 ns (ncu source page: 300-400 iterations per tile, 23 % of all instructions the kernel issues).  Polling with a real
// sleep in between was measured as an alternative (V224_PROTO_SLEEP_NS = 100 .. 1600): no gain with 3 decoders in lockstep
// (8.94 - 9.16 vs 9.01 us per pass), 1 - 8 % slower with one -- the spin fills issue slots nobody else wants, the compute
// rounds are bound by the ALU/FMA pipes (profiles/r01_ab_proto_sleep.txt).  Default 0 = the hardware wait.
#ifndef V224_PROTO_SLEEP_NS
#define V224_PROTO_SLEEP_NS 0
#endif
__device__ __forceinline__ void mbar_wait_proto(uint64_t *bar, unsigned parity)
{
#if V224_PROTO_SLEEP_NS > 0
    while (!mbar_test(bar, parity)) __nanosleep(V224_PROTO_SLEEP_NS);
#else
    mbar_wait(bar, parity);
#endif
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
// bulk asynchronous copy global -> shared (TMA engine, no tensor map), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
// one tile = one 2-D tensor copy: box FUSED_TILE_COLS columns x 256 rows at column x0 (TMA engine; bytes counted on the mbarrier)
__device__ __forceinline__ void tma_load_tile(void *smem_dst, const void *tmap, int x0, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_addr(smem_dst)), "l"(tmap), "r"(x0), "r"(0), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async()        { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem()   { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bar_compute()     // the compute warps' own barrier (the protocol warp never joins)
{
    asm volatile("bar.sync 1, %0;" ::"n"(FUSED_THREADS) : "memory");
}

// ------------------------------------------------------------------------------------------
// shared memory of one CTA
// ------------------------------------------------------------------------------------------
struct TileInfo {
    const uint16_t *oldm;
    uint16_t *newm;
    uint32_t *ring;
    PassStats *st;
    uint32_t tau;
    uint32_t sub2;       // sub in both halves
    int go;              // 1 run the tile, 0 skip it (decoder stopped), -1 no more work
    int careful;
    int n, s;            // pass and decoder (the protocol warps' own bookkeeping)
    uint32_t lab;        // packed_tile_labels(tau): the tile's part of every stage's branch label
    int measure;         // reduce min / max of the tile's output (pass_word)
    int discard;         // the pass cannot be invalidated: its input lines may be dropped from the L2 once they are in shared memory
};
// Tile k of a CTA uses bookkeeping slot k % NSLOT (info, pass table, full/done barriers) and data buffer k % XCHG_BUFS.
// A data buffer is free again as soon as the tile's round-2 reads are over (`freeb`), long before its stores are out
// (`done`), so the protocol warp can have the input of tile k+2 in flight while tile k is still in its second round.
// NSLOT = 2 (one tile of run-ahead) is the measured optimum: with 3-4 slots the protocol warps pre-claim most of
// the tiles that are in flight and the dynamic queue stops balancing (3 decoders: 9.2 us per pass at 2 slots, 10.0 at 4).
#ifndef V224_NSLOT
#define V224_NSLOT 2
#endif
// Two small savings in the tile body, measured together with the pipe balance of the butterfly (profiles/r02_ab_pipe_balance.txt):
// V224_SUB_SKIP: most passes carry sub == 0 (a subtraction follows a MEASURED pass two passes later), and the 16 * NQ subtractions
// of the load are skipped then (warp-uniform branch); V224_OPERAND_PREFETCH: 1 = a stage's operands are loaded by the PREVIOUS stage,
// right after its butterflies (the registers are free there), so that a stage does not open with two shared loads its first adds wait
// for; n >= 2 = before pair index n - 2 of the butterflies (ptxas sinks those loads back next to the stores: no gain).  Together
// 8.15 -> 8.06 us per pass with 4 decoders in lockstep.
#ifndef V224_SUB_SKIP
#define V224_SUB_SKIP 1
#endif
#ifndef V224_OPERAND_PREFETCH
#define V224_OPERAND_PREFETCH 1
#endif
constexpr int NSLOT = V224_NSLOT;
constexpr int CTA_THREADS = FUSED_THREADS + 64;              // compute warps + producer warp + retirer warp
struct __align__(128) FusedSmem {
    uint32_t tile[XCHG_BUFS][256 * FUSED_TILE_COLS / 2];   // 256 rows x 64 columns of uint16 (32 KiB each): tile input (bulk mode) and round-1 -> round-2 exchange
    uint32_t tab[NSLOT][PASSTAB_WORDS];                      // operand table + ring rows of the tile's pass
    TileInfo info[NSLOT];
    uint32_t s0[FK + 4];                                     // state-0 metric after each stage (meaningful in tile 0 only)
    uint64_t full[NSLOT], done[NSLOT];                       // mbarriers: tile handed over (and its input landed) / tile's stores issued
    uint64_t slotfree[NSLOT];                                // mbarriers: the tile is retired from its bookkeeping slot
    uint64_t freeb[2];                                       // mbarriers: the data buffer's last reader is through
};

// Per-pass tables of a whole launch: the operand table from the pass's 8 symbol pairs, and the ring row of each
// of its 8 stages ((T0 + 8*pass + t) mod len, done once here instead of per thread and stage).  The rows' format
// tags are set here as well (only by the last stage of the launch that writes a given ring row).
__global__ void __launch_bounds__(PASSTAB_WORDS) k_build_passtab(uint32_t *tab, const uint8_t *syms, int npasses, long long T0, int len,
                                                                 uint8_t *row_fmt)
{
    const int pass = blockIdx.x, e = threadIdx.x;
    if (pass >= npasses) return;
    uint32_t v;
    if (e < OPTAB_WORDS) {
        v = optab_entry(e, syms + 2 * (size_t)pass * FK);
    } else if (e >= PASSTAB_CZ) {
        // upper bound of state 0's metric growth over the pass (PASSTAB_CZ); the other words of the tail are padding
        v = 0;
        if (e == PASSTAB_CZ)
            for (int t = 0; t < FK; t++) {
                const uint8_t *sp = syms + 2 * ((size_t)pass * FK + t);
                v += (uint32_t)((G1FLIP ? 255 - sp[0] : sp[0]) + (G2FLIP ? 255 - sp[1] : sp[1]));
            }
    } else {
        const int t = e - OPTAB_WORDS;
        const long long stage = (long long)pass * FK + t;
        v = (uint32_t)((T0 + stage) % len);
        if (stage + len >= (long long)npasses * FK) row_fmt[v] = (uint8_t)(ROWFMT_FUSED_BASE + t + 1);
    }
    tab[(size_t)pass * PASSTAB_WORDS + e] = v;
}

// ------------------------------------------------------------------------------------------
// compute warps: one tile, eight stages
// ------------------------------------------------------------------------------------------
// Xv / Kv: the stage's operands, loaded by the previous stage (V224_OPERAND_PREFETCH above); on return those of the next stage.
template <int T, bool CAREFUL>
__device__ __forceinline__ void fused_stage(uint32_t (&A)[16][NQ], uint32_t labels, const uint32_t *tab, uint32_t *s0, uint32_t *ring_chunk,
                                            PassStats *st, uint32_t (&Xv)[4], uint32_t (&Kv)[4])
{
    uint32_t dw[NQ];
    if (V224_OPERAND_PREFETCH >= 2) {
        // ... or already before pair index V224_OPERAND_PREFETCH - 2 of the butterflies (eight more live registers from there on)
        uint32_t Xn[4], Kn[4];
        acs_stage_body<T, V224_OPERAND_PREFETCH - 2>(A, Xv, Kv, dw, labels, tab, Xn, Kn);
        if (T < FK) {
#pragma unroll
            for (int i = 0; i < 4; i++) { Xv[i] = Xn[i]; Kv[i] = Kn[i]; }
        }
    } else if (V224_OPERAND_PREFETCH) {
        acs_stage_body<T>(A, Xv, Kv, dw);
        if (T < FK) stage_operands<(T < FK ? T + 1 : T)>(labels, tab, Xv, Kv);
    } else {
        acs_stage<T>(A, labels, tab, dw);
    }
    // row * 1 MiB + this thread's chunk: one 32 x 32 + 64 multiply-add
    uint32_t *dst;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(dst) : "r"(tab[OPTAB_WORDS + T - 1]), "n"((unsigned)ROWBYTES), "l"(ring_chunk));
    if (NQ == 4)      st_cs_v4(dst, make_uint4(dw[0], dw[1], dw[NQ - 2], dw[NQ - 1]));
    else if (NQ == 2) st_cs_v2(dst, dw[0], dw[NQ - 1]);
    else              st_cs_b32(dst, dw[0]);
    if (threadIdx.x == 0) s0[T] = A[0][0] & 0xffffu;               // slot 0 (tile 0, thread 0) always holds state 0
    if (CAREFUL && T < FK) {
        uint32_t mn = __reduce_min_sync(0xffffffffu, tile_min(A));
        if ((threadIdx.x & 31) == 0) atomicMin(&st->minP[T][(blockIdx.x * (FUSED_THREADS / 32) + (threadIdx.x >> 5)) % STAT_BUCKETS][0], mn);
    }
}

// Tile `tau` = columns [FUSED_TILE_COLS tau, FUSED_TILE_COLS (tau + 1)) of all 256 rows.  Metrics are read through L2 only (another SM wrote
// them, possibly within this launch).  `xbuf` = this tile's exchange buffer.
template <bool CAREFUL>
__device__ __forceinline__ void fused_tile(uint32_t *xbuf, const uint32_t *tab, uint32_t *s0, const TileInfo &ti, uint64_t *freeb, uint32_t labthr,
                                           int trace_n)
{
    const uint32_t tid = threadIdx.x, tau = ti.tau, sub = ti.sub2;
    const uint32_t labels = labthr ^ ti.lab;               // this thread's branch label of every stage, two bits each
    PassStats *st = ti.st;
    uint32_t *ring_chunk = ti.ring + (size_t)(tau * FUSED_THREADS + tid) * NQ;      // see fused_bit_address()
    uint32_t A[16][NQ];
    uint32_t Xv[4], Kv[4];
    if (V224_OPERAND_PREFETCH) stage_operands<1>(labels, tab, Xv, Kv);
    {
        // ---- round 1: thread = (ml = thr, g); registers = 16 mh rows x 2*NQ columns ----
        uint32_t thr, g;
        round1_map(tid, thr, g);
        const uint32_t G = tau * FUSED_COLGROUPS + g;              // global column group: columns COLW*G ..
        if (BULK_LOAD && V224_SUB_SKIP) {
            // sub == 0 in most passes: the subtractions are skipped then (V224_SUB_SKIP above)
#pragma unroll
            for (int mh = 0; mh < 16; mh++) {
                const uint32_t e = (mh * 16 + thr) * FUSED_COLGROUPS + g;
                if (NQ == 4) {
                    const uint4 v = reinterpret_cast<const uint4 *>(xbuf)[e];
                    A[mh][0] = v.x; A[mh][1] = v.y; A[mh][NQ - 2] = v.z; A[mh][NQ - 1] = v.w;
                } else if (NQ == 2) {
                    const uint2 v = reinterpret_cast<const uint2 *>(xbuf)[e];
                    A[mh][0] = v.x; A[mh][NQ - 1] = v.y;
                } else {
                    A[mh][0] = xbuf[e];
                }
            }
            if (sub != 0) {
#pragma unroll
                for (int mh = 0; mh < 16; mh++)
#pragma unroll
                    for (int q = 0; q < NQ; q++) A[mh][q] -= sub;
            }
        } else if (BULK_LOAD) {
            // the protocol warp's bulk copies put the tile into this buffer (row m at 128 m bytes) before the hand-over
#pragma unroll
            for (int mh = 0; mh < 16; mh++) {
                const uint32_t e = (mh * 16 + thr) * FUSED_COLGROUPS + g;
                if (NQ == 4) {
                    const uint4 v = reinterpret_cast<const uint4 *>(xbuf)[e];
                    A[mh][0] = v.x - sub; A[mh][1] = v.y - sub; A[mh][NQ - 2] = v.z - sub; A[mh][NQ - 1] = v.w - sub;
                } else if (NQ == 2) {
                    const uint2 v = reinterpret_cast<const uint2 *>(xbuf)[e];
                    A[mh][0] = v.x - sub; A[mh][NQ - 1] = v.y - sub;
                } else {
                    A[mh][0] = xbuf[e] - sub;
                }
            }
        } else {
            const uint8_t *src = reinterpret_cast<const uint8_t *>(ti.oldm) + ((size_t)thr * 32768 + (size_t)G * COLW) * 2;
#pragma unroll
            for (int mh = 0; mh < 16; mh++) {
                if (NQ == 4) {
                    const uint4 v = __ldcg(reinterpret_cast<const uint4 *>(src + (size_t)mh * 16 * 65536));
                    A[mh][0] = v.x - sub; A[mh][1] = v.y - sub; A[mh][NQ - 2] = v.z - sub; A[mh][NQ - 1] = v.w - sub;
                } else if (NQ == 2) {
                    const uint2 v = __ldcg(reinterpret_cast<const uint2 *>(src + (size_t)mh * 16 * 65536));
                    A[mh][0] = v.x - sub; A[mh][NQ - 1] = v.y - sub;
                } else {
                    A[mh][0] = __ldcg(reinterpret_cast<const uint32_t *>(src + (size_t)mh * 16 * 65536)) - sub;
                }
            }
        }
        TRACE(trace_n, tau, 2);
        fused_stage<1, CAREFUL>(A, labels, tab, s0, ring_chunk, st, Xv, Kv);
        fused_stage<2, CAREFUL>(A, labels, tab, s0, ring_chunk, st, Xv, Kv);
        fused_stage<3, CAREFUL>(A, labels, tab, s0, ring_chunk, st, Xv, Kv);
        fused_stage<4, CAREFUL>(A, labels, tab, s0, ring_chunk, st, Xv, Kv);
        // ---- exchange: rows m = mh*16 + ml ----
        // The exchange happens in place: a thread writes exactly the rows it read (the swizzle only moves elements
        // between lanes of its own warp).  With a single buffer the previous tile's round-2 reads must be over.
        if (BULK_LOAD) __syncwarp();
        else if (XCHG_BUFS == 1) bar_compute();
#pragma unroll
        for (int mh = 0; mh < 16; mh++) {
            const uint32_t e = xchg_index(mh * 16 + thr, g);
            if (NQ == 4)      reinterpret_cast<uint4 *>(xbuf)[e] = make_uint4(A[mh][0], A[mh][1], A[mh][NQ - 2], A[mh][NQ - 1]);
            else if (NQ == 2) reinterpret_cast<uint2 *>(xbuf)[e] = make_uint2(A[mh][0], A[mh][NQ - 1]);
            else              xbuf[e] = A[mh][0];
        }
    }
    bar_compute();
    {
        // ---- round 2: thread = (mh = thr, g); registers = 16 ml rows ----
        uint32_t thr, g;
        round2_map(tid, thr, g);
        const uint32_t G = tau * FUSED_COLGROUPS + g;
#pragma unroll
        for (int ml = 0; ml < 16; ml++) {
            const uint32_t e = xchg_index(thr * 16 + ml, g);
            if (NQ == 4) {
                const uint4 v = reinterpret_cast<const uint4 *>(xbuf)[e];
                A[ml][0] = v.x; A[ml][1] = v.y; A[ml][NQ - 2] = v.z; A[ml][NQ - 1] = v.w;
            } else if (NQ == 2) {
                const uint2 v = reinterpret_cast<const uint2 *>(xbuf)[e];
                A[ml][0] = v.x; A[ml][NQ - 1] = v.y;
            } else {
                A[ml][0] = xbuf[e];
            }
        }
        {
            // this warp is through with the data buffer: the producer warp may bulk-copy the tile after next into it
            if (BULK_LOAD) fence_proxy_async_smem();
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(freeb);
        }
        TRACE(trace_n, tau, 3);
        fused_stage<5, CAREFUL>(A, labels, tab, s0, ring_chunk, st, Xv, Kv);
        fused_stage<6, CAREFUL>(A, labels, tab, s0, ring_chunk, st, Xv, Kv);
        fused_stage<7, CAREFUL>(A, labels, tab, s0, ring_chunk, st, Xv, Kv);
        fused_stage<8, CAREFUL>(A, labels, tab, s0, ring_chunk, st, Xv, Kv);
        TRACE(trace_n, tau, 4);
        // ---- output: slot (m, j) holds state (j << 8) | m; per column 16 consecutive ml = 32 B, four lanes = one line ----
        {
            uint16_t *dst = ti.newm + ((size_t)(G * COLW) << 8) + thr * 16;
#pragma unroll
            for (int q = 0; q < NQ; q++) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    uint32_t w[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) w[i] = __byte_perm(A[2 * i][q], A[2 * i + 1][q], h ? 0x7632 : 0x5410);
                    st_v8(dst + ((q * 2 + h) << 8), w);
                }
            }
        }
        // ---- statistics of the final stage (only in the passes that measure; state 0 always) ----
        if (CAREFUL || ti.measure) {
            const uint32_t mn = __reduce_min_sync(0xffffffffu, tile_min(A));
            const uint32_t mx = __reduce_max_sync(0xffffffffu, tile_max(A));
            if ((tid & 31) == 0) {
                const uint32_t b = (blockIdx.x * (FUSED_THREADS / 32) + (tid >> 5)) % STAT_BUCKETS;
                atomicMin(&st->minP[FK][b][0], mn);
                atomicMax(&st->maxP[b][0], mx);
            }
        }
        __syncwarp();
        if (tau == 0 && tid < FK) st->s0[tid + 1] = s0[tid + 1];              // tid 0 wrote them (same warp)
        TRACE(trace_n, tau, 5);
    }
}

// ------------------------------------------------------------------------------------------
// protocol warp
// ------------------------------------------------------------------------------------------
// 64-column build only: a decoder alone (32-column build) keeps its three buffers in the L2 anyway
#ifndef V224_DISCARD_INPUT
#define V224_DISCARD_INPUT (V224_TILE_COLS_LOG2 == 6)
#endif
constexpr bool DISCARD_INPUT = V224_DISCARD_INPUT;
static_assert(!DISCARD_INPUT || FUSED_TILE_COLS == 64, "a discarded line is 128 bytes = one tile's piece of a metric row");
constexpr unsigned SPIN_LIMIT = 20u * 1000u * 1000u;      // polls of >= 20-40 ns: seconds.  A wait that long means a broken invariant.

__device__ __forceinline__ void slot_reset(PassSlot &s, int pass)
{
    stats_reset(s.st);
    s.done_word = done_word_fresh(pass);          // no tile done yet, tagged with the pass the slot now serves
}

__global__ void k_persist_begin(Ctl *c, int npasses, int force_careful, long long expected_T, int no_discard, const uint32_t *passtab, int measure_all)
{
    PersistCtl &pc = c->pc;
    pc.next_item = 0;
    pc.resolved_upto = 0;
    pc.npasses = npasses;
    pc.force_careful = force_careful;
    pc.no_discard = no_discard;
    pc.measure_all = measure_all;
    pc.lbP = (unsigned)c->sub;                 // the buffer the launch starts from: min >= sub, max <= sub + spread (exact after a measured pass)
    pc.ubP = (unsigned)(c->sub + c->spread);
    pc.Ostore = c->O - c->sub;                 // external convention: R = (P - sub) + O
    pc.maxR_prev = c->maxR;
    for (int i = 0; i < PSLOTS; i++) { slot_reset(pc.slot[i], i); pc.slot[i].pass_word = 0; }
    pc.passtab = passtab;
    // state 0 can gain at most cz per pass (PASSTAB_CZ): a pass that cannot reach the trigger need not record per-stage minima
    const long long cz0 = passtab[PASSTAB_CZ], cz1 = npasses > 1 ? passtab[PASSTAB_WORDS + PASSTAB_CZ] : 510ll * FK;
    const int careful0 = force_careful || (c->R0 + cz0 >= RENORM_TRIGGER);
    const int careful1 = force_careful || (c->R0 + cz0 + cz1 >= RENORM_TRIGGER);
    pc.slot[0].pass_word = make_pass_word(0, careful0, c->sub, !no_discard && discard_ok(c->maxR, c->spread, FK));
    pc.slot[1].pass_word = make_pass_word(1, careful1, 0, !no_discard && discard_ok(c->maxR, c->spread, 2 * FK));
    int stop = npasses;
    if (c->spread > MAX_FAST_SPREAD || c->error || c->T != expected_T) stop = 0;
    // non-careful passes are not validated stage by stage: keep them well away from saturation
    if (!careful0 && c->maxR + 510ll * FK > 32767) stop = 0;
    if (stop > 1 && !careful1 && c->maxR + 510ll * 2 * FK > 32767) stop = 1;
    pc.stop_pass = stop;
}

// Resolve pass n (run by the protocol thread that completed the pass's last tile).  Replays the reference's
// renormalisation test per stage, validates that the reference could not have saturated, commits
// the pass (or invalidates it), and publishes the parameters of pass n+2.
__device__ void resolve_persist(Ctl *c, int n)
{
    PersistCtl &pc = c->pc;
    for (unsigned spins = 0; (int)ld_acquire(&pc.resolved_upto) != n;) {
        __nanosleep(64);
        if (++spins > SPIN_LIMIT) { atomicOr((unsigned *)&c->error, 16u); atomicMin(&pc.stop_pass, 0); return; }
    }
    PassSlot &sl = pc.slot[n % PSLOTS];
    const unsigned long long pw = *(volatile unsigned long long *)&sl.pass_word;
    const bool careful = pass_word_careful(pw);
    const int sub = pass_word_sub(pw);
    bool valid = !c->error && n < *(volatile int *)&pc.stop_pass;
    long long O = pc.Ostore + sub;                     // offset of the values this pass loaded
    long long maxR = pc.maxR_prev;                     // exact at pass start; +510 per stage bounds it inside
    long long renormals = 0;
    int count = 0;
    for (int t = 1; valid && t <= FK; t++) {
        // the adds of stage t clip in the reference iff some R + branch metric exceeds SHRT_MAX (:296-299)
        if (maxR + 510 > 32767) { valid = false; break; }
        maxR += 510;
        const long long R0 = (long long)*(volatile unsigned *)&sl.st.s0[t] + O;
        if (R0 >= RENORM_TRIGGER) {                                        // viterbi224_sse2.c:351
            const unsigned mnt = stats_min(sl.st, t);
            if (!(careful || t == FK) || mnt == 0xffffffffu) { c->error |= 1; valid = false; break; }
            const long long minR = (long long)mnt + O;                    // :358-366
            renormals += (minR < 0 ? minR + 65536 : minR) + 32768;        // :354,:366,:367 (uint16 read of the minimum)
            count++;
            O -= minR + 32768;                                             // :373
            maxR -= minR + 32768;
        }
    }
    const unsigned z = *(volatile unsigned *)&sl.st.s0[FK];
    unsigned mn, mx;
    if (pass_word_measure(pw) || careful) {
        mn = stats_min(sl.st, FK);
        mx = stats_max(sl.st);
        if (valid && (mn == 0xffffffffu || mx < mn)) { c->error |= 2; valid = false; }
    } else {
        // not measured: the pass loaded P - sub (sub <= lbP by construction); a metric never decreases, the largest grows by <= 255 per stage
        mn = pc.lbP - (unsigned)sub;
        mx = pc.ubP - (unsigned)sub + (unsigned)(MAX_GROWTH_PER_STAGE * FK);
    }
    if (valid && (long long)mx - mn > MAX_FAST_SPREAD) valid = false;
    if (valid) {
        pc.Ostore = O;
        pc.lbP = mn;
        pc.ubP = mx;
        pc.maxR_prev = (long long)mx + O;
        c->renormals += renormals;
        c->renorm_count += count;
        // external view (what the host and the single-stage kernel see between launches)
        c->sub = (int)mn;
        c->O = O + mn;
        c->R0 = (long long)z + O;
        c->maxR = (long long)mx + O;
        c->spread = (long long)mx - mn;
        c->T += FK;
        c->cur = (c->cur + 1) % NBUF;
        c->n_fused++;
        if (careful) c->n_careful++;
        // parameters of pass n+2 (its slot is free: pass n-2 is long resolved)
        PassSlot &nx = pc.slot[(n + 2) % PSLOTS];
        slot_reset(nx, n + 2);
        const int sub1 = pass_word_sub(*(volatile unsigned long long *)&pc.slot[(n + 1) % PSLOTS].pass_word);
        long long cz = 510ll * 2 * FK;
        if (n + 2 < pc.npasses) cz = (long long)pc.passtab[(size_t)(n + 1) * PASSTAB_WORDS + PASSTAB_CZ] + pc.passtab[(size_t)(n + 2) * PASSTAB_WORDS + PASSTAB_CZ];
        const int careful2 = pc.force_careful || ((long long)z + O + cz >= RENORM_TRIGGER);
        if (!careful2 && (long long)mx + O + 510ll * 2 * FK > 32767) {
            if (n + 2 < pc.stop_pass) pc.stop_pass = n + 2;
        }
        __threadfence();                                                   // the reset slot before the word that opens it
        // (the statistics are pass n's: pass n+2 ends 2 * FK stages later)
        const int discard2 = !pc.no_discard && discard_ok((long long)mx + O, (long long)mx - mn, 2 * FK);
        // Pass n+2 reduces min / max if it is careful, periodically, at the end of the launch (the host and the one-stage kernel
        // then see exact values), and whenever its output could come near anything the values are compared with -- the
        // saturation watch, the spread limit, the 16 bits of P: the bounds below hold for pass n+2's output (two passes of
        // growth), so every comparison that can go the other way is made on measured values and a loose bound never costs a decision.
        const long long slack = MAX_GROWTH_PER_STAGE * FK * 2;
        const int measure2 = pc.measure_all || careful2 || (n + 2) % MEASURE_EVERY == MEASURE_EVERY - 1 || n + 2 >= pc.npasses - 1 ||
                             (long long)mx + O + 510ll * 2 * FK + slack > 32767 || (long long)mx - mn + 510ll * 2 * FK + slack > MAX_FAST_SPREAD ||
                             (long long)mx + slack > 40000;
        *(volatile unsigned long long *)&nx.pass_word = make_pass_word(n + 2, careful2, (int)mn - sub1, discard2, measure2);   // sub <= min of pass n+1's output
    } else {
        // the pass (and anything that already consumed its output) is discarded; its input buffer is intact
        if (n < pc.stop_pass) { pc.stop_pass = n; c->n_invalidated++; }
    }
    __threadfence();
    st_release(&pc.resolved_upto, (unsigned)(n + 1));
}

// Two protocol warps per CTA, so that the two chains of L2 round trips run side by side instead of one after the other
// (measured: the serial chain -- publish, claim, poll, fence, tensor copy -- was longer than a tile's compute time and set
// the CTA's cycle):
//   producer: for tile k: wait until bookkeeping slot k % NSLOT is free, claim the next (pass, decoder, tile) item, poll its
//             pass parameters and the previous pass's completion counters, fence, issue the tile's tensor copy, hand over.
//   retirer : for tile k: wait until the compute warps have issued its stores, free the slot (the tile's identity moves into
//             registers), publish the tile with one release-add on the pass's done word; the pass's last tile runs the resolver.
// Neither waits for the other except through the slot hand-back, and a dependency wait of the producer never delays a
// publication (a tile this CTA still has to publish may be exactly what its next tile depends on).
__device__ void producer_warp(FusedSmem &sm, const MultiArgs &m)
{
    const unsigned lane = threadIdx.x & 31;
    unsigned *queue = &m.ctx[0].ctl->pc.next_item;
    const unsigned per_pass = (unsigned)m.nctx * FUSED_TILES;
    for (unsigned k = 0;; k++) {
        const unsigned b = k % NSLOT;
        if (k >= (unsigned)NSLOT) mbar_wait_proto(&sm.slotfree[b], (k / NSLOT - 1) & 1);       // tile k - NSLOT is done with the slot
        if (BULK_LOAD && k >= 2) mbar_wait_proto(&sm.freeb[k & 1], ((k - 2) >> 1) & 1);      // tile k - 2 has read the data buffer
        unsigned item = 0;
        if (lane == 0) item = atomicAdd(queue, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        const int n = (int)(item / per_pass);
        const unsigned r = item % per_pass, w = r % FUSED_TILES, s = r / FUSED_TILES;
#if defined(V224_ORDER_DIAG) && V224_TILE_COLS_LOG2 == 5
        // A/B: subgroups (h = tile >> 8, p = tile & 3) in anti-diagonal order h + p instead of class by class
        uint32_t tau;
        {
            constexpr unsigned char HP[16][2] = {{0,0},{0,1},{1,0},{0,2},{1,1},{2,0},{0,3},{1,2},{2,1},{3,0},{1,3},{2,2},{3,1},{2,3},{3,2},{3,3}};
            const unsigned sg = w >> 6, i = w & 63u;
            tau = HP[sg][0] * 256u + i * 4u + HP[sg][1];
        }
#else
        const uint32_t tau = (w % 256u) * TILE_CLASSES + w / 256u;          // a pass emits its tile classes in turn
#endif
        if (n >= m.npasses) {
            // no more work: tell the compute warps and the retirer
            if (lane == 0) sm.info[b].go = -1;
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.full[b]);
            return;
        }
        TRACE(TKEY(n, s), tau, 0);
        const PersistArgs &a = m.ctx[s];
        PersistCtl &pc = a.ctl->pc;
        if (!BULK_LOAD) {
            // the pass table does not depend on anything that is still running (bulk mode: it travels with the tile)
            const uint4 *src = reinterpret_cast<const uint4 *>(a.passtab + (size_t)n * PASSTAB_WORDS);
            for (unsigned e = lane; e < PASSTAB_WORDS / 4; e += 32) reinterpret_cast<uint4 *>(sm.tab[b])[e] = __ldcg(src + e);
        }
        unsigned long long pw = 0;
        bool stopped = false;
        for (unsigned spins = 0;; spins++) {
            // one round of parallel loads: pass parameters, stop mark, the previous pass's completion counters
            unsigned long long v = 0;
            if (lane == 0) v = ld_relaxed64(&pc.slot[n % PSLOTS].pass_word);
            if (lane == 1) v = ld_relaxed(&pc.stop_pass);
            if (lane == 2 && n > 0) v = ld_relaxed64(&pc.slot[(n - 1) % PSLOTS].done_word);
            pw = __shfl_sync(0xffffffffu, v, 0);
            const int stop = (int)(unsigned)__shfl_sync(0xffffffffu, v, 1);
            const unsigned long long dwd = __shfl_sync(0xffffffffu, v, 2);
            stopped = n >= stop;
            // tile tau reads only the 256 tiles of class tau >> 8 (== tile index mod TILE_CLASSES) of the previous pass
            // (the slots are recycled every PSLOTS passes: the done word carries the pass it counts for, so a stale word of
            // pass n - 1 - PSLOTS can never read as "complete")
            const bool ready = (unsigned)(pw >> 32) == (unsigned)(n + 1) &&
                               (n == 0 || (done_word_pass(dwd) == done_word_pass(done_word_fresh(n - 1)) && done_class_count(dwd, tau >> 8) >= 256u));
            if (stopped || ready) break;
            if (spins > SPIN_LIMIT) {            // a wait that lasts seconds means a broken invariant: flag it, never hang the GPU
                if (lane == 0) { atomicOr((unsigned *)&a.ctl->error, 16u); atomicMin(&pc.stop_pass, 0); }
                stopped = true;
                break;
            }
            __nanosleep(60);
        }
        fence_acq_rel();                           // acquire: the producers' metric stores, the pass parameters
        if (BULK_LOAD) fence_proxy_async();        // ... which the bulk-copy engine (async proxy) is about to read
        const int cur = (a.cur0 + n) % NBUF;       // buffers advance by one per resolved pass
        if (lane == 0) {
            TileInfo &ti = sm.info[b];
            ti.oldm = a.metrics[cur];
            ti.newm = a.metrics[(cur + 1) % NBUF];
            ti.ring = a.ring;
            ti.st = &pc.slot[n % PSLOTS].st;
            ti.tau = tau;
            ti.sub2 = (uint32_t)pass_word_sub(pw) * 0x10001u;
            ti.careful = (int)pass_word_careful(pw);
            ti.discard = (int)pass_word_discard(pw);
            ti.measure = (int)pass_word_measure(pw);
            ti.go = stopped ? 0 : 1;
            ti.n = n;
            ti.s = (int)s;
            ti.lab = packed_tile_labels(tau << FUSED_COLS_LOG2);
            TRACE(TKEY(n, s), tau, 1);
        }
        __syncwarp();
        if (lane == 0) {
            if (BULK_LOAD && !stopped) {
                // pass table (1 KiB) and tile (256 rows x 2 * FUSED_TILE_COLS bytes, rows 64 KiB apart, one tensor copy) -> shared memory;
                // the hand-over completes when the bytes have landed
                mbar_arrive_expect_tx(&sm.full[b], 256u * FUSED_TILE_COLS * 2u + PASSTAB_WORDS * 4u);
                bulk_g2s(sm.tab[b], a.passtab + (size_t)n * PASSTAB_WORDS, PASSTAB_WORDS * 4u, &sm.full[b]);
                tma_load_tile(sm.tile[k & 1], reinterpret_cast<const uint8_t *>(a.tmaps) + (size_t)cur * TMAP_BYTES, (int)(tau * FUSED_TILE_COLS), &sm.full[b]);
            } else {
                mbar_arrive(&sm.full[b]);
            }
        }
        __syncwarp();
    }
}

__device__ void retirer_warp(FusedSmem &sm, const MultiArgs &m)
{
    const unsigned lane = threadIdx.x & 31;
    for (unsigned k = 0;; k++) {
        const unsigned b = k % NSLOT, par = (k / NSLOT) & 1;
        mbar_wait_proto(&sm.full[b], par);
        const TileInfo &ti = sm.info[b];
        const int go = ti.go, n = ti.n, s = ti.s;
        const unsigned tau = ti.tau;
        if (go < 0) return;
        if (DISCARD_INPUT && BULK_LOAD && go > 0 && ti.discard) {
            // The tile's input is in shared memory now (the hand-over completed with the tensor copy's bytes), and every
            // 2 * FUSED_TILE_COLS-byte piece of a metric row is read by exactly ONE tile: the 256 lines are dead.  With several
            // decoders in lockstep their buffers do not fit the L2 together, and dead lines would be written back to HBM when
            // they are evicted -- 16.6 MB of the 25 MB a pass writes, 100 W of the board's 1000 W (profiles/r02_ab_discard.txt).
            // Only for passes that cannot be invalidated (pass_word): an invalidated pass is run again from this very input.
            const uint8_t *base = reinterpret_cast<const uint8_t *>(ti.oldm) + (size_t)tau * FUSED_TILE_COLS * 2;
#pragma unroll
            for (int k = 0; k < 8; k++)
                asm volatile("discard.global.L2 [%0], 128;" ::"l"(base + (size_t)(k * 32 + lane) * 65536) : "memory");
        }
        mbar_wait_proto(&sm.done[b], par);         // every compute warp has issued the tile's stores and is done with the slot
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.slotfree[b]);
        if (go > 0 && lane == 0) {
            Ctl *c = m.ctx[s].ctl;
            PassSlot &sl = c->pc.slot[n % PSLOTS];
            // release: the compute warps' stores (ordered before this thread by the mbarrier) become visible GPU-wide
            // before the tile counts as done
            const unsigned long long old = atom_acq_rel_add64(&sl.done_word, done_increment(tau % TILE_CLASSES));
            TRACE(TKEY(n, s), tau, 6);
            if (done_total(old) == FUSED_TILES - 1) resolve_persist(c, n);
            TRACE(TKEY(n, s), tau, 7);
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(CTA_THREADS, FUSED_CTAS_PER_SM) k_acs_persist(MultiArgs m)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    FusedSmem &sm = *reinterpret_cast<FusedSmem *>(smem_raw);
    const uint32_t tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < NSLOT; i++) { mbar_init(&sm.full[i], 1); mbar_init(&sm.done[i], FUSED_THREADS / 32); }
        for (int i = 0; i < NSLOT; i++) mbar_init(&sm.slotfree[i], 1);
        mbar_init(&sm.freeb[0], FUSED_THREADS / 32); mbar_init(&sm.freeb[1], FUSED_THREADS / 32);
    }
    __syncthreads();
    if (tid >= FUSED_THREADS) {
        if (tid < FUSED_THREADS + 32) producer_warp(sm, m);
        else retirer_warp(sm, m);
        return;
    }
    // the thread's own part of the branch labels (its row / column group in round 1 and in round 2), fixed for the kernel
    uint32_t labthr;
    {
        uint32_t t1, g1, t2, g2;
        round1_map(tid, t1, g1);
        round2_map(tid, t2, g2);
        labthr = packed_thread_labels((t1 << 15) | (g1 << COLW_LOG2), (t2 << 19) | (g2 << COLW_LOG2));
    }
    for (unsigned k = 0;; k++) {
        const unsigned b = k % NSLOT;
        mbar_wait(&sm.full[b], (k / NSLOT) & 1);
        const TileInfo &ti = sm.info[b];
        const int go = ti.go;
        if (go < 0) break;
        if (go > 0) {
            uint32_t *xbuf = sm.tile[XCHG_BUFS == 2 ? (k & 1) : 0];
            if (ti.careful) fused_tile<true>(xbuf, sm.tab[b], sm.s0, ti, &sm.freeb[k & 1], labthr, TKEY(ti.n, ti.s));
            else            fused_tile<false>(xbuf, sm.tab[b], sm.s0, ti, &sm.freeb[k & 1], labthr, TKEY(ti.n, ti.s));
        } else {
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&sm.freeb[k & 1]);     // a skipped tile still hands its buffer on
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&sm.done[b]);      // this warp's stores of tile k are issued; it is done with slot b
    }
}

// ------------------------------------------------------------------------------------------
// launch
// ------------------------------------------------------------------------------------------
size_t passtab_bytes(int npasses) { return (size_t)npasses * PASSTAB_WORDS * sizeof(uint32_t); }

cudaError_t build_metric_tensor_maps(uint16_t *const *metrics, void *dev_out, cudaStream_t st, const char **why)
{
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                 const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    static std::mutex encode_mu;
    std::lock_guard<std::mutex> encode_lk(encode_mu);
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) { if (why) *why = "cuTensorMapEncodeTiled not available"; return e != cudaSuccess ? e : cudaErrorNotSupported; }
        encode = reinterpret_cast<EncodeFn>(fn);
    }
    static_assert(sizeof(CUtensorMap) == TMAP_BYTES, "tensor map size");
    alignas(64) CUtensorMap maps[NBUF];
    for (int i = 0; i < NBUF; i++) {
        const cuuint64_t dims[2] = {32768, 256};                 // innermost first: columns j, rows m
        const cuuint64_t strides[1] = {65536};                   // bytes between rows
        const cuuint32_t box[2] = {FUSED_TILE_COLS, 256};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = encode(&maps[i], CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, metrics[i], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { if (why) *why = "cuTensorMapEncodeTiled failed"; return cudaErrorInvalidValue; }
    }
    cudaError_t e = cudaMemcpyAsync(dev_out, maps, sizeof maps, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(st);                            // `maps` lives on this stack frame
}

// One persistent launch over m.nctx decoders x m.npasses passes (+ one table build and one begin kernel per decoder).
cudaError_t launch_persist(const MultiArgs &m, cudaStream_t st)
{
    int dev = 0;
    cudaGetDevice(&dev);
    // per-device launch geometry, looked up once; callers may come from several host threads (one per GPU)
    constexpr int MAXDEV = 64;
    static std::mutex mu;
    static int checked[MAXDEV], slots[MAXDEV], sm_count[MAXDEV];
    if (dev < 0 || dev >= MAXDEV) return cudaErrorInvalidDevice;
    int nslots = 0;
    {
        std::lock_guard<std::mutex> lk(mu);
        if (!checked[dev]) {
            cudaError_t e = cudaFuncSetAttribute(k_acs_persist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FusedSmem));
            if (e != cudaSuccess) return e;
            cudaFuncSetAttribute(k_acs_persist, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            int per_sm = 0, sms = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_acs_persist, CTA_THREADS, sizeof(FusedSmem));
            if (e != cudaSuccess) return e;
            e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            if (e != cudaSuccess) return e;
            slots[dev] = per_sm * sms;
            sm_count[dev] = sms;
            checked[dev] = 1;
        }
        nslots = slots[dev];
    }
    if (m.grid_limit > 0 && m.grid_limit < nslots) nslots = m.grid_limit;
    if (m.grid_limit < 0 && -m.grid_limit * sm_count[dev] < nslots) nslots = -m.grid_limit * sm_count[dev];          // that many CTAs per SM
    for (int s = 0; s < m.nctx; s++) {
        const PersistArgs &a = m.ctx[s];
        k_build_passtab<<<m.npasses, PASSTAB_WORDS, 0, st>>>(a.passtab, a.syms + 2 * (size_t)a.pos0, m.npasses, a.T0, a.len, a.row_fmt);
        k_persist_begin<<<1, 1, 0, st>>>(a.ctl, m.npasses, a.force_careful, a.T0, m.no_discard, a.passtab, m.measure_all);
    }
    const long long items = (long long)m.npasses * m.nctx * FUSED_TILES;
    const int grid = (int)(items < nslots ? items : nslots);
    k_acs_persist<<<grid, CTA_THREADS, sizeof(FusedSmem), st>>>(m);
    return cudaGetLastError();
}

} // namespace V224_NS

#ifdef V224_BRIDGE
// The 32-column builds are reached from the runtime (compiled against the 64-column headers) through two entries each,
// <V224_BRIDGE>_launch_persist and <V224_BRIDGE>_build_metric_tensor_maps; the argument structures have the same layout in
// every namespace (nothing in them depends on the tile shape).
#define V224_CAT2(a, b) a##b
#define V224_CAT(a, b) V224_CAT2(a, b)
extern "C" cudaError_t V224_CAT(V224_BRIDGE, _launch_persist)(const void *multi_args, cudaStream_t st)
{
    return V224_NS::launch_persist(*static_cast<const V224_NS::MultiArgs *>(multi_args), st);
}
extern "C" cudaError_t V224_CAT(V224_BRIDGE, _build_metric_tensor_maps)(uint16_t *const *metrics, void *dev_out, cudaStream_t st, const char **why)
{
    return V224_NS::build_metric_tensor_maps(metrics, dev_out, st, why);
}
#endif

#ifdef V224_TRACE
extern "C" int v224_debug_read_trace(unsigned long long *host, unsigned long long n)
{
    return (int)cudaMemcpyFromSymbol(host, V224_NS::g_trace, n * sizeof(unsigned long long));
}
extern "C" int v224_debug_read_smid(unsigned *host, unsigned long long n)
{
    return (int)cudaMemcpyFromSymbol(host, V224_NS::g_smid, n * sizeof(unsigned));
}
#endif
