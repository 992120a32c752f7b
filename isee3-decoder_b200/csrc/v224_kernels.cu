// v224_kernels.cu -- sm_100a kernels of the B200 viterbi224 decoder.
//
//   k_init            metric fill + control-block reset               (viterbi224_sse2.c:37-53)
//   k_acs_fused       8 trellis stages per HBM pass                   (viterbi224_sse2.c:264-328, x8)
//   k_acs_single      1 trellis stage; SAT variant = exact int16-saturating arithmetic
//   k_chainback_*     frame traceback, speculative segments + verify  (viterbi224_sse2.c:113-161)
//   k_walk            decodebit / decodeword walk                     (viterbi224_sse2.c:164-243)
//   k_stream_trace    one walk per output bit (vdecode.c:145-152 pattern, batched)
//   k_argmin, k_minmax, k_export_row, k_export_metrics, k_import_metrics
//
// Renormalisation (viterbi224_sse2.c:351-377) never touches HBM: the last CTA of every pass
// ("resolver") replays the reference's test on state 0 and folds the adjustment into the
// 64-bit offset Ctl::O (R_reference = P_hbm + O).
#include "v224_common.cuh"
#include "v224_fused_core.cuh"
#include "v224_kernels.h"
#include <cstdio>

namespace v224 {

#ifdef V224_TRACE
__device__ unsigned long long g_trace[64 * 1024 * 8];      // [pass < 64][tile][event]
__device__ unsigned g_smid[64 * 1024];
__device__ __forceinline__ unsigned long long gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define TRACE(n, tau, ev) do { if (threadIdx.x == 0 && (n) < 64) { g_trace[((n) * 1024 + (tau)) * 8 + (ev)] = gtime(); if ((ev) == 1) { unsigned sm_; asm volatile("mov.u32 %0, %smid;" : "=r"(sm_)); g_smid[(n) * 1024 + (tau)] = sm_; } } } while (0)
#else
#define TRACE(n, tau, ev) do { } while (0)
#endif


// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_cs_v4(void *p, uint4 v)   // streaming store: decision rows are write-once
{
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_v8(void *p, const uint32_t (&v)[8])   // 256-bit store (sm_100+)
{
    asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
                 "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}

// Reset the per-pass statistics (resolver, and k_init).
__device__ __forceinline__ void reset_stats(Ctl *c)
{
#pragma unroll
    stats_reset(c->st);
    c->ticket = 0;
}

// The resolver: executed by ONE thread after every CTA of a pass has published its statistics.
// Replays viterbi224_sse2.c:351-377 for each of the pass's `ks` stages on the virtual
// reference metric R = P + O, then prepares the next pass (sub, R0, maxR) and advances T/pos.
__device__ void resolve_pass(Ctl *c, int ks, bool careful, bool sat)
{
    long long O = sat ? -32768ll : c->O;      // a SAT stage stores P = R + 32768
    for (int t = 1; t <= ks; t++) {
        const long long R0 = (long long)*(volatile unsigned *)&c->st.s0[t] + O;
        if (R0 >= RENORM_TRIGGER) {                                   // :351 state 0 only
            const unsigned mnt = stats_min(c->st, t);
            if (!(careful || t == ks) || mnt == 0xffffffffu) { c->error |= 1; break; }
            const long long minR = (long long)mnt + O;                // :358-366 global minimum
            // :354,:366 the minimum is read through a uint16_t: a negative one counts 65536 more
            const long long adjust = (minR < 0 ? minR + 65536 : minR) + 32768;
            c->renormals += adjust;                                   // :367
            c->renorm_count += 1;
            O -= minR + 32768;                                        // :373 (mod 2^16 == this, min -> SHRT_MIN)
        }
    }
    const unsigned mn = stats_min(c->st, ks), mx = stats_max(c->st), z = *(volatile unsigned *)&c->st.s0[ks];
    if (mn == 0xffffffffu || mx < mn) c->error |= 2;
    O += mn;                         // the next pass subtracts mn from every P while loading
    c->sub = (int)mn;
    c->O = O;
    c->R0 = (long long)z - mn + O;
    c->maxR = (long long)mx - mn + O;
    c->spread = (long long)mx - mn;
    c->T += ks;
    c->cur = (c->cur + 1) % NBUF;
    reset_stats(c);
}

// ------------------------------------------------------------------------------------------
// init
// ------------------------------------------------------------------------------------------
// P = bias everywhere, 0 in the start state, O = SHRT_MIN  <=>  R = SHRT_MIN+bias / SHRT_MIN.
__global__ void __launch_bounds__(256) k_init(uint16_t *m0, Ctl *c, uint32_t start_state, int bias, int start_value)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;     // one uint4 (8 metrics) per thread
    const uint32_t b = (uint32_t)bias | ((uint32_t)bias << 16);
    uint4 v = make_uint4(b, b, b, b);
    reinterpret_cast<uint4 *>(m0)[i] = v;
    __syncthreads();
    if (start_state / 8 == i && start_value >= 0) m0[start_state] = (uint16_t)start_value;
    if (i == 0) {
        c->O = -32768;
        c->renormals = 0;
        c->T = 0;
        c->renorm_count = 0;
        c->sub = 0;
        c->R0 = (start_state == 0 && start_value >= 0 ? start_value : bias) - 32768;
        c->maxR = bias - 32768;
        c->spread = bias;
        c->cur = 0;
        c->error = 0;
        reset_stats(c);
    }
}

// ------------------------------------------------------------------------------------------
// fused 8-stage pass
// ------------------------------------------------------------------------------------------
struct __align__(16) FusedSmem {
    uint32_t tile[256 * FUSED_TILE_COLS / 2];   // 256 rows x FUSED_TILE_COLS columns of uint16 (32 / 16 KiB)
    uint32_t optab[OPTAB_WORDS];                 // 1 KiB
};

// Operand tables of a whole launch: one 1 KiB table per pass, from that pass's 8 symbol pairs.
__global__ void __launch_bounds__(256) k_build_optab(uint32_t *optab, const uint8_t *syms, int npasses)
{
    const int pass = blockIdx.x;
    if (pass < npasses) optab[(size_t)pass * OPTAB_WORDS + threadIdx.x] = optab_entry(threadIdx.x, syms + 2 * (size_t)pass * FK);
}

template <int T>
__device__ __forceinline__ void fused_stage(uint32_t (&A)[16][4], uint32_t pbase, const uint32_t *optab, uint32_t *ring, uint8_t *row_fmt,
                                            int len, PassStats &st, long long T0, bool careful, bool first, uint32_t chunk,
                                            int fmt_base)
{
    uint32_t dw[4];
    acs_stage<T>(A, pbase, optab, dw);
    const long long row = (T0 + T - 1) % len;
    st_cs_v4(reinterpret_cast<uint8_t *>(ring) + (size_t)row * ROWBYTES + (size_t)chunk * 16, make_uint4(dw[0], dw[1], dw[2], dw[3]));
    if (first) {
        st.s0[T] = A[0][0] & 0xffffu;            // slot 0 always holds state 0
        row_fmt[row] = (uint8_t)(fmt_base + T);
    }
    if (careful && T < FK) {
        uint32_t mn = __reduce_min_sync(0xffffffffu, tile_min(A));
        if ((threadIdx.x & 31) == 0) atomicMin(&st.minP[T][(blockIdx.x * 4 + (threadIdx.x >> 5)) % STAT_BUCKETS][0], mn);
    }
}

// The body shared by the per-pass kernel and the persistent kernel: one tile, eight stages.
// The tile is the column groups [g0, g0 + ncg) (8 columns each; ncg = 8, or 6 in the balanced partition);
// thread tid = thr * ncg + g handles row group thr and column group g0 + g; threads >= 16 * ncg only keep
// the CTA barriers company (a 6-group tile leaves its fourth warp idle).
// LDCG: metrics are read through L2 only (another SM wrote them, possibly within this launch).
__device__ __forceinline__ void fused_tile(FusedSmem &sm, const uint16_t *oldm, uint16_t *newm, uint32_t *ring, uint8_t *row_fmt, int len,
                                           const uint32_t *optab_g, PassStats &st, long long T0, uint32_t sub,
                                           bool careful, uint32_t g0, uint32_t ncg, int fmt_base, int trace_n = 1 << 30, uint32_t trace_id = 0)
{
    const uint32_t tid = threadIdx.x;
    const bool active = tid < 16u * ncg;
    const uint32_t thr = tid / ncg, g = tid % ncg;                 // row group (16), column group within the tile
    const uint32_t G = g0 + g;                                     // global column group: columns 8G .. 8G+7
    const uint32_t chunk = g0 * 16u + tid;                         // = g0*16 + thr*ncg + g, see fused_bit_address()
    const bool first = G == 0 && thr == 0;
    uint4 *t4 = reinterpret_cast<uint4 *>(sm.tile);

    // operand table of this pass (precomputed by k_build_optab): 1 KiB -> shared memory.  The persistent
    // kernel fetches it before it waits for the previous pass (optab_g == nullptr here).
    if (optab_g != nullptr)
        for (int e = tid; e < OPTAB_WORDS / 4; e += FUSED_THREADS)
            reinterpret_cast<uint4 *>(sm.optab)[e] = __ldcg(reinterpret_cast<const uint4 *>(optab_g) + e);

    uint32_t A[16][4];
    if (active) {
        // ---- round 1: thread = (ml = thr, g); registers = 16 mh rows x 8 columns ----
        const uint4 *src = reinterpret_cast<const uint4 *>(oldm) + (size_t)thr * 4096 + G;
#pragma unroll
        for (int mh = 0; mh < 16; mh++) {
            const uint4 v = __ldcg(src + (size_t)mh * 16 * 4096);
            A[mh][0] = v.x - sub; A[mh][1] = v.y - sub; A[mh][2] = v.z - sub; A[mh][3] = v.w - sub;
        }
    }
    __syncthreads();                                               // optab ready
#ifdef V224_TRACE
    if (tid == 0 && trace_n < 64) { unsigned keep = A[0][0] ^ A[15][3]; if (keep == 0x12345678u) g_trace[0] = 0; g_trace[(trace_n * 1024 + trace_id) * 8 + 2] = gtime(); }
#endif
    if (active) {
        const uint32_t pbase = (thr << 15) | (G << 3);
        fused_stage<1>(A, pbase, sm.optab, ring, row_fmt, len, st, T0, careful, first, chunk, fmt_base);
        fused_stage<2>(A, pbase, sm.optab, ring, row_fmt, len, st, T0, careful, first, chunk, fmt_base);
        fused_stage<3>(A, pbase, sm.optab, ring, row_fmt, len, st, T0, careful, first, chunk, fmt_base);
        fused_stage<4>(A, pbase, sm.optab, ring, row_fmt, len, st, T0, careful, first, chunk, fmt_base);
        // ---- exchange: rows m = mh*16 + ml; element (row, g) at 16-byte index row*ncg + g.  A quarter-warp
        // touches 8 consecutive 16-byte slots when ncg = 8 (conflict-free); ncg = 6 costs a few 2-way conflicts ----
#pragma unroll
        for (int mh = 0; mh < 16; mh++) t4[(mh * 16 + thr) * ncg + g] = make_uint4(A[mh][0], A[mh][1], A[mh][2], A[mh][3]);
    }
    __syncthreads();
    if (active) {
#pragma unroll
        for (int ml = 0; ml < 16; ml++) {
            const uint4 v = t4[(thr * 16 + ml) * ncg + g];
            A[ml][0] = v.x; A[ml][1] = v.y; A[ml][2] = v.z; A[ml][3] = v.w;
        }
#ifdef V224_TRACE
        if (tid == 0 && trace_n < 64) g_trace[(trace_n * 1024 + trace_id) * 8 + 3] = gtime();
#endif
        // ---- round 2: thread = (mh = thr, g); registers = 16 ml rows ----
        const uint32_t pbase = (thr << 19) | (G << 3);
        fused_stage<5>(A, pbase, sm.optab, ring, row_fmt, len, st, T0, careful, first, chunk, fmt_base);
        fused_stage<6>(A, pbase, sm.optab, ring, row_fmt, len, st, T0, careful, first, chunk, fmt_base);
        fused_stage<7>(A, pbase, sm.optab, ring, row_fmt, len, st, T0, careful, first, chunk, fmt_base);
        fused_stage<8>(A, pbase, sm.optab, ring, row_fmt, len, st, T0, careful, first, chunk, fmt_base);
#ifdef V224_TRACE
        if (tid == 0 && trace_n < 64) g_trace[(trace_n * 1024 + trace_id) * 8 + 4] = gtime();
#endif
        // ---- statistics of the final stage ----
        {
            const uint32_t mn = __reduce_min_sync(0xffffffffu, tile_min(A));
            const uint32_t mx = __reduce_max_sync(0xffffffffu, tile_max(A));
            if ((tid & 31) == 0) {
                const uint32_t b = (blockIdx.x * 4 + (tid >> 5)) % STAT_BUCKETS;
                atomicMin(&st.minP[FK][b][0], mn);
                atomicMax(&st.maxP[b][0], mx);
            }
        }
        // ---- output: slot (m, j) holds state (j << 8) | m; per column 16 consecutive ml = 32 B ----
        {
            const uint32_t jbase = G * 8;
#pragma unroll
            for (int q = 0; q < 4; q++) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    uint32_t w[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) w[i] = __byte_perm(A[2 * i][q], A[2 * i + 1][q], h ? 0x7632 : 0x5410);
                    st_v8(newm + ((size_t)(jbase + q * 2 + h) << 8) + thr * 16, w);
                }
            }
        }
    }
}

__global__ void __launch_bounds__(FUSED_THREADS, FUSED_CTAS_PER_SM) k_acs_fused(FusedArgs a)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    FusedSmem &sm = *reinterpret_cast<FusedSmem *>(smem_raw);
    Ctl *c = a.ctl;

    // Every CTA takes the same go / no-go decision from the (quiescent) control block.
    if (c->T != a.expected_T || c->error) return;
    if (c->maxR + 510ll * FK > 32767 || c->spread > MAX_FAST_SPREAD) return;   // reference could saturate: host runs SAT stages
    const bool careful = a.force_careful || (c->R0 + 510ll * FK >= RENORM_TRIGGER);
    fused_tile(sm, a.metrics[c->cur], a.metrics[(c->cur + 1) % NBUF], a.ring, a.row_fmt, a.len, a.optab,
               c->st, c->T, (uint32_t)c->sub * 0x10001u, careful, blockIdx.x * FUSED_COLGROUPS, FUSED_COLGROUPS, 0);

    // ---- last CTA resolves the pass ----
    __shared__ unsigned s_ticket;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_ticket = atomicAdd(&c->ticket, 1u);
    }
    __syncthreads();
    if (s_ticket == gridDim.x - 1 && threadIdx.x == 0) {
        __threadfence();
        if (careful) c->n_careful++;
        c->n_fused++;
        resolve_pass(c, FK, careful, false);
    }
}

// ------------------------------------------------------------------------------------------
// persistent multi-pass kernel: one launch runs `npasses` 8-stage passes as a dataflow
// ------------------------------------------------------------------------------------------
// Work item = (pass n, tile tau), taken from an atomic queue in pass-major order.  Tile tau of pass
// n+1 reads, for every row m, 64 states out of pass n's output tile 2m + (tau >> 8): it depends on
// exactly the even (tau < 256) or the odd (tau >= 256) tiles of pass n.  Each pass therefore runs
// its even tiles first; the next pass starts as soon as those are out, while the odd tiles are still
// in flight -- no grid-wide barrier, no launch gap.  The statistics of pass n are only complete when
// the pass is, so pass n+2 is the first that can use them: `sub` and the careful flag lag two passes.
__device__ __forceinline__ unsigned ld_acquire(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(unsigned *p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void slot_reset(PassSlot &s)
{
#pragma unroll
    stats_reset(s.st);
    for (int k = 0; k < TILE_CLASSES; k++) s.done[k] = 0;
    s.done_total = 0;
}

__global__ void k_persist_begin(Ctl *c, int npasses, int force_careful, long long expected_T)
{
    PersistCtl &pc = c->pc;
    pc.next_item = 0;
    pc.next_rank = 0;
    for (int i = 0; i < 256; i++) { pc.sm_rank[i] = -1; pc.sm_slots[i] = 0; }
    pc.resolved_upto = 0;
    pc.npasses = npasses;
    pc.force_careful = force_careful;
    pc.Ostore = c->O - c->sub;                 // external convention: R = (P - sub) + O
    pc.maxR_prev = c->maxR;
    for (int i = 0; i < PSLOTS; i++) slot_reset(pc.slot[i]);
    pc.slot[0].sub = c->sub;
    pc.slot[1].sub = 0;
    pc.slot[0].careful = force_careful || (c->R0 + 510ll * FK >= RENORM_TRIGGER);
    pc.slot[1].careful = force_careful || (c->R0 + 510ll * 2 * FK >= RENORM_TRIGGER);
    int stop = npasses;
    if (c->spread > MAX_FAST_SPREAD || c->error || c->T != expected_T) stop = 0;
    // non-careful passes are not validated stage by stage: keep them well away from saturation
    if (!pc.slot[0].careful && c->maxR + 510ll * FK > 32767) stop = 0;
    if (stop > 1 && !pc.slot[1].careful && c->maxR + 510ll * 2 * FK > 32767) stop = 1;
    pc.stop_pass = stop;
}

// Resolve pass n (run by the thread that completed the pass's last tile).  Replays the reference's
// renormalisation test per stage, validates that the reference could not have saturated, commits
// the pass (or invalidates it), and publishes the parameters of pass n+2.
__device__ void resolve_persist(Ctl *c, int n)
{
    PersistCtl &pc = c->pc;
    while ((int)ld_acquire(&pc.resolved_upto) != n) __nanosleep(64);
    PassSlot &sl = pc.slot[n % PSLOTS];
    const bool careful = sl.careful != 0;
    bool valid = !c->error && n < *(volatile int *)&pc.stop_pass;
    long long O = pc.Ostore + sl.sub;                  // offset of the values this pass loaded
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
    const unsigned mn = stats_min(sl.st, FK), mx = stats_max(sl.st), z = *(volatile unsigned *)&sl.st.s0[FK];
    if (valid && (mn == 0xffffffffu || mx < mn)) { c->error |= 2; valid = false; }
    if (valid && (long long)mx - mn > MAX_FAST_SPREAD) valid = false;
    if (valid) {
        pc.Ostore = O;
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
        slot_reset(nx);
        nx.sub = (int)mn - pc.slot[(n + 1) % PSLOTS].sub;                  // <= min of pass n+1's output, >= 0
        nx.careful = pc.force_careful || ((long long)z + O + 510ll * 2 * FK >= RENORM_TRIGGER);
        if (!nx.careful && (long long)mx + O + 510ll * 2 * FK > 32767) {
            if (n + 2 < pc.stop_pass) pc.stop_pass = n + 2;
        }
    } else {
        // the pass (and anything that already consumed its output) is discarded; its input buffer is intact
        if (n < pc.stop_pass) { pc.stop_pass = n; c->n_invalidated++; }
    }
    __threadfence();
    st_release(&pc.resolved_upto, (unsigned)(n + 1));
}

__device__ __forceinline__ unsigned ld_relaxed(const void *p)
{
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Work distribution modes.
//   DYNAMIC : CTAs take (pass, tile) items from an atomic queue in pass-major order, a pass's tile classes in turn.
//             Nothing has to be co-resident because only running CTAs hold items.
//   STATIC  : CTA b owns uniform tile b in every pass; all FUSED_TILES CTAs must be co-resident, so the kernel is
//             launched cooperatively and the driver refuses instead of deadlocking.
//   BALANCED: 592 CTAs = 148 SMs x 4.  Each CTA finds out which SM it landed on and takes one of that SM's four
//             tiles of the balanced partition (8,8,6,6 or 8,6,6,6 column groups): every SM carries 14 or 13 warps
//             of work per pass instead of 16 or 12.  Cooperative launch as well.
enum { MODE_DYNAMIC = 0, MODE_STATIC = 1, MODE_BALANCED = 2 };

template <int MODE>
__global__ void __launch_bounds__(FUSED_THREADS, FUSED_CTAS_PER_SM) k_acs_persist(PersistArgs a)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    FusedSmem &sm = *reinterpret_cast<FusedSmem *>(smem_raw);
    Ctl *c = a.ctl;
    PersistCtl &pc = c->pc;
    __shared__ int s_go, s_sub, s_careful;
    __shared__ unsigned s_item, s_g0, s_ncg;
    const uint32_t tid = threadIdx.x;
    constexpr unsigned NTILES = MODE == MODE_BALANCED ? BAL_TILES : FUSED_TILES;
    constexpr int FMT = MODE == MODE_BALANCED ? ROWFMT_BALANCED : 0;

    if (MODE == MODE_BALANCED) {
        if (tid == 0) {
            unsigned smid;
            asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
            smid &= 255u;
            const unsigned slot = atomicAdd(&pc.sm_slots[smid], 1u);
            int rank;
            if (slot == 0) {
                rank = (int)atomicAdd(&pc.next_rank, 1u);
                st_release(reinterpret_cast<unsigned *>(&pc.sm_rank[smid]), (unsigned)rank);
            } else {
                while ((rank = (int)ld_acquire(reinterpret_cast<const unsigned *>(&pc.sm_rank[smid]))) < 0) __nanosleep(50);
            }
            uint32_t g0 = 0, ncg = 0;
            if (rank < BAL_SMS && slot < (unsigned)BAL_CTAS_PER_SM) balanced_tile((uint32_t)rank, slot, g0, ncg);
            else c->error |= 8;                                    // not 148 SMs x 4 CTAs: the host must not use this mode
            s_g0 = g0;
            s_ncg = ncg;
        }
        __syncthreads();
        if (s_ncg == 0) return;
    }

    for (int k = 0;; k++) {
        int n;
        uint32_t g0, ncg, tau;
        if (MODE == MODE_DYNAMIC) {
            if (tid == 0) s_item = atomicAdd(&pc.next_item, 1u);
            __syncthreads();
            const unsigned item = s_item;
            n = (int)(item / FUSED_TILES);
            const unsigned w = item % FUSED_TILES;                             // a pass emits its tile classes in turn
            tau = (w % 256u) * TILE_CLASSES + w / 256u;
            g0 = tau * FUSED_COLGROUPS; ncg = FUSED_COLGROUPS;
        } else if (MODE == MODE_STATIC) {
            n = k; tau = blockIdx.x;
            g0 = tau * FUSED_COLGROUPS; ncg = FUSED_COLGROUPS;
        } else {
            n = k; tau = blockIdx.x;
            g0 = s_g0; ncg = s_ncg;
        }
        if (n >= a.npasses) break;
        TRACE(n, tau, 0);
        PassSlot &sl = pc.slot[n % PSLOTS];
        // Everything that does not depend on the previous pass's data is fetched BEFORE waiting for it:
        // the operand table (symbols are known) and the pass parameters (published two passes ago).
        for (int e = tid; e < OPTAB_WORDS / 4; e += FUSED_THREADS)
            reinterpret_cast<uint4 *>(sm.optab)[e] = __ldcg(reinterpret_cast<const uint4 *>(a.optab + (size_t)n * OPTAB_WORDS) + e);
        if (tid == 0) {
            while ((int)ld_relaxed(&pc.resolved_upto) < n - 1) __nanosleep(200);      // parameters of pass n exist
            s_sub = (int)ld_relaxed(&sl.sub);
            s_careful = (int)ld_relaxed(&sl.careful);
            int go = n < (int)ld_relaxed(&pc.stop_pass);
            if (go && n > 0) {
                // uniform tile tau reads only the 256 tiles == (tau >> 8) mod TILE_CLASSES of the previous pass;
                // balanced tiles do not line up with that structure and wait for the whole previous pass
                const PassSlot &pv = pc.slot[(n - 1) % PSLOTS];
                const unsigned *dep = MODE == MODE_BALANCED ? &pv.done_total : &pv.done[tau >> 8];
                const unsigned need = MODE == MODE_BALANCED ? NTILES : 256u;
                unsigned spins = 0;
                while (ld_acquire(dep) < need) {
                    __nanosleep(32);
                    if ((++spins & 15u) == 0 && n >= (int)ld_relaxed(&pc.stop_pass)) { go = 0; break; }
                }
            }
            s_go = go;
        }
        __syncthreads();
        if (!s_go) break;
        TRACE(n, tau, 1);
        // buffer and stage counter advance by one per resolved pass: pass n sits at a fixed offset from the launch state
        const int cur = (a.cur0 + n) % NBUF;
        fused_tile(sm, a.metrics[cur], a.metrics[(cur + 1) % NBUF], a.ring, a.row_fmt, a.len, nullptr,
                   sl.st, a.T0 + (long long)n * FK, (uint32_t)s_sub * 0x10001u, s_careful != 0, g0, ncg, FMT, n, tau);
        __syncthreads();                       // every thread's stores and statistics are issued
        TRACE(n, tau, 5);
        if (tid == 0) {
            __threadfence();                   // ... and visible GPU-wide before the tile counts as done
            TRACE(n, tau, 6);
            if (MODE != MODE_BALANCED) atomicAdd(&sl.done[tau % TILE_CLASSES], 1u);
            if (atomicAdd(&sl.done_total, 1u) == NTILES - 1) resolve_persist(c, n);
            TRACE(n, tau, 7);
        }
    }
}

// Multi-context variant of the dynamic queue: item = (pass n, context s, tile), pass-major, contexts in turn.
// Every context keeps its own control block, buffers and resolver; only the queue head (context 0's) is shared.
__global__ void __launch_bounds__(FUSED_THREADS, FUSED_CTAS_PER_SM) k_acs_persist_multi(MultiArgs m)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    FusedSmem &sm = *reinterpret_cast<FusedSmem *>(smem_raw);
    __shared__ int s_go, s_sub, s_careful;
    __shared__ unsigned s_item;
    const uint32_t tid = threadIdx.x;
    unsigned *queue = &m.ctx[0].ctl->pc.next_item;
    const unsigned per_pass = (unsigned)m.nctx * FUSED_TILES;

    for (;;) {
        if (tid == 0) s_item = atomicAdd(queue, 1u);
        __syncthreads();
        const unsigned item = s_item;
        const int n = (int)(item / per_pass);
        if (n >= m.npasses) break;
        const unsigned r = item % per_pass, s = r / FUSED_TILES, w = r % FUSED_TILES;
        const uint32_t tau = (w % 256u) * TILE_CLASSES + w / 256u;             // a pass emits its tile classes in turn
        const PersistArgs &a = m.ctx[s];
        Ctl *c = a.ctl;
        PersistCtl &pc = c->pc;
        PassSlot &sl = pc.slot[n % PSLOTS];
        for (int e = tid; e < OPTAB_WORDS / 4; e += FUSED_THREADS)
            reinterpret_cast<uint4 *>(sm.optab)[e] = __ldcg(reinterpret_cast<const uint4 *>(a.optab + (size_t)n * OPTAB_WORDS) + e);
        if (tid == 0) {
            while ((int)ld_relaxed(&pc.resolved_upto) < n - 1) __nanosleep(200);      // parameters of pass n exist
            s_sub = (int)ld_relaxed(&sl.sub);
            s_careful = (int)ld_relaxed(&sl.careful);
            int go = n < (int)ld_relaxed(&pc.stop_pass);
            if (go && n > 0) {
                const unsigned *dep = &pc.slot[(n - 1) % PSLOTS].done[tau >> 8];
                unsigned spins = 0;
                while (ld_acquire(dep) < 256u) {
                    __nanosleep(32);
                    if ((++spins & 15u) == 0 && n >= (int)ld_relaxed(&pc.stop_pass)) { go = 0; break; }
                }
            }
            s_go = go;
        }
        __syncthreads();
        if (!s_go) continue;                   // this context stopped (saturation watch); the others go on
        const int cur = (a.cur0 + n) % NBUF;
        fused_tile(sm, a.metrics[cur], a.metrics[(cur + 1) % NBUF], a.ring, a.row_fmt, a.len, nullptr,
                   sl.st, a.T0 + (long long)n * FK, (uint32_t)s_sub * 0x10001u, s_careful != 0, tau * FUSED_COLGROUPS, FUSED_COLGROUPS, 0);
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            atomicAdd(&sl.done[tau % TILE_CLASSES], 1u);
            if (atomicAdd(&sl.done_total, 1u) == FUSED_TILES - 1) resolve_persist(c, n);
        }
    }
}

// ------------------------------------------------------------------------------------------
// single stage (remainders, per-bit streaming, exact saturating fallback)
// ------------------------------------------------------------------------------------------
// One thread = 8 butterflies b0..b0+7 -> 16 new states 2*b0 .. 2*b0+15 (canonical row layout).
template <bool SAT>
__global__ void __launch_bounds__(256) k_acs_single(SingleArgs a)
{
    Ctl *c = a.ctl;
    if (c->T != a.expected_T || c->error) return;
    if (!SAT && (c->maxR + 510 > 32767 || c->spread > MAX_FAST_SPREAD)) return;   // reference could saturate: use the SAT variant
    const long long T0 = c->T;
    const int sub = c->sub;
    const long long O = c->O;
    const uint16_t *oldm = a.metrics[c->cur];
    uint16_t *newm = a.metrics[(c->cur + 1) % NBUF];
    const int s0 = a.use_arg_syms ? a.sym0 : a.syms[2 * (size_t)a.expected_pos];
    const int s1 = a.use_arg_syms ? a.sym1 : a.syms[2 * (size_t)a.expected_pos + 1];

    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t b0 = gid * 8;
    const uint4 va = reinterpret_cast<const uint4 *>(oldm)[gid];
    const uint4 vc = reinterpret_cast<const uint4 *>(oldm)[gid + NBFLY / 8];
    const uint32_t wa[4] = {va.x, va.y, va.z, va.w}, wc[4] = {vc.x, vc.y, vc.z, vc.w};
    uint32_t out[8];
    uint32_t dec = 0;
    int mn = 0x7fffffff, mx = -0x7fffffff;
#pragma unroll
    for (int e = 0; e < 8; e++) {
        const uint32_t b = b0 + e;
        // expected symbols, viterbi224_sse2.c:75-76
        const int e1 = G1FLIP ^ (__popc((2u * b) & POLY1) & 1);
        const int e2 = G2FLIP ^ (__popc((2u * b) & POLY2) & 1);
        const int x = (e1 ? 255 - s0 : s0) + (e2 ? 255 - s1 : s1);     // :292
        const int y = 510 - x;                                        // :293
        int pa = (int)((wa[e >> 1] >> ((e & 1) * 16)) & 0xffff) - sub;
        int pc = (int)((wc[e >> 1] >> ((e & 1) * 16)) & 0xffff) - sub;
        int m0, m1, m2, m3;
        if (SAT) {
            // exact reference arithmetic on R = P + O (int16, saturating adds, :296-299)
            const int ra = (int)(pa + O), rc = (int)(pc + O);
            m0 = min(ra + x, 32767); m1 = min(rc + y, 32767);
            m2 = min(ra + y, 32767); m3 = min(rc + x, 32767);
        } else {
            m0 = pa + x; m1 = pc + y; m2 = pa + y; m3 = pc + x;
        }
        const int d0 = m0 > m1, d1 = m2 > m3;                         // :316-317 strict
        int n0 = min(m0, m1), n1 = min(m2, m3);                       // :319-320
        if (SAT) { n0 += 32768; n1 += 32768; }                        // store P = R + 32768
        dec |= (uint32_t)d0 << (2 * e) | (uint32_t)d1 << (2 * e + 1); // :324
        out[e] = (uint32_t)n0 | ((uint32_t)n1 << 16);                 // :326-327 states 2b, 2b+1
        mn = min(mn, min(n0, n1));
        mx = max(mx, max(n0, n1));
    }
    uint4 *dst = reinterpret_cast<uint4 *>(newm) + (size_t)gid * 2;
    dst[0] = make_uint4(out[0], out[1], out[2], out[3]);
    dst[1] = make_uint4(out[4], out[5], out[6], out[7]);
    const long long row = T0 % a.len;
    // 16 decision bits per thread; pair lanes into 32-bit words
    const uint32_t other = __shfl_xor_sync(0xffffffffu, dec, 1);
    if ((threadIdx.x & 1) == 0)
        __stcs(a.ring + (size_t)row * ROWWORDS + gid / 2, dec | (other << 16));

    const uint32_t wmn = __reduce_min_sync(0xffffffffu, (uint32_t)mn);
    const uint32_t wmx = __reduce_max_sync(0xffffffffu, (uint32_t)mx);
    __shared__ uint32_t s_mn[8], s_mx[8];
    if ((threadIdx.x & 31) == 0) { s_mn[threadIdx.x >> 5] = wmn; s_mx[threadIdx.x >> 5] = wmx; }
    if (gid == 0) { c->st.s0[1] = out[0] & 0xffffu; a.row_fmt[row] = ROWFMT_CANON; }

    __shared__ unsigned s_ticket;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t bmn = s_mn[0], bmx = s_mx[0];
        for (int w = 1; w < 8; w++) { bmn = min(bmn, s_mn[w]); bmx = max(bmx, s_mx[w]); }
        atomicMin(&c->st.minP[1][blockIdx.x % STAT_BUCKETS][0], bmn);
        atomicMax(&c->st.maxP[blockIdx.x % STAT_BUCKETS][0], bmx);
        __threadfence();
        s_ticket = atomicAdd(&c->ticket, 1u);
    }
    __syncthreads();
    if (s_ticket == gridDim.x - 1 && threadIdx.x == 0) {
        __threadfence();
        if (SAT) c->n_sat++; else c->n_single++;
        resolve_pass(c, 1, true, SAT);
    }
}
template __global__ void k_acs_single<false>(SingleArgs);
template __global__ void k_acs_single<true>(SingleArgs);

// ------------------------------------------------------------------------------------------
// traceback
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t read_decision(const uint32_t *ring, const uint8_t *row_fmt, long long row, uint32_t state)
{
    const uint8_t f = row_fmt[row];
    const uint32_t bit = f == ROWFMT_CANON ? state : fused_bit_address(f, state);
    return (ring[(size_t)row * ROWWORDS + (bit >> 5)] >> (bit & 31)) & 1u;     // viterbi224_sse2.c:141
}

// Speculative segment walk.  Segment i covers bits [i*L, min((i+1)*L, nbits)).  Its end state is
// guessed by walking `warm` extra stages back from state 0 (the true endstate for the last one).
__global__ void k_chainback_seg(TraceArgs a, uint32_t nbits, uint32_t endstate, int L, int warm, uint8_t *out,
                                uint32_t *seg_guess, uint32_t *seg_final)
{
    const uint32_t nseg = (nbits + L - 1) / L;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nseg) return;
    const uint32_t lo = i * L, hi = min(nbits, (i + 1) * L);
    uint32_t st;
    if (hi == nbits) {
        st = endstate & STATEMASK;
    } else {
        uint32_t w = min(nbits, hi + (uint32_t)warm);
        st = (w == nbits) ? (endstate & STATEMASK) : 0u;
        while (w > hi) {
            --w;
            const uint32_t bit = read_decision(a.ring, a.row_fmt, w % (uint32_t)a.len, st);
            st = (bit << (K - 2)) | (st >> 1);
        }
    }
    seg_guess[i] = st;
    uint32_t n = hi;
    uint32_t dbyte = 0;
    while (n > lo) {
        --n;
        dbyte = ((st & 1u) << 7) | (dbyte >> 1);                              // :137
        if ((n & 7) == 0) out[n >> 3] = (uint8_t)dbyte;                       // :138-139
        const uint32_t bit = read_decision(a.ring, a.row_fmt, n % (uint32_t)a.len, st);   // :140-141
        st = (bit << (K - 2)) | (st >> 1);                                    // :142
    }
    seg_final[i] = st;
}

// Verify the guesses from the last segment backwards; re-walk a segment serially when its guess
// was wrong.  Result: `out` equals the reference's serial chainback bit for bit.
__global__ void k_chainback_fix(TraceArgs a, uint32_t nbits, int L, uint8_t *out, uint32_t *seg_guess, uint32_t *seg_final,
                                unsigned *redo_count)
{
    const uint32_t nseg = (nbits + L - 1) / L;
    if (nseg < 2) return;
    for (int i = (int)nseg - 2; i >= 0; i--) {
        const uint32_t truth = seg_final[i + 1];
        if (seg_guess[i] == truth) continue;
        atomicAdd(redo_count, 1u);
        const uint32_t lo = (uint32_t)i * L, hi = (uint32_t)(i + 1) * L;
        uint32_t st = truth, n = hi, dbyte = 0;
        while (n > lo) {
            --n;
            dbyte = ((st & 1u) << 7) | (dbyte >> 1);
            if ((n & 7) == 0) out[n >> 3] = (uint8_t)dbyte;
            const uint32_t bit = read_decision(a.ring, a.row_fmt, n % (uint32_t)a.len, st);
            st = (bit << (K - 2)) | (st >> 1);
        }
        seg_guess[i] = truth;
        seg_final[i] = st;
    }
}

// decodebit / decodeword: walk `delay` rows back from ring position dp (viterbi224_sse2.c:164-243).
// result[0] = last bit (or -1), result[1..2] = the 64-bit shift register of decodeword.
__global__ void k_walk(TraceArgs a, long long dp, int delay, uint32_t endstate, int use_argmin, const unsigned long long *argmin_key,
                       unsigned long long *result)
{
    uint32_t st = use_argmin ? (uint32_t)(*argmin_key & 0xffffffffu) : endstate;
    st &= STATEMASK;
    long long row = dp;
    int bit = -1;
    unsigned long long word = 0;
    while (delay-- > 0) {
        if (--row < 0) row = a.len - 1;                                        // :190-191
        bit = (int)read_decision(a.ring, a.row_fmt, row, st);
        st = ((uint32_t)bit << (K - 2)) | (st >> 1);
        word = ((unsigned long long)bit << 63) | (word >> 1);                  // :237
    }
    result[0] = (unsigned long long)(long long)bit;
    result[1] = word;
}

// Batched streaming traceback: output i is what decodebit(delay, 0) returns right after stage
// T_first + i has been appended (vdecode.c:145-152).  Rows older than the last init read as 0.
__global__ void k_stream_trace(TraceArgs a, long long T_first, int nout, int delay, uint8_t *bits_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nout) return;
    long long t = T_first + i + 1;          // stages appended so far; newest row = t-1
    uint32_t st = 0;
    uint32_t bit = 0;
    for (int s = 0; s < delay; s++) {
        --t;
        if (t < 0) { bit = 0; break; }
        bit = read_decision(a.ring, a.row_fmt, t % a.len, st);
        st = (bit << (K - 2)) | (st >> 1);
    }
    bits_out[i] = (uint8_t)bit;
}

// ------------------------------------------------------------------------------------------
// reductions / export
// ------------------------------------------------------------------------------------------
// argmin with lowest index on ties (viterbi224_sse2.c:173-182): key = (P << 32) | index.
__global__ void __launch_bounds__(256) k_argmin(const uint16_t *m, unsigned long long *key)
{
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint4 v = reinterpret_cast<const uint4 *>(m)[gid];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    unsigned long long best = ~0ull;
#pragma unroll
    for (int e = 0; e < 8; e++) {
        const unsigned long long p = (w[e >> 1] >> ((e & 1) * 16)) & 0xffff;
        const unsigned long long k = (p << 32) | (gid * 8 + e);
        best = k < best ? k : best;
    }
    for (int o = 16; o; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other < best ? other : best;
    }
    if ((threadIdx.x & 31) == 0) atomicMin(key, best);
}

__global__ void __launch_bounds__(256) k_minmax(const uint16_t *m, unsigned *mnmx)
{
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint4 v = reinterpret_cast<const uint4 *>(m)[gid];
    uint32_t lo = __vminu2(__vminu2(v.x, v.y), __vminu2(v.z, v.w));
    uint32_t hi = __vmaxu2(__vmaxu2(v.x, v.y), __vmaxu2(v.z, v.w));
    uint32_t mn = min(lo & 0xffff, lo >> 16), mx = max(hi & 0xffff, hi >> 16);
    mn = __reduce_min_sync(0xffffffffu, mn);
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0) { atomicMin(&mnmx[0], mn); atomicMax(&mnmx[1], mx); }
}

// Test hook: one decision row in the reference's canonical layout.
__global__ void __launch_bounds__(256) k_export_row(TraceArgs a, long long row, uint32_t *out)
{
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;     // output word
    const uint8_t f = a.row_fmt[row];
    const uint32_t *r = a.ring + (size_t)row * ROWWORDS;
    if (f == ROWFMT_CANON) { out[w] = r[w]; return; }
    uint32_t v = 0;
    for (int b = 0; b < 32; b++) {
        const uint32_t addr = fused_bit_address(f, w * 32 + b);
        v |= ((r[addr >> 5] >> (addr & 31)) & 1u) << b;
    }
    out[w] = v;
}

// Test hooks: metrics in the reference's int16 domain (R = P - sub + O) and back.
__global__ void __launch_bounds__(256) k_export_metrics(const uint16_t *m, const Ctl *c, int16_t *out, int *range_error)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const long long r = (long long)m[i] - c->sub + c->O;
    if (r < -32768 || r > 32767) *range_error = 1;
    out[i] = (int16_t)r;
}
__global__ void __launch_bounds__(256) k_import_metrics(uint16_t *m, const int16_t *in)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    m[i] = (uint16_t)((int)in[i] + 32768);
}
// after k_import_metrics + k_minmax: make the control block describe the imported state
__global__ void k_import_ctl(Ctl *c, const uint16_t *m, const unsigned *mnmx, long long renormals, long long T)
{
    c->O = -32768;
    c->sub = 0;
    c->renormals = renormals;
    c->T = T;
    c->R0 = (long long)m[0] - 32768;
    c->maxR = (long long)mnmx[1] - 32768;
    c->spread = (long long)mnmx[1] - (long long)mnmx[0];
    c->error = 0;
    reset_stats(c);
}

// ------------------------------------------------------------------------------------------
// launch wrappers (called from the runtime; all asynchronous on `st`)
// ------------------------------------------------------------------------------------------
static bool g_fused_attr_set[64];
cudaError_t launch_init(uint16_t *m0, Ctl *c, uint32_t start_state, int bias, int start_value, cudaStream_t st)
{
    k_init<<<NSTATES / 8 / 256, 256, 0, st>>>(m0, c, start_state, bias, start_value);
    return cudaGetLastError();
}
cudaError_t launch_fused(const FusedArgs &a, cudaStream_t st)
{
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !g_fused_attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_acs_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FusedSmem));
        if (e != cudaSuccess) return e;
        g_fused_attr_set[dev] = true;
    }
    k_build_optab<<<1, OPTAB_WORDS, 0, st>>>(a.optab, a.syms + 2 * (size_t)a.expected_pos, 1);
    k_acs_fused<<<FUSED_TILES, FUSED_THREADS, sizeof(FusedSmem), st>>>(a);
    return cudaGetLastError();
}
// mode: 0 dynamic queue, 1 static uniform tiles, 2 balanced tiles (needs 148 SMs x 4 CTAs); -1 = best available
cudaError_t launch_persist(const PersistArgs &a, int mode, cudaStream_t st)
{
    int dev = 0;
    cudaGetDevice(&dev);
    static int checked[64], slots[64], sms_of[64], per_sm_of[64];
    if (dev >= 0 && dev < 64 && !checked[dev]) {
        const void *fns[3] = {(const void *)k_acs_persist<MODE_DYNAMIC>, (const void *)k_acs_persist<MODE_STATIC>, (const void *)k_acs_persist<MODE_BALANCED>};
        for (const void *f : fns) {
            cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FusedSmem));
            if (e != cudaSuccess) return e;
            // 4 x 33 KiB or 8 x 17 KiB (+1 KiB reserved each) per SM: ask for the large shared-memory carve-out
            cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        }
        int per_sm = 0, sms = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_acs_persist<MODE_BALANCED>, FUSED_THREADS, sizeof(FusedSmem));
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        slots[dev] = per_sm * sms; sms_of[dev] = sms; per_sm_of[dev] = per_sm;
        checked[dev] = 1;
    }
    const bool can_balance = FUSED_TILE_COLS == 64 && sms_of[dev] == BAL_SMS && per_sm_of[dev] == BAL_CTAS_PER_SM;
    // default: the dynamic queue -- measured 16.5 us per pass against 20.6 (static) and 20.3 (balanced) on B200
    // (tools/ab_kernels.py, profiles/); the balanced partition's 3-warp tiles still leave 4 warps on three of
    // the four schedulers, and cooperative launches place CTAs less evenly than the queue does.
    if (mode < 0) mode = MODE_DYNAMIC;
    if (mode == MODE_BALANCED && !can_balance) mode = MODE_DYNAMIC;
    if (mode == MODE_STATIC && slots[dev] < FUSED_TILES) mode = MODE_DYNAMIC;
    k_build_optab<<<a.npasses, OPTAB_WORDS, 0, st>>>(a.optab, a.syms + 2 * (size_t)a.pos0, a.npasses);
    k_persist_begin<<<1, 1, 0, st>>>(a.ctl, a.npasses, a.force_careful, a.T0);
    PersistArgs args = a;
    void *params[] = {&args};
    if (mode == MODE_BALANCED)
        return cudaLaunchCooperativeKernel((const void *)k_acs_persist<MODE_BALANCED>, dim3(BAL_TILES), dim3(FUSED_THREADS), params, sizeof(FusedSmem), st);
    if (mode == MODE_STATIC)
        return cudaLaunchCooperativeKernel((const void *)k_acs_persist<MODE_STATIC>, dim3(FUSED_TILES), dim3(FUSED_THREADS), params, sizeof(FusedSmem), st);
    const long long items = (long long)a.npasses * FUSED_TILES;
    const int grid = (int)(items < slots[dev] ? items : slots[dev]);
    k_acs_persist<MODE_DYNAMIC><<<grid, FUSED_THREADS, sizeof(FusedSmem), st>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_persist_multi(const MultiArgs &m, cudaStream_t st)
{
    int dev = 0;
    cudaGetDevice(&dev);
    static int checked[64], slots[64];
    if (dev >= 0 && dev < 64 && !checked[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_acs_persist_multi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FusedSmem));
        if (e != cudaSuccess) return e;
        cudaFuncSetAttribute(k_acs_persist_multi, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        int per_sm = 0, sms = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_acs_persist_multi, FUSED_THREADS, sizeof(FusedSmem));
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        slots[dev] = per_sm * sms;
        checked[dev] = 1;
    }
    for (int s = 0; s < m.nctx; s++) {
        const PersistArgs &a = m.ctx[s];
        k_build_optab<<<m.npasses, OPTAB_WORDS, 0, st>>>(a.optab, a.syms + 2 * (size_t)a.pos0, m.npasses);
        k_persist_begin<<<1, 1, 0, st>>>(a.ctl, m.npasses, a.force_careful, a.T0);
    }
    const long long items = (long long)m.npasses * m.nctx * FUSED_TILES;
    const int grid = (int)(items < slots[dev] ? items : slots[dev]);
    k_acs_persist_multi<<<grid, FUSED_THREADS, sizeof(FusedSmem), st>>>(m);
    return cudaGetLastError();
}
cudaError_t launch_single(const SingleArgs &a, bool sat, cudaStream_t st)
{
    if (sat) k_acs_single<true><<<NBFLY / 8 / 256, 256, 0, st>>>(a);
    else     k_acs_single<false><<<NBFLY / 8 / 256, 256, 0, st>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_chainback(const TraceArgs &a, uint32_t nbits, uint32_t endstate, int L, int warm, uint8_t *out, uint32_t *seg_guess,
                             uint32_t *seg_final, unsigned *redo_count, cudaStream_t st)
{
    const uint32_t nseg = (nbits + L - 1) / L;
    k_chainback_seg<<<(nseg + 31) / 32, 32, 0, st>>>(a, nbits, endstate, L, warm, out, seg_guess, seg_final);
    k_chainback_fix<<<1, 1, 0, st>>>(a, nbits, L, out, seg_guess, seg_final, redo_count);
    return cudaGetLastError();
}
cudaError_t launch_walk(const TraceArgs &a, long long dp, int delay, uint32_t endstate, int use_argmin, const unsigned long long *argmin_key,
                        unsigned long long *result, cudaStream_t st)
{
    k_walk<<<1, 1, 0, st>>>(a, dp, delay, endstate, use_argmin, argmin_key, result);
    return cudaGetLastError();
}
cudaError_t launch_stream_trace(const TraceArgs &a, long long T_first, int nout, int delay, uint8_t *bits_out, cudaStream_t st)
{
    if (nout <= 0) return cudaSuccess;
    k_stream_trace<<<(nout + 63) / 64, 64, 0, st>>>(a, T_first, nout, delay, bits_out);
    return cudaGetLastError();
}
cudaError_t launch_argmin(const uint16_t *m, unsigned long long *key, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(key, 0xff, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    k_argmin<<<NSTATES / 8 / 256, 256, 0, st>>>(m, key);
    return cudaGetLastError();
}
cudaError_t launch_minmax(const uint16_t *m, unsigned *mnmx, cudaStream_t st)
{
    const unsigned init[2] = {0xffffffffu, 0u};
    cudaError_t e = cudaMemcpyAsync(mnmx, init, sizeof init, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return e;
    k_minmax<<<NSTATES / 8 / 256, 256, 0, st>>>(m, mnmx);
    return cudaGetLastError();
}
cudaError_t launch_export_row(const TraceArgs &a, long long row, uint32_t *out, cudaStream_t st)
{
    k_export_row<<<ROWWORDS / 256, 256, 0, st>>>(a, row, out);
    return cudaGetLastError();
}
cudaError_t launch_export_metrics(const uint16_t *m, const Ctl *c, int16_t *out, int *range_error, cudaStream_t st)
{
    k_export_metrics<<<NSTATES / 256, 256, 0, st>>>(m, c, out, range_error);
    return cudaGetLastError();
}
cudaError_t launch_import_metrics(uint16_t *m, const int16_t *in, Ctl *c, unsigned *mnmx, long long renormals, long long T, cudaStream_t st)
{
    k_import_metrics<<<NSTATES / 256, 256, 0, st>>>(m, in);
    cudaError_t e = launch_minmax(m, mnmx, st);
    if (e != cudaSuccess) return e;
    k_import_ctl<<<1, 1, 0, st>>>(c, m, mnmx, renormals, T);
    return cudaGetLastError();
}

} // namespace v224

#ifdef V224_TRACE
extern "C" int v224_debug_read_trace(unsigned long long *host, unsigned long long n)
{
    return (int)cudaMemcpyFromSymbol(host, v224::g_trace, n * sizeof(unsigned long long));
}
extern "C" int v224_debug_read_smid(unsigned *host, unsigned long long n)
{
    return (int)cudaMemcpyFromSymbol(host, v224::g_smid, n * sizeof(unsigned));
}
#endif
