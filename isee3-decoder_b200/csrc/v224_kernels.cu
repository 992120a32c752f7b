// v224_kernels.cu -- sm_100a kernels of the B200 viterbi224 decoder.
//
//   k_init            metric fill + control-block reset               (viterbi224_sse2.c:37-53)
//   (the fused 8-stage pass, k_acs_persist, lives in v224_acs_persist.cu)
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
#include "v224_kernels.h"
#include <cstdio>
#include <cstdlib>

namespace v224 {

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
// Reset the per-pass statistics (resolver, and k_init).
__device__ __forceinline__ void reset_stats(Ctl *c)
{
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
        c->spec_first = 0;
        reset_stats(c);
    }
}

__device__ __forceinline__ uint32_t read_decision(const uint32_t *ring, const uint8_t *row_fmt, long long row, uint32_t state)
{
    const uint8_t f = row_fmt[row];
    const uint32_t bit = f == ROWFMT_CANON ? state : fused_bit_address(f, state);
    const uint32_t flip = f == ROWFMT_CANON ? 0u : FUSED_ROWS_COMPLEMENTED;
    return ((ring[(size_t)row * ROWWORDS + (bit >> 5)] >> (bit & 31)) & 1u) ^ flip;     // viterbi224_sse2.c:141
}

// Incremental decodebit walk (see k_walk_incremental below): `delay` rows back from ring head T, stopping where the path
// rejoins the previous call's (cached) path.  Returns the bit decodebit(delay, endstate) returns.
// first_bit >= 0: the decision of (row T-1, endstate) is known to the caller (the one-stage kernel walks while the rest of
// its grid is still writing that row) and is not read from the ring.
// all_canon: the host knows that every row in the ring is in the canonical layout (no fused pass has written into it):
// the row tags are not read, which halves the chain of dependent loads.
__device__ int incremental_walk(const uint32_t *ring, const uint8_t *row_fmt, int len, long long T, long long prev_T, int delay, uint32_t endstate,
                                uint32_t *cache, unsigned *steps_out, int first_bit = -1, bool all_canon = false)
{
    uint32_t st = endstate & STATEMASK;
    int bit = -1;
    unsigned steps = 0;
    for (long long t = T - 1; t >= T - delay; t--) {
        const long long row = ((t % len) + len) % len;
        if (t == T - 1 && first_bit >= 0) bit = first_bit;
        else if (all_canon) bit = (int)((__ldcg(ring + (size_t)row * ROWWORDS + (st >> 5)) >> (st & 31)) & 1u);      // viterbi224_sse2.c:141
        else bit = (int)read_decision(ring, row_fmt, row, st);
        st = ((uint32_t)bit << (K - 2)) | (st >> 1);
        steps++;
        const bool cached = t <= prev_T - 1 && t >= prev_T - delay && t > T - delay;
        if (cached && cache[row] == st) {
            // merged with the previous path: the state at time T - delay is the cached one
            const long long r2 = (((T - delay) % len) + len) % len;
            bit = (int)((cache[r2] >> (K - 2)) & 1u);
            break;
        }
        cache[row] = st;
    }
    if (steps_out) atomicAdd(steps_out, steps);
    return bit;
}

// ------------------------------------------------------------------------------------------
// single stage (remainders, per-bit streaming, exact saturating fallback)
// ------------------------------------------------------------------------------------------
// One thread = 8 butterflies b0..b0+7 -> 16 new states 2*b0 .. 2*b0+15 (canonical row layout).
template <bool SAT>
__global__ void __launch_bounds__(256) k_acs_single(SingleArgs a)
{
    Ctl *c = a.ctl;
    if (c->T != a.expected_T || c->error || (!SAT && (c->maxR + 510 > 32767 || c->spread > MAX_FAST_SPREAD))) {
        // declined: stale stage counter, sticky error, or the reference could saturate (the host then uses the SAT variant)
        if (a.mailbox && blockIdx.x == 0 && threadIdx.x == 0) {
            a.mailbox->declined = 1;
            __threadfence_system();
            a.mailbox->seq = a.seq;
        }
        return;
    }
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

    // Per-bit streaming: the decodebit walk the caller is about to ask for (vdecode.c:152) starts at new state spec_end of
    // THIS row and from there on only reads older rows: the thread that computed that decision walks now, while the rest
    // of the grid is still working, and leaves the answer for the resolver to post.
    if (a.mailbox && a.spec_walk && (a.spec_end >> 4) == gid) {
        const int first = (int)((dec >> (a.spec_end & 15u)) & 1u);
        a.mailbox->walk_bit = incremental_walk(a.ring, a.row_fmt, a.len, T0 + 1, a.spec_prev_T, a.spec_delay, a.spec_end, a.walk_cache, a.walk_steps, first);
        __threadfence_system();
    }
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
        if (a.mailbox) {
            // per-bit streaming: everything the host needs after this stage, in mapped host memory, sequence number last
            // (the speculative walk's answer is already there: its thread wrote it before its block took a ticket)
            if (!a.spec_walk) a.mailbox->walk_bit = -2;
            const unsigned *src = reinterpret_cast<const unsigned *>(c);
            unsigned *dst = reinterpret_cast<unsigned *>(a.mailbox->ctl_head);
            for (unsigned i = 0; i < CTL_HOST_BYTES / 4; i++) dst[i] = src[i];
            a.mailbox->declined = 0;
            __threadfence_system();
            a.mailbox->seq = a.seq;
        }
    }
}
template __global__ void k_acs_single<false>(SingleArgs);
template __global__ void k_acs_single<true>(SingleArgs);

// ------------------------------------------------------------------------------------------
// single stage, fast form: packed 2 x uint16 arithmetic, one wave, every SM the same number of bytes
// ------------------------------------------------------------------------------------------
// Same stage as k_acs_single<false> (viterbi224_sse2.c:277-328), restructured for latency: this is the kernel behind the
// per-bit ABI pattern (vdecode.c:145: update(1) per decoded bit), where one launch is all there is to hide latency in.
//   * unit = 8 butterflies b0 .. b0+7 (two 16-byte loads, one 32-byte metric store, 16 decision bits); a block owns a
//     contiguous range of units -- NUNITS / gridDim.x of them, the same for every SM -- and its threads go through it two
//     units at a time with all four loads in flight before the first add
//   * a register holds the metrics of two neighbouring butterflies; both halves run the butterfly with their own branch
//     metric.  The expected symbols of butterfly b0 + e are parity(16 u & POLY) ^ parity(2 e & POLY) (:75-76; the two
//     polynomials differ in register bit 1 only, which 16 u does not have): one population count per unit decides whether
//     the unit uses the kernel's four packed branch metrics or their complements 510 - x (:293)
//   * the block takes ONE ticket; 2 * SMs blocks instead of 2048
// The metrics stay below 2^16 in every half (spread <= MAX_FAST_SPREAD, checked on entry), so packed adds cannot carry.
constexpr uint32_t NUNITS = NBFLY / 8;            // 2^19
constexpr int SINGLE_FAST_THREADS = 512;
#ifndef V224_SINGLE_DEPTH
#define V224_SINGLE_DEPTH 2
#endif
constexpr int SINGLE_FAST_DEPTH = V224_SINGLE_DEPTH;   // units a thread has in flight (all their loads issued before the first add)

struct FastUnit { uint4 va, vc; uint32_t u; };

__device__ __forceinline__ void fast_unit(const FastUnit &f, uint32_t sub2, const uint32_t (&X0)[4], uint16_t *newm, uint32_t *ring_row,
                                          uint32_t &mnp, uint32_t &mxp, uint32_t &dec_out, uint32_t &n00)
{
    const uint32_t wa[4] = {f.va.x, f.va.y, f.va.z, f.va.w}, wc[4] = {f.vc.x, f.vc.y, f.vc.z, f.vc.w};
    // expected symbols of the unit's first butterfly, linear part (flips are folded into X0)
    const bool t = (__popc((16u * f.u) & POLY1) & 1u) != 0;
    uint32_t out[8];
    uint32_t dec = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t x0 = X0[i], y0 = 0x01fe01feu - x0;
        const uint32_t X = t ? y0 : x0, Y = t ? x0 : y0;                  // :292-293
        const uint32_t pa = wa[i] - sub2, pc = wc[i] - sub2;
        const uint32_t m0 = pa + X, m1 = pc + Y, m2 = pa + Y, m3 = pc + X;  // :296-299
        const uint32_t g0 = __vcmpgtu2(m0, m1), g1 = __vcmpgtu2(m2, m3);    // :316-317 strict
        const uint32_t n0 = __vminu2(m0, m1), n1 = __vminu2(m2, m3);        // :319-320
        // decision bits of butterflies 2i (low halves) and 2i+1 (high halves): d0 at bit 2e, d1 at bit 2e+1 (:324)
        const uint32_t tt = (g0 & 0x00010001u) | ((g1 & 0x00010001u) << 1);
        dec |= ((tt | (tt >> 14)) & 0xfu) << (4 * i);
        out[2 * i] = __byte_perm(n0, n1, 0x5410);                            // states 2b, 2b+1 of butterfly 2i (:326-327)
        out[2 * i + 1] = __byte_perm(n0, n1, 0x7632);                        // ... of butterfly 2i+1
        mnp = __vminu2(mnp, __vminu2(n0, n1));
        mxp = __vmaxu2(mxp, __vmaxu2(n0, n1));
        if (i == 0) n00 = n0 & 0xffffu;
    }
    uint4 *dst = reinterpret_cast<uint4 *>(newm) + (size_t)f.u * 2;
    dst[0] = make_uint4(out[0], out[1], out[2], out[3]);
    dst[1] = make_uint4(out[4], out[5], out[6], out[7]);
    // 16 decision bits per unit; neighbouring lanes hold neighbouring units: pair them into 32-bit words
    const uint32_t other = __shfl_xor_sync(0xffffffffu, dec, 1);
    if ((f.u & 1u) == 0) __stcs(ring_row + f.u / 2, dec | (other << 16));
    dec_out = dec;
}

// Run by the thread that took the stage's last ticket: resolve the stage and, in per-bit streaming, post the report.
__device__ void finish_single_fast(Ctl *c, const SingleArgs &a)
{
    __threadfence();
    c->n_single++;
    resolve_pass(c, 1, true, false);
    if (a.mailbox) {
        // (a speculative walk's answer is already there: the walker wrote it before it took its ticket)
        if (!a.spec_walk) a.mailbox->walk_bit = -2;
        const unsigned *src = reinterpret_cast<const unsigned *>(c);
        unsigned *dst = reinterpret_cast<unsigned *>(a.mailbox->ctl_head);
        for (unsigned i = 0; i < CTL_HOST_BYTES / 4; i++) dst[i] = src[i];
        a.mailbox->declined = 0;
        __threadfence_system();
        a.mailbox->seq = a.seq;
    }
}

__global__ void __launch_bounds__(SINGLE_FAST_THREADS, 2) k_acs_single_fast(SingleArgs a)
{
    Ctl *c = a.ctl;
    if (c->T != a.expected_T || c->error || c->maxR + 510 > 32767 || c->spread > MAX_FAST_SPREAD) {
        // declined: stale stage counter, sticky error, or the reference could saturate (the host then uses the SAT variant)
        if (a.mailbox && blockIdx.x == 0 && threadIdx.x == 0) {
            a.mailbox->declined = 1;
            __threadfence_system();
            a.mailbox->seq = a.seq;
        }
        return;
    }
    const long long T0 = c->T;
    if (blockIdx.x == 0 && a.spec_walk) {
        // Walker block (per-bit streaming; block 0, so that it is resident from the start): the decodebit walk the caller is about to ask for (vdecode.c:152) starts at new
        // state spec_end of THIS row and from there on only reads older rows.  It waits for that one decision (the thread
        // that computes it publishes it, tagged with the stage) and walks while the other blocks are still working.
        if (threadIdx.x == 0) {
            const unsigned want = (unsigned)(T0 + 1);
            unsigned v = 0;
            for (unsigned spins = 0; spins < 50u * 1000u * 1000u; spins++) {
                v = *(volatile unsigned *)&c->spec_first;
                if ((v >> 1) == (want & 0x7fffffffu)) break;
            }
            long long bit = -3;                                   // the decision never came: the host falls back to its own walk
            if ((v >> 1) == (want & 0x7fffffffu))
                bit = incremental_walk(a.ring, a.row_fmt, a.len, T0 + 1, a.spec_prev_T, a.spec_delay, a.spec_end, a.walk_cache, a.walk_steps,
                                       (int)(v & 1u), a.all_canon != 0);
            a.mailbox->walk_bit = bit;
            __threadfence_system();
            if (atomicAdd(&c->ticket, 1u) == gridDim.x - 1) finish_single_fast(c, a);     // the walk outlasted the stage
        }
        return;
    }
    const uint32_t nwork = gridDim.x - (a.spec_walk ? 1u : 0u);    // blocks that share the stage's units
    const uint32_t wblock = blockIdx.x - (a.spec_walk ? 1u : 0u);  // this block's index among them
    const uint32_t sub2 = (uint32_t)c->sub * 0x10001u;
    const uint16_t *oldm = a.metrics[c->cur];
    uint16_t *newm = a.metrics[(c->cur + 1) % NBUF];
    const int s0 = a.use_arg_syms ? a.sym0 : a.syms[2 * (size_t)a.expected_pos];
    const int s1 = a.use_arg_syms ? a.sym1 : a.syms[2 * (size_t)a.expected_pos + 1];
    const long long row = T0 % a.len;
    uint32_t *ring_row = a.ring + (size_t)row * ROWWORDS;
    // the four packed branch metrics of a unit whose linear label part is 0: butterflies (2i, 2i+1) in the halves
    uint32_t X0[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint32_t x[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint32_t e = 2 * i + h;
            const int e1 = G1FLIP ^ (__popc((2u * e) & POLY1) & 1), e2 = G2FLIP ^ (__popc((2u * e) & POLY2) & 1);   // :75-76
            x[h] = (uint32_t)((e1 ? 255 - s0 : s0) + (e2 ? 255 - s1 : s1));
        }
        X0[i] = x[0] | (x[1] << 16);
    }
    // this block's units: warp-aligned, the same count (+-32) for every block
    const uint32_t wu0 = (uint32_t)(((unsigned long long)wblock * (NUNITS / 32)) / nwork) * 32u;
    const uint32_t wu1 = (uint32_t)(((unsigned long long)(wblock + 1) * (NUNITS / 32)) / nwork) * 32u;
    uint32_t mnp = 0xffffffffu, mxp = 0;
    const uint4 *old4 = reinterpret_cast<const uint4 *>(oldm);
    for (uint32_t base = wu0; base < wu1; base += SINGLE_FAST_DEPTH * SINGLE_FAST_THREADS) {
        FastUnit f[SINGLE_FAST_DEPTH];
        bool on[SINGLE_FAST_DEPTH];
#pragma unroll
        for (int k = 0; k < SINGLE_FAST_DEPTH; k++) {
            f[k].u = base + k * SINGLE_FAST_THREADS + threadIdx.x;
            on[k] = f[k].u < wu1;                       // whole warps switch on and off together (the range is warp-aligned)
            if (on[k]) { f[k].va = old4[f[k].u]; f[k].vc = old4[f[k].u + NUNITS]; }
        }
#pragma unroll
        for (int k = 0; k < SINGLE_FAST_DEPTH; k++) {
            if (!on[k]) continue;
            uint32_t dec, n00;
            fast_unit(f[k], sub2, X0, newm, ring_row, mnp, mxp, dec, n00);
            if (f[k].u == 0) { c->st.s0[1] = n00; a.row_fmt[row] = ROWFMT_CANON; }
            // per-bit streaming: the walk's first decision (new state spec_end of THIS row), for the walker block
            if (a.spec_walk && (a.spec_end >> 4) == f[k].u)
                *(volatile unsigned *)&c->spec_first = ((unsigned)(T0 + 1) << 1) | ((dec >> (a.spec_end & 15u)) & 1u);
        }
    }
    uint32_t mn = min(mnp & 0xffffu, mnp >> 16), mx = max(mxp & 0xffffu, mxp >> 16);
    mn = __reduce_min_sync(0xffffffffu, mn);
    mx = __reduce_max_sync(0xffffffffu, mx);
    __shared__ uint32_t s_mn[SINGLE_FAST_THREADS / 32], s_mx[SINGLE_FAST_THREADS / 32];
    if ((threadIdx.x & 31) == 0) { s_mn[threadIdx.x >> 5] = mn; s_mx[threadIdx.x >> 5] = mx; }
    __shared__ unsigned s_ticket;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t bmn = s_mn[0], bmx = s_mx[0];
        for (int w = 1; w < SINGLE_FAST_THREADS / 32; w++) { bmn = min(bmn, s_mn[w]); bmx = max(bmx, s_mx[w]); }
        atomicMin(&c->st.minP[1][blockIdx.x % STAT_BUCKETS][0], bmn);
        atomicMax(&c->st.maxP[blockIdx.x % STAT_BUCKETS][0], bmx);
        __threadfence();
        s_ticket = atomicAdd(&c->ticket, 1u);
    }
    __syncthreads();
    if (s_ticket == gridDim.x - 1 && threadIdx.x == 0) finish_single_fast(c, a);
}

// ------------------------------------------------------------------------------------------
// traceback
// ------------------------------------------------------------------------------------------
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

// Incremental form of the same walk for the per-bit streaming pattern of vdecode.c:145-152 (update(1) + decodebit(delay, 0)
// after every stage): survivors merge, so the walk from the new ring head rejoins the previous call's path after a
// few steps; from there on the states are the ones cached last time and the answer is looked up.  cache[t % len] = state
// of the path at time t (after the row of stage t was applied) for the previous walk, which started at head prev_T.
// T = stages appended so far; rows T-delay .. T-1 must still be in the ring (len > delay) and unchanged since.
// The result is identical to k_walk's; only the number of dependent loads differs.
__global__ void k_walk_incremental(TraceArgs a, long long T, long long prev_T, int delay, uint32_t endstate, uint32_t *cache,
                                   unsigned long long *result, unsigned *steps_out)
{
    const int bit = incremental_walk(a.ring, a.row_fmt, a.len, T, prev_T, delay, endstate, cache, steps_out);
    result[0] = (unsigned long long)(long long)bit;
    result[1] = 0;
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

// The same for up to MAX_CTX decoders in ONE launch.  A walk is a chain of `delay` dependent loads (latency bound, ~0.7 us per
// step): the lockstep decoders' tracebacks of one chunk of stages are independent chains, so they share one launch instead of
// running one after the other (4 launches of 140 us each at delay 200 -> one).
__global__ void k_stream_trace_multi(StreamTraceMulti m)
{
    const StreamTraceJob &j = m.job[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= j.nout) return;
    long long t = j.T_first + i + 1;
    uint32_t st = 0;
    uint32_t bit = 0;
    for (int s = 0; s < m.delay; s++) {
        --t;
        if (t < 0) { bit = 0; break; }
        bit = read_decision(j.a.ring, j.a.row_fmt, t % j.a.len, st);
        st = (bit << (K - 2)) | (st >> 1);
    }
    j.bits_out[i] = (uint8_t)bit;
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

// Hand-over check of the segmented stream decode: two decoders make identical decisions from here on iff their
// path-metric vectors differ by a constant.  out[0] = min, out[1] = max of a[s] - b[s] over all states.
__global__ void __launch_bounds__(256) k_metric_diff(const uint16_t *a, const uint16_t *b, int *out)
{
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint4 va = __ldcg(reinterpret_cast<const uint4 *>(a) + gid), vb = __ldcg(reinterpret_cast<const uint4 *>(b) + gid);
    const uint32_t wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
    int mn = 0x7fffffff, mx = -0x7fffffff;
#pragma unroll
    for (int e = 0; e < 8; e++) {
        const int d = (int)((wa[e >> 1] >> ((e & 1) * 16)) & 0xffff) - (int)((wb[e >> 1] >> ((e & 1) * 16)) & 0xffff);
        mn = min(mn, d);
        mx = max(mx, d);
    }
    mn = __reduce_min_sync(0xffffffffu, mn);
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0) { atomicMin(&out[0], mn); atomicMax(&out[1], mx); }
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
        v |= (((r[addr >> 5] >> (addr & 31)) & 1u) ^ FUSED_ROWS_COMPLEMENTED) << b;
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
    c->spec_first = 0;
    reset_stats(c);
}

// ------------------------------------------------------------------------------------------
// launch wrappers (called from the runtime; all asynchronous on `st`)
// ------------------------------------------------------------------------------------------
cudaError_t launch_init(uint16_t *m0, Ctl *c, uint32_t start_state, int bias, int start_value, cudaStream_t st)
{
    k_init<<<NSTATES / 8 / 256, 256, 0, st>>>(m0, c, start_state, bias, start_value);
    return cudaGetLastError();
}
cudaError_t launch_single(const SingleArgs &a, bool sat, cudaStream_t st)
{
    if (sat) { k_acs_single<true><<<NBFLY / 8 / 256, 256, 0, st>>>(a); return cudaGetLastError(); }
    if (a.slow_form) { k_acs_single<false><<<NBFLY / 8 / 256, 256, 0, st>>>(a); return cudaGetLastError(); }
    // two blocks of 512 threads per SM, all resident at once (with a speculative walk block 0 walks and the others share the units)
    static int sms_of[64];
    int dev = 0;
    cudaGetDevice(&dev);
    int sms = (dev >= 0 && dev < 64) ? sms_of[dev] : 0;
    if (sms == 0) {
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
        if (dev >= 0 && dev < 64) sms_of[dev] = sms;                 // (a benign race: every thread stores the same value)
    }
    k_acs_single_fast<<<2 * sms, SINGLE_FAST_THREADS, 0, st>>>(a);     // all resident at once, the walker block (if any) included
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
cudaError_t launch_walk_incremental(const TraceArgs &a, long long T, long long prev_T, int delay, uint32_t endstate, uint32_t *cache,
                                    unsigned long long *result, unsigned *steps_out, cudaStream_t st)
{
    k_walk_incremental<<<1, 1, 0, st>>>(a, T, prev_T, delay, endstate, cache, result, steps_out);
    return cudaGetLastError();
}
cudaError_t launch_stream_trace(const TraceArgs &a, long long T_first, int nout, int delay, uint8_t *bits_out, cudaStream_t st)
{
    if (nout <= 0) return cudaSuccess;
    k_stream_trace<<<(nout + 63) / 64, 64, 0, st>>>(a, T_first, nout, delay, bits_out);
    return cudaGetLastError();
}
cudaError_t launch_stream_trace_multi(const StreamTraceMulti &m, cudaStream_t st)
{
    int most = 0;
    for (int k = 0; k < m.njobs; k++) most = m.job[k].nout > most ? m.job[k].nout : most;
    if (most <= 0 || m.njobs <= 0) return cudaSuccess;
    k_stream_trace_multi<<<dim3((most + 63) / 64, m.njobs), 64, 0, st>>>(m);
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
cudaError_t launch_metric_diff(const uint16_t *a, const uint16_t *b, int *out2, cudaStream_t st)
{
    const int init[2] = {0x7fffffff, -0x7fffffff};
    cudaError_t e = cudaMemcpyAsync(out2, init, sizeof init, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return e;
    k_metric_diff<<<NSTATES / 8 / 256, 256, 0, st>>>(a, b, out2);
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

