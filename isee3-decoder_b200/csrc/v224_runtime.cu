// v224_runtime.cu -- host runtime and C ABI of libviterbi224_b200.
//
// Exports exactly the nine symbols of the reference's viterbi224.h:8-16 (see
// include/viterbi224.h) plus the v224x_* extensions (include/viterbi224_b200.h).
// Everything that touches a metric or a decision bit runs in the CUDA kernels of
// v224_kernels.cu; there is no CPU arithmetic path in this file.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdarg>
#include <mutex>
#include <vector>
#include <string>
#include <thread>
#include <atomic>
#include <algorithm>
#include "v224_kernels.h"
#include "../../include/viterbi224.h"
#include "../../include/viterbi224_b200.h"

using namespace v224;

// the fused pass compiled for 32-column tiles (v224_acs_persist.cu with -DV224_TILE_COLS_LOG2=5, namespace v224t32)
extern "C" cudaError_t v224_t32_launch_persist(const void *multi_args, cudaStream_t st);
extern "C" cudaError_t v224_t32_build_metric_tensor_maps(uint16_t *const *metrics, void *dev_out, cudaStream_t st, const char **why);
#ifdef V224_WITH_Q1
// A/B builds only: 32-column tiles with one packed register per row and thread (-DV224_NQ=1, namespace v224t32q1); same tensor maps
extern "C" cudaError_t v224_t32q1_launch_persist(const void *multi_args, cudaStream_t st);
#endif

namespace {

thread_local char g_err[512] = "";
thread_local int g_device = -1;      // -1: leave the CUDA current device alone (device 0 by default)

void set_err(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    if (getenv("V224_DEBUG")) fprintf(stderr, "[viterbi224_b200] %s\n", g_err);
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            set_err("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);   \
            return -1;                                                                             \
        }                                                                                          \
    } while (0)

// ---- pooled ring allocations: callers such as decode.c:216-229 create and delete a 1 GiB
// decoder per frame; keep the last few rings around instead of going back to the driver. ----
struct PoolEntry { int dev; size_t bytes; void *ptr; };
std::mutex g_pool_mu;
std::vector<PoolEntry> g_pool;
constexpr size_t POOL_MAX_ENTRIES = 4;
constexpr size_t POOL_MAX_BYTES = (size_t)24 << 30;      // cached rings never hold more than this (oldest are released first)

void release_parked_decoders();      // defined below the Decoder type

void *pool_get(int dev, size_t bytes)
{
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        for (size_t i = 0; i < g_pool.size(); i++)
            if (g_pool[i].dev == dev && g_pool[i].bytes == bytes) {
                void *p = g_pool[i].ptr;
                g_pool.erase(g_pool.begin() + i);
                return p;
            }
    }
    void *p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
        // release everything that is only parked -- recycled decoders (each holds a ring) and cached rings -- and retry once
        cudaGetLastError();
        release_parked_decoders();
        std::vector<PoolEntry> drop;
        {
            std::lock_guard<std::mutex> lk(g_pool_mu);
            drop.swap(g_pool);
        }
        for (auto &e : drop) { cudaSetDevice(e.dev); cudaFree(e.ptr); }
        cudaSetDevice(dev);
        cudaGetLastError();
        if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    }
    return p;
}
void pool_put(int dev, size_t bytes, void *ptr)
{
    std::vector<PoolEntry> evict;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        g_pool.push_back({dev, bytes, ptr});
        size_t total = 0;
        for (auto &e : g_pool) total += e.bytes;
        while (!g_pool.empty() && (g_pool.size() > POOL_MAX_ENTRIES || total > POOL_MAX_BYTES)) {
            total -= g_pool.front().bytes;
            evict.push_back(g_pool.front());
            g_pool.erase(g_pool.begin());
        }
    }
    for (auto &e : evict) { cudaSetDevice(e.dev); cudaFree(e.ptr); }
    if (!evict.empty()) cudaSetDevice(dev);
}

struct Decoder {
    uint32_t magic;
    int dev;
    int len;
    cudaStream_t stream;
    uint16_t *metrics[NBUF];
    uint32_t *ring;
    size_t ring_bytes;
    uint8_t *row_fmt;
    Ctl *ctl;
    Ctl *h_ctl;                 // pinned mirror, refreshed at the end of every update
    uint8_t *dsyms; size_t dsyms_cap;
    void *tmaps;                            // NBUF tensor maps of the metric buffers (device memory), box = one 64-column tile
    void *tmaps32;                          // the same with box = one 32-column tile (the fused pass of a decoder running alone)
    uint32_t *optab; size_t optab_cap;     // per-pass tables (operands + ring rows) of the running batch
    uint8_t *dout;  size_t dout_cap;       // chainback / stream output staging
    uint32_t *seg;  size_t seg_cap;        // chainback segment bookkeeping
    unsigned *d_redo;
    unsigned long long *d_key;             // argmin key
    unsigned *d_mnmx;
    unsigned long long *d_result, *h_result;   // walk results (pinned host copy)
    int *d_flag;
    // per-bit streaming (vdecode.c:145-152): the previous decodebit walk, so that the next one stops where it rejoins it
    uint32_t *walk_cache;      // [len] state of the cached path at stage t, at index t % len
    long long cache_T;         // ring head (stages since init) of the cached walk
    int cache_delay, cache_valid;
    uint32_t cache_end;
    unsigned *d_walk_steps;    // dependent loads spent in decodebit walks (test / tuning counter)
    // per-bit streaming fast path: the one-stage kernel reports into mapped host memory and carries the decodebit walk
    Mailbox *mb, *mb_dev;      // pinned mapped host memory / its device address
    unsigned long long mb_seq; // sequence number of the last launch that reports through the mailbox
    int mb_pending;            // that launch has not been waited for yet (h_ctl is stale until it has)
    int spec_want, spec_delay; // the caller's decodebit pattern (set by decodebit): speculate on it in the next update(1)
    uint32_t spec_end;
    int spec_valid;            // the last stage carried the walk for (spec_delay, spec_end) at stage counter spec_T
    long long spec_T, spec_bit;
    int no_mailbox;            // option: always take the synchronous path
    int no_discard;            // option: fused passes never drop their consumed input lines from the L2
    int measure_all;           // option: every fused pass reduces the min / max of its output
    int slow_single;           // option: the scalar form of the one-stage kernel
    int fused_rows;            // a fused pass has written rows into the ring since it was last cleared (their layout differs)
    // segmented stream decode: auxiliary decoders (owned, cached), snapshot of this decoder's metrics at its hand-over point
    struct Decoder *aux[2 * MAX_CTX - 1];   // [0, MAX_CTX-1): lockstep partners of this decoder; [MAX_CTX-1, ..): second lane set of the frame batches
    uint16_t *snap;
    int *d_segdiff;            // [2 * MAX_CTX]: min / max of the metric difference at each hand-over check
    cudaEvent_t ev0, ev1, kev0, kev1;
    // options
    int force_single, force_sat, force_careful, per_pass_launch, chain_seg, chain_warm, no_walk_cache, grid_limit, tile32;
    // counters
    unsigned long long launches, acs_launches_timed, acs_passes_timed, chainback_redo;
    long long stages_total;                // trellis stages ever run on this ring (how many rows a recycled decoder has to clear)
    int ring_dirty_all;                    // v224x_set_state moved the stage counter: rows were not written from row 0 upwards
    double acs_ms;
    int time_kernels;
};
constexpr uint32_t MAGIC = 0x56323234u;   // "V224"
constexpr int TILE32_DEFAULT = 1;         // which build of the fused pass a decoder running alone uses (option "tile32")
constexpr int NAUX = 2 * MAX_CTX - 1;

int mailbox_wait(struct Decoder *d);

Decoder *as_dec(void *p)
{
    Decoder *d = static_cast<Decoder *>(p);
    if (!d) return nullptr;
    if (d->magic != MAGIC) { set_err("not a viterbi224_b200 handle"); return nullptr; }
    if (d->mb_pending && mailbox_wait(d)) return nullptr;      // an asynchronous per-bit stage is still in flight: h_ctl is stale until it reports
    return d;
}

int bind(Decoder *d)
{
    CU(cudaSetDevice(d->dev));
    return 0;
}

int grow(void **buf, size_t *cap, size_t need)
{
    if (*cap >= need) return 0;
    if (*buf) cudaFree(*buf);
    *buf = nullptr; *cap = 0;
    size_t n = std::max(need, (size_t)4096);
    CU(cudaMalloc(buf, n));
    *cap = n;
    return 0;
}

int sync_ctl(Decoder *d)
{
    CU(cudaMemcpyAsync(d->h_ctl, d->ctl, CTL_HOST_BYTES, cudaMemcpyDeviceToHost, d->stream));
    CU(cudaStreamSynchronize(d->stream));
    if (d->h_ctl->error) { set_err("device control block reports invariant violation %d", d->h_ctl->error); return -1; }
    return 0;
}

// Wait for the one-stage launch that reports through the mailbox (polling mapped host memory: no copy, no stream
// synchronisation), then refresh the host's mirror of the control block from it.
int mailbox_wait(Decoder *d)
{
    if (!d->mb_pending) return 0;
    d->mb_pending = 0;
    Mailbox *mb = d->mb;
    bool arrived = false;
    for (unsigned long long spins = 0; spins < 400ull * 1000 * 1000; spins++) {
        if (mb->seq == d->mb_seq) { arrived = true; break; }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
        if ((spins & 0xfffff) == 0xfffff && cudaStreamQuery(d->stream) != cudaErrorNotReady) {
            // the stream is idle (or broken): the report is there now or never
            arrived = mb->seq == d->mb_seq;
            break;
        }
    }
    if (!arrived) {
        const cudaError_t e = cudaStreamSynchronize(d->stream);
        if (mb->seq != d->mb_seq) { set_err("per-bit stage did not report (%s)", cudaGetErrorString(e)); cudaGetLastError(); return -1; }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    if (mb->declined) { set_err("per-bit stage declined to run (stage %lld)", d->h_ctl->T); return -1; }
    memcpy(d->h_ctl, mb->ctl_head, CTL_HOST_BYTES);
    d->spec_bit = mb->walk_bit;
    if (d->spec_valid && d->spec_bit < -1) { d->spec_valid = 0; d->cache_valid = 0; }      // the walker gave up: its cache entries cannot be trusted
    if (d->h_ctl->error) { set_err("device control block reports invariant violation %d", d->h_ctl->error); return -1; }
    return 0;
}

// update_viterbi224_blk(p, syms, 1) of the per-bit streaming pattern (vdecode.c:145): ONE asynchronous launch.  The host
// knows from its (fresh) mirror of the control block whether the stage can renormalise or needs the saturating variant; if
// neither, the call returns 0 -- the reference's return value -- without waiting, and the next call picks the report up.
// The launch also carries the decodebit walk the caller asked for after the previous stage (vdecode.c:152).
// Returns -2 when the fast path does not apply (the caller then takes the synchronous path).
int update_one_fast(Decoder *d, int s0, int s1)
{
    if (d->no_mailbox || d->force_sat || d->time_kernels || !d->mb) return -2;
    const Ctl *h = d->h_ctl;
    if (h->error || h->maxR + 510 > 32767 || h->spread > MAX_FAST_SPREAD) return -2;      // what k_acs_single<false> checks itself
    const long long T = h->T;
    SingleArgs a{d->ctl, {d->metrics[0], d->metrics[1], d->metrics[2]}, d->ring, d->row_fmt, nullptr, d->len, 0, T, 1, s0, s1};
    a.mailbox = d->mb_dev;
    a.seq = ++d->mb_seq;
    a.slow_form = d->slow_single;
    d->spec_valid = 0;
    if (d->spec_want && d->spec_delay < d->len && !d->no_walk_cache) {
        if (!d->walk_cache) {
            CU(cudaMalloc(&d->walk_cache, (size_t)d->len * sizeof(uint32_t)));
            CU(cudaMalloc(&d->d_walk_steps, sizeof(unsigned)));
            CU(cudaMemsetAsync(d->d_walk_steps, 0, sizeof(unsigned), d->stream));
        }
        const long long Tn = T + 1;
        const bool inc = d->cache_valid && d->cache_delay == d->spec_delay && d->cache_end == d->spec_end && Tn > d->cache_T && Tn - d->cache_T < d->spec_delay;
        a.spec_walk = 1;
        a.spec_delay = d->spec_delay;
        a.spec_end = d->spec_end;
        a.spec_prev_T = inc ? d->cache_T : Tn - 4ll * d->len - 4ll * d->spec_delay;
        a.walk_cache = d->walk_cache;
        a.walk_steps = d->d_walk_steps;
        a.all_canon = d->fused_rows ? 0 : 1;
        d->cache_valid = 1; d->cache_T = Tn; d->cache_delay = d->spec_delay; d->cache_end = d->spec_end;
        d->spec_valid = 1;
        d->spec_T = Tn;
    }
    CU(launch_single(a, false, d->stream));
    d->launches++;
    d->stages_total++;
    d->mb_pending = 1;
    const int ren0 = h->renorm_count;
    if (h->R0 + 510 < RENORM_TRIGGER) return 0;        // state 0 cannot reach the trigger in this stage (viterbi224_sse2.c:351)
    if (mailbox_wait(d)) return -1;
    return d->h_ctl->renorm_count - ren0;
}

// ---- recycled decoders: decode.c:216-229 creates and deletes a 1024-row decoder PER FRAME.  A create / delete pair
// costs ~6 ms of allocations (pinned host memory, streams, events, tensor maps); a recycled decoder only clears the ring
// rows it ever wrote and re-runs init.  At most DEC_POOL_MAX decoders wait here, none with auxiliary decoders attached. ----
std::mutex g_dec_mu;
std::vector<Decoder *> g_dec_pool;
constexpr size_t DEC_POOL_MAX = 2;

void destroy(Decoder *d)
{
    if (!d) return;
    for (int i = 0; i < NAUX; i++) if (d->aux[i]) { destroy(d->aux[i]); d->aux[i] = nullptr; }
    cudaSetDevice(d->dev);
    if (d->stream) cudaStreamSynchronize(d->stream);
    if (d->ring) pool_put(d->dev, d->ring_bytes, d->ring);
    for (int i = 0; i < NBUF; i++) cudaFree(d->metrics[i]);
    cudaFree(d->row_fmt); cudaFree(d->ctl);
    cudaFree(d->optab); cudaFree(d->tmaps); cudaFree(d->tmaps32);
    cudaFree(d->dsyms); cudaFree(d->dout); cudaFree(d->seg); cudaFree(d->d_redo); cudaFree(d->d_key);
    cudaFree(d->d_mnmx); cudaFree(d->d_result); cudaFree(d->d_flag); cudaFree(d->snap); cudaFree(d->d_segdiff); cudaFree(d->walk_cache); cudaFree(d->d_walk_steps);
    if (d->h_ctl) cudaFreeHost(d->h_ctl);
    if (d->h_result) cudaFreeHost(d->h_result);
    if (d->mb) cudaFreeHost(d->mb);
    if (d->ev0) cudaEventDestroy(d->ev0);
    if (d->ev1) cudaEventDestroy(d->ev1);
    if (d->kev0) cudaEventDestroy(d->kev0);
    if (d->kev1) cudaEventDestroy(d->kev1);
    if (d->stream) cudaStreamDestroy(d->stream);
    d->magic = 0;
    cudaGetLastError();
    free(d);
}

void release_parked_decoders()
{
    std::vector<Decoder *> drop;
    {
        std::lock_guard<std::mutex> lk(g_dec_mu);
        drop.swap(g_dec_pool);
    }
    for (Decoder *d : drop) { d->magic = MAGIC; destroy(d); }
}

int do_init(Decoder *d, int bias, int start_state)
{
    if (bind(d)) return -1;
    const uint32_t ss = start_state < 0 ? 0u : ((uint32_t)start_state & STATEMASK);
    d->cache_valid = 0;
    d->spec_valid = 0;
    CU(launch_init(d->metrics[0], d->ctl, ss, bias, start_state < 0 ? -1 : 0, d->stream));
    d->launches++;
    if (sync_ctl(d)) return -1;
    return 0;
}

TraceArgs trace_args(Decoder *d) { return TraceArgs{d->ring, d->row_fmt, d->len}; }

// A decoder from the recycle pool becomes indistinguishable from a freshly created one: default options, zero counters,
// the ring rows it ever wrote read as zero again, init(0).
int recycle(Decoder *d)
{
    if (bind(d)) return -1;
    d->magic = MAGIC;
    d->force_single = d->force_sat = d->force_careful = d->per_pass_launch = d->no_walk_cache = d->grid_limit = d->no_mailbox = 0;
    d->spec_want = d->spec_valid = 0;
    d->slow_single = 0;
    d->no_discard = 0;
    d->measure_all = 0;
    d->fused_rows = 0;                       // every row it ever wrote is cleared below, tags included
    d->tile32 = TILE32_DEFAULT;
    d->chain_seg = 64;
    d->chain_warm = 192;
    d->launches = d->acs_launches_timed = d->acs_passes_timed = d->chainback_redo = 0;
    d->acs_ms = 0;
    d->time_kernels = 0;
    // rows are written from row 0 upwards unless v224x_set_state moved the stage counter: then any row may be dirty
    const size_t rows = d->ring_dirty_all ? (size_t)d->len : (size_t)std::min<long long>(d->len, std::max<long long>(0, d->stages_total));
    d->ring_dirty_all = 0;
    if (rows) {
        CU(cudaMemsetAsync(d->ring, 0, rows * ROWBYTES, d->stream));
        CU(cudaMemsetAsync(d->row_fmt, 0, rows, d->stream));
    }
    CU(cudaMemsetAsync(d->ctl, 0, sizeof(Ctl), d->stream));
    CU(cudaMemsetAsync(d->d_redo, 0, sizeof(unsigned), d->stream));
    if (d->d_walk_steps) CU(cudaMemsetAsync(d->d_walk_steps, 0, sizeof(unsigned), d->stream));
    d->stages_total = 0;
    return do_init(d, INIT_BIAS, 0);
}

// Run nbits trellis stages on device-resident symbols.  Returns renormalisation count or -1.
// Every launch carries the stage counter it was issued for; a kernel whose predecessor declined (saturation
// watch) finds a different counter in the control block and declines too, so the host can enqueue a whole
// batch without looking and sort it out afterwards.
int update_core(Decoder *d, const uint8_t *dev_syms, int nbits, int arg_s0 = -1, int arg_s1 = -1)
{
    if (nbits <= 0) return 0;
    d->stages_total += nbits;
    d->spec_valid = 0;
    const long long T_start = d->h_ctl->T;
    const int ren_start = d->h_ctl->renorm_count;
    constexpr int BATCH_STAGES = 8192;
    int pos = 0;
    while (pos < nbits) {
        const int end = std::min(nbits, pos + BATCH_STAGES);
        if (d->time_kernels) CU(cudaEventRecord(d->kev0, d->stream));
        int passes_in_batch = 0;
        unsigned long long n = 0;
        int p = pos;
        const bool fuse = !d->force_single && !d->force_sat;
        if (fuse && end - p >= FK) {
            // one persistent launch runs all full passes of this batch as a dataflow ("per_pass_launch": one launch per pass)
            const int total = (end - p) / FK;
            d->fused_rows = 1;
            // two passes of one persistent launch may be in flight at once: with fewer than 2 * FK ring rows they would
            // write the same row, so such rings get one launch per pass
            const int per_launch = (d->per_pass_launch || d->len < 2 * FK) ? 1 : total;
            if (grow((void **)&d->optab, &d->optab_cap, passtab_bytes(per_launch))) return -1;
            for (int done_p = 0; done_p < total; done_p += per_launch) {
                const int npasses = std::min(per_launch, total - done_p);
                MultiArgs m;
                m.nctx = 1;
                m.npasses = npasses;
                m.no_discard = d->no_discard;
                m.measure_all = d->measure_all;
                // a decoder alone is latency bound (pass n+1 needs all of pass n).  64-column tiles: one CTA per SM finishes a tile
                // sooner than three sharing the SM, and the pass with it (13.6 instead of 15.3 us); 32-column tiles (the default
                // for a lone decoder): half the tile latency and finer dependencies, 12.9 us (profiles/r02_probe_*.txt)
                m.grid_limit = d->grid_limit > 0 ? d->grid_limit : -1;
                // a decoder alone runs the 32-column-tile build of the pass ("tile32" = 0: the 64-column one)
                const bool t32 = d->tile32 != 0;
                m.ctx[0] = PersistArgs{d->ctl, {d->metrics[0], d->metrics[1], d->metrics[2]}, d->ring, d->row_fmt, dev_syms, t32 ? d->tmaps32 : d->tmaps, d->optab,
                                       d->len, p, (int)((d->h_ctl->cur + (p - pos) / FK) % NBUF), T_start + p, npasses, d->force_careful};
#ifdef V224_WITH_Q1
                if (d->tile32 == 3) { m.grid_limit = d->grid_limit; CU(v224_t32q1_launch_persist(&m, d->stream)); }
                else
#endif
                if (t32) { m.grid_limit = d->grid_limit > 0 ? d->grid_limit : -3; CU(v224_t32_launch_persist(&m, d->stream)); }   // 3 of 5 possible CTAs per SM: measured optimum
                else CU(launch_persist(m, d->stream));
                p += npasses * FK;
                n += 3;
            }
            passes_in_batch = total;
        }
        while (p < end) {
            SingleArgs a{d->ctl, {d->metrics[0], d->metrics[1], d->metrics[2]}, d->ring, d->row_fmt, dev_syms, d->len, p, T_start + p,
                         arg_s0 >= 0, arg_s0, arg_s1};
            a.slow_form = d->slow_single;
            CU(launch_single(a, d->force_sat != 0, d->stream));
            p += 1;
            n++;
        }
        if (d->time_kernels) CU(cudaEventRecord(d->kev1, d->stream));
        d->launches += n;
        if (sync_ctl(d)) return -1;
        if (d->time_kernels) {
            float ms = 0;
            CU(cudaEventElapsedTime(&ms, d->kev0, d->kev1));
            d->acs_ms += ms;
            d->acs_launches_timed += n;
            d->acs_passes_timed += passes_in_batch;
        }
        const int done = (int)(d->h_ctl->T - T_start);
        if (done >= end) { pos = end; continue; }
        // A pass declined to run: the reference's metrics are within 510*k of int16 saturation.
        // Do that stage with the exact saturating kernel and carry on from there.
        pos = done;
        SingleArgs a{d->ctl, {d->metrics[0], d->metrics[1], d->metrics[2]}, d->ring, d->row_fmt, dev_syms, d->len, pos, T_start + pos,
                     arg_s0 >= 0, arg_s0, arg_s1};
        CU(launch_single(a, true, d->stream));
        d->launches++;
        if (sync_ctl(d)) return -1;
        if (d->h_ctl->T != T_start + pos + 1) { set_err("saturating stage did not run (stage %lld)", d->h_ctl->T); return -1; }
        pos += 1;
    }
    return d->h_ctl->renorm_count - ren_start;
}

// Lockstep update of nctx decoders on ds[0]'s stream (everything they did before must be complete or on that stream).
int multi_update_core(Decoder **ds, const unsigned char *const *dev_syms, int nctx, int nbits, int *renorms_out)
{
    Decoder *d0 = ds[0];
    cudaStream_t st = d0->stream;
    long long T_start[MAX_CTX];
    int ren_start[MAX_CTX];
    for (int s = 0; s < nctx; s++) {
        T_start[s] = ds[s]->h_ctl->T;
        ren_start[s] = ds[s]->h_ctl->renorm_count;
        if (renorms_out) renorms_out[s] = 0;
    }
    constexpr int BATCH_STAGES = 8192;
    const int fused_total = nbits / FK * FK;
    int pos = 0;
    bool lockstep = true;
    for (int s = 0; s < nctx; s++) { ds[s]->stages_total += nbits; ds[s]->spec_valid = 0; }
    for (int s = 0; s < nctx; s++) if (ds[s]->force_single || ds[s]->force_sat || ds[s]->len < 2 * FK) lockstep = false;
    while (lockstep && pos < fused_total) {
        const int end = std::min(fused_total, pos + BATCH_STAGES);
        const int npasses = (end - pos) / FK;
        MultiArgs m;
        m.nctx = nctx;
        m.npasses = npasses;
        m.no_discard = d0->no_discard;
        m.measure_all = d0->measure_all;
        m.grid_limit = d0->grid_limit > 0 ? d0->grid_limit : (nctx == 1 ? -1 : 0);
        const bool mt32 = d0->tile32 == 2 || d0->tile32 == 4;        // measurement knobs: lockstep decoders on the 32-column-tile builds
        for (int s = 0; s < nctx; s++) {
            Decoder *d = ds[s];
            d->fused_rows = 1;
            if (grow((void **)&d->optab, &d->optab_cap, passtab_bytes(npasses))) return -1;
            m.ctx[s] = PersistArgs{d->ctl, {d->metrics[0], d->metrics[1], d->metrics[2]}, d->ring, d->row_fmt, dev_syms[s], mt32 ? d->tmaps32 : d->tmaps, d->optab,
                                   d->len, pos, d->h_ctl->cur, T_start[s] + pos, npasses, d->force_careful};
        }
        if (d0->time_kernels) CU(cudaEventRecord(d0->kev0, st));
#ifdef V224_WITH_Q1
        if (d0->tile32 == 4) { m.grid_limit = d0->grid_limit; CU(v224_t32q1_launch_persist(&m, st)); }
        else
#endif
        if (mt32) { m.grid_limit = d0->grid_limit; CU(v224_t32_launch_persist(&m, st)); }
        else CU(launch_persist(m, st));
        if (d0->time_kernels) CU(cudaEventRecord(d0->kev1, st));
        d0->launches += 2 * nctx + 1;
        for (int s = 0; s < nctx; s++) {
            Decoder *d = ds[s];
            CU(cudaMemcpyAsync(d->h_ctl, d->ctl, CTL_HOST_BYTES, cudaMemcpyDeviceToHost, st));
        }
        CU(cudaStreamSynchronize(st));
        if (d0->time_kernels) {
            float ms = 0;
            CU(cudaEventElapsedTime(&ms, d0->kev0, d0->kev1));
            d0->acs_ms += ms;
            d0->acs_launches_timed += 2 * nctx + 1;
            d0->acs_passes_timed += (unsigned long long)npasses * nctx;
        }
        for (int s = 0; s < nctx; s++) {
            if (ds[s]->h_ctl->error) { set_err("device control block reports invariant violation %d", ds[s]->h_ctl->error); return -1; }
            if (ds[s]->h_ctl->T - T_start[s] < end) lockstep = false;      // a decoder declined a pass: finish everyone one by one
        }
        pos = end;
    }
    // remainders (and the rare decoder that left lockstep) go through the single-decoder path, which resumes at the decoder's stage counter
    for (int s = 0; s < nctx; s++) {
        Decoder *d = ds[s];
        const int done = (int)(d->h_ctl->T - T_start[s]), ren = d->h_ctl->renorm_count - ren_start[s];
        int r = 0;
        if (done < nbits) {
            // the decoder's own stream takes the rest; everything so far ran on d0's stream and is complete
            d->stages_total -= nbits - done;           // update_core counts them itself
            r = update_core(d, dev_syms[s] + 2 * (size_t)done, nbits - done);
            if (r < 0) return -1;
        }
        if (renorms_out) renorms_out[s] = ren + r;
    }
    return 0;
}

// A metric snapshot request of a range decode: after local stage `at` of the range (stages [0, at) applied) the exact
// decoder's current metric buffer is copied to `dst` (16 MiB of device memory on this GPU).  dst == nullptr: none.
struct Snap { int at; uint16_t *dst; };

int take_snaps(Decoder *d, int pos, const Snap *snaps, int nsnaps, cudaStream_t st)
{
    for (int k = 0; k < nsnaps; k++)
        if (snaps[k].dst && snaps[k].at == pos)
            CU(cudaMemcpyAsync(snaps[k].dst, d->metrics[d->h_ctl->cur], METRICBYTES, cudaMemcpyDeviceToDevice, st));
    return 0;
}

// Stream decode of nbits stages starting at local stage `base` of a range whose first `lead` stages are warm-up
// (no output): the output of local stage t >= lead goes to dev_bits[t - lead].  Snapshot positions are range-local.
int stream_core(Decoder *d, const uint8_t *dev_syms, int nbits, int delay, uint8_t *dev_bits, int base = 0, int lead = 0,
                const Snap *snaps = nullptr, int nsnaps = 0)
{
    if (delay <= 0 || delay >= d->len) { set_err("stream decode needs 0 < delay < len (delay %d, len %d)", delay, d->len); return -1; }
    const int chunk_max = d->len - delay;
    int renorms = 0;
    if (take_snaps(d, base, snaps, nsnaps, d->stream)) return -1;
    for (int done = 0; done < nbits;) {
        int n = std::min(chunk_max, nbits - done);
        for (int k = 0; k < nsnaps; k++) {
            const int rel = snaps[k].at - base - done;
            if (snaps[k].dst && rel > 0 && rel < n) n = rel;
        }
        const long long T_first = d->h_ctl->T;
        const int r = update_core(d, dev_syms + 2 * (size_t)done, n);
        if (r < 0) return -1;
        renorms += r;
        const int o0 = std::max(0, lead - base - done);           // first stage of this chunk that emits
        if (o0 < n) {
            CU(launch_stream_trace(trace_args(d), T_first + o0, n - o0, delay, dev_bits + (base + done + o0 - lead), d->stream));
            d->launches++;
        }
        done += n;
        if (take_snaps(d, base + done, snaps, nsnaps, d->stream)) return -1;
    }
    CU(cudaStreamSynchronize(d->stream));
    return renorms;
}

// stream_core on another decoder of the same call: its own stream, after everything queued on `st` so far.
int stream_core_on(Decoder *h, cudaStream_t st, const uint8_t *dev_syms, int nbits, int delay, uint8_t *dev_bits, int base, int lead,
                   const Snap *snaps, int nsnaps)
{
    CU(cudaStreamSynchronize(st));
    return stream_core(h, dev_syms, nbits, delay, dev_bits, base, lead, snaps, nsnaps);
}

// Exchange everything but identity (the handle address the caller holds, the list of auxiliary decoders, options).
void swap_bodies(Decoder *d, Decoder *o)
{
    Decoder td = *d, to = *o;
    *d = to;
    *o = td;
    for (int i = 0; i < NAUX; i++) { d->aux[i] = td.aux[i]; o->aux[i] = nullptr; }
    d->d_segdiff = td.d_segdiff; o->d_segdiff = to.d_segdiff;
    // streams and events stay with the handle too (a timer started on the handle must stop on the same events)
    d->stream = td.stream; o->stream = to.stream;
    d->ev0 = td.ev0; d->ev1 = td.ev1; d->kev0 = td.kev0; d->kev1 = td.kev1;
    o->ev0 = to.ev0; o->ev1 = to.ev1; o->kev0 = to.kev0; o->kev1 = to.kev1;
    d->force_single = td.force_single; d->force_sat = td.force_sat; d->force_careful = td.force_careful;
    d->per_pass_launch = td.per_pass_launch; d->chain_seg = td.chain_seg; d->chain_warm = td.chain_warm; d->no_walk_cache = td.no_walk_cache;
    d->grid_limit = td.grid_limit; d->tile32 = td.tile32;
    d->time_kernels = td.time_kernels; d->acs_ms = td.acs_ms; d->acs_launches_timed = td.acs_launches_timed;
    d->acs_passes_timed = td.acs_passes_timed; d->launches = td.launches;
    o->time_kernels = to.time_kernels; o->acs_ms = to.acs_ms; o->acs_launches_timed = to.acs_launches_timed;
    o->acs_passes_timed = to.acs_passes_timed; o->launches = to.launches;
}

// ---- segmented stream decode -----------------------------------------------------------------
// One stream, nseg contiguous segments, nseg decoders advanced in lockstep by the persistent kernel (the CTAs that
// would wait at one decoder's pass boundary work on another decoder's tiles).  Decoder 0 (the caller's handle)
// continues its current state over segment 0; decoder i >= 1 starts W = delay + conv stages before its segment from
// uniform metrics.  A decoder started late makes the same decisions as the sequential one from the stage at which
// the two path-metric vectors differ by a constant.  That is CHECKED, not assumed: decoder i's metrics `conv`
// stages after its start are compared with decoder i-1's metrics at the same stream position (`delay` stages
// before decoder i's first output, so that every row its walks touch lies after the check).  If a check fails,
// the stream from that segment on is decoded again sequentially by the last exact decoder.  The output therefore
// always equals v224x_stream_decode's.
Decoder *make_aux(Decoder *d)
{
    const int saved = g_device;
    g_device = d->dev;
    Decoder *a = static_cast<Decoder *>(create_viterbi224(d->len));
    g_device = saved;
    if (a) { a->force_careful = d->force_careful; a->per_pass_launch = d->per_pass_launch; }
    return a;
}

// `lead` leading stages are warm-up of the whole range (multi-GPU time segments: the caller started the handle from uniform
// metrics); the output of local stage t >= lead goes to dev_bits[t - lead].  nbits = all local stages (lead included).
// Snapshots (range-local positions) are taken from the decoder that is exact at that position.
int seg_core(Decoder *d, const uint8_t *dev_syms, int nbits, int delay, uint8_t *dev_bits, int nseg, int conv, v224x_seg_report *rep,
             int lead = 0, const Snap *snaps = nullptr, int nsnaps = 0)
{
    if (delay <= 0 || delay >= d->len) { set_err("stream decode needs 0 < delay < len (delay %d, len %d)", delay, d->len); return -1; }
    if (conv < 0) conv = 2048;
    const int W = delay + conv;                          // warm-up of segments 1..
    int S = std::max(1, std::min(nseg, MAX_CTX));
    constexpr int MIN_SEG = 4096;                        // not worth a warm-up below this
    while (S > 1 && (long long)nbits < (long long)S * (W + MIN_SEG)) S--;
    // the first decoder carries the range's own warm-up and its early snapshot; the last one the late snapshot
    while (S > 1) {
        const int A_ = ((nbits - W) / S) & ~7;
        bool ok = A_ >= delay && lead < A_ + W;
        for (int k = 0; k < nsnaps; k++)
            if (snaps[k].dst && !(snaps[k].at <= A_ + W || snaps[k].at >= (S - 1) * A_ + W)) ok = false;   // a decoder's own warm-up is not exact
        if (ok) break;
        S--;
    }
    if (rep) { rep->segments = S; rep->warm = S > 1 ? W : 0; rep->verified = 0; rep->redone = 0; rep->extra_stages = 0; rep->worst_spread = 0; }
    if (S == 1) return stream_core(d, dev_syms, nbits, delay, dev_bits, 0, lead, snaps, nsnaps);

    const int A = ((nbits - W) / S) & ~7;                 // lockstep length of a segment's own range
    const int Ltot = A + W;                               // local stages every decoder runs in lockstep
    Decoder *D[MAX_CTX] = {d};
    const uint8_t *sy[MAX_CTX];
    long long T0[MAX_CTX];
    for (int i = 1; i < S; i++) {
        if (!d->aux[i - 1]) d->aux[i - 1] = make_aux(d);
        D[i] = d->aux[i - 1];
        if (!D[i]) { set_err("segmented decode: cannot create decoder %d of %d: %s", i, S, g_err); return -1; }
        if (!D[i]->snap && cudaMalloc(&D[i]->snap, METRICBYTES) != cudaSuccess) { set_err("segmented decode: snapshot allocation failed"); cudaGetLastError(); return -1; }
        if (do_init(D[i], INIT_BIAS, -1)) return -1;      // uniform metrics: no state is favoured
    }
    if (!d->d_segdiff) CU(cudaMalloc(&d->d_segdiff, 2 * MAX_CTX * sizeof(int)));
    for (int i = 0; i < S; i++) {
        sy[i] = dev_syms + 2 * (size_t)i * A;             // decoder i runs range stages [i*A, i*A + Ltot)
        T0[i] = D[i]->h_ctl->T;
        CU(cudaStreamSynchronize(D[i]->stream));
    }
    cudaStream_t st = d->stream;
    const int chunk_max = d->len - delay;
    const int check_early = conv, check_late = Ltot - delay;       // local stage of the two sides of a hand-over check
    // range snapshots inside the lockstep part: positions of decoder 0 (early) and of decoder S-1 (late), in loop-local stages
    int stops[2 + 2 * 4];
    int nstops = 0;
    stops[nstops++] = check_early;
    stops[nstops++] = check_late;
    for (int k = 0; k < nsnaps && k < 4; k++) {
        if (!snaps[k].dst) continue;
        if (snaps[k].at <= Ltot) stops[nstops++] = snaps[k].at;                                   // decoder 0 is exact there
        else if (snaps[k].at <= (S - 1) * A + Ltot) stops[nstops++] = snaps[k].at - (S - 1) * A;  // the last decoder's own range
    }
    if (take_snaps(D[0], 0, snaps, nsnaps, st)) return -1;
    int a = 0;
    while (a < Ltot) {
        int b = std::min(Ltot, a + chunk_max);
        for (int k = 0; k < nstops; k++) if (a < stops[k] && b > stops[k]) b = stops[k];
        const uint8_t *sp[MAX_CTX];
        for (int i = 0; i < S; i++) sp[i] = sy[i] + 2 * (size_t)a;
        if (multi_update_core(D, sp, S, b - a, nullptr)) return -1;
        StreamTraceMulti tm;
        tm.njobs = 0;
        tm.delay = delay;
        for (int i = 0; i < S; i++) {
            // decoder i >= 1 emits nothing inside its warm-up, nobody inside the range's
            const int o0 = std::max(std::max(a, i ? W : 0), lead - i * A);
            if (o0 < b) tm.job[tm.njobs++] = StreamTraceJob{trace_args(D[i]), T0[i] + o0, b - o0, dev_bits + ((size_t)i * A + o0 - lead)};
        }
        if (tm.njobs) {
            // the decoders' tracebacks are independent chains of dependent loads: one launch for all of them
            CU(launch_stream_trace_multi(tm, st));
            d->launches++;
        }
        if (b == check_early)
            for (int i = 1; i < S; i++) CU(cudaMemcpyAsync(D[i]->snap, D[i]->metrics[D[i]->h_ctl->cur], METRICBYTES, cudaMemcpyDeviceToDevice, st));
        if (b == check_late)
            for (int i = 0; i + 1 < S; i++) {
                CU(launch_metric_diff(D[i]->metrics[D[i]->h_ctl->cur], D[i + 1]->snap, d->d_segdiff + 2 * i, st));
                d->launches++;
            }
        // range snapshots: the first decoder is exact by definition, the last one if every hand-over check passes
        // (otherwise the sequential redo below passes the position again and overwrites the snapshot)
        if (take_snaps(D[0], b, snaps, nsnaps, st)) return -1;
        for (int k = 0; k < nsnaps; k++)
            if (snaps[k].dst && snaps[k].at > Ltot && snaps[k].at - (S - 1) * A == b)
                CU(cudaMemcpyAsync(snaps[k].dst, D[S - 1]->metrics[D[S - 1]->h_ctl->cur], METRICBYTES, cudaMemcpyDeviceToDevice, st));
        a = b;
    }
    int diff[2 * MAX_CTX];
    CU(cudaMemcpyAsync(diff, d->d_segdiff, sizeof(int) * 2 * (S - 1), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    int head = S - 1;                                     // the decoder that holds the exact state at the end of its range
    for (int i = 0; i + 1 < S; i++) {
        const int spread = diff[2 * i + 1] - diff[2 * i];
        if (rep && spread > rep->worst_spread) rep->worst_spread = spread;
        if (spread != 0) { head = i; break; }               // decoder i+1 had not converged: everything after decoder i is redone
        if (rep) rep->verified++;
    }
    // the tail: what is left after the lockstep part, or -- after a failed check -- everything from decoder `head`'s end
    const int done_upto = (head + 1) * A + W;              // range stages decoded exactly so far
    if (rep) { rep->redone = S - 1 - head; rep->extra_stages = (long long)(S - 1) * W + (long long)(S - 1 - head) * A; }
    if (done_upto < nbits) {
        Decoder *h = D[head];
        if (stream_core_on(h, st, dev_syms + 2 * (size_t)done_upto, nbits - done_upto, delay, dev_bits, done_upto, lead, snaps, nsnaps) < 0) return -1;
    }
    CU(cudaStreamSynchronize(st));
    if (head != 0) swap_bodies(d, D[head]);               // the caller's handle continues the stream from its end
    return 0;
}

// ---- batched frame decode -------------------------------------------------------------------
// The reference's frame callers (vtest224.c:116-118, hybridtest.c:186-193, decode.c:220-222) decode one frame at a
// time: init(start) / update(framebits) / chainback(framebits, end).  Frames are independent, so a batch of them is
// another natural axis for the lockstep launch: up to MAX_CTX decoders run one frame each side by side (the CTAs that
// would wait at one frame's pass boundary work on another frame's tiles).  Every frame is decoded exactly as the
// three-call sequence would decode it.
int frames_core(Decoder *d, const uint8_t *host_syms, int nframes, int framebits, const unsigned *start_states, const unsigned *end_states,
                uint8_t *host_data, int nlock)
{
    if (framebits <= 0 || framebits > d->len) { set_err("frame decode needs 0 < framebits <= len (framebits %d, len %d)", framebits, d->len); return -1; }
    const int S = std::max(1, std::min(std::min(nlock, MAX_CTX), nframes));
    const size_t fsyms = 2 * (size_t)framebits, fbytes = ((size_t)framebits + 7) / 8;
    // Two lane sets: while one set's frames are in their traceback (latency-bound chains of dependent ring loads, each on
    // its own decoder's stream) the other set's frames run their ACS passes.  Set 0 = this decoder + its lockstep partners,
    // set 1 = S more decoders (only when there is a second group of frames).
    const int nsets = nframes > S ? 2 : 1;
    Decoder *D[2][MAX_CTX] = {{d}, {nullptr}};
    for (int i = 1; i < S; i++) {
        if (!d->aux[i - 1]) d->aux[i - 1] = make_aux(d);
        D[0][i] = d->aux[i - 1];
        if (!D[0][i]) { set_err("frame decode: cannot create decoder %d of %d: %s", i, S, g_err); return -1; }
    }
    for (int i = 0; nsets == 2 && i < S; i++) {
        Decoder *&a = d->aux[MAX_CTX - 1 + i];
        if (!a) a = make_aux(d);
        D[1][i] = a;
        if (!a) { set_err("frame decode: cannot create decoder %d of the second lane set: %s", i, g_err); return -1; }
    }
    if (grow((void **)&d->dsyms, &d->dsyms_cap, fsyms * (size_t)nframes)) return -1;
    if (grow((void **)&d->dout, &d->dout_cap, fbytes * (size_t)nframes)) return -1;
    cudaStream_t st = d->stream;
    CU(cudaMemcpyAsync(d->dsyms, host_syms, fsyms * (size_t)nframes, cudaMemcpyHostToDevice, st));
    const int L = std::max(8, d->chain_seg & ~7);
    const uint32_t nseg = ((uint32_t)framebits + L - 1) / L;
    for (int k = 0; k < nsets; k++)
        for (int i = 0; i < S; i++) {
            if (grow((void **)&D[k][i]->seg, &D[k][i]->seg_cap, 2 * (size_t)nseg * sizeof(uint32_t))) return -1;
            CU(cudaStreamSynchronize(D[k][i]->stream));
        }
    CU(cudaStreamSynchronize(st));                       // the symbols are on the device before any other stream reads them
    int group = 0;
    for (int f0 = 0; f0 < nframes; f0 += S, group++) {
        Decoder **G = D[nsets == 2 ? (group & 1) : 0];
        cudaStream_t sg = G[0]->stream;                  // the set's ACS stream (its first decoder's)
        const int nb = std::min(S, nframes - f0);
        const uint8_t *sp[MAX_CTX];
        // the set's previous tracebacks (two groups back) are through before its rings and metrics are written again;
        // they ran while the other set was in its ACS passes
        for (int i = 1; i < S; i++) CU(cudaStreamSynchronize(G[i]->stream));
        for (int i = 0; i < nb; i++) {
            G[i]->cache_valid = 0;
            const uint32_t ss = (start_states ? start_states[f0 + i] : 0u) & STATEMASK;
            CU(launch_init(G[i]->metrics[0], G[i]->ctl, ss, INIT_BIAS, 0, sg));                 // init_viterbi224(start), viterbi224_sse2.c:37-53
            CU(cudaMemcpyAsync(G[i]->h_ctl, G[i]->ctl, CTL_HOST_BYTES, cudaMemcpyDeviceToHost, sg));
            sp[i] = d->dsyms + fsyms * (size_t)(f0 + i);
        }
        CU(cudaStreamSynchronize(sg));
        d->launches += nb;
        if (multi_update_core(G, sp, nb, framebits, nullptr)) return -1;         // ends with a synchronisation: the rows are written
        for (int i = 0; i < nb; i++) {
            const uint32_t es = end_states ? end_states[f0 + i] : 0u;
            CU(launch_chainback(trace_args(G[i]), (uint32_t)framebits, es, L, d->chain_warm, d->dout + fbytes * (size_t)(f0 + i), G[i]->seg, G[i]->seg + nseg,
                                G[i]->d_redo, G[i]->stream));
            d->launches += 2;
        }
    }
    for (int k = 0; k < nsets; k++)
        for (int i = 0; i < S; i++) CU(cudaStreamSynchronize(D[k][i]->stream));
    CU(cudaMemcpyAsync(host_data, d->dout, fbytes * (size_t)nframes, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}


// ---- one time segment of a longer stream (multi-GPU decode) -----------------------------------
// Range-local stages [0, lead + nout); the first `lead` are warm-up from uniform metrics (lead == 0: the handle's
// current state continues).  Snapshots: early after stage lead - delay, late after stage lead + nout - delay.
int range_core(Decoder *d, const uint8_t *dev_syms, int lead, int nout, int delay, uint8_t *dev_bits, int nseg, int conv,
               uint16_t *snap_early, uint16_t *snap_late, v224x_seg_report *rep)
{
    if (lead < 0 || nout < 0) { set_err("range decode: negative length"); return -1; }
    if (snap_early && lead < delay) { set_err("range decode: the early snapshot needs lead >= delay (lead %d, delay %d)", lead, delay); return -1; }
    if (snap_late && nout < delay) { set_err("range decode: the late snapshot needs nout >= delay (nout %d, delay %d)", nout, delay); return -1; }
    if (lead > 0 && do_init(d, INIT_BIAS, -1)) return -1;        // mid-stream start: no state is favoured
    Snap snaps[2] = {{lead - delay, snap_early}, {lead + nout - delay, snap_late}};
    if (lead + nout == 0) return 0;
    return seg_core(d, dev_syms, lead + nout, delay, dev_bits, nseg, conv, rep, lead, snaps, 2) < 0 ? -1 : 0;
}

int spread_core(Decoder *d, const uint16_t *a, const uint16_t *b, int *spread_out)
{
    if (!d->d_segdiff) CU(cudaMalloc(&d->d_segdiff, 2 * MAX_CTX * sizeof(int)));
    CU(launch_metric_diff(a, b, d->d_segdiff, d->stream));
    d->launches++;
    int diff[2];
    CU(cudaMemcpyAsync(diff, d->d_segdiff, sizeof diff, cudaMemcpyDeviceToHost, d->stream));
    CU(cudaStreamSynchronize(d->stream));
    *spread_out = diff[1] - diff[0];
    return 0;
}

} // namespace

// ---- the multi-GPU context ------------------------------------------------------------------
constexpr int MULTI_MAX = 16;
struct v224x_multi {
    int n;
    int devs[MULTI_MAX];
    Decoder *dec[MULTI_MAX];
    uint16_t *snap_early[MULTI_MAX], *snap_late[MULTI_MAX], *snap_peer[MULTI_MAX];   // on the slot's own GPU
    uint8_t *dsyms[MULTI_MAX], *dbits[MULTI_MAX];
    size_t dsyms_cap[MULTI_MAX], dbits_cap[MULTI_MAX];
    int head;                      // slot whose decoder holds the state at the end of the stream decoded so far
    int ring_rows;
};

namespace {

struct RangeJob {
    int slot;
    long long first;               // first output stage of the range (stream position inside this call)
    int lead, nout;
    bool want_early, want_late;
    v224x_seg_report rep;
    int rc;
    std::string err;
};

// Decode one range on its slot's GPU: symbols in, bits out (host memory of the caller), snapshots left on the GPU.
void run_range(v224x_multi *m, RangeJob *j, const unsigned char *syms, int delay, unsigned char *bits_out, int nseg, int conv)
{
    j->rc = -1;
    const int k = j->slot;
    Decoder *d = m->dec[k];
    do {
        if (bind(d)) break;
        const size_t nsym = 2 * ((size_t)j->lead + (size_t)j->nout);
        if (grow((void **)&m->dsyms[k], &m->dsyms_cap[k], nsym)) break;
        if (grow((void **)&m->dbits[k], &m->dbits_cap[k], (size_t)j->nout)) break;
        if (cudaMemcpyAsync(m->dsyms[k], syms + 2 * (size_t)(j->first - j->lead), nsym, cudaMemcpyHostToDevice, d->stream) != cudaSuccess) {
            set_err("range %lld: copying the symbols to device %d failed: %s", j->first, d->dev, cudaGetErrorString(cudaGetLastError()));
            break;
        }
        if (range_core(d, m->dsyms[k], j->lead, j->nout, delay, m->dbits[k], nseg, conv, j->want_early ? m->snap_early[k] : nullptr,
                       j->want_late ? m->snap_late[k] : nullptr, &j->rep))
            break;
        if (cudaMemcpyAsync(bits_out + j->first, m->dbits[k], (size_t)j->nout, cudaMemcpyDeviceToHost, d->stream) != cudaSuccess ||
            cudaStreamSynchronize(d->stream) != cudaSuccess) {
            set_err("range %lld: copying the decoded bits back from device %d failed: %s", j->first, d->dev, cudaGetErrorString(cudaGetLastError()));
            break;
        }
        j->rc = 0;
    } while (0);
    if (j->rc) j->err = g_err;                  // g_err is per thread: hand the text to the caller's thread
}

// Is the decoder of slot `later` (early snapshot) in step with the decoder of slot `earlier` (late snapshot)?
int handover_spread(v224x_multi *m, int earlier, int later, int *spread)
{
    Decoder *d = m->dec[earlier];
    if (bind(d)) return -1;
    const uint16_t *other = m->snap_early[later];
    if (m->devs[earlier] != m->devs[later]) {
        // 16 MiB over NVLink (peer copy; the runtime stages it through the host where peer access is not possible)
        CU(cudaMemcpyPeerAsync(m->snap_peer[earlier], m->devs[earlier], m->snap_early[later], m->devs[later], METRICBYTES, d->stream));
        other = m->snap_peer[earlier];
    }
    return spread_core(d, m->snap_late[earlier], other, spread);
}

int multi_core(v224x_multi *m, const unsigned char *syms, long long nbits, int delay, unsigned char *bits_out, int nseg, int conv,
               v224x_multi_report *rep)
{
    if (delay <= 0 || delay >= m->ring_rows) { set_err("multi decode needs 0 < delay < ring rows (delay %d, rows %d)", delay, m->ring_rows); return -1; }
    if (nbits > 0x7fffffffll - 65536) { set_err("multi decode: at most 2^31 - 65536 bits per call"); return -1; }
    if (conv < 0) conv = 2048;
    const int W = delay + conv;
    // a range pays its warm-up; below a few warm-ups of output it is not worth a GPU
    int G = m->n;
    const long long min_range = std::max<long long>(4ll * W, 16384);
    while (G > 1 && nbits / G < min_range) G--;
    if (rep) { memset(rep, 0, sizeof *rep); rep->gpus = G; }
    RangeJob jobs[MULTI_MAX];
    for (int g = 0; g < G; g++) {
        RangeJob &j = jobs[g];
        j.slot = (m->head + g) % m->n;
        j.first = nbits * g / G;
        const long long last = nbits * (g + 1) / G;
        j.lead = g ? W : 0;
        j.nout = (int)(last - j.first);
        j.want_early = g > 0;
        j.want_late = g + 1 < G;
        j.rc = 0;
        memset(&j.rep, 0, sizeof j.rep);
    }
    {
        // one host thread per GPU for the duration of the call (the calling thread takes the first range)
        std::vector<std::thread> th;
        for (int g = 1; g < G; g++) th.emplace_back(run_range, m, &jobs[g], syms, delay, bits_out, nseg, conv);
        run_range(m, &jobs[0], syms, delay, bits_out, nseg, conv);
        for (auto &t : th) t.join();
    }
    for (int g = 0; g < G; g++)
        if (jobs[g].rc) { set_err("multi decode, range %d on device %d: %s", g, m->devs[jobs[g].slot], jobs[g].err.c_str()); return -1; }
    // hand-overs, in stream order; a range whose decoder had not converged is decoded again by the decoder that is
    // exact at the range's start (the one that produced the previous range's accepted output)
    int exact = jobs[0].slot;
    for (int g = 1; g < G; g++) {
        int spread = 0;
        if (handover_spread(m, exact, jobs[g].slot, &spread)) return -1;
        if (rep && spread > rep->worst_spread) rep->worst_spread = spread;
        if (spread == 0) {
            if (rep) rep->handovers_verified++;
            exact = jobs[g].slot;
            continue;
        }
        RangeJob redo = jobs[g];
        redo.slot = exact;
        redo.lead = 0;
        redo.want_early = false;
        run_range(m, &redo, syms, delay, bits_out, nseg, conv);
        if (redo.rc) { set_err("multi decode, range %d again on device %d: %s", g, m->devs[exact], redo.err.c_str()); return -1; }
        if (rep) {
            rep->ranges_redone++;
            rep->extra_stages += redo.nout + redo.rep.extra_stages;
            rep->inner_verified += redo.rep.verified;
            rep->inner_redone += redo.rep.redone;
        }
    }
    m->head = exact;
    if (rep)
        for (int g = 0; g < G; g++) {
            rep->extra_stages += jobs[g].lead + jobs[g].rep.extra_stages;
            rep->inner_verified += jobs[g].rep.verified;
            rep->inner_redone += jobs[g].rep.redone;
        }
    return 0;
}

} // namespace

// ==========================================================================================
// the reference's nine entry points
// ==========================================================================================
extern "C" {

void *create_viterbi224(int len)
{
    if (len <= 0) { set_err("create_viterbi224: len must be positive"); return nullptr; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_err("create_viterbi224: no CUDA device (this library has no CPU path)");
        return nullptr;
    }
    int dev = g_device;
    if (dev < 0) { if (cudaGetDevice(&dev) != cudaSuccess) dev = 0; }
    if (cudaSetDevice(dev) != cudaSuccess) { set_err("cudaSetDevice(%d) failed", dev); cudaGetLastError(); return nullptr; }

    {
        Decoder *r = nullptr;
        {
            std::lock_guard<std::mutex> lk(g_dec_mu);
            for (size_t i = 0; i < g_dec_pool.size(); i++)
                if (g_dec_pool[i]->dev == dev && g_dec_pool[i]->len == len) { r = g_dec_pool[i]; g_dec_pool.erase(g_dec_pool.begin() + i); break; }
        }
        if (r) {
            if (recycle(r) == 0) return r;
            r->magic = MAGIC;
            destroy(r);                                  // something is wrong with it: fall through to a fresh one
        }
    }
    Decoder *d = static_cast<Decoder *>(calloc(1, sizeof(Decoder)));
    if (!d) return nullptr;
    d->magic = MAGIC;
    d->dev = dev;
    d->len = len;
    d->chain_seg = 64;
    d->chain_warm = 192;
    d->tile32 = TILE32_DEFAULT;
    d->ring_bytes = (size_t)len * ROWBYTES;
    bool ok = true;
    ok = ok && cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < NBUF; i++) ok = ok && cudaMalloc(&d->metrics[i], METRICBYTES) == cudaSuccess;
    ok = ok && cudaMalloc(&d->tmaps, NBUF * TMAP_BYTES) == cudaSuccess && cudaMalloc(&d->tmaps32, NBUF * TMAP_BYTES) == cudaSuccess;
    if (ok) {
        const char *why = nullptr;
        if (build_metric_tensor_maps(d->metrics, d->tmaps, d->stream, &why) != cudaSuccess ||
            v224_t32_build_metric_tensor_maps(d->metrics, d->tmaps32, d->stream, &why) != cudaSuccess) {
            set_err("create_viterbi224(%d): tensor maps: %s", len, why ? why : cudaGetErrorString(cudaGetLastError()));
            destroy(d);
            return nullptr;
        }
    }
    ok = ok && cudaMalloc(&d->row_fmt, (size_t)len) == cudaSuccess;
    ok = ok && cudaMalloc(&d->ctl, sizeof(Ctl)) == cudaSuccess;
    ok = ok && cudaMalloc(&d->d_redo, sizeof(unsigned)) == cudaSuccess;
    ok = ok && cudaMalloc(&d->d_key, sizeof(unsigned long long)) == cudaSuccess;
    ok = ok && cudaMalloc(&d->d_mnmx, 2 * sizeof(unsigned)) == cudaSuccess;
    ok = ok && cudaMalloc(&d->d_result, 2 * sizeof(unsigned long long)) == cudaSuccess;
    ok = ok && cudaMalloc(&d->d_flag, sizeof(int)) == cudaSuccess;
    ok = ok && cudaMallocHost(&d->h_ctl, sizeof(Ctl)) == cudaSuccess;
    ok = ok && cudaMallocHost(&d->h_result, 2 * sizeof(unsigned long long)) == cudaSuccess;
    ok = ok && cudaHostAlloc(&d->mb, sizeof(Mailbox), cudaHostAllocMapped) == cudaSuccess &&
         cudaHostGetDevicePointer(reinterpret_cast<void **>(&d->mb_dev), d->mb, 0) == cudaSuccess;
    if (ok) memset(d->mb, 0, sizeof(Mailbox));
    ok = ok && cudaEventCreate(&d->ev0) == cudaSuccess && cudaEventCreate(&d->ev1) == cudaSuccess;
    ok = ok && cudaEventCreate(&d->kev0) == cudaSuccess && cudaEventCreate(&d->kev1) == cudaSuccess;
    if (ok) {
        d->ring = static_cast<uint32_t *>(pool_get(dev, d->ring_bytes));
        ok = d->ring != nullptr;
    }
    // a fresh ring reads as zero (the reference's malloc'd ring is unspecified; zero is what a
    // fresh mapping holds and what bitsync.c:245 style early reads rely on)
    ok = ok && cudaMemsetAsync(d->ring, 0, d->ring_bytes, d->stream) == cudaSuccess;
    ok = ok && cudaMemsetAsync(d->row_fmt, 0, (size_t)len, d->stream) == cudaSuccess;
    ok = ok && cudaMemsetAsync(d->ctl, 0, sizeof(Ctl), d->stream) == cudaSuccess;
    ok = ok && cudaMemsetAsync(d->d_redo, 0, sizeof(unsigned), d->stream) == cudaSuccess;
    if (!ok) {
        set_err("create_viterbi224(%d): device allocation failed: %s", len, cudaGetErrorString(cudaGetLastError()));
        destroy(d);
        return nullptr;
    }
    if (do_init(d, INIT_BIAS, 0)) { destroy(d); return nullptr; }     // viterbi224_sse2.c:78
    return d;
}

int init_viterbi224(void *p, int starting_state)
{
    Decoder *d = as_dec(p);
    if (!d) return -1;
    return do_init(d, INIT_BIAS, (int)((uint32_t)starting_state & STATEMASK));
}

int update_viterbi224_blk(void *p, const unsigned char *syms, int nbits)
{
    Decoder *d = as_dec(p);
    if (!d) return -1;
    if (nbits <= 0) return 0;
    if (bind(d)) return -1;
    if (nbits == 1) {    // vdecode.c:145 -- the two symbols ride in the kernel arguments
        const int r = update_one_fast(d, syms[0], syms[1]);
        if (r != -2) return r;
        d->spec_valid = 0;
        return update_core(d, nullptr, 1, syms[0], syms[1]);
    }
    d->spec_valid = 0;
    if (grow((void **)&d->dsyms, &d->dsyms_cap, 2 * (size_t)nbits)) return -1;
    CU(cudaMemcpyAsync(d->dsyms, syms, 2 * (size_t)nbits, cudaMemcpyHostToDevice, d->stream));
    return update_core(d, d->dsyms, nbits);
}

int chainback_viterbi224(void *p, unsigned char *data, unsigned int nbits, unsigned int endstate)
{
    Decoder *d = as_dec(p);
    if (!d) return -1;
    if (nbits == 0) return 0;
    if (bind(d)) return -1;
    const size_t nbytes = ((size_t)nbits + 7) / 8;
    const int L = std::max(8, d->chain_seg & ~7);
    const uint32_t nseg = (nbits + L - 1) / L;
    if (grow((void **)&d->dout, &d->dout_cap, nbytes)) return -1;
    if (grow((void **)&d->seg, &d->seg_cap, 2 * (size_t)nseg * sizeof(uint32_t))) return -1;
    CU(launch_chainback(trace_args(d), nbits, endstate, L, d->chain_warm, d->dout, d->seg, d->seg + nseg, d->d_redo, d->stream));
    d->launches += 2;
    CU(cudaMemcpyAsync(data, d->dout, nbytes, cudaMemcpyDeviceToHost, d->stream));
    CU(cudaStreamSynchronize(d->stream));
    return 0;
}

static int walk(Decoder *d, int delay, int endstate)
{
    if (bind(d)) return -1;
    const int use_argmin = endstate < 0;
    if (use_argmin) {
        CU(launch_argmin(d->metrics[d->h_ctl->cur], d->d_key, d->stream));
        d->launches++;
    }
    const long long dp = d->h_ctl->T % d->len;
    CU(launch_walk(trace_args(d), dp, delay, (uint32_t)endstate, use_argmin, d->d_key, d->d_result, d->stream));
    d->launches++;
    CU(cudaMemcpyAsync(d->h_result, d->d_result, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, d->stream));
    CU(cudaStreamSynchronize(d->stream));
    return 0;
}

// decodebit with a fixed end state: same result as walk(), but the walk stops where it rejoins the previous call's
// path (cached on the device).  Needs the `delay` newest rows in the ring, i.e. delay < len (vdecode.c:94 allocates delay + 1).
static int walk_incremental(Decoder *d, int delay, uint32_t endstate)
{
    if (bind(d)) return -1;
    if (!d->walk_cache) {
        CU(cudaMalloc(&d->walk_cache, (size_t)d->len * sizeof(uint32_t)));
        CU(cudaMalloc(&d->d_walk_steps, sizeof(unsigned)));
        CU(cudaMemsetAsync(d->d_walk_steps, 0, sizeof(unsigned), d->stream));
    }
    const long long T = d->h_ctl->T;
    const bool inc = d->cache_valid && d->cache_delay == delay && d->cache_end == endstate && T > d->cache_T && T - d->cache_T < delay;
    const long long prev_T = inc ? d->cache_T : T - 4ll * d->len - 4ll * delay;      // no overlap: a full walk that fills the cache
    CU(launch_walk_incremental(trace_args(d), T, prev_T, delay, endstate, d->walk_cache, d->d_result, d->d_walk_steps, d->stream));
    d->launches++;
    CU(cudaMemcpyAsync(d->h_result, d->d_result, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, d->stream));
    CU(cudaStreamSynchronize(d->stream));
    d->cache_valid = 1; d->cache_T = T; d->cache_delay = delay; d->cache_end = endstate;
    return 0;
}

int decodebit_viterbi224(void *p, int delay, int endstate)
{
    Decoder *d = as_dec(p);
    if (!d) return -1;
    if (delay <= 0) return -1;
    if (endstate >= 0 && delay < d->len && !d->no_walk_cache) {
        const uint32_t es = (uint32_t)endstate & STATEMASK;
        // the stage that just ran carried exactly this walk (asked for after the previous stage): the answer is in the mailbox
        if (d->spec_valid && d->spec_T == d->h_ctl->T && d->spec_delay == delay && d->spec_end == es && d->spec_bit >= -1)
            return (int)d->spec_bit;
        d->spec_want = 1; d->spec_delay = delay; d->spec_end = es;      // the next update(1) brings the walk along
        if (walk_incremental(d, delay, es)) return -1;
    } else if (walk(d, delay, endstate)) return -1;
    return (int)(long long)d->h_result[0];
}

unsigned long long decodeword_viterbi224(void *p, int delay, int endstate)
{
    Decoder *d = as_dec(p);
    if (!d || delay <= 0) return 0;
    if (walk(d, delay, endstate)) return 0;
    return d->h_result[1];
}

static int metric_extreme(void *p, int want_max)
{
    Decoder *d = as_dec(p);
    if (!d) return -1;
    if (bind(d)) return -1;
    CU(launch_minmax(d->metrics[d->h_ctl->cur], d->d_mnmx, d->stream));
    d->launches++;
    unsigned mnmx[2];
    CU(cudaMemcpyAsync(mnmx, d->d_mnmx, sizeof mnmx, cudaMemcpyDeviceToHost, d->stream));
    CU(cudaStreamSynchronize(d->stream));
    const long long r = (long long)mnmx[want_max] - d->h_ctl->sub + d->h_ctl->O;    // reference-domain metric
    return (int)(r + d->h_ctl->renormals);                                           // viterbi224_sse2.c:95,108
}
int max_metric_viterbi224(void *p) { return metric_extreme(p, 1); }
int min_metric_viterbi224(void *p) { return metric_extreme(p, 0); }

void delete_viterbi224(void *p)
{
    Decoder *d = as_dec(p);
    if (!d) return;
    // the next create_viterbi224 of the same size gets this decoder back (decode.c:216-229 deletes and creates per frame)
    for (int i = 0; i < NAUX; i++) if (d->aux[i]) { destroy(d->aux[i]); d->aux[i] = nullptr; }
    Decoder *evict = nullptr;
    bool pooled = false;
    if (d->ring_bytes <= POOL_MAX_BYTES / 2 && cudaSetDevice(d->dev) == cudaSuccess && cudaStreamSynchronize(d->stream) == cudaSuccess) {
        d->magic = 0;                                    // a stale handle is rejected while the decoder waits in the pool
        std::lock_guard<std::mutex> lk(g_dec_mu);
        g_dec_pool.push_back(d);
        pooled = true;
        if (g_dec_pool.size() > DEC_POOL_MAX) { evict = g_dec_pool.front(); g_dec_pool.erase(g_dec_pool.begin()); }
    }
    cudaGetLastError();
    if (evict) { evict->magic = MAGIC; destroy(evict); }
    if (!pooled) destroy(d);
}

// ==========================================================================================
// extensions
// ==========================================================================================
int v224x_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
int v224x_set_device(int dev)
{
    int n = v224x_device_count();
    if (dev < 0 || dev >= n) { set_err("v224x_set_device(%d): %d devices visible", dev, n); return -1; }
    g_device = dev;
    return 0;
}
const char *v224x_last_error(void) { return g_err; }
const char *v224x_version(void) { return "viterbi224_b200 0.1 (sm_100a)"; }

int v224x_init_uniform(void *p, int bias, int start_state)
{
    Decoder *d = as_dec(p);
    if (!d) return -1;
    if (bias < 0 || bias > 20000) { set_err("bias out of range"); return -1; }
    return do_init(d, bias, start_state < 0 ? -1 : (int)((uint32_t)start_state & STATEMASK));
}

int v224x_update_dev(void *p, const unsigned char *dev_syms, int nbits)
{
    Decoder *d = as_dec(p);
    if (!d) return -1;
    if (bind(d)) return -1;
    return update_core(d, dev_syms, nbits);
}

// Advance several decoders (same device) by nbits stages each in lockstep: one persistent launch per batch over
// all of them.  Returns 0 or -1; per-decoder renormalisation counts through renorms_out (may be NULL).
int v224x_update_multi_dev(void **handles, const unsigned char *const *dev_syms, int nctx, int nbits, int *renorms_out)
{
    if (nctx < 1 || nctx > MAX_CTX || !handles || !dev_syms) { set_err("v224x_update_multi_dev: 1..%d decoders", MAX_CTX); return -1; }
    Decoder *ds[MAX_CTX];
    for (int s = 0; s < nctx; s++) {
        ds[s] = as_dec(handles[s]);
        if (!ds[s]) return -1;
        if (ds[s]->dev != ds[0]->dev) { set_err("decoders must live on one device"); return -1; }
        for (int t = 0; t < s; t++) if (ds[t] == ds[s]) { set_err("the same decoder was passed twice"); return -1; }
    }
    if (nbits <= 0) return 0;
    if (bind(ds[0])) return -1;
    for (int s = 1; s < nctx; s++) CU(cudaStreamSynchronize(ds[s]->stream));
    return multi_update_core(ds, dev_syms, nctx, nbits, renorms_out);
}

int v224x_stream_decode_dev(void *p, const unsigned char *dev_syms, int nbits, int delay, unsigned char *dev_bits_out)
{
    Decoder *d = as_dec(p);
    if (!d) return -1;
    if (nbits <= 0) return 0;
    if (bind(d)) return -1;
    return stream_core(d, dev_syms, nbits, delay, dev_bits_out);
}

int v224x_stream_decode(void *p, const unsigned char *syms, int nbits, int delay, unsigned char *bits_out)
{
    Decoder *d = as_dec(p);
    if (!d) return -1;
    if (nbits <= 0) return 0;
    if (bind(d)) return -1;
    if (grow((void **)&d->dsyms, &d->dsyms_cap, 2 * (size_t)nbits)) return -1;
    if (grow((void **)&d->dout, &d->dout_cap, (size_t)nbits)) return -1;
    CU(cudaMemcpyAsync(d->dsyms, syms, 2 * (size_t)nbits, cudaMemcpyHostToDevice, d->stream));
    const int r = stream_core(d, d->dsyms, nbits, delay, d->dout);
    if (r < 0) return -1;
    CU(cudaMemcpyAsync(bits_out, d->dout, (size_t)nbits, cudaMemcpyDeviceToHost, d->stream));
    CU(cudaStreamSynchronize(d->stream));
    return r;
}

int v224x_stream_decode_seg_dev(void *p, const unsigned char *dev_syms, int nbits, int delay, unsigned char *dev_bits_out, int nseg, int conv,
                                v224x_seg_report *rep)
{
    Decoder *d = as_dec(p);
    if (!d) return -1;
    if (nbits <= 0) return 0;
    if (bind(d)) return -1;
    return seg_core(d, dev_syms, nbits, delay, dev_bits_out, nseg, conv, rep) < 0 ? -1 : 0;
}

int v224x_stream_decode_seg(void *p, const unsigned char *syms, int nbits, int delay, unsigned char *bits_out, int nseg, int conv,
                            v224x_seg_report *rep)
{
    Decoder *d = as_dec(p);
    if (!d) return -1;
    if (nbits <= 0) return 0;
    if (bind(d)) return -1;
    if (grow((void **)&d->dsyms, &d->dsyms_cap, 2 * (size_t)nbits)) return -1;
    if (grow((void **)&d->dout, &d->dout_cap, (size_t)nbits)) return -1;
    CU(cudaMemcpyAsync(d->dsyms, syms, 2 * (size_t)nbits, cudaMemcpyHostToDevice, d->stream));
    uint8_t *dsyms = d->dsyms, *dout = d->dout;          // the handle's body may be exchanged with the last segment's decoder
    if (seg_core(d, dsyms, nbits, delay, dout, nseg, conv, rep) < 0) return -1;
    CU(cudaMemcpyAsync(bits_out, dout, (size_t)nbits, cudaMemcpyDeviceToHost, d->stream));
    CU(cudaStreamSynchronize(d->stream));
    return 0;
}

int v224x_range_decode_dev(void *p, const unsigned char *dev_syms, int lead, int nout, int delay, unsigned char *dev_bits_out, int nseg,
                           int conv, void *snap_early_dev, void *snap_late_dev, v224x_seg_report *rep)
{
    Decoder *d = as_dec(p);
    if (!d) return -1;
    if (bind(d)) return -1;
    return range_core(d, dev_syms, lead, nout, delay, dev_bits_out, nseg, conv, static_cast<uint16_t *>(snap_early_dev),
                      static_cast<uint16_t *>(snap_late_dev), rep);
}

int v224x_range_decode(void *p, const unsigned char *syms, int lead, int nout, int delay, unsigned char *bits_out, int nseg, int conv,
                       void *snap_early_dev, void *snap_late_dev, v224x_seg_report *rep)
{
    Decoder *d = as_dec(p);
    if (!d) return -1;
    if (lead < 0 || nout < 0) { set_err("range decode: negative length"); return -1; }
    if (bind(d)) return -1;
    const size_t n = (size_t)lead + (size_t)nout;
    if (grow((void **)&d->dsyms, &d->dsyms_cap, 2 * n)) return -1;
    if (grow((void **)&d->dout, &d->dout_cap, (size_t)nout)) return -1;
    CU(cudaMemcpyAsync(d->dsyms, syms, 2 * n, cudaMemcpyHostToDevice, d->stream));
    uint8_t *dsyms = d->dsyms, *dout = d->dout;          // the handle's body may be exchanged with the last segment's decoder
    if (range_core(d, dsyms, lead, nout, delay, dout, nseg, conv, static_cast<uint16_t *>(snap_early_dev), static_cast<uint16_t *>(snap_late_dev), rep))
        return -1;
    CU(cudaMemcpyAsync(bits_out, dout, (size_t)nout, cudaMemcpyDeviceToHost, d->stream));
    CU(cudaStreamSynchronize(d->stream));
    return 0;
}

int v224x_metric_spread_dev(void *p, const void *dev_a, const void *dev_b, int *spread_out)
{
    Decoder *d = as_dec(p);
    if (!d || !dev_a || !dev_b || !spread_out) return -1;
    if (bind(d)) return -1;
    return spread_core(d, static_cast<const uint16_t *>(dev_a), static_cast<const uint16_t *>(dev_b), spread_out);
}

size_t v224x_snapshot_bytes(void) { return METRICBYTES; }

void v224x_multi_delete(v224x_multi *m)
{
    if (!m) return;
    for (int k = 0; k < m->n; k++) {
        if (cudaSetDevice(m->devs[k]) != cudaSuccess) { cudaGetLastError(); continue; }
        if (m->dec[k]) { m->dec[k]->magic = MAGIC; destroy(m->dec[k]); }
        cudaFree(m->snap_early[k]); cudaFree(m->snap_late[k]); cudaFree(m->snap_peer[k]);
        cudaFree(m->dsyms[k]); cudaFree(m->dbits[k]);
    }
    cudaGetLastError();
    free(m);
}

v224x_multi *v224x_multi_create(const int *devices, int ngpu, int ring_rows)
{
    const int ndev = v224x_device_count();
    if (ngpu < 1 || ngpu > MULTI_MAX || ndev < 1 || (!devices && ngpu > ndev)) { set_err("v224x_multi_create: %d GPUs asked for, %d visible (1..%d supported)", ngpu, ndev, MULTI_MAX); return nullptr; }
    v224x_multi *m = static_cast<v224x_multi *>(calloc(1, sizeof(v224x_multi)));
    if (!m) return nullptr;
    m->n = ngpu;
    m->ring_rows = ring_rows;
    const int saved = g_device;
    for (int k = 0; k < ngpu; k++) {
        m->devs[k] = devices ? devices[k] : k;
        // (a device may be listed more than once: its ranges then share the GPU -- how the one-GPU test tier runs this code)
        bool ok = m->devs[k] >= 0 && m->devs[k] < ndev;
        if (!ok) { set_err("v224x_multi_create: bad device %d", m->devs[k]); g_device = saved; v224x_multi_delete(m); return nullptr; }
        g_device = m->devs[k];
        m->dec[k] = static_cast<Decoder *>(create_viterbi224(ring_rows));
        ok = m->dec[k] != nullptr;
        ok = ok && cudaMalloc(&m->snap_early[k], METRICBYTES) == cudaSuccess && cudaMalloc(&m->snap_late[k], METRICBYTES) == cudaSuccess &&
             cudaMalloc(&m->snap_peer[k], METRICBYTES) == cudaSuccess;
        if (!ok) {
            if (m->dec[k]) set_err("v224x_multi_create: snapshot buffers on device %d: %s", m->devs[k], cudaGetErrorString(cudaGetLastError()));
            g_device = saved;
            v224x_multi_delete(m);
            return nullptr;
        }
    }
    g_device = saved;
    // peer access makes the snapshot copies direct NVLink transfers (without it the runtime stages them through the host)
    for (int a = 0; a < ngpu; a++)
        for (int b = 0; b < ngpu; b++) {
            int can = 0;
            if (a == b || cudaDeviceCanAccessPeer(&can, m->devs[a], m->devs[b]) != cudaSuccess || !can) continue;
            if (cudaSetDevice(m->devs[a]) == cudaSuccess) cudaDeviceEnablePeerAccess(m->devs[b], 0);
            cudaGetLastError();                  // "already enabled" is fine
        }
    m->head = 0;
    return m;
}

int v224x_multi_init(v224x_multi *m, int starting_state)
{
    if (!m) return -1;
    m->head = 0;
    return do_init(m->dec[0], INIT_BIAS, (int)((uint32_t)starting_state & STATEMASK));      // init_viterbi224 for the stream
}

int v224x_multi_stream_decode(v224x_multi *m, const unsigned char *syms, long long nbits, int delay, unsigned char *bits_out, int nseg,
                              int conv, v224x_multi_report *rep)
{
    if (!m || !syms || !bits_out) { set_err("v224x_multi_stream_decode: NULL argument"); return -1; }
    if (nbits <= 0) { if (rep) memset(rep, 0, sizeof *rep); return 0; }
    return multi_core(m, syms, nbits, delay, bits_out, nseg <= 0 ? 4 : nseg, conv, rep);
}

void v224x_trim(void)
{
    release_parked_decoders();
    std::vector<PoolEntry> drop;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        drop.swap(g_pool);
    }
    for (auto &e : drop) { cudaSetDevice(e.dev); cudaFree(e.ptr); }
    cudaGetLastError();
}

int v224x_decode_frames(void *p, const unsigned char *syms, int nframes, int framebits, const unsigned int *start_states,
                        const unsigned int *end_states, unsigned char *data_out, int nlock)
{
    Decoder *d = as_dec(p);
    if (!d) return -1;
    if (nframes <= 0) return 0;
    if (bind(d)) return -1;
    return frames_core(d, syms, nframes, framebits, start_states, end_states, data_out, nlock <= 0 ? 4 : nlock);
}

void *v224x_dev_alloc(void *p, size_t bytes)
{
    Decoder *d = as_dec(p);
    if (!d || bind(d)) return nullptr;
    void *q = nullptr;
    if (cudaMalloc(&q, bytes) != cudaSuccess) { set_err("cudaMalloc(%zu) failed", bytes); cudaGetLastError(); return nullptr; }
    return q;
}
void v224x_dev_free(void *p, void *dev_ptr)
{
    Decoder *d = as_dec(p);
    if (!d || bind(d)) return;
    cudaFree(dev_ptr);
}
int v224x_h2d(void *p, void *dev_dst, const void *host_src, size_t bytes)
{
    Decoder *d = as_dec(p);
    if (!d || bind(d)) return -1;
    CU(cudaMemcpyAsync(dev_dst, host_src, bytes, cudaMemcpyHostToDevice, d->stream));
    CU(cudaStreamSynchronize(d->stream));
    return 0;
}
int v224x_d2h(void *p, void *host_dst, const void *dev_src, size_t bytes)
{
    Decoder *d = as_dec(p);
    if (!d || bind(d)) return -1;
    CU(cudaMemcpyAsync(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost, d->stream));
    CU(cudaStreamSynchronize(d->stream));
    return 0;
}
void *v224x_host_alloc_pinned(size_t bytes)
{
    void *q = nullptr;
    if (cudaMallocHost(&q, bytes) != cudaSuccess) { set_err("cudaMallocHost(%zu) failed", bytes); cudaGetLastError(); return nullptr; }
    return q;
}
void v224x_host_free_pinned(void *host_ptr) { if (host_ptr) cudaFreeHost(host_ptr); }

int v224x_timer_start(void *p)
{
    Decoder *d = as_dec(p);
    if (!d || bind(d)) return -1;
    CU(cudaStreamSynchronize(d->stream));
    CU(cudaEventRecord(d->ev0, d->stream));
    return 0;
}
float v224x_timer_stop_ms(void *p)
{
    Decoder *d = as_dec(p);
    if (!d || bind(d)) return -1.f;
    if (cudaEventRecord(d->ev1, d->stream) != cudaSuccess || cudaEventSynchronize(d->ev1) != cudaSuccess) { cudaGetLastError(); return -1.f; }
    float ms = -1.f;
    if (cudaEventElapsedTime(&ms, d->ev0, d->ev1) != cudaSuccess) { cudaGetLastError(); return -1.f; }   // do not leave the error for the next launch
    return ms;
}
int v224x_kernel_time_reset(void *p)
{
    Decoder *d = as_dec(p);
    if (!d) return -1;
    d->acs_ms = 0; d->acs_launches_timed = 0; d->acs_passes_timed = 0;
    return 0;
}
int v224x_kernel_time_enable(void *p, int on)
{
    Decoder *d = as_dec(p);
    if (!d) return -1;
    d->time_kernels = on;
    return 0;
}
float v224x_kernel_time_ms(void *p, unsigned long long *n_acs_launches)
{
    Decoder *d = as_dec(p);
    if (!d) return -1.f;
    if (n_acs_launches) *n_acs_launches = d->acs_launches_timed;
    return (float)d->acs_ms;
}
unsigned long long v224x_kernel_time_passes(void *p)
{
    Decoder *d = as_dec(p);
    return d ? d->acs_passes_timed : 0;
}

int v224x_get_stats(void *p, v224x_stats *out)
{
    Decoder *d = as_dec(p);
    if (!d || !out) return -1;
    if (bind(d)) return -1;
    if (sync_ctl(d)) return -1;
    unsigned redo = 0;
    CU(cudaMemcpy(&redo, d->d_redo, sizeof redo, cudaMemcpyDeviceToHost));
    out->launches = d->launches;
    out->fused_passes = d->h_ctl->n_fused;
    out->careful_passes = d->h_ctl->n_careful;
    out->single_stages = d->h_ctl->n_single;
    out->sat_stages = d->h_ctl->n_sat;
    out->invalidated_passes = d->h_ctl->n_invalidated;
    out->chainback_redo = redo;
    out->renormals = d->h_ctl->renormals;
    out->stages = d->h_ctl->T;
    unsigned steps = 0;
    if (d->d_walk_steps) CU(cudaMemcpy(&steps, d->d_walk_steps, sizeof steps, cudaMemcpyDeviceToHost));
    out->walk_steps = steps;
    return 0;
}

int v224x_get_metrics(void *p, int16_t *host_out)
{
    Decoder *d = as_dec(p);
    if (!d || bind(d)) return -1;
    int16_t *tmp = nullptr;
    CU(cudaMalloc(&tmp, METRICBYTES));
    int rc = 0;
    do {
        if (cudaMemsetAsync(d->d_flag, 0, sizeof(int), d->stream) != cudaSuccess) { rc = -1; break; }
        if (launch_export_metrics(d->metrics[d->h_ctl->cur], d->ctl, tmp, d->d_flag, d->stream) != cudaSuccess) { rc = -1; break; }
        d->launches++;
        int flag = 0;
        if (cudaMemcpyAsync(host_out, tmp, METRICBYTES, cudaMemcpyDeviceToHost, d->stream) != cudaSuccess) { rc = -1; break; }
        if (cudaMemcpyAsync(&flag, d->d_flag, sizeof(int), cudaMemcpyDeviceToHost, d->stream) != cudaSuccess) { rc = -1; break; }
        if (cudaStreamSynchronize(d->stream) != cudaSuccess) { rc = -1; break; }
        if (flag) { set_err("metric outside the reference's int16 range"); rc = -1; }
    } while (0);
    cudaFree(tmp);
    if (rc) cudaGetLastError();
    return rc;
}

int v224x_set_state(void *p, const int16_t *host_metrics, long long renormals, long long stages)
{
    Decoder *d = as_dec(p);
    if (!d || bind(d)) return -1;
    d->cache_valid = 0;
    d->spec_valid = 0;
    d->ring_dirty_all = 1;
    int16_t *tmp = nullptr;
    CU(cudaMalloc(&tmp, METRICBYTES));
    int rc = 0;
    do {
        if (cudaMemcpyAsync(tmp, host_metrics, METRICBYTES, cudaMemcpyHostToDevice, d->stream) != cudaSuccess) { rc = -1; break; }
        if (launch_import_metrics(d->metrics[d->h_ctl->cur], tmp, d->ctl, d->d_mnmx, renormals, stages, d->stream) != cudaSuccess) { rc = -1; break; }
        d->launches += 3;
        if (cudaStreamSynchronize(d->stream) != cudaSuccess) { rc = -1; break; }
    } while (0);
    cudaFree(tmp);
    if (rc) { set_err("v224x_set_state failed: %s", cudaGetErrorString(cudaGetLastError())); return -1; }
    return sync_ctl(d);
}

int v224x_get_row(void *p, int row, uint32_t *host_out)
{
    Decoder *d = as_dec(p);
    if (!d || bind(d)) return -1;
    if (row < 0 || row >= d->len) { set_err("row out of range"); return -1; }
    uint32_t *tmp = nullptr;
    CU(cudaMalloc(&tmp, ROWBYTES));
    int rc = 0;
    if (launch_export_row(trace_args(d), row, tmp, d->stream) != cudaSuccess) rc = -1;
    d->launches++;
    if (!rc && cudaMemcpyAsync(host_out, tmp, ROWBYTES, cudaMemcpyDeviceToHost, d->stream) != cudaSuccess) rc = -1;
    if (!rc && cudaStreamSynchronize(d->stream) != cudaSuccess) rc = -1;
    cudaFree(tmp);
    if (rc) { set_err("v224x_get_row failed: %s", cudaGetErrorString(cudaGetLastError())); return -1; }
    return 0;
}

int v224x_set_option(void *p, const char *key, long long value)
{
    Decoder *d = as_dec(p);
    if (!d || !key) return -1;
    if (!strcmp(key, "force_single")) d->force_single = (int)value;
    else if (!strcmp(key, "force_sat")) d->force_sat = (int)value;
    else if (!strcmp(key, "force_careful")) d->force_careful = (int)value;
    else if (!strcmp(key, "per_pass_launch")) d->per_pass_launch = (int)value;
    else if (!strcmp(key, "no_walk_cache")) d->no_walk_cache = (int)value;
    else if (!strcmp(key, "grid_limit")) d->grid_limit = (int)std::max(0ll, value);
    else if (!strcmp(key, "no_mailbox")) d->no_mailbox = (int)value;
    else if (!strcmp(key, "tile32")) d->tile32 = (int)value;
    else if (!strcmp(key, "slow_single")) d->slow_single = (int)value;
    else if (!strcmp(key, "no_discard")) d->no_discard = (int)value;
    else if (!strcmp(key, "measure_all")) d->measure_all = (int)value;
    else if (!strcmp(key, "chain_seg")) d->chain_seg = (int)std::max(8ll, value);
    else if (!strcmp(key, "chain_warm")) d->chain_warm = (int)std::max(0ll, value);
    else { set_err("unknown option %s", key); return -1; }
    return 0;
}

} // extern "C"
