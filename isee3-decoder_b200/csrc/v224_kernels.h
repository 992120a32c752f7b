// v224_kernels.h -- launch interface between the host runtime and the CUDA kernels.
#pragma once
#include "v224_common.cuh"

namespace V224_NS {

struct PersistArgs {
    Ctl *ctl;
    uint16_t *metrics[NBUF];
    uint32_t *ring;
    uint8_t *row_fmt;
    const uint8_t *syms;     // symbols of the running update call; this launch starts at stage pos0
    const void *tmaps;       // device array of NBUF tensor maps (128 bytes each): metrics[i] as a [256][32768] uint16 tensor, box 256 x 64
    uint32_t *passtab;       // npasses x PASSTAB_WORDS words (operand table + ring rows), filled by k_build_passtab at launch
    int len;
    int pos0;                // stages of this call already done when the launch starts
    int cur0;                // metric buffer holding the launch's input
    long long T0;            // stages since init when the launch starts
    int npasses;             // 8-stage passes to run
    int force_careful;
};

// 1..MAX_CTX independent decoders advanced in lockstep by ONE persistent launch: while the tiles of one decoder's
// pass drain, the CTAs already work on another decoder's pass, so nobody waits at a pass boundary.
constexpr int MAX_CTX = 4;
struct MultiArgs {
    int nctx;
    int npasses;             // per context
    int measure_all;         // every pass reduces the min / max of its output (option "measure_all")
    int no_discard;          // never drop consumed metric lines from the L2 (option "no_discard")
    int grid_limit;          // 0: one CTA per resident slot (148 SMs x CTAs per SM); > 0: at most this many CTAs; < 0: -grid_limit CTAs per SM
    PersistArgs ctx[MAX_CTX];
};

// Mapped pinned host memory the one-stage kernel reports into (per-bit ABI pattern, vdecode.c:145-152): the control-block
// head the host mirrors, the result of the speculative decodebit walk, and -- written last, after a system-wide fence --
// the sequence number of the launch.  The host polls `seq` instead of paying a copy and a stream synchronisation per call.
struct Mailbox {
    unsigned char ctl_head[256];         // the first CTL_HOST_BYTES of the control block after the stage
    long long walk_bit;                  // decodebit(delay, endstate) at the new ring head (-2: no walk was asked for)
    int declined;                        // the stage did not run (stale stage counter / saturation watch / error)
    int pad;
    volatile unsigned long long seq;
};
static_assert(CTL_HOST_BYTES <= 256, "mailbox control-block copy");

struct SingleArgs {
    Ctl *ctl;
    uint16_t *metrics[NBUF];
    uint32_t *ring;
    uint8_t *row_fmt;
    const uint8_t *syms;
    int len;
    int expected_pos;        // index of this stage's symbol pair in syms
    long long expected_T;    // the control block's stage counter this launch was issued for (else it declines)
    int use_arg_syms;        // per-bit streaming: the two symbols travel as kernel arguments
    int sym0, sym1;
    // per-bit streaming: report through the mailbox (NULL: no report), optionally with the decodebit walk the caller is about to ask for
    Mailbox *mailbox;
    unsigned long long seq;
    int spec_walk, spec_delay;
    uint32_t spec_end;
    long long spec_prev_T;
    uint32_t *walk_cache;
    unsigned *walk_steps;
    int all_canon;           // every row of the ring is in the canonical layout (the walk need not read the row tags)
    int slow_form;           // test knob: the one-thread-per-8-butterflies scalar form of the stage instead of the packed one
};

struct TraceArgs {
    const uint32_t *ring;
    const uint8_t *row_fmt;
    int len;
};

cudaError_t launch_init(uint16_t *m0, Ctl *c, uint32_t start_state, int bias, int start_value, cudaStream_t st);
cudaError_t launch_persist(const MultiArgs &m, cudaStream_t st);
size_t passtab_bytes(int npasses);
// tensor maps of a decoder's NBUF metric buffers ([256 rows][32768 columns] of uint16, box = one tile) into dev_out (NBUF x 128 bytes)
cudaError_t build_metric_tensor_maps(uint16_t *const *metrics, void *dev_out, cudaStream_t st, const char **why);
constexpr size_t TMAP_BYTES = 128;
cudaError_t launch_single(const SingleArgs &a, bool sat, cudaStream_t st);
cudaError_t launch_chainback(const TraceArgs &a, uint32_t nbits, uint32_t endstate, int L, int warm, uint8_t *out, uint32_t *seg_guess,
                             uint32_t *seg_final, unsigned *redo_count, cudaStream_t st);
cudaError_t launch_walk(const TraceArgs &a, long long dp, int delay, uint32_t endstate, int use_argmin, const unsigned long long *argmin_key,
                        unsigned long long *result, cudaStream_t st);
cudaError_t launch_walk_incremental(const TraceArgs &a, long long T, long long prev_T, int delay, uint32_t endstate, uint32_t *cache,
                                    unsigned long long *result, unsigned *steps_out, cudaStream_t st);
cudaError_t launch_stream_trace(const TraceArgs &a, long long T_first, int nout, int delay, uint8_t *bits_out, cudaStream_t st);
// the streaming tracebacks of several lockstep decoders in one launch
struct StreamTraceJob { TraceArgs a; long long T_first; int nout; uint8_t *bits_out; };
struct StreamTraceMulti { int njobs, delay; StreamTraceJob job[MAX_CTX]; };
cudaError_t launch_stream_trace_multi(const StreamTraceMulti &m, cudaStream_t st);
cudaError_t launch_argmin(const uint16_t *m, unsigned long long *key, cudaStream_t st);
cudaError_t launch_minmax(const uint16_t *m, unsigned *mnmx, cudaStream_t st);
cudaError_t launch_metric_diff(const uint16_t *a, const uint16_t *b, int *out2, cudaStream_t st);
cudaError_t launch_export_row(const TraceArgs &a, long long row, uint32_t *out, cudaStream_t st);
cudaError_t launch_export_metrics(const uint16_t *m, const Ctl *c, int16_t *out, int *range_error, cudaStream_t st);
cudaError_t launch_import_metrics(uint16_t *m, const int16_t *in, Ctl *c, unsigned *mnmx, long long renormals, long long T, cudaStream_t st);

} // namespace V224_NS
