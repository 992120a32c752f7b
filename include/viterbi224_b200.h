/*
 * viterbi224_b200.h -- extensions of libviterbi224_b200 beyond the reference's nine entry
 * points (include/viterbi224.h).  Plain C ABI: pointers and sizes only.
 *
 * Why they exist: the reference's streaming caller, vdecode.c:142-158, forces one synchronous
 * update(1) + decodebit(delay, 0) pair per decoded bit through the ABI.  That works here as a
 * drop-in, but a GPU needs work in blocks.  v224x_stream_decode() is that loop in block form;
 * its output is byte-for-byte what the per-bit loop returns.  The *_dev variants take device
 * pointers so that a caller (bench.py's `value` leg) can keep the symbol stream resident in HBM.
 */
#ifndef VITERBI224_B200_EXT_H
#define VITERBI224_B200_EXT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- device selection / diagnostics -------------------------------------------------- */
int         v224x_device_count(void);            /* CUDA devices visible, 0 if none              */
int         v224x_set_device(int dev);           /* device used by subsequent create calls       */
const char *v224x_last_error(void);              /* text of the last failure in this thread      */
const char *v224x_version(void);

/* ---- block-mode equivalents of the reference call patterns ---------------------------- */

/* Equivalent to, for i in [0, nbits):
 *     update_viterbi224_blk(p, syms + 2*i, 1);
 *     bits_out[i] = decodebit_viterbi224(p, delay, 0);          (vdecode.c:145,152)
 * with decision rows older than the last init reading as zero (a freshly created reference
 * ring).  bits_out[i] is 0 or 1.  The ring given to create_viterbi224() must hold more than
 * `delay` rows; the call works through the stream in chunks of (len - delay) stages.
 * Returns the number of renormalisations (sum of the per-bit update return values), -1 on error. */
int v224x_stream_decode(void *p, const unsigned char *syms, int nbits, int delay, unsigned char *bits_out);

/* Same, with syms and bits_out in device memory of the decoder's GPU. */
int v224x_stream_decode_dev(void *p, const unsigned char *dev_syms, int nbits, int delay, unsigned char *dev_bits_out);

/* update_viterbi224_blk with the symbols already in device memory. */
int v224x_update_dev(void *p, const unsigned char *dev_syms, int nbits);

/* Advance nctx (1..4) independent decoders of one GPU by nbits stages each, in lockstep: one persistent launch
 * works through all of them, so the GPU does not idle at any decoder's pass boundary.  Each decoder ends in exactly
 * the state nctx separate v224x_update_dev() calls would leave.  renorms_out[i] (optional) = decoder i's return value. */
int v224x_update_multi_dev(void **handles, const unsigned char *const *dev_syms, int nctx, int nbits, int *renorms_out);

/* v224x_stream_decode with the stream cut into nseg (1..4) contiguous segments that nseg decoders on this GPU
 * advance in lockstep (one persistent kernel works through all of them, so no SM idles at a decoder's pass
 * boundary).  The handle continues its current state over the first segment; each further decoder starts
 * delay + conv stages early from uniform metrics (conv < 0: 2048).  bits_out is ALWAYS what v224x_stream_decode
 * returns: every hand-over is checked on the device (the two decoders' path-metric vectors must differ by a
 * constant `delay` stages before the later one's first output -- from there on their decisions are identical),
 * and when a check fails the rest of the stream is decoded again sequentially.  After the call the handle holds
 * the end-of-stream state up to a constant metric offset: further stream/decodebit calls continue exactly, while
 * update return values and min/max metrics are relative to that offset.  Returns 0, -1 on error. */
typedef struct {
    int segments;            /* segments used (1 = stream too short: plain sequential decode)          */
    int warm;                /* warm-up stages of every segment but the first (delay + conv)           */
    int verified;            /* hand-over checks that passed                                           */
    int redone;              /* segments decoded again sequentially because a check failed             */
    long long extra_stages;  /* trellis stages run on top of nbits (warm-ups + redone segments)        */
    int worst_spread;        /* largest (max - min) metric difference seen at a check (0 = converged)  */
} v224x_seg_report;
int v224x_stream_decode_seg(void *p, const unsigned char *syms, int nbits, int delay, unsigned char *bits_out, int nseg, int conv,
                            v224x_seg_report *report);
int v224x_stream_decode_seg_dev(void *p, const unsigned char *dev_syms, int nbits, int delay, unsigned char *dev_bits_out, int nseg,
                                int conv, v224x_seg_report *report);

/* A batch of independent frames, each decoded exactly as the reference's three-call sequence would decode it
 *     init_viterbi224(p, start_states[f]); update_viterbi224_blk(p, syms + 2*framebits*f, framebits);
 *     chainback_viterbi224(p, data_out + ceil(framebits/8)*f, framebits, end_states[f]);
 * (vtest224.c:116-118, hybridtest.c:186-193, decode.c:220-222), with up to nlock (1..4, <= 0: 4) frames side by side in
 * one persistent launch.  start_states / end_states may be NULL (all 0).  Needs framebits <= the handle's len.
 * The tracebacks of a group run concurrently (one stream per decoder) and overlap the next group's ACS passes, which use a
 * second set of decoders; the library owns up to 2 * nlock - 1 extra decoders of the handle's size for this.
 * Afterwards the handle holds the state of the last frame of its lane (as after that frame's chainback).
 * Returns 0, -1 on error. */
int v224x_decode_frames(void *p, const unsigned char *syms, int nframes, int framebits, const unsigned int *start_states,
                        const unsigned int *end_states, unsigned char *data_out, int nlock);

/* init variant for time-segmented decoding: every metric = SHRT_MIN + bias and no state is
 * favoured (start_state < 0), or init_viterbi224 semantics (start_state >= 0). */
int v224x_init_uniform(void *p, int bias, int start_state);

/* ---- host side of the streaming driver (vdecode.c:101-140,186) ---------------------------------------------
 * Which received symbols form the pairs vdecode hands to update_viterbi224_blk: the 34-tap sync correlator over the
 * last 4096 symbols, the once-per-frame comparison of the in-phase and out-of-phase peaks, and the one-symbol slip of
 * a phase flip (the symbol at the decision is dropped, the next one is paired with the stale even-slot symbol).
 * Host arithmetic only (north_star keeps the phase flip on the host); no GPU needed.
 *   soft / nsyms : received soft symbols (symdemod byte format, symdemod.c:240-251)
 *   start_phase  : vdecode -p (vdecode.c:77);   dontflip : vdecode -F (vdecode.c:71)
 *   pairs_out    : 2 bytes per pair, room for nsyms / 2 + 1 pairs
 *   cmp_out      : NULL, or 2 bytes per pair: the hard-sliced symbols vdecode compares the re-encoded pair with
 *                  (vdecode.c:176-177, for decode delay `delay`)
 *   flip_at      : NULL, or room for flip_cap entries: index of the first pair after each phase flip; *nflips = how many
 * Returns the number of pairs, -1 on bad arguments. */
long long v224x_pair_symbols(const unsigned char *soft, long long nsyms, int start_phase, int dontflip, int delay,
                             unsigned char *pairs_out, unsigned char *cmp_out, long long *flip_at, int flip_cap, int *nflips);

/* ---- time segments of one stream on several GPUs (vdecode.c:145-152 per range) --------------------
 * A stream of pairs is cut into contiguous output ranges, one per GPU.  The decoder of a range that does not begin the
 * stream starts `lead` = delay + conv stages early from uniform metrics; it makes the decisions of the sequential
 * decoder from the stage at which the two path-metric vectors differ by a constant.  That is CHECKED: the later range
 * saves its metrics `delay` stages before its first output (every row its tracebacks touch lies after that point), the
 * earlier range saves its metrics at the same stream position, and the two snapshots must differ by a constant
 * (v224x_metric_spread_dev == 0).  If they do not, the earlier range's decoder -- exact at its end -- decodes the later
 * range again (lead = 0).  The stitched output is therefore always what one sequential decoder produces.
 *
 * v224x_range_decode: one range on this handle's GPU.  syms = the pairs of stream stages [first - lead, first + nout).
 *   lead == 0 : the handle continues from its current state (stream start after init_viterbi224, or the previous call)
 *   lead  > 0 : the handle restarts from uniform metrics, the first `lead` stages give no output
 *   bits_out[i] = what decodebit(delay, 0) returns after stage lead + i            (nout entries)
 *   snap_early_dev : NULL or 16 MiB of device memory: metrics after stage lead - delay        (needs lead >= delay)
 *   snap_late_dev  : NULL or 16 MiB of device memory: metrics after stage lead + nout - delay (needs nout >= delay)
 *   nseg, conv, report : as for v224x_stream_decode_seg (the range itself runs as nseg lockstep segments on its GPU)
 * The _dev variant takes syms / bits_out in device memory of the handle's GPU.  Returns 0, -1 on error. */
int v224x_range_decode(void *p, const unsigned char *syms, int lead, int nout, int delay, unsigned char *bits_out, int nseg, int conv,
                       void *snap_early_dev, void *snap_late_dev, v224x_seg_report *report);
int v224x_range_decode_dev(void *p, const unsigned char *dev_syms, int lead, int nout, int delay, unsigned char *dev_bits_out, int nseg,
                           int conv, void *snap_early_dev, void *snap_late_dev, v224x_seg_report *report);
/* max - min over all 2^23 states of (a[s] - b[s]) for two metric snapshots in device memory of the handle's GPU:
 * 0 <=> the vectors differ by a constant <=> the two decoders make identical decisions from there on. */
int v224x_metric_spread_dev(void *p, const void *dev_a, const void *dev_b, int *spread_out);
size_t v224x_snapshot_bytes(void);

/* The same inside one process: one host thread and one CUDA stream per GPU, snapshots moved with peer copies.
 * devices[ngpu] = CUDA device ordinals (NULL: 0 .. ngpu-1; an ordinal listed twice makes two ranges share that GPU); ring_rows = decision-ring rows per decoder (> delay; the
 * stream is worked through in chunks of ring_rows - delay stages).  The context owns one decoder per GPU (plus the
 * lockstep partners of nseg > 1) and keeps the stream's state between calls: v224x_multi_init = init_viterbi224 for the
 * stream, every v224x_multi_stream_decode call continues where the last one ended (block-wise callers such as
 * vdecode_block -G N); the range that begins a block runs on the GPU that holds that state. */
typedef struct v224x_multi v224x_multi;
typedef struct {
    int gpus;                 /* ranges (= GPUs) used for this call (short inputs use fewer)                        */
    int handovers_verified;   /* GPU-to-GPU hand-overs whose snapshot check passed                                 */
    int ranges_redone;        /* ranges decoded again by the previous range's decoder because a check failed       */
    int worst_spread;         /* largest snapshot spread seen at a GPU-to-GPU check (0 = every range had converged) */
    int inner_verified;       /* sums of the per-GPU reports (lockstep segments inside each range)                 */
    int inner_redone;
    long long extra_stages;   /* trellis stages run on top of nbits (warm-ups and redone ranges / segments)        */
    long long residual_diffs; /* output bits that can differ from the sequential decode: 0 by construction         */
} v224x_multi_report;
v224x_multi *v224x_multi_create(const int *devices, int ngpu, int ring_rows);
int  v224x_multi_init(v224x_multi *m, int starting_state);
int  v224x_multi_stream_decode(v224x_multi *m, const unsigned char *syms, long long nbits, int delay, unsigned char *bits_out,
                               int nseg, int conv, v224x_multi_report *report);
void v224x_multi_delete(v224x_multi *m);

/* Give pooled device memory (recycled decoders and rings of deleted handles) back to the driver. */
void v224x_trim(void);

/* ---- device memory helpers for callers without a CUDA runtime of their own ------------ */
void *v224x_dev_alloc(void *p, size_t bytes);
void  v224x_dev_free(void *p, void *dev_ptr);
int   v224x_h2d(void *p, void *dev_dst, const void *host_src, size_t bytes);
int   v224x_d2h(void *p, void *host_dst, const void *dev_src, size_t bytes);
void *v224x_host_alloc_pinned(size_t bytes);
void  v224x_host_free_pinned(void *host_ptr);

/* ---- timing on the decoder's own CUDA stream ------------------------------------------ */
int   v224x_timer_start(void *p);                /* record an event on the decoder's stream      */
float v224x_timer_stop_ms(void *p);              /* record + synchronise, elapsed milliseconds   */
/* Accumulated device time of the ACS pass kernels alone since the last reset (events around
 * every ACS launch batch), and the number of launches. */
int   v224x_kernel_time_reset(void *p);
int   v224x_kernel_time_enable(void *p, int on);
float v224x_kernel_time_ms(void *p, unsigned long long *n_acs_launches);
unsigned long long v224x_kernel_time_passes(void *p);   /* 8-stage passes inside the timed launches */

/* ---- counters --------------------------------------------------------------------------- */
typedef struct {
    unsigned long long launches;       /* kernels launched by this handle since create          */
    unsigned long long fused_passes;   /* 8-stage passes executed                               */
    unsigned long long careful_passes; /* ... of which recorded per-stage minima                */
    unsigned long long single_stages;  /* 1-stage passes (fast arithmetic)                      */
    unsigned long long sat_stages;     /* 1-stage passes (exact saturating arithmetic)          */
    unsigned long long invalidated_passes; /* fused passes discarded by the saturation validation  */
    unsigned long long chainback_redo; /* speculative chainback segments that had to be redone  */
    long long          renormals;      /* the reference's `renormals` accumulator               */
    long long          stages;         /* trellis stages since the last init                    */
    unsigned long long walk_steps;     /* dependent ring loads spent in decodebit(delay, state>=0) walks */
} v224x_stats;
int v224x_get_stats(void *p, v224x_stats *out);

/* ---- test hooks (state inspection; used by tests/ for parity, not by applications) ------ */
/* Path metrics as the reference holds them (int16, reference domain), 2^23 values. */
int v224x_get_metrics(void *p, int16_t *host_out);
/* Load a mid-stream state: metrics (reference domain), renormals, stages since init. */
int v224x_set_state(void *p, const int16_t *host_metrics, long long renormals, long long stages);
/* One decision row (0 <= row < len) in the reference's layout: 2^18 words, bit s of the row =
 * decision of new state s (viterbi224_sse2.c:141). */
int v224x_get_row(void *p, int row, uint32_t *host_out);
/* Knobs: "force_single"=1 never fuse, "force_sat"=1 exact saturating single stages only,
 * "force_careful"=1 always record per-stage minima, "per_pass_launch"=1 one kernel launch per
 * 8-stage pass instead of the persistent multi-pass kernel, "chain_seg"/"chain_warm" chainback
 * segment length / warm-up depth, "no_walk_cache"=1 decodebit always walks all `delay` rows (default: it stops where the
 * walk rejoins the previous call's path, same result). */
int v224x_set_option(void *p, const char *key, long long value);

#ifdef __cplusplus
}
#endif
#endif /* VITERBI224_B200_EXT_H */
