/*
 * viterbi224.h -- the drop-in boundary of libviterbi224_b200.
 *
 * These nine entry points are the libfec-style C ABI of the ISEE-3/ICE K=24, rate-1/2
 * Viterbi decoder.  Their names, argument lists and return conventions are the ones the
 * reference declares in its viterbi224.h:8-16, so that the reference's own callers
 * (vtest224.c, vdecode.c, hybridtest.c, decode.c) compile against this header unchanged and
 * link against libviterbi224_b200.so instead of viterbi224_sse2.o.  Behaviour follows the
 * reference's SSE2 build (viterbi224_sse2.c), bit for bit; every call is synchronous:
 * when it returns, its outputs are in the caller's host buffers.
 *
 * The handle is opaque.  Underneath it: path metrics and the decision ring live in B200
 * HBM, the add-compare-select butterflies and the tracebacks are sm_100a CUDA kernels.
 * There is no CPU fallback: without a usable CUDA device create_viterbi224() returns NULL.
 */
#ifndef VITERBI224_B200_ABI_H
#define VITERBI224_B200_ABI_H

#ifdef __cplusplus
extern "C" {
#endif

/* Allocate a decoder whose decision ring holds `len` trellis stages (1 MiB each, in HBM)
 * and initialise it as init_viterbi224(p, 0) does.  NULL on failure.
 * Replaces viterbi224_sse2.c:56-80. */
void *create_viterbi224(int len);

/* Start a new frame/stream: every path metric = SHRT_MIN+5000, the metric of
 * (starting_state mod 2^23) = SHRT_MIN, ring position and renormalisation total = 0.
 * 0 on success, -1 if p is NULL.  Replaces viterbi224_sse2.c:37-53. */
int init_viterbi224(void *p, int starting_state);

/* Run `nbits` trellis stages over syms[0 .. 2*nbits): 8-bit offset-128 soft symbols, two
 * per data bit (POLY1 symbol first).  `syms` is a caller-owned HOST buffer, consumed before
 * return.  Appends nbits decision rows to the ring (wrapping at len) and returns the number
 * of path-metric renormalisations that happened during this call (the SSE2 build's return
 * value); -1 if p is NULL or a device error occurred.  Replaces viterbi224_sse2.c:259-389. */
int update_viterbi224_blk(void *p, const unsigned char *syms, int nbits);

/* Trace the survivor that ends in (endstate mod 2^23) after bit nbits-1 back to bit 0 and
 * write the decoded bits MSB-first into data[0 .. ceil(nbits/8)) (HOST buffer).  Row n of the
 * walk is ring row n % len, counted from the last init.  0 on success, -1 on NULL handle or
 * device error.  Replaces viterbi224_sse2.c:113-161. */
int chainback_viterbi224(void *p, unsigned char *data, unsigned int nbits, unsigned int endstate);

/* Walk `delay` stages back from the newest ring row, starting in `endstate` (endstate < 0:
 * start from the state with the smallest path metric, lowest index on ties), and return the
 * last decision bit read; -1 if p is NULL or delay <= 0.  Replaces viterbi224_sse2.c:164-203. */
int decodebit_viterbi224(void *p, int delay, int endstate);

/* Same walk, returning the last (up to) 64 decision bits, oldest in bit 63.
 * Replaces viterbi224_sse2.c:206-243. */
unsigned long long decodeword_viterbi224(void *p, int delay, int endstate);

/* Largest / smallest current path metric plus the running renormalisation total,
 * truncated to int; -1 if p is NULL.  Replace viterbi224_sse2.c:82-109. */
int max_metric_viterbi224(void *p);
int min_metric_viterbi224(void *p);

/* Release the decoder (NULL is allowed).  Replaces viterbi224_sse2.c:248-255. */
void delete_viterbi224(void *p);

#ifdef __cplusplus
}
#endif
#endif /* VITERBI224_B200_ABI_H */
