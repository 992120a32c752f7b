/*
 * ref_shim_sse2.c -- TEST INFRASTRUCTURE ONLY.
 * Compiles the UNMODIFIED reference SSE2 decoder (included by path from the
 * read-only reference checkout, -I$(REF); nothing is copied into this repo) into one
 * translation unit together with a few accessors, so that tests can see inside
 * the reference's opaque `struct v224` (viterbi224_sse2.c:26-34): metrics,
 * decision rows, renormals, ring position, and can load a mid-stream state
 * (the checkpointed-window verification of SURVEY.md section 8c).
 *
 * Built only where the reference checkout exists (this container); the result
 * lives in oracle/_ref/ (git-ignored, shipped to the GPU box as a binary).
 */
#include "viterbi224_sse2.c"

const int16_t *refshim_metrics(void *p)   { return ((struct v224 *)p)->old_metrics->s; }
int16_t *refshim_metrics_mut(void *p)     { return ((struct v224 *)p)->old_metrics->s; }
const uint32_t *refshim_row(void *p, int row) { return ((struct v224 *)p)->decisions[row].w; }
long long refshim_renormals(void *p)      { return ((struct v224 *)p)->renormals; }
void refshim_set_renormals(void *p, long long r) { ((struct v224 *)p)->renormals = r; }
int refshim_dp(void *p)                   { struct v224 *vp = p; return (int)(vp->dp - vp->decisions); }
void refshim_set_dp(void *p, int row)     { struct v224 *vp = p; vp->dp = &vp->decisions[row % vp->len]; }
int refshim_len(void *p)                  { return ((struct v224 *)p)->len; }
/* The reference mallocs the ring without clearing it; tests that read rows
 * before they are written want a defined value. */
void refshim_zero_ring(void *p)           { struct v224 *vp = p; memset(vp->decisions, 0, (size_t)vp->len * sizeof(decision_t)); }
