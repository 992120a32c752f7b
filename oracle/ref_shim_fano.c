/* TEST INFRASTRUCTURE ONLY.  metrics.c of the reference reads a global `Verbose` that its main programs define
 * (metrics.c:14); this one-line translation unit supplies it so that fano.c + metrics.c link into oracle/_ref/libv224_reffano.so
 * (the unmodified sources, compiled where they lie). */
int Verbose = 0;
