#!/bin/bash
# Drop-in proof: compile the reference's own callers UNCHANGED from the read-only checkout and
# link them against libviterbi224_b200.so in place of viterbi224_sse2.o (reference Makefile:74,43,68).
# Only runs where /root/reference exists; outputs go to oracle/_ref (git-ignored, shipped as binaries).
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF=${REF:-/root/reference}
OUT="$ROOT/oracle/_ref"
LIBDIR="$ROOT/isee3-decoder_b200"
mkdir -p "$OUT"
CFLAGS="-O3 -march=x86-64-v3 -mtune=generic -g -Wall -I$REF"
LINK="-L$LIBDIR -lviterbi224_b200 -Wl,-rpath,\$ORIGIN/../../isee3-decoder_b200 -lm"
build() { # name, sources...
  local name=$1; shift
  if [ ! -e "$OUT/$name" ] || [ "$LIBDIR/libviterbi224_b200.so" -nt "$OUT/$name" ]; then
    gcc $CFLAGS -o "$OUT/$name" "$@" $LINK
  fi
}
build vtest224_b200   $REF/vtest224.c $REF/encode.c $REF/sim.c
build vdecode_b200    $REF/vdecode.c $REF/timeformat.c
build hybridtest_b200 $REF/hybridtest.c $REF/encode.c $REF/fano.c $REF/metrics.c $REF/sim.c
build decode_b200     $REF/decode.c $REF/timeformat.c $REF/metrics.c $REF/fano.c
