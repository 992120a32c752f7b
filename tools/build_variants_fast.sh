#!/bin/bash
# A/B builds that differ only in the fused pass (v224_acs_persist.cu, both tile shapes): the two objects are recompiled with the
# variant's flags and linked with the in-tree objects of the other translation units (python isee3-decoder_b200/build.py first).
# usage: tools/build_variants_fast.sh name1:"-DFLAG ..." name2:"..."   (four variants at a time)
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/_bin
one() {
  set -e
  C=isee3-decoder_b200/csrc
  spec="$1"; name="${spec%%:*}"; flags="${spec#*:}"; [ "$flags" = "$spec" ] && flags=""
  F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --cudart static -Xcompiler -fPIC,-O2,-Wall -Xptxas -v -Xptxas -O1"
  o=tools/_bin/_$name
  nvcc $F $flags -c -o $o.a.o $C/v224_acs_persist.cu 2> $o.log
  nvcc $F $flags -DV224_TILE_COLS_LOG2=5 -DV224_NS=v224t32 -DV224_BRIDGE=v224_t32 -c -o $o.b.o $C/v224_acs_persist.cu 2>> $o.log
  nvcc -shared --cudart static -gencode arch=compute_100a,code=sm_100a -o tools/_bin/libv224_$name.so $o.a.o $o.b.o $C/v224_kernels.o $C/v224_runtime.o $C/v224_pairing.o
  echo "$(grep -E "spill" $o.log | sed -n '1p;3p' | sed -E 's/ +/ /g' | tr '\n' '|') <- $name"
  rm -f $o.a.o $o.b.o $o.log
}
export -f one
printf '%s\n' "$@" | xargs -P 4 -I{} bash -c 'one "$@"' _ {}
