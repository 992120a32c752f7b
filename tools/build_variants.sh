#!/bin/bash
# A/B builds of the library with different kernel shapes into tools/_bin/libv224_<name>.so (used by tools/ab_variants.py).
# usage: tools/build_variants.sh name1:"-DFLAG ..." name2:"..."
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/_bin
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  python isee3-decoder_b200/build.py --out tools/_bin/libv224_$name.so $flags > /dev/null
  grep -A2 "k_acs_persist" tools/_bin/libv224_$name.so.ptxas.log | grep -E "spill|Used" | tr '\n' ' '; echo " <- $name"
done
