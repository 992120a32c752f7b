#!/bin/bash
# Trace build of the library (-DV224_TRACE: per-tile phase timestamps) into tools/_bin; used by tools/trace_passes.py only.
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/_bin
python isee3-decoder_b200/build.py --out tools/_bin/libv224_trace.so -DV224_TRACE
