#!/bin/bash
mkdir -p gpurun_out


timeout 300 python tools/ab_variants.py default nslot3 nslot4 > gpurun_out/ab_variants.log 2>&1
cat gpurun_out/ab_variants.log

