#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/ab_variants.py default nq2c3 nq4c4 nq4c3 > gpurun_out/ab_variants.log 2>&1
cat gpurun_out/ab_variants.log
