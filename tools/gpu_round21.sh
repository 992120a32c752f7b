#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 200 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for s in 3 4; do timeout 300 python bench.py --steps 2 --warmup 3 --segments $s --no-cpu-baseline 2>/dev/null | python -c "import json,sys; b=json.loads(sys.stdin.read()); print('discard', b[\"check\"][\"segments_rank0\"][\"segments\"], round(b[\"value\"]), round(b[\"roofline\"][\"mean_pass_us\"],3), b[\"clocks\"][\"sm_mhz\"], b[\"clocks\"][\"power_w_max\"], b[\"check\"][\"bit_errors_vs_transmitted\"], b[\"check\"][\"segmented_output_identical_to_sequential_rank0\"])"; done
timeout 200 python tools/ab_variants.py default nodiscard
