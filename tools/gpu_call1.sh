#!/bin/bash
# first GPU call of round 2: variants of the fused spot pass side by side, then the GPU test suite
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/call1_smi.txt 2>&1
timeout 600 python tools/rev_variants.py 296 > gpurun_out/rev_variants.json 2> gpurun_out/rev_variants.err
echo "rev_variants rc=$?"
tail -8 gpurun_out/rev_variants.err
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/gpu_tests_call1.log 2>&1
echo "pytest rc=$?"
tail -25 gpurun_out/gpu_tests_call1.log
