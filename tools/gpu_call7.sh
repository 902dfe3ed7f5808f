#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_paraxial.py tests/test_abi.py -m gpu -q -s -p no:cacheprovider > gpurun_out/gpu_paraxial_tests.log 2>&1
echo "pytest rc=$?"; grep -v Warning gpurun_out/gpu_paraxial_tests.log | tail -40
timeout 600 python tools/profile_optical_loss.py GAGAGA 64 1024 4096 > gpurun_out/optical_loss_profile.json 2> gpurun_out/optical_loss_profile.err
cat gpurun_out/optical_loss_profile.json; tail -5 gpurun_out/optical_loss_profile.err
