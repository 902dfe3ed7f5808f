#!/bin/bash
# quick check of a kernel change: parity tests of the fused pass, then the default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py tests/test_gpu_random.py -m gpu -q -x -p no:cacheprovider > gpurun_out/gpu_quick_tests.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/gpu_quick_tests.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print('value %.1f G  ms %.4f  e2e %.1f G (%.4f ms) kernel_ms %.4f  frac %.4f drop-in %.1f eager %.1f' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['drop_in_api_value']/1e9, d['e2e']['eager_api_value']/1e9), d['roofline']['kernel'])
print('sustained', d.get('sustained'))
PY
