#!/bin/bash
# GPU test suite, bench line, and an ncu capture of the hot kernel
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1
echo "pytest rc=$?"
tail -30 gpurun_out/gpu_tests.log
timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err
echo "bench rc=$?"
tail -3 gpurun_out/bench_r2_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['roofline']['frac'], d['roofline']['kernel_ms'], d.get('cpu_baseline'))
PY
