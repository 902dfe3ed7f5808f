#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/gpu_tests.log
for v in tmem12c2 tmem16 reg8; do
  TL_REV=$v timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_$v.json'))
print('$v', 'value %.1f G  ms %.4f  e2e %.1f G (%.4f ms) kernel_ms %.4f  frac %.4f drop-in %.1f eager %.1f' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['drop_in_api_value']/1e9, d['e2e']['eager_api_value']/1e9))
PY
done
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_raw.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
grep -E "k_stage_ref|k_spot_rev|k_reduce_owner|k_lens_finalize" gpurun_out/r2_launches_raw.csv | head -8 | awk -F'","' '{print $5, $(NF)}' | cut -c1-120
