#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/rev_variants.py 296 > gpurun_out/rev_variants.json 2> gpurun_out/rev_variants.err
grep -o '^[a-z0-9]* \|"ms_mean": [0-9.]*\|"ms_min": [0-9.]*' gpurun_out/rev_variants.err | paste - - - 
for v in reg8 tmem12c2 tmem8 tmem12 tmem16; do
  TL_REV=$v timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_$v.json'))
print('$v', 'value %.1f G  ms %.4f  e2e %.1f G  kernel_ms %.4f  frac %.4f' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['roofline']['kernel_ms'], d['roofline']['frac']), d['roofline']['kernel'])
PY
done
