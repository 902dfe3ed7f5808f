#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/gpu_tests.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_default.json'))
print('value %.1f G  ms %.4f  e2e %.1f G (%.4f ms) kernel_ms %.4f  frac %.4f drop-in %.1f eager %.1f' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['drop_in_api_value']/1e9, d['e2e']['eager_api_value']/1e9), d['roofline']['kernel'])
PY
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_raw.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python - <<'PY'
import csv,re
rows=list(csv.reader(open('gpurun_out/r2_launches_raw.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
hdr=rows[hi]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
n=0
for r in rows[hi+1:]:
    if len(r)<=vi: continue
    name=re.sub(r'\(.*','',r[ki]).replace('void ','').replace('<unnamed>::','')
    if name.startswith('k_'):
        print(f'{float(r[vi]):9.0f} ns {name}'); n+=1
        if n>8: break
PY
