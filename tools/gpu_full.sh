#!/bin/bash
# what the driver runs at round end, on one GPU: the GPU tests, smoke(), the reference arm, the default bench
mkdir -p gpurun_out
t0=$(date +%s)
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1
echo "pytest rc=$? ($(( $(date +%s) - t0 )) s)"; tail -3 gpurun_out/gpu_tests.log
t0=$(date +%s)
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$? ($(( $(date +%s) - t0 )) s)"; tail -2 gpurun_out/smoke.log
t0=$(date +%s)
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo "reference arm rc=$? ($(( $(date +%s) - t0 )) s)"; cut -c1-600 gpurun_out/bench_reference.json
t0=$(date +%s)
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_driver_like.json 2> gpurun_out/bench_driver_like.err
echo "bench --steps 20 rc=$? ($(( $(date +%s) - t0 )) s)"
t0=$(date +%s)
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench default rc=$? ($(( $(date +%s) - t0 )) s)"
python - <<PY
import json
for f in ('bench_driver_like','bench_default'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    print(f, 'value %.1f G  ms %.4f  e2e %.1f G (%.4f ms) kernel_ms %.4f  frac %.4f drop-in %.1f eager %.1f' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['drop_in_api_value']/1e9, d['e2e']['eager_api_value']/1e9), d['roofline']['kernel'], d['clocks'], d['gpu_launches'])
    print('   sustained', d['sustained']['value']/1e9, d['sustained']['clocks'], ' cpu', d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline',{}).get('kind'))
    for r in (d.get('batched_lenses') or {}).get('runs', []):
        print('   batched', r['lenses'], 'eager ms', round(r['ms_per_step'],3), 'graphed ms', r['graphed_ms_per_step'], 'lenses/s', r['graphed_lenses_per_s'])
PY
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_raw.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
