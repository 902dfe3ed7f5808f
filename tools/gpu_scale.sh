#!/bin/bash
# scaling evidence on N GPUs of one box: the driver's bench command (20 steps) and a longer one,
# the world-2 peer-exchange test (N >= 2) and, on 8 GPUs, BASELINE.json config 5 at full size
N=$1
mkdir -p gpurun_out
run() {   # steps warmup tag
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps $1 --warmup $2 > gpurun_out/bench_r2_n${N}_$3.json 2> gpurun_out/bench_r2_n${N}_$3.err
  echo "bench N=$N steps=$1 rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_r2_n${N}_$3.json').read().strip().splitlines()[-1])
    print('N=$N $3: value %.1f G  ms %.4f  e2e %.1f G (%.4f ms)  step_ms %s  %s' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['ms_per_step'], d.get('step_ms'), d.get('warning','')))
except Exception as e:
    print('no line', e); print(open('gpurun_out/bench_r2_n${N}_$3.err').read()[-1500:])
PY
}
run 20 3 steps20
run 100 5 steps100
if [ "$N" -ge 2 ]; then
  timeout 600 python -m pytest tests/test_gpu_peer.py -q -p no:cacheprovider > gpurun_out/peer_tests_n$N.log 2>&1
  echo "peer tests rc=$?"; tail -3 gpurun_out/peer_tests_n$N.log
fi
if [ "$N" -eq 8 ] && [ -z "$SKIP_CONFIG5" ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 \
      tools/optimize_lens.py --steps 500 --side 2310 > gpurun_out/r2_config5_8gpu_500steps.json 2> gpurun_out/config5_8gpu.err
  echo "config5 rc=$?"; cat gpurun_out/r2_config5_8gpu_500steps.json; tail -3 gpurun_out/config5_8gpu.err
fi
