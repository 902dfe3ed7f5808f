#!/bin/bash
# N-GPU bench: peer-memory exchange vs NCCL all-reduce (run under gpurun --gpus N)
N=${1:-2}
STEPS=${2:-200}
mkdir -p gpurun_out
for mode in ${MODES:-peer nccl}; do
  TL_BENCH_COLLECTIVE=$mode TL_BENCH_WATCHDOG_S=150 timeout 240 python -m torch.distributed.run --nnodes=1 \
    --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) bench.py --gpus $N \
    --steps $STEPS --warmup 5 --no-cpu-baseline > gpurun_out/bench_n${N}_${mode}.json 2> gpurun_out/bench_n${N}_${mode}.err
  echo "$mode rc=$?"
  tail -1 gpurun_out/bench_n${N}_${mode}.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$mode', round(d['value']/1e9,1), 'G ev/s', round(d['ms_per_step'],4), 'ms/step, e2e', round(d['e2e']['value']/1e9,1), round(d['e2e']['ms_per_step'],4))
except Exception as e: print('no line', e)"
  tail -5 gpurun_out/bench_n${N}_${mode}.err
done
