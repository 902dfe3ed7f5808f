#!/bin/bash
# 8-GPU bench under a few NCCL small-message settings (latency of the 27 KB moment all-reduce)
run() {
  label=$1; shift
  env "$@" TL_BENCH_WATCHDOG_S=120 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
      --master-addr 127.0.0.1 --master-port $((29520 + RANDOM % 100)) bench.py --gpus $N --steps 100 --warmup 5 \
      --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$label', round(d['value']/1e9,1), 'G ev/s', round(d['ms_per_step'],4), 'ms/step, e2e', round(d['e2e']['value']/1e9,1))"
}
N=${1:-8}
run default X=1
run proto_LL NCCL_PROTO=LL
run algo_tree NCCL_ALGO=Tree
run nvls_off NCCL_NVLS_ENABLE=0
run ll_ring NCCL_PROTO=LL NCCL_ALGO=Ring
