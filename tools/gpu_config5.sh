#!/bin/bash
# BASELINE.json config 5 alone on N GPUs of one box: tools/gpu_config5.sh N side tag   (8 GPUs: side 2310 = 256 M rays/step)
N=$1; SIDE=$2; TAG=$3
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
    tools/optimize_lens.py --steps 500 --side $SIDE > gpurun_out/r2_config5_${TAG}.json 2> gpurun_out/config5_${TAG}.err
echo "config5 rc=$?"; cat gpurun_out/r2_config5_${TAG}.json; tail -3 gpurun_out/config5_${TAG}.err
