#!/bin/bash
# ncu summary of k_peer_allreduce (world 1: ncu must not run a multi-rank command, so this is the kernel's
# local path -- push into its own window, flag, rank-order sum -- without the NVLink hop)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_peer.py -q -p no:cacheprovider -k world1 > gpurun_out/peer_plain.log 2>&1 && tail -2 gpurun_out/peer_plain.log &&
ncu --set full --clock-control none --import-source on -k regex:k_peer_allreduce -s 1 -c 1 -f -o gpurun_out/prof_r2f_peer_allreduce \
    python -m pytest tests/test_gpu_peer.py -q -p no:cacheprovider -k world1 > gpurun_out/ncu_r2f.log 2>&1 &&
python tools/ncu_summary.py gpurun_out/prof_r2f_peer_allreduce.ncu-rep 1 > gpurun_out/prof_r2f_peer_allreduce.txt
echo "rc=$?"; head -34 gpurun_out/prof_r2f_peer_allreduce.txt; rm -f gpurun_out/prof_r2f_peer_allreduce.ncu-rep
