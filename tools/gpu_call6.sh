#!/bin/bash
# full ncu capture (with SASS / source counters) of the shipped hot kernel, after its plain run
mkdir -p gpurun_out
python tools/profile_spot.py 2 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_spot_rev -s 2 -c 1 -f -o gpurun_out/prof_r2d_spot_rev_tmem12 \
    python tools/profile_spot.py 2 > gpurun_out/ncu_r2c.log 2>&1 &&
python tools/ncu_summary.py gpurun_out/prof_r2d_spot_rev_tmem12.ncu-rep 46261248 > gpurun_out/prof_r2d_spot_rev_tmem12.txt
echo "rc=$?"; head -60 gpurun_out/prof_r2d_spot_rev_tmem12.txt
