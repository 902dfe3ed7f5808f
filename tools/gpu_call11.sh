#!/bin/bash
# round 2, last session: early-exit Newton in the asphere fast policy -- tests, timing, ncu summary, full-size configs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -s -p no:cacheprovider > gpurun_out/asphere_tests.log 2>&1
echo "pytest rc=$?"; grep "^f/" gpurun_out/asphere_tests.log; tail -2 gpurun_out/asphere_tests.log
python tools/profile_general.py 296 > gpurun_out/general_plain.log 2>&1 && cat gpurun_out/general_plain.log &&
ncu --set full --clock-control none --import-source on -k regex:k_trace_gen -s 3 -c 1 -f -o gpurun_out/prof_r2j_trace_gen \
    python tools/profile_general.py 296 > gpurun_out/ncu_r2j.log 2>&1 &&
python tools/ncu_summary.py gpurun_out/prof_r2j_trace_gen.ncu-rep 50466816 > gpurun_out/prof_r2j_trace_gen.txt
echo "ncu rc=$?"; head -12 gpurun_out/prof_r2j_trace_gen.txt; rm -f gpurun_out/prof_r2j_trace_gen.ncu-rep
timeout 600 python tools/full_size_configs.py > gpurun_out/full_size_configs.json 2> gpurun_out/full_size.err; echo "full size rc=$?"; grep -E "fwd_bwd_ms|events_per_s|ms_per_step|\"ms\"" gpurun_out/full_size_configs.json | head -12
