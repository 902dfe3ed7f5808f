#!/bin/bash
# ncu summary of the asphere fused pass k_trace_gen (config 3's hot kernel), after its plain run
mkdir -p gpurun_out
python tools/profile_general.py 296 > gpurun_out/general_plain.log 2>&1 && cat gpurun_out/general_plain.log &&
ncu --set full --clock-control none --import-source on -k regex:k_trace_gen -s 3 -c 1 -f -o gpurun_out/prof_r2e_trace_gen \
    python tools/profile_general.py 296 > gpurun_out/ncu_r2e.log 2>&1 &&
python tools/ncu_summary.py gpurun_out/prof_r2e_trace_gen.ncu-rep 50466816 > gpurun_out/prof_r2e_trace_gen.txt
echo "rc=$?"; head -40 gpurun_out/prof_r2e_trace_gen.txt; rm -f gpurun_out/prof_r2e_trace_gen.ncu-rep
