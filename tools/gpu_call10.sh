#!/bin/bash
# ncu summaries of the two forward-only instantiations of k_spot_rev (config 2), after the plain run:
# eval16 = <16,16,ACC_NONE,0> (fused sweep: moments only), out16 = <16,16,ACC_OUT,0> (tl_trace_fwd: writes its outputs)
mkdir -p gpurun_out
timeout 300 python tools/profile_forward.py > gpurun_out/plain_forward.log 2>&1 || { tail -5 gpurun_out/plain_forward.log; exit 1; }
tail -2 gpurun_out/plain_forward.log
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_spot_revILi16ELi16ELi2 -s 3 -c 1 -f -o gpurun_out/prof_r2g_eval16 \
    python tools/profile_forward.py > gpurun_out/ncu_r2g.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_r2g_eval16.ncu-rep 46261248 > gpurun_out/prof_r2g_spot_rev_eval16.txt
echo "rc=$?"; head -60 gpurun_out/prof_r2g_spot_rev_eval16.txt; rm -f gpurun_out/prof_r2g_eval16.ncu-rep
