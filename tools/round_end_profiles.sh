#!/bin/bash
# round-end evidence on ONE GPU: bench line, ncu launch list of the same command, full captures of
# the hot kernels (each only after its plain run exited 0).  Reports are condensed on the box
# (tools/ncu_summary.py); only the hot kernel's .ncu-rep travels back (gpurun_out is capped at 64 MiB).
mkdir -p gpurun_out
python bench.py --steps 100 --warmup 5 > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err || exit 1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/final_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
capture() {   # name, kernel regex, skip, events per launch, command...
  name=$1; regex=$2; skip=$3; events=$4; shift 4
  "$@" > gpurun_out/plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f -o gpurun_out/$name "$@" \
      > gpurun_out/ncu_$name.log 2>&1 &&
  python tools/ncu_summary.py gpurun_out/$name.ncu-rep $events > gpurun_out/$name.txt
}
capture prof_r1e k_trace_adj 2 46261248 python tools/profile_spot.py 2
capture prof_r1_rows k_spot_rows 2 11010048 python tools/profile_batched.py 1024
capture prof_r1_pensum k_trace_adj 1 46261248 python tools/profile_penalty_kernel.py
capture prof_r1_fwdpw k_trace_fwd_pw 1 46261248 python tools/profile_penalty_kernel.py
rm -f gpurun_out/prof_r1_rows.ncu-rep gpurun_out/prof_r1_pensum.ncu-rep gpurun_out/prof_r1_fwdpw.ncu-rep
ls -la gpurun_out | tail -20
