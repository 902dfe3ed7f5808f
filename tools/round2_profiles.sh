#!/bin/bash
# round-2 evidence on ONE GPU: ncu launch list of the bench command and a full capture of the hot
# kernel (each only after its plain run exited 0).  Reports are condensed on the box
# (tools/ncu_summary.py); the hot kernel's .ncu-rep travels back too.
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_raw.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
capture() {   # name, kernel regex, skip, events per launch, command...
  name=$1; regex=$2; skip=$3; events=$4; shift 4
  "$@" > gpurun_out/plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f -o gpurun_out/$name "$@" \
      > gpurun_out/ncu_$name.log 2>&1 &&
  python tools/ncu_summary.py gpurun_out/$name.ncu-rep $events > gpurun_out/$name.txt
  echo "$name rc=$?"
}
capture prof_r2a_spot_rev_reg8 k_spot_rev 2 46261248 python tools/profile_spot.py 2
TL_REV=tmem12c2 capture prof_r2b_spot_rev_tmem12c2 k_spot_rev 2 46261248 env TL_REV=tmem12c2 python tools/profile_spot.py 2
rm -f gpurun_out/prof_r2b_spot_rev_tmem12c2.ncu-rep
ls -la gpurun_out | tail -12
head -40 gpurun_out/prof_r2a_spot_rev_reg8.txt
