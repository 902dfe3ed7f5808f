#!/bin/bash
# the bench line's headline numbers only (kernel change experiments)
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c '
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print("value %.1f G  ms %.4f  kernel_ms %.4f  frac %.4f  fwd_sweep_ms %.4f" % (d["value"] / 1e9, d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["forward"]["fused_sweep"]["ms"]))'
