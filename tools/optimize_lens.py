"""BASELINE.json config 5 driver: Adam on the 12-surface asphere lens, rays sharded over the ranks.

    python tools/optimize_lens.py [--steps 500] [--side 296]
    torchrun --nproc-per-node 8 tools/optimize_lens.py --steps 500 --side 1184   # 256 M rays / step
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchoptics_b200 import RayTracer, prescriptions   # noqa: E402
from torchoptics_b200.optimize import optimize_spot     # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--steps', type=int, default=500)
ap.add_argument('--side', type=int, default=296)
ap.add_argument('--lr', type=float, default=5e-5)
args = ap.parse_args()
rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = f'cuda:{local}'
if world > 1:
    torch.distributed.init_process_group('nccl', device_id=torch.device(dev))
specs, lens = prescriptions.asphere_12(dev)
tracer = RayTracer(mode='circular', n_rays=(args.side, args.side), rel_fields=tuple(np.linspace(0, 1, 16).tolist()),
                   wavelengths=('C', 'd', 'F'), default_device=dev)
rays = 16 * 3 * args.side * args.side
torch.cuda.synchronize()
t0 = time.perf_counter()
best, history = optimize_spot(tracer, specs, lens, steps=args.steps, lr=args.lr, shard=(rank, world))
torch.cuda.synchronize()
secs = time.perf_counter() - t0
if rank == 0:
    print(f'{args.steps} Adam steps, {rays} rays/step on {world} GPU(s): rms {history[0]:.5f} -> {history[-1]:.5f} '
          f'in {secs:.2f} s ({rays * 12 * args.steps / secs / 1e9:.1f} G asphere events/s incl. host)')
if world > 1:
    sys.stdout.flush()
    os._exit(0)
