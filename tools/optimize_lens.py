"""BASELINE.json config 5 driver: Adam on the 12-surface asphere lens, rays sharded over the ranks.

    python tools/optimize_lens.py [--steps 500] [--side 296]
    torchrun --nproc-per-node 8 tools/optimize_lens.py --steps 500 --side 2310   # 256 M rays / step
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
group = None
if world > 1:
    torch.distributed.init_process_group('nccl', device_id=torch.device(dev))
    from torchoptics_b200.peer import PeerExchange
    group = PeerExchange(capacity=1 << 16)        # the per-step exchange of the moment sums
specs, lens = prescriptions.asphere_12(dev)
tracer = RayTracer(mode='circular', n_rays=(args.side, args.side), rel_fields=tuple(np.linspace(0, 1, 16).tolist()),
                   wavelengths=('C', 'd', 'F'), default_device=dev)
rays = 16 * 3 * args.side * args.side
torch.cuda.synchronize()
t0 = time.perf_counter()
stamps = []


def stamp(step, loss):
    stamps.append(time.perf_counter())


best, history = optimize_spot(tracer, specs, lens, steps=args.steps, lr=args.lr, shard=(rank, world), group=group,
                              callback=stamp)
torch.cuda.synchronize()
secs = time.perf_counter() - t0
if rank == 0:
    import json
    step_s = float(np.median(np.diff(stamps))) if len(stamps) > 2 else secs / args.steps
    print(json.dumps({'config': 'BASELINE.json config 5', 'steps': args.steps, 'rays_per_step': rays, 'gpus': world,
                      'collective': 'peer-memory exchange' if group is not None else 'none',
                      'total_s': secs, 'median_s_per_step': step_s,
                      'events_per_s_incl_host_and_optimizer': rays * 12 / step_s,
                      'rms_first': history[0], 'rms_min': min(history), 'rms_last': history[-1]}))
if world > 1:
    sys.stdout.flush()
    os._exit(0)
