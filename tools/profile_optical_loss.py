"""The reference's real workload shape end to end on one B200: a mini-batch of B
network outputs -> Optical_Loss.optical_loss_unsupervised (decode, last curvature, staging, ray aiming, fused
spot pass + fused penalty pass over B lenses x 8 fields x 3 wavelengths x 64 pupil points) -> backward to the
network outputs.  The designs are the golden samples of tests/golden/optical_loss/<type>.npz tiled with a
+-1 % jitter.   python tools/profile_optical_loss.py [lens_type] [B ...]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torchoptics_b200 import _native                      # noqa: E402
from torchoptics_b200.optical_loss import Optical_Loss    # noqa: E402


def batch_of(lens_type, B, device, seed=0):
    with np.load(os.path.join(ROOT, 'tests', 'golden', 'optical_loss', lens_type + '.npz')) as z:
        inputs, outputs = z['inputs'], z['outputs']
    rng = np.random.default_rng(seed)
    pick = rng.integers(0, inputs.shape[0], B)
    x = inputs[pick].copy()
    y = outputs[pick] * rng.uniform(0.99, 1.01, (B, outputs.shape[1])).astype(np.float32)
    return torch.from_numpy(x).to(device), torch.from_numpy(y.astype(np.float32)).to(device)


def measure(lens_type, B, device='cuda:0', reps=20, graph=True):
    loss_fn = Optical_Loss(lens_type)
    x, y = batch_of(lens_type, B, device)
    y.requires_grad_(True)
    sequence, stop_idx = lens_type, int(x[0, -3])

    def step():
        loss, rms, pen = loss_fn.optical_loss_unsupervised(x, y, 0.2, device, sequence=sequence, stop_idx=stop_idx)
        g, = torch.autograd.grad(loss, y)
        return loss, rms, pen, g
    for _ in range(3):
        out = step()
    torch.cuda.synchronize()
    before = _native.launch_count()
    step()
    launches = _native.launch_count() - before
    graphed_ms = None
    # the whole step -- decode, kernels, autograd backward to the network outputs -- as one CUDA graph.  Results of
    # earlier EAGER steps must not be alive across the capture: with a previous step's loss (and its autograd graph)
    # still referenced, torch's capture of the backward pass is invalidated ("operation failed due to a previous
    # error during capture"; bisected on the B200: decode + autograd.grad alone shows it, forward-only captures do not)
    out = None
    for attempt in range(2 if graph else 0):
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    step()
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                g_out = step()
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize()
            ga, gb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ga.record()
            for _ in range(reps):
                g.replay()
            gb.record()
            torch.cuda.synchronize()
            graphed_ms = ga.elapsed_time(gb) / reps
            eager = step()
            assert torch.allclose(g_out[0], eager[0], rtol=1e-6) and torch.allclose(g_out[3], eager[3], rtol=1e-5, atol=1e-7)
            break
        except Exception as exc:      # report, do not hide the eager numbers
            graphed_ms = None
            print(f'[profile_optical_loss] graph capture failed (attempt {attempt + 1}): {str(exc).splitlines()[0]}',
                  file=sys.stderr)
            torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a.record()
    for _ in range(reps):
        out = step()
    b.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps
    ms = a.elapsed_time(b) / reps
    S = loss_fn.numsurf
    rays = B * Optical_Loss.N_FIELDS * 3 * Optical_Loss.N_PUPIL_RINGS ** 2
    events = rays * S * 2          # two fused passes (spot, penalty), each forward + adjoint over every ray-surface event
    return {'lens_type': lens_type, 'lenses': B, 'rays': rays, 'ms_per_step': ms, 'wall_ms_per_step': wall * 1e3,
            'lenses_per_s': B / (ms * 1e-3), 'events_per_s': events / (ms * 1e-3), 'library_launches_per_step': launches,
            'graphed_ms_per_step': graphed_ms,
            'graphed_lenses_per_s': None if graphed_ms is None else B / (graphed_ms * 1e-3),
            'graphed_events_per_s': None if graphed_ms is None else events / (graphed_ms * 1e-3),
            'loss': float(out[0]), 'rms': float(out[1]), 'penalty': float(out[2]),
            'finite_grads': bool(torch.isfinite(out[3]).all())}


if __name__ == '__main__':
    lens_type = sys.argv[1] if len(sys.argv) > 1 else 'GAGAGA'
    sizes = [int(v) for v in sys.argv[2:]] or [64, 1024, 4096]
    for B in sizes:
        print(json.dumps(measure(lens_type, B)), flush=True)
