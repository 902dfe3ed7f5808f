"""compute_psf (tl_psf_bin) at a spot-sweep size on one B200: G grids x 3 channels
x R rays into n x n bins.  Reports rays/s and the fraction of the FP32 FMA peak at the algorithmic work of
the separable soft histogram: (n_xh + n_y) exponentials (3 flop + 1 MUFU each) + n_xh * n_y multiply-adds per
ray.   python tools/profile_psf.py [rays_per_channel] [bins]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchoptics_b200 import ops   # noqa: E402
from torchoptics_b200 import ray_tracing_lite as rt   # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
    bins = int(sys.argv[2]) if len(sys.argv) > 2 else 21
    dev = 'cuda:0'
    gen = torch.Generator(device='cpu').manual_seed(0)
    G, C = 16, 3
    x = (torch.randn((1, G, C, n), generator=gen) * 0.006).abs().to(dev)
    y = (torch.randn((1, G, C, n), generator=gen) * 0.008 + torch.linspace(0, 20, G).reshape(1, G, 1, 1)).to(dev)
    target = y.reshape(G, -1).mean(dim=1)
    fn = lambda: rt.compute_psf(x, y, (bins, bins), 0.002, target)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    reps = 10
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    # the binning launch alone
    xs, ys = x.reshape(G, C, n), y.reshape(G, C, n)
    incr = torch.full((G,), 0.002, device=dev)
    win = torch.full((G,), 0.002 * bins, device=dev)
    k = lambda: ops.psf_bin(xs, ys, target, incr, incr, win, win, (bins, bins))
    for _ in range(3):
        k()
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        k()
    b.record()
    torch.cuda.synchronize()
    k_ms = a.elapsed_time(b) / reps
    rays = G * C * n
    n_xh = bins // 2 + 1 if bins % 2 else bins // 2
    flop_per_ray = 2 * n_xh * bins + 4 * (n_xh + bins)
    print(json.dumps({'rays': rays, 'grids': G, 'channels': C, 'bins': [bins, bins], 'compute_psf_ms': ms,
                      'psf_bin_ms': k_ms, 'rays_per_s': rays / (k_ms * 1e-3), 'flop_per_ray': flop_per_ray,
                      'mufu_per_ray': n_xh + bins,
                      'tflops': rays * flop_per_ray / (k_ms * 1e-3) / 1e12,
                      'frac_fp32_peak': rays * flop_per_ray / (k_ms * 1e-3) / 74.45e12,
                      'input_gb_per_s': rays * 8 / (k_ms * 1e-3) / 1e9,
                      'unit_mass': float(out[3].double().sum(dim=(-1, -2)).mean())}))


if __name__ == '__main__':
    main()
