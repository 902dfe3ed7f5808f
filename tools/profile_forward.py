"""Forward-only fused sweep (trace + per-field spot moments, nothing materialised) on config 2: the forward-only
k_spot_rev and the round-1 kernel (TL_NO_REV=1), reference heights + kernel + row reduction as a CUDA graph, L2
flushed.  python tools/profile_forward.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchoptics_b200 import RayTracer, ops, prescriptions   # noqa: E402

dev = 'cuda:0'
specs, lens = prescriptions.double_gauss(dev)
tracer = RayTracer(mode='circular', n_rays=(296, 296), rel_fields=tuple(np.linspace(0, 1, 16).tolist()),
                   wavelengths=('C', 'd', 'F'), default_device=dev)
args = [a.detach() for a in tracer._ray_set(specs, lens)]
events = 16 * 3 * 296 * 296 * 11
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
ref = None
for name, env in (('round1', {'TL_NO_REV': '1'}), ('eval16', {})):
    for k in ('TL_NO_REV', 'TL_REV_EVAL'):
        os.environ.pop(k, None)
    os.environ.update(env)
    fn = lambda: ops.spot_moments(*args, want_grad=False)
    for _ in range(3):
        m, _ = fn()
    torch.cuda.synchronize()
    m = m.cpu().numpy()
    if ref is None:
        ref = m
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    a = [torch.cuda.Event(enable_timing=True) for _ in range(30)]
    b = [torch.cuda.Event(enable_timing=True) for _ in range(30)]
    for i in range(30):
        flush.zero_()
        a[i].record()
        g.replay()
        b[i].record()
    torch.cuda.synchronize()
    ms = float(np.mean([x.elapsed_time(y) for x, y in zip(a, b)]))
    scale = np.abs(ref).max(axis=(0, 1, 2))
    print(json.dumps({'variant': name, 'kernel': ops.spot_kernel_name(*args, want_grad=False), 'ms': ms,
                      'g_events_per_s': events / ms / 1e6, 'frac_fp32_peak_at_61_flop': events * 61 / (ms * 1e-3) / 74.45e12,
                      'max_rel_diff_vs_round1': float((np.abs(m - ref).max(axis=(0, 1, 2)) / scale).max()),
                      'n_ok_equal': bool(np.array_equal(m[..., 2], ref[..., 2]))}), flush=True)
