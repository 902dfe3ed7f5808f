"""Variants of the fused spot pass side by side on one B200: the round-1 kernel (TL_NO_REV=1,
k_trace_adj<12,SPOT_GRAD,f4>) and every TL_REV variant of k_spot_rev (csrc/spot_rev.cuh).

For each: moments against the round-1 kernel's on (a) the config-2 lens at full size (all rays
clear) and (b) an over-filled Cooke pupil (misses, total reflection, backward rays: the exact-policy
re-trace and the dead-lane mirroring), then ms per launch with L2 flushed between launches, as a
CUDA graph.   python tools/rev_variants.py [side] > gpurun_out/rev_variants.json"""
import json
import os
import statistics
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchoptics_b200 import RayTracer, ops, prescriptions   # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 else 296
dev = 'cuda:0'
VARIANTS = [('round1', {'TL_NO_REV': '1'})] + [(v, {'TL_REV': v}) for v in
                                                ('reg8', 'tmem8', 'tmem10', 'tmem12', 'tmem12c2', 'tmem14c2', 'tmem16')]


def set_env(env):
    for k in ('TL_NO_REV', 'TL_REV'):
        os.environ.pop(k, None)
    os.environ.update(env)


def rays(lens_name, n_side, fields, scale=1.0):
    if lens_name == 'double_gauss':
        specs, lens = prescriptions.double_gauss(dev)
    else:
        specs, lens = prescriptions.load_yaml(lens_name, dev)
    if scale != 1.0:
        specs = specs.scale(scale)
    tracer = RayTracer(mode='circular', n_rays=(n_side, n_side), rel_fields=fields, wavelengths=('C', 'd', 'F'),
                       default_device=dev)
    return [a.detach() for a in tracer._ray_set(specs, lens)], lens.c.shape[1]


def timed(fn, reps, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    side_stream = torch.cuda.Stream()
    side_stream.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side_stream):
        fn()
    torch.cuda.current_stream().wait_stream(side_stream)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        fn()
    a = [torch.cuda.Event(enable_timing=True) for _ in range(reps)]
    b = [torch.cuda.Event(enable_timing=True) for _ in range(reps)]
    for i in range(reps):
        flush.zero_()
        a[i].record()
        graph.replay()
        b[i].record()
    torch.cuda.synchronize()
    ms = [x.elapsed_time(y) for x, y in zip(a, b)]
    return statistics.mean(ms), min(ms)


def main():
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    big, S = rays('double_gauss', side, tuple(np.linspace(0, 1, 16).tolist()))
    wild, _ = rays('baseline_cooke.yml', 48, (0., 0.5, 0.707, 1.0), scale=2.6)
    events = 16 * 3 * side * side * S
    out = {'events_per_launch': events, 'variants': {}}
    ref = {}
    for name, env in VARIANTS:
        set_env(env)
        row = {}
        try:
            for tag, args in (('config2', big), ('wild_cooke', wild)):
                m, ref_y = ops.spot_moments(*args)
                torch.cuda.synchronize()
                m = m.cpu().numpy()
                if name == 'round1':
                    ref[tag] = m
                    row[tag + '_n_ok'] = float(m[..., -1].sum())
                    row[tag + '_rays'] = int(np.prod(args[0].shape[2:3])) * m.shape[1] * m.shape[2]
                else:
                    scale = np.abs(ref[tag]).max(axis=(0, 1, 2), keepdims=True) + 1e-30
                    per_slot = (np.abs(m - ref[tag]) / scale).max(axis=(0, 1, 2))
                    row[tag + '_max_rel_diff_vs_round1'] = float(per_slot.max())
                    worst = int(per_slot.argmax())
                    row[tag + '_worst_slot'] = [worst, float(scale.ravel()[worst]), float(np.sort(per_slot)[-4])]
                    row[tag + '_n_ok_equal'] = bool(np.array_equal(m[..., -1], ref[tag][..., -1]))
                    row[tag + '_finite'] = bool(np.isfinite(m).all())
            mean_ms, min_ms = timed(lambda: ops.spot_moments(*big), 30, flush)
            row.update(ms_mean=mean_ms, ms_min=min_ms, g_events_per_s=events / mean_ms / 1e6,
                       frac_fp32_peak=events * 166 / (mean_ms * 1e-3) / 74.45e12)
        except Exception as exc:      # a variant that fails must not hide the others
            row['error'] = repr(exc)
        out['variants'][name] = row
        print(name, json.dumps(row), file=sys.stderr, flush=True)
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
