#!/usr/bin/env python
"""Benchmark of the ray-trace hot path: ray-surface events/s, forward+backward.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): Double-Gauss 50 mm f/3, 11 surfaces,
16 fields x 3 wavelengths x 296^2 pupil points = 4 205 568 rays per GPU,
i.e. 46.3 M ray-surface events per step.  One step = trace -> RMS spot loss ->
gradients w.r.t. c, t, mu, z (the fused `spot_rms` pass + its backward).
With N > 1 (torchrun) every rank traces its own 4.2 M-ray slice of an N x larger
pupil grid (weak scaling) and the per-field sums are all-reduced over NCCL.

Printed keys follow the driver contract: `value` is device-timed with inputs
resident in HBM (CUDA events on the launching stream, L2 flushed between steps);
`e2e` goes through the public RayTracer API from pinned host buffers and reads
the loss and gradients back; `roofline` is the dominant kernel against the FP32
FMA peak (this path has no dense contraction and ~0 HBM traffic);
`cpu_baseline` is the oracle port timed on the host cores on a bounded sample.
`--impl reference` runs only that CPU arm.
"""
import argparse
import json
import os
import statistics
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FLOPS_FWD, FLOPS_BWD = 61, 105          # per spherical ray-surface event
N_FIELDS, N_SIDE = 16, 296
WAVELENGTHS = ('C', 'd', 'F')
METRIC = 'ray-surface events/sec fwd+bwd'
# dram__bytes_read.sum + dram__bytes_write.sum of one k_trace_adj launch of this workload, from the
# ncu --set full capture summarised in profiles/r1d_spot_grad_f4_geometric.txt (805 120 B + 0 B):
# the pupil grid and the surface tables; the algorithmic HBM traffic of the fused pass is ~0.
NCU_DRAM_BYTES_PER_LAUNCH = 805120


def workload_name(n_theta):
    return (f'double_gauss_S11_F{N_FIELDS}_W{len(WAVELENGTHS)}_pupil{N_SIDE}x{n_theta}'
            f'_fwd+bwd_rms')


# ----------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------
_SAMPLER_CODE = r"""
import json, select, sys, time
try:
    import pynvml as nv
    nv.nvmlInit()
    h = nv.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
    max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
except Exception as exc:
    print(json.dumps({'error': repr(exc)}), flush=True)
    sys.exit(0)
print('ready', flush=True)
samples, mask = [], 0
while True:
    try:
        samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
        mask |= nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
    except Exception:
        pass
    if select.select([sys.stdin], [], [], 0.0005)[0]:      # the parent closed our stdin: done
        break
print(json.dumps({'samples': samples, 'mask': mask, 'max_mhz': max_mhz}), flush=True)
"""


class ClockSampler:
    """SM clock and throttle reasons of one GPU, sampled through NVML by a CHILD PROCESS for as long as the `with`
    block runs.  (A sampling thread of this process gets one or two samples out of a 5 ms timed region: the timing
    loop holds the GIL.)  `index` is the NVML device index = CUDA_VISIBLE_DEVICES-relative local rank on these boxes."""
    REASONS = {0x2: 'applications_clocks_setting', 0x4: 'sw_power_cap', 0x8: 'hw_slowdown',
               0x20: 'sw_thermal_slowdown', 0x40: 'hw_thermal_slowdown', 0x80: 'hw_power_brake',
               0x100: 'display_clock_setting'}

    def __init__(self, index):
        self.index, self.proc, self.result = index, None, None

    def __enter__(self):
        import subprocess
        try:
            self.proc = subprocess.Popen([sys.executable, '-c', _SAMPLER_CODE, str(self.index)], stdin=subprocess.PIPE,
                                         stdout=subprocess.PIPE, text=True)
            first = self.proc.stdout.readline().strip()
            if first != 'ready':
                self.result = json.loads(first) if first else {'error': 'sampler did not start'}
                self.proc.stdin.close()
                self.proc.wait(timeout=10)
                self.proc = None
        except Exception as exc:
            self.result, self.proc = {'error': repr(exc)}, None
        return self

    def __exit__(self, *exc):
        if self.proc is not None:
            try:
                self.proc.stdin.close()
                self.result = json.loads(self.proc.stdout.readline())
                self.proc.wait(timeout=10)
            except Exception as err:
                self.result = {'error': repr(err)}
                self.proc.kill()

    def summary(self):
        res = self.result or {'error': 'not run'}
        if not res.get('samples'):
            return {'sm_mhz': None, 'sm_max_mhz': res.get('max_mhz'), 'reasons': ['nvml_unavailable: ' + str(res.get('error'))]}
        reasons = [name for bit, name in self.REASONS.items() if res['mask'] & bit]
        return {'sm_mhz': statistics.median(res['samples']), 'sm_min_mhz': min(res['samples']), 'sm_max_mhz': res['max_mhz'],
                'reasons': reasons, 'samples': len(res['samples'])}


# ----------------------------------------------------------------------------
# CPU arm: the reference's own path on the host cores
# ----------------------------------------------------------------------------
def cpu_arm(steps, warmup, side=N_SIDE):
    """trace_rays -> compute_rms2d -> backward of the config-2 lens at the configured pupil
    (side^2 points, 16 fields, 3 wavelengths) on the host cores, all threads.  Runs the UNMODIFIED
    reference staged under oracle/_ref (oracle/make_ref.py; kind "reference") when it travelled with
    the snapshot, else the oracle port of it (kind "port")."""
    n_threads = os.cpu_count() or 1
    torch.set_num_threads(n_threads)          # torchrun exports OMP_NUM_THREADS=1: undo it for this arm
    fields = tuple(np.linspace(0, 1, N_FIELDS).tolist())
    from oracle import make_ref
    if make_ref.available():
        from torchoptics_b200.prescriptions import DOUBLE_GAUSS as d      # plain prescription data
        rtl, lm = make_ref.load_reference()
        kind = 'reference'
        structure = lm.Structure(np.array(d['stop_idx']), sequence=np.array(d['sequence']), default_device='cpu')
        lens = lm.Lens(structure, *(torch.tensor(d[k], dtype=torch.float32) for k in ('c', 't', 'nd', 'v')))
        epd = lens.efl.detach() / torch.tensor(d['f_number'])
        specs = lm.Specs(structure, epd, torch.deg2rad(torch.tensor(d['hfov'])))
        for name in ('c', 't', 'nd'):
            getattr(lens, name).requires_grad_(True)
        tracer = rtl.RayTracer(mode='circular', n_rays=(side, side), rel_fields=fields, wavelengths=WAVELENGTHS,
                               default_device='cpu')

        def one_step():
            x, y, cx, cy, ok, bw = tracer.trace_rays(specs, lens)
            rms = rtl.compute_rms2d(x, y, ok)
            torch.autograd.grad(rms, [lens.c, lens.t, lens.nd])
            return float(rms)
        what = 'the unmodified reference (oracle/_ref: torchlens.ray_tracing_lite RayTracer.trace_rays + compute_rms2d + autograd)'
    else:
        from oracle import trace_oracle as oracle
        from torchoptics_b200 import RayTracer, prescriptions
        kind = 'port'
        specs, lens = prescriptions.double_gauss('cpu')
        for name in ('c', 't', 'nd'):
            getattr(lens, name).requires_grad_(True)
        tracer = RayTracer(mode='circular', n_rays=(side, side), rel_fields=fields, wavelengths=WAVELENGTHS,
                           default_device='cpu')

        def one_step():
            out = oracle.trace(*tracer._ray_set(specs, lens))
            rms = oracle.spot_rms(out[0], out[1], out[4])
            torch.autograd.grad(rms, [lens.c, lens.t, lens.nd])
            return float(rms)
        what = 'oracle (torch CPU eager port of trace_skew + compute_rms2d + autograd; oracle/_ref not staged)'
    S = int(lens.c.shape[1])
    events = N_FIELDS * len(WAVELENGTHS) * side * side * S
    times, rms = [], None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        rms = one_step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    best = min(times)
    return {'value': events / best, 'unit': 'events/s', 'cores': torch.get_num_threads(), 'kind': kind,
            'sample': f'{what}, double-gauss S{S}, {N_FIELDS} fields x {len(WAVELENGTHS)} wavelengths x '
                      f'{side}^2 pupil = {events // S} rays ({events} events) per step, best of {steps} after '
                      f'{warmup} warm-up, rms {rms:.6f}',
            'ms_per_step': statistics.mean(times) * 1e3, 'host_cpus': os.cpu_count(), 'rays': events // S}


def run_reference(args, rank):
    if rank != 0:
        return
    steps = max(1, min(args.steps, 5))
    warmup = max(1, min(args.warmup, 1))
    base = cpu_arm(steps, warmup)
    line = {'impl': 'reference', 'metric': METRIC, 'value': base['value'], 'unit': 'events/s',
            'n_gpus': args.gpus, 'steps': steps, 'warmup': warmup,
            'ms_per_step': base['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': workload_name(N_SIDE), 'lens': 'double_gauss_50mm_f3', 'fields': N_FIELDS,
                       'wavelengths': len(WAVELENGTHS), 'rays': base['rays'],
                       'note': 'one host runs one CPU copy of the per-GPU workload (4.2 M rays) whatever --gpus is'},
            'cpu_baseline': {k: base[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
            'e2e': {'value': base['value'], 'unit': 'events/s', 'h2d_bytes_per_step': 0,
                    'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-graph', action='store_true', help='time eager launches instead of a CUDA graph')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3

    import faulthandler
    # never hang a box: dump every thread's stack and exit if the whole run exceeds the limit
    faulthandler.dump_traceback_later(float(os.environ.get('TL_BENCH_WATCHDOG_S', 600)), exit=True)

    def note(msg):
        if os.environ.get('TL_BENCH_VERBOSE'):
            print(f'[bench rank {rank}] {msg}', file=sys.stderr, flush=True)

    import torch.distributed as dist
    from torchoptics_b200 import RayTracer, _native, ops, prescriptions
    from torchoptics_b200 import ray_tracing_lite as rt

    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local_rank)
    dev = f'cuda:{local_rank}'
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device(dev))
    _native.load()

    n_theta = N_SIDE * world                     # weak scaling: 4.2 M rays per rank
    fields = tuple(np.linspace(0, 1, N_FIELDS).tolist())
    specs, lens = prescriptions.double_gauss(dev)
    S = lens.c.shape[1]
    tracer = RayTracer(mode='circular', n_rays=(N_SIDE, n_theta), rel_fields=fields,
                       wavelengths=WAVELENGTHS, default_device=dev)
    rays_total = N_FIELDS * len(WAVELENGTHS) * N_SIDE * n_theta
    events_total = rays_total * S
    shard = (rank, world)
    # the one collective of the data path: our peer-memory exchange kernel (NVLink P2P stores,
    # rank-ordered sum) unless TL_BENCH_COLLECTIVE=nccl asks for torch.distributed.all_reduce
    exchange = None
    collective = 'none (1 rank)'
    if world > 1:
        if os.environ.get('TL_BENCH_COLLECTIVE', 'peer') == 'nccl':
            collective = 'nccl all_reduce (27 KB fp64) inside the CUDA graph'
        else:
            from torchoptics_b200.peer import PeerExchange
            try:       # (raises on EVERY rank together if any rank cannot create / map a window)
                exchange = PeerExchange(capacity=1 << 16)
                collective = ('k_peer_allreduce: one-shot all-reduce over NVLink peer memory (CUDA IPC '
                              'windows), inside the CUDA graph')
            except _native.NativeLibraryError as exc:
                print(f'[bench rank {rank}] {exc}; using the NCCL all_reduce', file=sys.stderr)
                collective = f'nccl all_reduce inside the CUDA graph (peer exchange unavailable: {exc})'

    # ---- device-resident step: inputs already in HBM ------------------------
    ray_args = [a.detach() for a in tracer._ray_set(specs, lens)]
    for i in (2, 5, 6, 7):                       # z, c, t, mu carry gradients
        ray_args[i] = ray_args[i].clone().requires_grad_(True)
    leaves = [ray_args[i] for i in (2, 5, 6, 7)]

    def step():
        # the fused lens pass as its bare kernel sequence: staging (index model, pupil position,
        # field cosines) -> chief rays -> trace + adjoint -> row reduction -> (exchange) -> finalize ->
        # chain rule to (c, t, nd, v): RMS and its gradients, nothing per-ray materialised
        return tracer.spot_rms_and_grads(specs, lens, shard=shard, group=exchange)

    note('first eager step')
    before = _native.launch_count()
    step()
    torch.cuda.synchronize()
    launches_per_step = _native.launch_count() - before

    graph = None
    note('capturing the step graph')
    if not args.no_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    step()
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                graph_out = step()
        except Exception as exc:                 # fall back to eager launches
            print(f'[bench] CUDA graph capture failed, timing eager launches: {exc}', file=sys.stderr)
            graph = None
    run_step = graph.replay if graph is not None else step

    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # 2x L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def host_lead():
        """A ~3 ms spin kernel in front of a timed loop (outside every timed interval: the events bracket flush ->
        step).  The loops below enqueue far faster than the device executes, but they start level with it right after
        a synchronize: a host hiccup in the first iterations (GC, the clock sampler's child starting) would leave the
        device idle INSIDE an event pair and be billed to the kernel -- seen once as a 0.239 ms mean over 50 launches
        of a 0.204 ms kernel.  With the host 3 ms ahead from the first iteration, event pairs measure device time only."""
        torch.cuda._sleep(int(6.0e6))

    note('warm-up')
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    with ClockSampler(local_rank) as clocks:     # (sampled through the warm-up too: the timed region is milliseconds)
        for _ in range(args.warmup):
            flush.zero_()
            run_step()
        barrier()
        # the host barrier leaves the ranks tens of microseconds apart, and the first exchange would bill
        # that skew to the first timed steps: a few UNTIMED steps let the exchange itself align the ranks
        host_lead()                              # (in front of the aligning steps: the ranks' spins end apart)
        for _ in range(3 if world > 1 else 0):
            flush.zero_()
            run_step()
        note('timed region')
        for i in range(args.steps):
            flush.zero_()                        # evict L2 between timed iterations
            starts[i].record()
            run_step()
            stops[i].record()
        barrier()
    step_ms = [a.elapsed_time(b) for a, b in zip(starts, stops)]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    ms_per_step = total_ms / args.steps
    value = events_total / (ms_per_step * 1e-3)
    step_stats = {'min': min(step_ms), 'median': statistics.median(step_ms), 'max': max(step_ms), 'rank': rank}

    # ---- the same step back to back for ~1.5 s: the SUSTAINED clock and throughput of this kernel (the
    # timed region above is a burst of a few milliseconds in which NVML can be read only a handful of times) ----
    note('sustained run')
    n_sustained = max(args.steps, int(1.5 / max(ms_per_step * 1e-3, 1e-6)))
    if world > 1:
        count = torch.tensor([n_sustained], dtype=torch.int64, device=dev)
        dist.all_reduce(count, op=dist.ReduceOp.MIN)      # (every rank must run the exchange the same number of times)
        n_sustained = int(count.item())
    s_start, s_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with ClockSampler(local_rank) as sustained_clocks:
        s_start.record()
        for _ in range(n_sustained):
            run_step()
        s_stop.record()
        torch.cuda.synchronize()
    sustained_ms = torch.tensor([s_start.elapsed_time(s_stop)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(sustained_ms, op=dist.ReduceOp.MAX)
    sustained_ms = float(sustained_ms.item())
    sustained = {'value': events_total * n_sustained / (sustained_ms * 1e-3), 'unit': 'events/s', 'steps': n_sustained,
                 'seconds': sustained_ms * 1e-3, 'ms_per_step': sustained_ms / n_sustained,
                 'l2': 'not flushed (back-to-back replays; the step reads 0.8 MB)', 'clocks': sustained_clocks.summary()}

    note('kernel-only timing')
    # ---- dominant kernel ALONE (k_spot_rev: one launch, no reference-height or row-reduction launch
    # around it), live CUDA events on the launching stream, L2 flushed before every launch ----
    plain = [a.detach() for a in ray_args]
    k_run = ops.spot_kernel_runner(*plain, shard=shard)
    for _ in range(3):
        k_run()
    torch.cuda.synchronize()
    k_starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    k_stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    host_lead()
    for i in range(args.steps):
        flush.zero_()
        k_starts[i].record()
        k_run()
        k_stops[i].record()
    torch.cuda.synchronize()
    kernel_ms = statistics.mean(a.elapsed_time(b) for a, b in zip(k_starts, k_stops))
    events_rank = events_total // world

    # ---- forward-only figures (the metric is quoted forward and forward+backward) ----
    def timed_graph(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        run = fn
        try:
            side2 = torch.cuda.Stream()
            side2.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side2):
                fn()
            torch.cuda.current_stream().wait_stream(side2)
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph):
                fn()
            run = gph.replay
        except Exception as exc:
            print(f'[bench] forward graph capture failed: {exc}', file=sys.stderr)
        a = [torch.cuda.Event(enable_timing=True) for _ in range(reps)]
        b = [torch.cuda.Event(enable_timing=True) for _ in range(reps)]
        host_lead()
        for i in range(reps):
            flush.zero_()
            a[i].record()
            run()
            b[i].record()
        torch.cuda.synchronize()
        return statistics.mean(x.elapsed_time(y) for x, y in zip(a, b))

    note('forward-only timing')
    fwd_reps = max(5, min(args.steps, 20))
    sweep_ms = timed_graph(lambda: ops.spot_moments(*plain, want_grad=False, shard=shard), fwd_reps)
    p_lo, p_hi = ops.pupil_slice(plain[0].shape[2], rank, world)
    sliced = [plain[0][:, :, p_lo:p_hi], plain[1][:, :, p_lo:p_hi]] + plain[2:]
    trace_ms = timed_graph(lambda: ops.trace(*sliced), fwd_reps)
    forward = {'fused_sweep': {'value': events_rank * world / (sweep_ms * 1e-3), 'unit': 'events/s',
                               'ms': sweep_ms, 'frac_fp32_peak': None,
                               'what': 'trace + per-field spot moments, nothing materialised'},
               'trace_skew': {'value': events_rank * world / (trace_ms * 1e-3), 'unit': 'events/s',
                              'ms': trace_ms, 'what': 'tl_trace_fwd writing x, y, cx, cy, ok, backward (18 B/ray)'}}
    sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
    peaks = {}
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    sm_max_mhz = peaks.get('sm_max_mhz') or clocks.summary().get('sm_max_mhz') or 1965.0
    peak_tflops = sms * 128 * 2 * sm_max_mhz * 1e6 / 1e12
    achieved = events_rank * (FLOPS_FWD + FLOPS_BWD) / (kernel_ms * 1e-3) / 1e12
    for entry in forward.values():
        entry['frac_fp32_peak'] = entry['value'] / world * FLOPS_FWD / 1e12 / peak_tflops
    roofline = {'bound': 'fp32_fma', 'achieved': achieved, 'peak': peak_tflops, 'unit': 'TFLOP/s',
                'frac': achieved / peak_tflops,
                'traffic': NCU_DRAM_BYTES_PER_LAUNCH if world == 1 else None,
                'kernel': f'{ops.spot_kernel_name(*plain, shard=shard)} (that one launch alone)',
                'kernel_ms': kernel_ms,
                'flops_per_event': FLOPS_FWD + FLOPS_BWD,
                'peak_source': f'{sms} SMs x 128 lanes x 2 flop x {sm_max_mhz:.0f} MHz '
                               f'(sm_max_mhz of MEASURED_PEAKS.json); algorithmic HBM bytes ~0, '
                               f'hbm peak {peaks.get("hbm_gbs")} GB/s is not the bound'}

    # ---- end to end through the public API, host buffers in, loss+grads out ---
    # GraphedSpotStep: pinned host prescription -> H2D -> index model / pupil position / ray set
    # -> fused trace+adjoint -> finalize -> chain rule -> D2H, captured once as a CUDA graph.
    note('end-to-end timing')
    from torchoptics_b200 import GraphedSpotStep, lens_modeling as lm
    host_lens = {k: getattr(lens, k).detach().cpu().pin_memory() for k in ('c', 't', 'nd', 'v')}      # pinned once
    graphed = None
    try:
        graphed = GraphedSpotStep(tracer, specs, lens, shard=shard, group=exchange)
    except Exception as exc:
        print(f'[bench] GraphedSpotStep capture failed, e2e uses the eager API: {exc}', file=sys.stderr)

    result_host = torch.empty((1 + 3 * S,), dtype=torch.float32).pin_memory()

    def read_back(rms, dl):
        """loss and gradients to the host in ONE copy"""
        result_host.copy_(torch.cat([rms.detach().reshape(1)] + [dl[k].grad.reshape(-1) for k in ('c', 't', 'nd')]),
                          non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(result_host[0]), result_host[1:]

    def e2e_eager_step():
        dl = {k: v.to(dev, non_blocking=True) for k, v in host_lens.items()}
        for k in ('c', 't', 'nd'):
            dl[k].requires_grad_(True)
        lens_i = lm.Lens(lens.structure, dl['c'], dl['t'], dl['nd'], dl['v'])
        rms, _ = tracer.spot_rms(specs, lens_i, shard=shard, group=exchange)
        rms[0].backward()
        return read_back(rms[0], dl)

    def e2e_step():
        if graphed is None:
            return e2e_eager_step()
        rms, grads = graphed(**host_lens)
        return float(rms[0]), grads

    def time_e2e(fn, n):
        for _ in range(3):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        barrier()
        secs = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(secs, op=dist.ReduceOp.MAX)
        return float(secs.item())

    def drop_in_step():
        """The reference's own call sequence, unchanged: trace_rays -> compute_rms2d -> backward."""
        dl = {k: v.to(dev, non_blocking=True) for k, v in host_lens.items()}
        for k in ('c', 't', 'nd'):
            dl[k].requires_grad_(True)
        lens_i = lm.Lens(lens.structure, dl['c'], dl['t'], dl['nd'], dl['v'])
        out = tracer.trace_rays(specs, lens_i)
        rms = rt.compute_rms2d(out[0], out[1], out[4])
        rms.backward()
        return read_back(rms, dl)

    e2e_secs = time_e2e(e2e_step, args.steps)
    e2e_value = events_total * args.steps / e2e_secs
    eager_steps = max(3, min(args.steps, 20))
    eager_secs = time_e2e(e2e_eager_step, eager_steps)
    drop_in_secs = time_e2e(drop_in_step, eager_steps) if world == 1 else None
    # ---- the penalty of compute_loss_out (rtl:641-657, osl:430-450) ----
    # fused penalty pass alone (device resident) and the graphed loss step rms + 0.2 * penalty
    # from host buffers; all ranks run it (it contains the collective), rank 0 reports
    note('penalty pass timing')
    penalty_row = None
    try:
        pen_reps = max(5, min(args.steps, 20))
        pen_ms = timed_graph(lambda: ops.penalty_sum(*plain, S, shard=shard, group=exchange), pen_reps)
        loss_step = GraphedSpotStep(tracer, specs, lens, shard=shard, group=exchange, penalty_rate=0.2)
        loss_secs = time_e2e(lambda: loss_step(**host_lens), pen_reps)
        penalty_row = {'penalty_pass': {'value': events_total / (pen_ms * 1e-3), 'unit': 'events/s', 'ms': pen_ms,
                                        'what': 'tl_penalty_accumulate + finalize: value and gradient of sum(Q), '
                                                'no stack materialised'},
                       'loss_step_e2e': {'value': events_total * pen_reps / loss_secs, 'unit': 'events/s',
                                         'ms': loss_secs / pen_reps * 1e3,
                                         'what': 'GraphedSpotStep(penalty_rate=0.2): host prescription -> '
                                                 'rms + 0.2 penalty and its gradients -> host',
                                         'penalty': float(loss_step.host_penalty[0])}}
        loss_step = None
    except Exception as exc:
        print(f'[bench] penalty timing failed: {exc}', file=sys.stderr)

    # ---- the reference's real workload shape: a mini-batch of B lens designs ->
    # Optical_Loss.optical_loss_unsupervised (batched; the reference loops over samples) -> backward ----
    batched_row = None
    if world == 1:
        note('batched-lens loss')
        try:
            sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'tools'))
            import profile_optical_loss
            batched_row = {'what': 'Optical_Loss(GAGAGA).optical_loss_unsupervised on B designs (8 fields x 3 wavelengths x '
                                   '8x8 pupil = 1 536 rays per lens, one ray-aiming iteration) + backward to the network '
                                   'outputs; events = rays x surfaces x 2 fused passes (spot, penalty), each forward + adjoint',
                           'runs': [profile_optical_loss.measure('GAGAGA', n_lenses, device=dev, reps=10)
                                    for n_lenses in (1024, 4096)]}
        except Exception as exc:
            print(f'[bench] batched-lens timing failed: {exc}', file=sys.stderr)

    # ---- BASELINE.json configs 3 / 5: the 12-surface even-asphere lens through the general-surface fused pass
    # (k_trace_gen: Newton sag solve with early exit, 11 gradients per surface), same fields / wavelengths / pupil
    # as the headline workload; rank 0 at N = 1 only (tools/full_size_configs.py has the 67 M-ray size) ----
    asphere_row = None
    if world == 1:
        note('asphere fused pass')
        try:
            from torchoptics_b200 import prescriptions
            a_specs, a_lens = prescriptions.asphere_12(dev)
            a_args = [a.detach() for a in tracer._ray_set(a_specs, a_lens)]
            a_ext = {k: v.detach() for k, v in tracer._extension_tables(a_lens).items() if v is not None}
            a_events = rays_total * int(a_args[6].shape[-1])
            a_reps = max(5, min(args.steps, 20))
            grad_ms = timed_graph(lambda: ops.spot_moments(*a_args, want_grad=True, **a_ext), a_reps)
            eval_ms = timed_graph(lambda: ops.spot_moments(*a_args, want_grad=False, **a_ext), a_reps)
            a_mom, _ = ops.spot_moments(*a_args, want_grad=False, **a_ext)
            asphere_row = {'what': 'asphere_12 (12 even-asphere surfaces, a4..a16), fused spot pass of the general-surface '
                                   'kernel k_trace_gen, L2 flushed before every launch; oracle: 4 fixed Newton steps, fast '
                                   'policy: early exit (2 per event on this lens)',
                           'rays': rays_total, 'events': a_events,
                           'fwd_bwd': {'value': a_events / (grad_ms * 1e-3), 'unit': 'asphere events/s', 'ms': grad_ms,
                                       'frac_fp32_peak_at_oracle_566_flop': a_events * 566 / (grad_ms * 1e-3) / 1e12 / peak_tflops,
                                       'frac_fp32_peak_at_executed_505_flop': a_events * 505 / (grad_ms * 1e-3) / 1e12 / peak_tflops},
                           'forward_sweep': {'value': a_events / (eval_ms * 1e-3), 'unit': 'asphere events/s', 'ms': eval_ms},
                           'ok_fraction': float(a_mom[..., -1].sum()) / rays_total}
            del a_args, a_ext, a_mom
        except Exception as exc:
            print(f'[bench] asphere timing failed: {exc}', file=sys.stderr)

    if graphed is not None:
        h2d, d2h = graphed.h2d_bytes, graphed.d2h_bytes
    else:
        h2d = sum(v.numel() * 4 for v in host_lens.values())
        d2h = 4 + (2 * S + lens.nd.shape[1]) * 4
    # the graphed and the eager API must agree
    check_rms, _ = e2e_eager_step()
    got_rms, _ = e2e_step()
    assert abs(check_rms - got_rms) <= 1e-6 * abs(check_rms), (check_rms, got_rms)

    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': 'events/s', 'n_gpus': world,
                'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step,
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
                'data': 'synthetic',
                'config': {'workload': workload_name(n_theta), 'lens': 'double_gauss_50mm_f3',
                           'surfaces': S, 'fields': N_FIELDS, 'wavelengths': len(WAVELENGTHS),
                           'rays': rays_total, 'events_per_step': events_total,
                           'rays_per_gpu': rays_total // world,
                           'parallelism': f'pupil-sharded dp{world}', 'collective': collective,
                           'l2': 'flushed between timed steps (256 MiB write)',
                           'launch': 'cuda_graph' if graph is not None else 'eager',
                           'step': 'RayTracer.spot_rms_and_grads: staging -> chief rays -> fused trace+adjoint -> '
                                   'reduce -> (exchange) -> finalize -> chain rule; rms + d rms/d(c,t,nd,v)',
                           'arith': 'guarded'},
                'clocks': dict(clocks.summary(), under_sustained_load=sustained['clocks']),
                'e2e': {'value': e2e_value, 'unit': 'events/s', 'h2d_bytes_per_step': h2d,
                        'd2h_bytes_per_step': d2h, 'ms_per_step': e2e_secs / args.steps * 1e3,
                        'api': 'GraphedSpotStep (CUDA graph of the public RayTracer.spot_rms_and_grads path: one pinned H2D copy, the kernel sequence, one D2H copy)'
                               if graphed is not None else 'RayTracer.spot_rms (eager)',
                        'eager_api_value': events_total * eager_steps / eager_secs,
                        'eager_api_ms_per_step': eager_secs / eager_steps * 1e3,
                        'drop_in_api_value': (events_total * eager_steps / drop_in_secs) if drop_in_secs else None,
                        'drop_in_api': 'the reference\'s own sequence, unchanged: trace_rays (materialises [B,F,P,W]) -> compute_rms2d -> backward; compute_rms2d recognises untouched trace outputs and runs the fused pass on their inputs'},
                'step_ms': step_stats, 'sustained': sustained,
                'gpu_launches': launches_per_step * args.steps,
                'roofline': roofline, 'forward': forward, 'penalty': penalty_row, 'batched_lenses': batched_row,
                'asphere': asphere_row}
        if value < 0.97 * e2e_value:      # device-timed slower than host-timed end to end: timed wait (rank skew)
            line['warning'] = ('value < e2e.value: the device-timed steps include waiting for the slowest rank '
                               '(see step_ms min / median / max)')
            print(f"[bench] warning: {line['warning']}", file=sys.stderr)
        if world == 1 and not args.no_cpu_baseline:
            base = cpu_arm(3, 1)
            line['cpu_baseline'] = {k: base[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
        print(json.dumps(line), flush=True)
    faulthandler.cancel_dump_traceback_later()
    if world > 1:
        # tearing the NCCL communicator down while captured graphs still reference it can block
        # forever: drop the graphs, meet at a barrier and leave without the destructor chain
        graph = graphed = k_graph = None
        barrier()
        if exchange is not None:
            status, epoch = exchange.status()
            if status != 0:
                print(f'[bench rank {rank}] peer exchange timed out (epoch {epoch})', file=sys.stderr)
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == '__main__':
    main()
