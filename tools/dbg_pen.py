import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_penalty_cpu import load_aggregate, KEYS
from tests.test_gpu_penalty import _inputs, _args, _loss
from oracle import trace_oracle as oracle
from torchoptics_b200 import ray_tracing_lite as rt
name = sys.argv[1] if len(sys.argv) > 1 else 'cooke_8x8'
rec, agg = load_aggregate(name)
n_seq = int(agg['n_seq'])
names = ('x', 'y', 'z', 'c', 't', 'mu')
for key_sel in ('theta_prime_norm',):
    i = _inputs(rec, 'cuda:0', broadcast=True)
    i['z'] = i['z'].expand(rec['out_ok'].shape).contiguous()
    for k in names: i[k] = i[k].clone().requires_grad_(True)
    out = rt.trace_skew(*_args(i), aggregate=True, arith='exact')
    cpu = _inputs(rec, 'cpu', broadcast=True, dtype=torch.float64)
    cpu['z'] = cpu['z'].expand(rec['out_ok'].shape).contiguous()
    for k in names: cpu[k] = cpu[k].clone().requires_grad_(True)
    with oracle.finite_penalty_gradients():
        ref = oracle.trace(*_args(cpu), True)
    def pen(o):
        if key_sel is None:
            return _loss(o, True, n_seq)[1] if o is ref else _loss(o, False, n_seq)[1]
        return torch.stack(o[6][key_sel]).sum() / n_seq
    got = torch.autograd.grad(pen(out), [i[k] for k in names])
    want = torch.autograd.grad(pen(ref), [cpu[k] for k in names])
    print('==', key_sel, 'gz got', got[2].sum().item(), 'want', want[2].sum().item(),
          'gt err', (got[4].cpu().double() - want[4]).abs().max().item(), 'gc err', (got[3].cpu().double() - want[3]).abs().max().item())
    d = (got[2].cpu().double() - want[2]).abs()
    flat = torch.argsort(d.reshape(-1), descending=True)[:6]
    for f in flat:
        idx = np.unravel_index(int(f), d.shape)
        th = [float(torch.stack(ref[6]['theta_norm'])[(k,) + idx]) * np.pi / 2 for k in range(len(ref[6]['theta_norm']))]
        thp = [float(torch.stack(ref[6]['theta_prime_norm'])[(k,) + idx]) * np.pi / 2 for k in range(len(th))]
        mine = [float(torch.stack(out[6]['theta_prime_norm'])[(k,) + idx]) * np.pi / 2 for k in range(len(th))]
        print('  ray', tuple(int(v) for v in idx), 'gz got %.6f want %.6f' % (got[2][idx].item(), want[2][idx].item()), 'x,y', cpu['x'][idx].item(), cpu['y'][idx].item(), 'thetap(rad)', np.round(thp, 6), 'mine', np.round(mine, 6))
