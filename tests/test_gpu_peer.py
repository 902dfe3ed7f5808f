"""Peer-memory exchange (tl_peer_*): the SUM all-reduce of the moment sums as one kernel over
NVLink peer memory.  World 1 runs on any GPU box (same kernel, self window); world 2 needs two
GPUs and is skipped otherwise.  The N-rank result must equal the NCCL all-reduce and the
single-process spot pass."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def test_world1_exchange_is_identity_eager_and_graphed():
    from torchoptics_b200.peer import PeerExchange
    ex = PeerExchange(capacity=5000)
    g = torch.Generator(device='cpu').manual_seed(3)
    data = torch.randn(3408, dtype=torch.float64, generator=g).cuda()
    for _ in range(3):
        assert torch.equal(ex.all_reduce(data), data)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ex.all_reduce(data)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = ex.all_reduce(data)
    for i in range(5):
        data.add_(1.0)
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, data)
    status, epoch = ex.status()
    assert status == 0 and epoch == 3 + 1 + 5      # the capture itself launches nothing
    with pytest.raises(ValueError):
        ex.all_reduce(torch.zeros(6000, dtype=torch.float64, device='cuda'))
    ex.close()


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from torchoptics_b200 import RayTracer, prescriptions
    from torchoptics_b200.peer import PeerExchange
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = f'cuda:{rank}'
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device(dev))
    ex = PeerExchange(capacity=4096)
    # raw exchange against NCCL, repeated so that both slot parities and the epoch logic are used
    g = torch.Generator(device='cpu').manual_seed(10 + rank)
    for step in range(6):
        data = torch.randn(3408, dtype=torch.float64, generator=g).to(dev)
        got = ex.all_reduce(data)
        want = data.clone()
        dist.all_reduce(want)
        gathered = [torch.empty_like(data) for _ in range(world)]
        dist.all_gather(gathered, data)
        ordered = gathered[0].clone()
        for other in gathered[1:]:
            ordered += other                          # rank order: the exchange's summation order
        assert torch.equal(got, ordered), step
        assert torch.allclose(got, want, rtol=1e-14, atol=0)
    # the sharded spot pass through the exchange, graph-captured, against NCCL and against 1 rank
    specs, lens = prescriptions.double_gauss(dev)
    tracer = RayTracer(mode='circular', n_rays=(64, 64), rel_fields=(0., 0.5, 1.),
                       wavelengths=('C', 'd', 'F'), default_device=dev)

    def evaluate(shard, group):
        leaves = [getattr(lens, k).detach().clone().requires_grad_(True) for k in ('c', 't', 'nd', 'v')]
        from torchoptics_b200.lens_modeling import Lens
        res = tracer.loss_unsup(specs, Lens(lens.structure, *leaves), shard=shard, group=group)
        grads = torch.autograd.grad(res['loss_unsup'].sum(), leaves)     # spot pass + penalty pass
        return torch.cat([res['rms'].detach().reshape(-1), res['penalty'].detach().reshape(-1)] +
                         [x.reshape(-1) for x in grads])

    via_peer = evaluate((rank, world), ex)
    via_nccl = evaluate((rank, world), None)
    alone = evaluate((0, 1), None)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        evaluate((rank, world), ex)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        graphed = evaluate((rank, world), ex)
    for _ in range(4):
        graph.replay()
    torch.cuda.synchronize()
    status, _ = ex.status()
    np.savez(os.path.join(out_dir, f'rank{rank}.npz'), peer=via_peer.cpu().numpy(),
             nccl=via_nccl.cpu().numpy(), alone=alone.cpu().numpy(), graphed=graphed.cpu().numpy(),
             status=status)
    ex.close()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_world2_exchange_matches_nccl_and_single_rank(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    recs = [np.load(tmp_path / f'rank{r}.npz') for r in range(world)]
    for rec in recs:
        assert int(rec['status']) == 0
        np.testing.assert_array_equal(rec['peer'], rec['graphed'])
        np.testing.assert_allclose(rec['peer'], rec['nccl'], rtol=1e-6, atol=1e-9)
        scale = np.abs(rec['alone']).max()
        assert np.abs(rec['peer'] - rec['alone']).max() <= 1e-4 * scale
    np.testing.assert_array_equal(recs[0]['peer'], recs[1]['peer'])    # same bits on every rank


def _late_peer_worker(rank, world, port, out_dir):
    """Rank 1 skips one exchange: rank 0's wait runs into the spin limit (~4 s) and must say so LOUDLY."""
    import torch.distributed as dist
    from torchoptics_b200.peer import PeerExchange
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = f'cuda:{rank}'
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device(dev))
    ex = PeerExchange(capacity=1024)
    data = torch.full((700,), float(rank + 1), dtype=torch.float64, device=dev)
    first = ex.all_reduce(data).clone()                  # a healthy step: 1 + 2 on both ranks
    torch.cuda.synchronize()
    dist.barrier()
    rec = {'first': first.cpu().numpy()}
    if rank == 0:
        lonely = ex.all_reduce(data).clone()             # the peer never comes: spin limit, NaN, status 1
        torch.cuda.synchronize()
        rec['lonely'] = lonely.cpu().numpy()
        rec['status_after'] = ex.status()[0]
    dist.barrier()                                       # (rank 1 waits here while rank 0 spins)
    later = ex.all_reduce(data).clone()                  # both ranks call again: rank 0 stays poisoned
    torch.cuda.synchronize()
    rec['later'] = later.cpu().numpy()
    rec['status_end'] = ex.status()[0]
    np.savez(os.path.join(out_dir, f'late{rank}.npz'), **rec)
    dist.barrier()
    ex.close()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_world2_late_peer_poisons_the_result_and_stays_poisoned(tmp_path):
    """a wait that times out must not return a sum of stale slots.  The rank that
    waited in vain gets NaN in every element of that step AND of every later step, and status() == 1."""
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_late_peer_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (np.load(tmp_path / f'late{r}.npz') for r in range(world))
    np.testing.assert_array_equal(r0['first'], np.full(700, 3.0))
    np.testing.assert_array_equal(r1['first'], np.full(700, 3.0))
    assert np.isnan(r0['lonely']).all() and int(r0['status_after']) == 1
    assert np.isnan(r0['later']).all() and int(r0['status_end']) == 1      # sticky
