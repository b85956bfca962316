"""N>1 path on CPU: world_size-2 gloo processes shard a batch, run a stand-in per-crop function on their shards with no
collective, and (optionally) gather; the result must equal the single-process result for even, ragged and tiny
batches (a rank may own an empty shard)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from image_restoration_b200.sharding import gather_shards, micro_batches, run_sharded, shard_bounds


def per_crop(x):           # independent per crop, like the forward pass
    return x * 2 + x.flatten(1).sum(1).view(-1, 1, 1, 1)


def test_shard_bounds_cover_everything():
    for n in (0, 1, 2, 7, 64, 4096, 4099):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    assert micro_batches(0, 130, 64) == [(0, 64), (64, 128), (128, 130)]
    assert shard_bounds(4096, 3, 8) == (1536, 2048)


def _worker(rank, world, port, n, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)
    batch = torch.randn(n, 3, 4, 6)
    local = run_sharded(per_crop, batch, rank, world, micro_batch=3)
    lo, hi = shard_bounds(n, rank, world)
    ok = local.shape[0] == hi - lo and torch.equal(local, per_crop(batch)[lo:hi])
    full = gather_shards(local, n)
    ok = ok and torch.equal(full, per_crop(batch))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize('n', [8, 7, 1])
def test_two_rank_gloo_sharding(n):
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def _grad_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from image_restoration_b200.grad_sync import GradAllReducer
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4), torch.nn.Linear(4, 4))
    for p in net[3].parameters():          # an unused branch: no gradient (find_unused_parameters semantics)
        p.grad = None
    x = torch.full((3, 8), float(rank + 1))
    net[2](net[1](net[0](x))).sum().backward()
    red = GradAllReducer(net.parameters(), bucket_mb=0.0001)      # tiny buckets: several all-reduces
    assert len(red.buckets) > 1
    red.sync(average=True)
    q.put((rank, [p.grad.clone().numpy() for p in net.parameters()]))   # plain arrays: tensor fds die with the worker
    dist.destroy_process_group()


def test_gradient_allreduce_world2_gloo():
    """Flat-buffer bucketed gradient all-reduce (the DDP exchange of base_model.py:70-73): both ranks end with the mean
    of their gradients; parameters without a gradient get zeros."""
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # expected: average of the two single-rank gradients
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4), torch.nn.Linear(4, 4))
    exp = None
    for rank in range(2):
        net.zero_grad(set_to_none=True)
        net[2](net[1](net[0](torch.full((3, 8), float(rank + 1))))).sum().backward()
        g = [p.grad.clone() if p.grad is not None else torch.zeros_like(p) for p in net.parameters()]
        exp = g if exp is None else [a + b for a, b in zip(exp, g)]
    exp = [e / 2 for e in exp]
    for rank in range(2):
        for got, e in zip(res[rank], exp):
            assert torch.allclose(torch.from_numpy(got), e, atol=1e-6)
