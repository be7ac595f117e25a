"""CPU, world_size 2 over gloo: the row-sharding + single all-gather plumbing reproduces the unsharded run."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fake_sampler(x, y, noise):
    # any row-independent map stands in for the per-image DDNM trajectory
    acc = x.clone()
    for k in range(noise.shape[0]):
        acc = acc * 0.9 + noise[k] * 0.1 + y.mean(dim=1).reshape(-1, 1, 1, 1)
    return acc, acc * 2


def _worker(rank, world, port, B, q):
    sys.path.insert(0, ROOT)
    from ddnm_b200.parallel import sharded_sample
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, 3, 8, 8, generator=g)
    y = torch.randn(B, 12, generator=g)
    nz = torch.randn(4, B, 3, 8, 8, generator=g)
    a, b = sharded_sample(_fake_sampler, x, y, nz)
    ra, rb = _fake_sampler(x, y, nz)
    q.put((rank, torch.equal(a, ra) and torch.equal(b, rb), tuple(a.shape)))
    dist.barrier()
    dist.destroy_process_group()


def _run(world, B, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    [p.start() for p in ps]
    res = [q.get(timeout=120) for _ in range(world)]
    [p.join(timeout=60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    assert all(ok for _, ok, _ in res), res
    assert all(shape[0] == B for _, _, shape in res)


def test_sharded_equals_unsharded_even_split():
    _run(2, 8, 29541)


def test_sharded_equals_unsharded_ragged_split():
    _run(2, 7, 29542)


def test_single_rank_passthrough():
    sys.path.insert(0, ROOT)
    from ddnm_b200.parallel import sharded_sample
    g = torch.Generator().manual_seed(0)
    x, y, nz = torch.randn(3, 3, 8, 8, generator=g), torch.randn(3, 12, generator=g), torch.randn(2, 3, 3, 8, 8, generator=g)
    a, b = sharded_sample(_fake_sampler, x, y, nz)
    ra, rb = _fake_sampler(x, y, nz)
    assert torch.equal(a, ra) and torch.equal(b, rb)
