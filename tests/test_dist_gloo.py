"""N>1 host path on CPU (gloo, world_size 2): problems shard by batch index with
no exchange; the only collective is the all-reduce of the gradient scalars."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib
    par = importlib.import_module("differentiable-ilqr_b200.parallel")
    B = 10
    lo, hi = par.shard_range(B, rank, world)
    # fake per-shard gradient sums: theta-grad is a batch SUM, loss a global mean
    x = torch.arange(B, dtype=torch.float64)
    local = torch.stack((x[lo:hi].sum(), (x[lo:hi] ** 2).sum() / B))
    tot = par.allreduce_sum_(local.clone())
    q.put((rank, lo, hi, tot.tolist()))
    dist.destroy_process_group()


def test_shard_and_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    out = sorted(q.get(timeout=120) for _ in ps)
    [p.join(timeout=60) for p in ps]
    assert [(o[1], o[2]) for o in out] == [(0, 5), (5, 10)]
    x = torch.arange(10, dtype=torch.float64)
    want = [float(x.sum()), float((x ** 2).sum() / 10)]
    for o in out:
        assert o[3] == want


def test_shard_range_covers_batch_exactly():
    import importlib
    sys.path.insert(0, ROOT)
    par = importlib.import_module("differentiable-ilqr_b200.parallel")
    for B in (1, 7, 64, 65536, 100003):
        for W in (1, 2, 3, 8):
            r = [par.shard_range(B, k, W) for k in range(W)]
            assert r[0][0] == 0 and r[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1
