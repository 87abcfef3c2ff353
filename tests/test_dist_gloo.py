"""CPU, world_size 2 over gloo: the multi-rank host logic of the env path -- shard ownership,
the max-over-ranks timing reduction and the whole-job aggregation bench.py performs."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, total_envs, q):
    sys.path.insert(0, ROOT)
    from marllb_b200.shard import owner_of, shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, n = shard_range(total_envs, rank, world)
    # every env this rank claims is owned by it, and ownership is a partition
    mine = torch.zeros(total_envs, dtype=torch.int64)
    mine[first:first + n] = 1
    assert all(owner_of(e, total_envs, world) == rank for e in range(first, first + n))
    dist.all_reduce(mine, op=dist.ReduceOp.SUM)
    assert bool((mine == 1).all())
    # bench.py semantics: time = max over ranks, value = units of all ranks / that time
    ms = torch.tensor([10.0 + 5.0 * rank], dtype=torch.float64)
    units = torch.tensor([float(n * 7)], dtype=torch.float64)
    dist.barrier()
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.all_reduce(units, op=dist.ReduceOp.SUM)
    q.put((rank, first, n, float(ms.item()), float(units.item())))
    dist.destroy_process_group()


@pytest.mark.parametrize("total_envs", [16, 1001])
def test_two_rank_sharding_and_reductions(total_envs):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total_envs, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, f0, n0, ms0, u0), (r1, f1, n1, ms1, u1) = out
    assert f0 == 0 and f1 == n0 and n0 + n1 == total_envs and abs(n0 - n1) <= 1
    assert ms0 == ms1 == 15.0 and u0 == u1 == total_envs * 7


def test_shard_range_properties():
    from marllb_b200.shard import owner_of, shard_range
    for total, world in ((8, 8), (1048576, 8), (13, 4), (3, 8)):
        seen = 0
        for r in range(world):
            first, n = shard_range(total, r, world)
            assert first == seen
            seen += n
            for e in (first, first + n - 1):
                if n:
                    assert owner_of(e, total, world) == r
        assert seen == total
    with pytest.raises(ValueError):
        shard_range(8, 8, 8)


# ---- policy update: one flat gradient bucket per optimiser, all-reduced and divided by the world size
def _grad_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import numpy as np
    from marllb_b200.policy.nn import FlatBucket, Params, allreduce_mean_
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)
    W = torch.randn(6, 10, generator=g)
    b = torch.randn(6, generator=g)
    X = torch.randn(8, 10, generator=g)        # the GLOBAL batch; each rank owns half of it
    Y = torch.randn(8, 6, generator=g)
    P = Params("cpu")
    P.add("fc.weight", W)
    P.add("fc.bias", b)
    bucket = FlatBucket([P])
    assert P.p["fc.weight"].data_ptr() == bucket.flat_p.data_ptr()       # tensors are views of the bucket
    assert bucket.flat_p.numel() == 60 + 8                               # bias padded to a 16-byte multiple

    def grads(x, y):   # mean-squared-error gradient of y_hat = x W^T + b over the rows given
        d = 2.0 * (x @ W.T + b - y) / y.numel()
        return d.T @ x, d.sum(0)

    lo, hi = rank * 4, rank * 4 + 4
    gw, gb = grads(X[lo:hi], Y[lo:hi])
    P.g["fc.weight"].copy_(gw)
    P.g["fc.bias"].copy_(gb)
    allreduce_mean_(bucket.flat_g)
    gw_all, gb_all = grads(X, Y)
    ok = torch.allclose(P.g["fc.weight"], gw_all, rtol=1e-5, atol=1e-7) and \
        torch.allclose(P.g["fc.bias"], gb_all, rtol=1e-5, atol=1e-7) and float(bucket.flat_g[66:].abs().sum()) == 0.0
    q.put((rank, bool(ok), np.asarray(bucket.flat_g).tobytes()))
    dist.destroy_process_group()


def test_flat_bucket_allreduce_equals_global_batch_gradient():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out[0][1] and out[1][1]
    assert out[0][2] == out[1][2]          # both ranks hold bit-identical reduced gradients
