"""Batch-sharded logpdf over a process group: world_size-2 gloo on CPU (host logic), NCCL on the GPU box."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import gp_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, backend, use_cuda, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist

    import stheno_b200 as S
    from stheno_b200.dist import shard_bounds, sharded_logpdf

    if use_cuda:
        torch.cuda.set_device(rank % torch.cuda.device_count())
    else:
        import pytest as _pt

        from tests import _cpu_backend

        mpatch = _pt.MonkeyPatch()
        _cpu_backend.install(mpatch)
    dist.init_process_group(backend, rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    B, n, d = 5, 20, 2
    x = rng.standard_normal((B, n, d))
    y = rng.standard_normal((B, n, 1))
    make = lambda xl: S.GP(S.EQ().stretch(1.3))(xl, 0.2)
    total = sharded_logpdf(make, x, y, reduce="sum")
    vec = sharded_logpdf(make, x, y, reduce="gather")
    lo, hi = shard_bounds(B, world, rank)
    loc = sharded_logpdf(make, x[lo:hi], y[lo:hi], reduce="sum", presharded=True)
    q.put((rank, float(total), np.asarray(vec), float(loc)))
    dist.barrier()
    dist.destroy_process_group()


def _run(world, backend, use_cuda, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, backend, use_cuda, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(0)
    x = rng.standard_normal((5, 20, 2))
    y = rng.standard_normal((5, 20, 1))
    ref = np.array([O.fdd_logpdf(("stretched", 1.3, ("eq",)), x[b], 0.2, y[b]) for b in range(5)])
    for rank, total, vec, loc in res:
        np.testing.assert_allclose(total, ref.sum(), rtol=1e-10)
        np.testing.assert_allclose(vec, ref, rtol=1e-10)
        np.testing.assert_allclose(loc, ref.sum(), rtol=1e-10)


def test_shard_bounds():
    from stheno_b200.dist import shard_bounds

    assert [shard_bounds(512, 8, r) for r in (0, 7)] == [(0, 64), (448, 512)]
    b = [shard_bounds(5, 2, r) for r in range(2)]
    assert b == [(0, 3), (3, 5)]
    assert [shard_bounds(3, 4, r) for r in range(4)] == [(0, 1), (1, 2), (2, 3), (3, 3)]


def test_sharded_logpdf_gloo_world2():
    _run(2, "gloo", False, 29611)


@pytest.mark.gpu
def test_sharded_logpdf_nccl():
    n = torch.cuda.device_count()
    # world 2 even on a 1-GPU box (both ranks share the device): exercises the NCCL all-reduce path
    if n < 2:
        pytest.skip("needs 2 GPUs for NCCL (two ranks on one device are rejected by NCCL)")
    _run(2, "nccl", True, 29612)
