"""Multi-GPU: shard the BATCH axis of independent GP problems across ranks (one process per GPU, ``torch.distributed``
over NCCL / NVLink), with ONE collective on the data path: the all-reduce of the scalar log-marginal
(BASELINE.json north_star).  A single dense Cholesky does not shard -- single-GP runs are
"replicas only".

The reference has no multi-device code at all; the batched semantics it does have (leading batch dimensions broadcast
through everything, ``tests/model/test_cases.py:134-176``) are what gets sharded here.
"""
import torch
import torch.distributed as dist

from ._util import from_dev, origin_of, to_dev

__all__ = ["shard_bounds", "sharded_logpdf"]


def shard_bounds(batch, world, rank):
    """Contiguous block ``[lo, hi)`` of the batch owned by ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(int(batch), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def sharded_logpdf(make_fdd, x, y, *, reduce="sum", group=None, presharded=False):
    """Log-marginal likelihood of ``B`` independent GPs, batch-sharded over the process group.

    ``make_fdd(x_local) -> FDD`` builds the (batched) finite-dimensional distribution for a slice of the inputs
    (e.g. ``lambda xl: GP(EQ())(xl, 0.1)``); ``x [B, n, d]``, ``y [B, n, 1]`` are the full arrays on every rank
    (each rank touches only its slice) or, with ``presharded=True``, already the local slices.

    ``reduce="sum"``: returns ``sum_b logpdf_b`` -- local sum, then ONE ``all_reduce(SUM)`` of a single scalar.
    ``reduce="gather"``: returns the ``(B,)`` vector on every rank (``all_gather`` of the local values).
    Without an initialised process group this is the plain single-process computation."""
    distributed = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if distributed else 1
    rank = dist.get_rank(group) if distributed else 0
    org = origin_of(x)
    if presharded:
        xl, yl = x, y
    else:
        lo, hi = shard_bounds(x.shape[0], world, rank)
        xl, yl = x[lo:hi], y[lo:hi]
    if xl.shape[0] > 0:
        lp = to_dev(make_fdd(xl).logpdf(yl)).reshape(-1)
    else:
        lp = to_dev(torch.zeros(0, dtype=torch.float64))
    if reduce == "sum":
        total = lp.sum().reshape(1)
        if distributed:
            dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
        return from_dev(total[0], org)
    if reduce == "gather":
        if not distributed:
            return from_dev(lp, org)
        sizes = [shard_bounds(x.shape[0], world, r) if not presharded else None for r in range(world)]
        if presharded:
            n_loc = torch.tensor([lp.numel()], device=lp.device)
            counts = [torch.zeros_like(n_loc) for _ in range(world)]
            dist.all_gather(counts, n_loc, group=group)
            lens = [int(c.item()) for c in counts]
        else:
            lens = [hi - lo for lo, hi in sizes]
        m = max(lens) if lens else 0
        pad = torch.zeros(m, dtype=lp.dtype, device=lp.device)
        pad[: lp.numel()] = lp
        parts = [torch.zeros_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        return from_dev(torch.cat([p[:l] for p, l in zip(parts, lens)]), org)
    raise ValueError(f"unknown reduce {reduce!r}")
