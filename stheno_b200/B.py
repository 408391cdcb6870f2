"""Minimal stand-in for the ``lab`` namespace ``B`` that Stheno users touch: the global Cholesky jitter
``B.epsilon`` (``README.md:820-830``), ``B.dense`` and ``B.to_numpy``."""
import numpy as np
import torch

#: Diagonal jitter added before every dense Cholesky (``B.reg``).  Reference default 1e-12; the reference's
#: examples raise it to 1e-6 for float32 (``README.md:983``).
epsilon = 1e-12

#: Arithmetic of the Cholesky trailing update for float64 problems.  "fp64" (default): fp64 tensor cores (DMMA), meets
#: the 1e-10 parity bar.  "tf32x3" (opt-in): fp32 panel copy + 3xTF32 products on the tcgen05 tensor cores, fp64
#: accumulation into the matrix -- ~2x faster, agrees to ~1e-6 relative.
precision = "fp64"

pi = np.pi
log_2_pi = float(np.log(2 * np.pi))


def dense(a):
    """Strip matrix structure: a :class:`stheno_b200.matrix.AbstractMatrix` becomes a plain tensor/array."""
    from .matrix import AbstractMatrix

    if isinstance(a, AbstractMatrix):
        return a.dense_out()
    return a


def to_numpy(a):
    a = dense(a)
    if isinstance(a, torch.Tensor):
        return a.detach().cpu().numpy()
    if isinstance(a, (tuple, list)):
        return type(a)(to_numpy(ai) for ai in a)
    return np.asarray(a)
