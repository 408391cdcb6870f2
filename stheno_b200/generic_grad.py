"""Differentiable routes for the two places that have no analytic backward kernel: the sparse ELBO
(``PseudoObs*.elbo``, ``stheno/model/observations.py:279-336``) and the Woodbury log-pdf of ``Linear`` models.

The reference differentiates everything by generic autograd through its backend ops
(``readme_example13_optimisation_torch.py:46-53``).  ``f(x).logpdf(y)`` has an analytic backward here
(``autograd.py``: K1-backward kernel); these two functions do not, and round 1 returned results whose graph was silently
cut by the raw-pointer kernels.  When -- and only when -- a hyper-parameter, input, noise or observation
that feeds them requires grad, they are evaluated by THIS module: the same arithmetic restated with differentiable torch
ops on the same device, so ``elbo.backward()`` is correct for every parameter.  Nothing here is used when no gradient is
requested, and nothing here is on the benchmarked path."""
import math

import numpy as np
import torch

from ._util import batch_flatten

__all__ = ["kernel_needs_grad", "kernel_torch", "kernel_diag_torch", "sparse_compute_torch", "woodbury_terms_torch"]


def _t(v, like):
    if isinstance(v, torch.Tensor):
        return v.to(device=like.device, dtype=like.dtype)
    return torch.as_tensor(np.asarray(v, np.float64), device=like.device, dtype=like.dtype)


def kernel_needs_grad(k):
    """True if a coefficient or length scale of the (flattenable) kernel expression ``k`` is a tensor that requires grad."""
    terms = k.flat_terms() if k is not None else None
    if not terms:
        return False
    for coef, fs in terms:
        if isinstance(coef, torch.Tensor) and coef.requires_grad:
            return True
        for f in fs:
            s = f[1]
            if isinstance(s, torch.Tensor) and s.requires_grad:
                return True
    return False


def _factor(kind, xs, ys, elwise, param=None):
    if kind == "linear":
        return (xs * ys).sum(-1) if elwise else xs @ ys.transpose(-1, -2)
    if kind == "one":
        shp = xs.shape[:-1] if elwise else xs.shape[:-1] + (ys.shape[-2],)
        return torch.ones(shp, dtype=xs.dtype, device=xs.device)
    if elwise:
        d2 = ((xs - ys) ** 2).sum(-1)
    else:
        d2 = ((xs.unsqueeze(-2) - ys.unsqueeze(-3)) ** 2).sum(-1)  # direct form, like K1
    if kind == "eq":
        return torch.exp(-0.5 * d2)
    if kind == "delta":
        return (d2 < 1e-10).to(xs.dtype)
    r = torch.sqrt(torch.clamp_min(d2, 1e-30))
    if kind == "matern12":
        return torch.exp(-r)
    if kind == "matern32":
        s = math.sqrt(3.0) * r
        return (1 + s) * torch.exp(-s)
    if kind == "matern52":
        s = math.sqrt(5.0) * r
        return (1 + s + 5.0 / 3.0 * d2) * torch.exp(-s)
    if kind == "rq":
        return torch.exp(-param * torch.log1p(d2 / (2.0 * param)))
    raise NotImplementedError(f"no differentiable restatement of kernel kind {kind!r}")


def _eval(k, x, y, elwise):
    terms = k.flat_terms()
    if terms is None:
        raise NotImplementedError(
            f"gradients through {type(k).__name__} inside a sparse approximation are not implemented "
            "(only sums / products / stretches of the elementary kernels)")
    out = None
    for coef, fs in terms:
        t = None
        for fac in fs:
            kind, s = fac[0], fac[1]
            xs, ys = (x, y) if s is None else (x / _t(s, x), y / _t(s, y))
            f = _factor(kind, xs, ys, elwise, fac[2] if len(fac) > 2 else None)
            t = f if t is None else t * f
        t = t * _t(coef, x)
        out = t if out is None else out + t
    if out is None:
        shp = x.shape[:-1] if elwise else x.shape[:-1] + (y.shape[-2],)
        out = torch.zeros(shp, dtype=x.dtype, device=x.device)
    return out


def kernel_torch(k, x, y):
    """``k(x, y)`` ``[..., n, m]`` with an autograd graph to the kernel's tensor hyper-parameters and to ``x``, ``y``."""
    return _eval(k, x, y, False)


def kernel_diag_torch(k, x):
    return _eval(k, x, x, True)


def sparse_compute_torch(method, k_z, k_zx, k_x, z, x, kn, noise_z_diag, y_bar, mean_z, eps):
    """``AbstractPseudoObservations._compute`` in differentiable torch ops.  ``z [.., m, d]``, ``x [.., n, d]``, ``kn [.., n]``,
    ``y_bar [.., n, 1]``, ``mean_z [.., m, 1]``.  Returns ``(K_z, LAL, mu, elbo)``."""
    m = z.shape[-2]
    eye = torch.eye(m, dtype=z.dtype, device=z.device)
    K_z = kernel_torch(k_z, z, z)  # :286
    if noise_z_diag is not None:
        K_z = K_z + torch.diag_embed(noise_z_diag)
    L_z = torch.linalg.cholesky(K_z + eps * eye)  # :300
    K_zx = kernel_torch(k_zx, z, x)  # :285
    W = torch.linalg.solve_triangular(L_z, K_zx, upper=False)  # :301
    trace_part = 0.0
    if method in ("vfe", "fitc"):
        corr = kernel_diag_torch(k_x, x) - (W * W).sum(-2)  # :304-306
        if method == "vfe":
            trace_part = (corr / kn).sum(-1)  # :308-310
        else:
            kn = kn + corr  # :311-313
    Ws = W / kn.unsqueeze(-2)
    A = eye + Ws @ W.transpose(-1, -2)  # :322
    L_A = torch.linalg.cholesky(A + eps * eye)
    prod = Ws @ y_bar  # :327
    t = torch.linalg.solve_triangular(L_A, prod, upper=False)
    sol = torch.linalg.solve_triangular(L_A.transpose(-1, -2), t, upper=True)
    mu = mean_z + L_z @ sol  # :329
    LAL = L_z @ A @ L_z.transpose(-1, -2)  # :323
    det_part = torch.log(2 * math.pi * kn).sum(-1) + 2 * torch.log(torch.diagonal(L_A, dim1=-2, dim2=-1)).sum(-1)  # :334
    iqf_part = (y_bar[..., 0] ** 2 / kn).sum(-1) - (t * t).sum((-1, -2))  # :335
    elbo = -0.5 * (det_part + iqf_part + trace_part)  # :336
    return K_z, LAL, mu, elbo


def woodbury_terms_torch(U, d, diff, eps):
    """``(logdet, diag(diff^T (D + U U^T)^-1 diff))`` by the determinant / inversion lemmas in differentiable torch ops.
    ``U [.., n, r]``, ``d [.., n]``, ``diff [.., n, k]``."""
    r = U.shape[-1]
    dinv = 1.0 / d.unsqueeze(-1)
    S = torch.eye(r, dtype=U.dtype, device=U.device) + U.transpose(-1, -2) @ (U * dinv)
    L = torch.linalg.cholesky(S + eps * torch.eye(r, dtype=U.dtype, device=U.device))
    logdet = torch.log(d).sum(-1) + 2 * torch.log(torch.diagonal(L, dim1=-2, dim2=-1)).sum(-1)
    ub = U.transpose(-1, -2) @ (diff * dinv)
    h = torch.linalg.solve_triangular(L, ub, upper=False)
    q = (diff * diff * dinv).sum(-2) - (h * h).sum(-2)
    return logdet, q
