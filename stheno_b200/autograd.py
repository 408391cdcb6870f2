"""Differentiable log-marginal likelihood: ``torch.autograd.Function`` around the fused K1 -> K2 -> K4 path with an
ANALYTIC backward:

    d logpdf / dK = 1/2 (alpha alpha^T - K^-1),   alpha = K^-1 (y - mu)

``K^-1 = L^-T L^-1`` comes from one tensor-core TRSM on the identity and one SYRK; the contraction with ``dK/dtheta``
(kernel scales, length scales through the pre-stretched inputs, inputs, noise) happens inside the K1-backward kernel
(``csrc/kernel_matrix_bwd.cu``) so no ``n x n`` gradient tensor per hyper-parameter is ever formed.  The reference gets
these gradients from torch autograd through ``exp`` / ``cholesky`` / ``triangular_solve``
(``readme_example13_optimisation_torch.py:46-53``)."""
import ctypes

import torch

from . import _lib, ops

__all__ = ["kernel_logpdf", "dense_logpdf", "kernel_matrix_grad"]


def _bwd_kernel(flat, xg, G, n):
    """``(term_sum [B, T], grad_xg like xg, diag [B, n])`` from the K1-backward kernel."""
    Bn, d = xg.shape[1], xg.shape[3]
    term_sum = torch.zeros(Bn, _lib.GPK_MAX_TERMS, dtype=xg.dtype, device=xg.device)
    grad_xg = torch.zeros_like(xg)
    diag = torch.zeros(Bn, n, dtype=xg.dtype, device=xg.device)
    desc = flat.desc()
    rc = ops._fn("gpk_kernel_matrix_bwd", xg.dtype)(
        ctypes.byref(desc), ops._ptr(xg), xg.stride(0), xg.stride(1), n, d, ops._ptr(G), G.stride(1), G.stride(0),
        ops._ptr(term_sum), ops._ptr(grad_xg), ops._ptr(diag), Bn, ops._stream(),
    )
    _lib.check(rc, "gpk_kernel_matrix_bwd")
    return term_sum, grad_xg, diag


class _KernelLogpdf(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coefs, xg, noise_scalar, noise_vec, rhs_t, structure, jitter):
        # coefs [T], xg [G, B, n, d], noise_scalar [] , noise_vec [B, n] or None, rhs_t [B, k, n]
        flat = ops.FlatKernel([(float(c), fs) for c, fs in zip(coefs.tolist(), structure)], xg.shape[0])
        ch = ops.chol_from_kernel(flat, xg.detach().contiguous(), noise_scalar=float(noise_scalar),
                                  noise_vec=None if noise_vec is None else noise_vec.detach(), jitter=jitter,
                                  rhs_t=rhs_t.detach())
        ctx.ch, ctx.flat = ch, flat
        ctx.xg = xg.detach().contiguous()
        ctx.has_nv = noise_vec is not None
        return ch.logpdf()

    @staticmethod
    def backward(ctx, g):
        ch, flat, xg = ctx.ch, ctx.flat, ctx.xg
        alpha, Gm = _alpha_and_G(ch, g)
        term_sum, grad_xg, diag = _bwd_kernel(flat, xg, Gm, ch.n)
        T = len(flat.terms)
        grad_coefs = term_sum[:, :T].sum(0)
        grad_noise_scalar = diag.sum()
        grad_noise_vec = diag if ctx.has_nv else None
        grad_rhs = -g.unsqueeze(-1) * alpha
        return grad_coefs, grad_xg, grad_noise_scalar, grad_noise_vec, grad_rhs, None, None


def _alpha_and_G(ch, g):
    """``alpha = K^-1 ybar`` (rows) and ``G = d(sum_c g_c logpdf_c)/dK = 1/2 (sum_c g_c alpha_c alpha_c^T - (sum g) K^-1)``
    as a full symmetric padded ``[B, n_pad, n_pad]`` tensor, from the factor ``ch`` with fused right-hand sides."""
    with ops.product_slices(7):  # gradients are checked at 1e-8, not at the 1e-10 bar of the forward quantities
        return _alpha_and_G_impl(ch, g)


def _alpha_and_G_impl(ch, g):
    Bn, n, n_pad, k = ch.batch, ch.n, ch.n_pad, ch.k
    dtype, dev = ch.dtype, ch.device
    arows = ch.new_rows(k)
    arows[:, :k, :n] = ch.rhs_half()
    ch.solve_rows_t_(arows)
    alpha = arows[:, :k, :n]
    V = torch.zeros(Bn, n_pad, n_pad, dtype=dtype, device=dev)
    V.diagonal(dim1=1, dim2=2).fill_(1.0)
    ch.solve_rows_(V)  # rows (L^-1 e_r)^T, i.e. V = L^-T ; K^-1 = V V^T
    Gm = torch.empty(Bn, n_pad, n_pad, dtype=dtype, device=dev)
    kp = ops.round_up(k, 16)
    for b in range(Bn):
        s = float(g[b].sum())
        ops.gemm_nt(V[b : b + 1], V[b : b + 1], Gm[b : b + 1], alpha=-0.5 * s, beta=0.0, lower=True)
        A = torch.zeros(1, n_pad, kp, dtype=dtype, device=dev)
        Bm = torch.zeros(1, n_pad, kp, dtype=dtype, device=dev)
        A[0, :n, :k] = (alpha[b] * (0.5 * g[b]).unsqueeze(-1)).t()
        Bm[0, :n, :k] = alpha[b].t()
        ops.gemm_nt(A, Bm, Gm[b : b + 1], alpha=1.0, beta=1.0, lower=True)
    del V
    ops.symmetrize_(Gm, n_pad)
    return alpha, Gm


class _DenseLogpdf(torch.autograd.Function):
    """``logpdf`` of ``N(0, K + jitter I)`` for an explicit dense ``K [B, n, n]`` with a gradient w.r.t. ``K`` itself --
    the route for covariances assembled from several kernel matrices (multi-output joints, ``stheno/mo``)."""

    @staticmethod
    def forward(ctx, K, rhs_t, jitter):
        ch = ops.chol_from_dense(K.detach(), jitter=jitter, rhs_t=rhs_t.detach())
        ctx.ch = ch
        return ch.logpdf()

    @staticmethod
    def backward(ctx, g):
        ch = ctx.ch
        alpha, Gm = _alpha_and_G(ch, g)
        n = ch.n
        return Gm[:, :n, :n].contiguous(), -g.unsqueeze(-1) * alpha, None


def dense_logpdf(K, rhs_t, jitter):
    return _DenseLogpdf.apply(K, rhs_t, jitter)


class _KernelMatrix(torch.autograd.Function):
    """Differentiable ``k(x, x)`` (same points) built by K1; backward = K1-backward on the symmetrised upstream gradient."""

    @staticmethod
    def forward(ctx, coefs, xg, structure):
        flat = ops.FlatKernel([(float(c), fs) for c, fs in zip(coefs.tolist(), structure)], xg.shape[0])
        ctx.flat, ctx.xg = flat, xg.detach().contiguous()
        return ops.kernel_matrix(flat, ctx.xg)

    @staticmethod
    def backward(ctx, G):
        flat, xg = ctx.flat, ctx.xg
        n = xg.shape[2]
        Gs = (0.5 * (G + G.transpose(1, 2))).contiguous()  # K is symmetric: only the symmetric part of G matters
        term_sum, grad_xg, _ = _bwd_kernel(flat, xg, Gs, n)
        return term_sum[:, : len(flat.terms)].sum(0), grad_xg, None


def kernel_matrix_grad(flat, xg):
    """``k(x, x) [B, n, n]`` with an autograd graph to the kernel's tensor hyper-parameters and to ``xg``."""
    raw = getattr(flat, "coef_raw", None) or [c for c, _ in flat.terms]
    coefs = torch.stack([
        (c if isinstance(c, torch.Tensor) else torch.tensor(float(c))).to(device=xg.device, dtype=xg.dtype).reshape(())
        for c in raw
    ])
    return _KernelMatrix.apply(coefs, xg, [fs for _, fs in flat.terms])


def kernel_logpdf(coefs, xg, noise_scalar, noise_vec, rhs_t, structure, jitter):
    """Differentiable ``logpdf`` ``[B, k]`` of ``N(0, sum_t coefs[t] prod phi(xg) + noise + jitter I)`` at ``rhs_t``."""
    return _KernelLogpdf.apply(coefs, xg, noise_scalar, noise_vec, rhs_t, structure, jitter)
