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

__all__ = ["kernel_logpdf"]


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
        Bn, n, n_pad, k = ch.batch, ch.n, ch.n_pad, ch.k
        dtype, dev = ch.dtype, ch.device
        # alpha rows: K^-1 ybar = L^-T (L^-1 ybar)
        arows = ch.new_rows(k)
        arows[:, :k, :n] = ch.rhs_half()
        ch.solve_rows_t_(arows)
        alpha = arows[:, :k, :n]  # [B, k, n]
        # K^-1 = V V^T with V = I L^-T (rows (L^-1 e_r)^T)
        V = torch.zeros(Bn, n_pad, n_pad, dtype=dtype, device=dev)
        V.diagonal(dim1=1, dim2=2).fill_(1.0)
        ch.solve_rows_(V)
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
        term_sum, grad_xg, diag = _bwd_kernel(flat, xg, Gm, n)
        T = len(flat.terms)
        grad_coefs = term_sum[:, :T].sum(0)
        grad_noise_scalar = diag.sum()
        grad_noise_vec = diag if ctx.has_nv else None
        grad_rhs = -g.unsqueeze(-1) * alpha
        return grad_coefs, grad_xg, grad_noise_scalar, grad_noise_vec, grad_rhs, None, None


def kernel_logpdf(coefs, xg, noise_scalar, noise_vec, rhs_t, structure, jitter):
    """Differentiable ``logpdf`` ``[B, k]`` of ``N(0, sum_t coefs[t] prod phi(xg) + noise + jitter I)`` at ``rhs_t``."""
    return _KernelLogpdf.apply(coefs, xg, noise_scalar, noise_vec, rhs_t, structure, jitter)
