"""Tensor-level operations of the GP hot path, executed by the hand-written sm_100a kernels of ``libgpk``.

Everything here takes and returns CUDA ``torch`` tensors (PyTorch = device memory + streams; the arithmetic is
in ``csrc/*.cu``).  There is no CPU implementation: calling an op with a CPU tensor raises.

Storage conventions (see ``include/gpk.h``): row-major, batch-major ``[B, rows, cols]``; matrices that get
factorised live in a workspace ``W[B, n_pad + extra, n_pad]`` padded to multiples of 128 with the identity, the
``extra`` rows below the matrix carrying right-hand sides ``b^T`` that leave the factorisation as ``(L^-1 b)^T``.
"""
import ctypes
import math

import torch

from . import _lib
from ._lib import KIND, KM_LOWER, KM_PAD_IDENTITY, KM_PAD_ZERO, KM_SAME, GpkError, KernelDesc, check

__all__ = [
    "FlatKernel",
    "Chol",
    "round_up",
    "kernel_matrix",
    "kernel_diag",
    "chol_from_kernel",
    "chol_from_dense",
    "gemm_nt",
    "launch_count",
]

TILE = 128


def round_up(n, m=TILE):
    return (int(n) + m - 1) // m * m


def _suffix(dtype):
    if dtype == torch.float64:
        return "f64"
    if dtype == torch.float32:
        return "f32"
    raise TypeError(f"stheno_b200 computes in float64 or float32, got {dtype}")


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "stheno_b200.ops: CUDA tensor required -- the hot path runs only on the sm_100a kernels "
                "(there is no CPU fallback)"
            )


def _fn(name, dtype):
    return getattr(_lib.load(), f"{name}_{_suffix(dtype)}")


def launch_count(reset=False):
    lib = _lib.load()
    c = int(lib.gpk_launch_count())
    if reset:
        lib.gpk_launch_count_reset()
    return c


class FlatKernel:
    """A kernel flattened to a sum of products of elementary kernels on pre-stretched inputs:
    ``sum_t coef_t * prod_f kind_f(x / scale[group_f], y / scale[group_f])``.

    ``terms`` is a list of ``(coef: float, [(kind: str, group: int), ...])``."""

    def __init__(self, terms, n_groups):
        # a factor is (kind, group) or (kind, group, param) -- param = the shape parameter of kinds that have one (rq: alpha)
        self.terms = [(float(c), [tuple([str(f[0]), int(f[1])] + [float(v) for v in f[2:3]]) for f in fs]) for c, fs in terms]
        self.n_groups = max(int(n_groups), 1)
        if len(self.terms) > _lib.GPK_MAX_TERMS:
            raise GpkError(f"kernel expands to {len(self.terms)} product terms (limit {_lib.GPK_MAX_TERMS})")
        if sum(len(fs) for _, fs in self.terms) > _lib.GPK_MAX_FACTORS:
            raise GpkError(f"kernel has more than {_lib.GPK_MAX_FACTORS} elementary factors")
        if self.n_groups > _lib.GPK_MAX_GROUPS:
            raise GpkError(f"kernel uses more than {_lib.GPK_MAX_GROUPS} distinct length scales")

    def desc(self):
        d = KernelDesc()
        d.n_terms = len(self.terms)
        d.n_groups = self.n_groups
        f = 0
        for t, (coef, fs) in enumerate(self.terms):
            d.term_begin[t] = f
            d.coef[t] = coef
            for fac in fs:
                d.fac_kind[f] = KIND[fac[0]]
                d.fac_group[f] = fac[1]
                d.fac_param[f] = fac[2] if len(fac) > 2 else 0.0
                f += 1
        d.term_begin[len(self.terms)] = f
        return d


def _check_groups(xg, flat):
    if xg.dim() != 4:
        raise ValueError("scaled inputs must have shape [groups, batch, n, d]")
    if xg.shape[0] < flat.n_groups:
        raise ValueError("not enough input groups for the kernel")


def _km_launch(flat, xg, yg, n, n2, d, flags, noise_scalar, noise_vec, jitter, out, ldo, o_bstride, batch):
    _require_cuda(xg, yg, out, noise_vec)
    xg = xg.contiguous()
    yg = xg if yg is xg else yg.contiguous()
    if noise_vec is not None:
        noise_vec = noise_vec.contiguous()
    desc = flat.desc()
    rc = _fn("gpk_kernel_matrix", out.dtype)(
        ctypes.byref(desc), _ptr(xg), xg.stride(0), xg.stride(1), n, _ptr(yg), yg.stride(0), yg.stride(1), n2, d,
        float(noise_scalar), _ptr(noise_vec), (noise_vec.stride(0) if noise_vec is not None else 0), float(jitter),
        flags, _ptr(out), ldo, o_bstride, batch, _stream(),
    )
    check(rc, "gpk_kernel_matrix")


def kernel_matrix(flat, xg, yg=None, *, same=None, noise_scalar=0.0, noise_vec=None, jitter=0.0):
    """Full ``[B, n, n2]`` kernel matrix ``k(x, y)`` (+ noise on the diagonal when ``same``).

    ``xg``/``yg``: ``[G, B, n, d]`` pre-stretched inputs; ``yg=None`` means the same object as ``xg``."""
    _check_groups(xg, flat)
    if yg is None:
        yg = xg
        same = True if same is None else same
    same = bool(same)
    B, n, d = xg.shape[1], xg.shape[2], xg.shape[3]
    n2 = yg.shape[2]
    out = torch.empty(B, n, n2, device=xg.device, dtype=xg.dtype)
    if n == 0 or n2 == 0:
        return out
    _km_launch(flat, xg, yg, n, n2, d, KM_SAME if same else 0, noise_scalar, noise_vec, jitter, out, n2, n * n2, B)
    return out


def kernel_diag(flat, xg, yg=None, *, same=None):
    """``k.elwise(x, y)`` -> ``[B, n]``."""
    _check_groups(xg, flat)
    if yg is None:
        yg, same = xg, (True if same is None else same)
    _require_cuda(xg, yg)
    xg = xg.contiguous()
    yg = xg if yg is xg else yg.contiguous()
    B, n, d = xg.shape[1], xg.shape[2], xg.shape[3]
    out = torch.empty(B, n, device=xg.device, dtype=xg.dtype)
    if n == 0:
        return out
    desc = flat.desc()
    rc = _fn("gpk_kernel_diag", xg.dtype)(
        ctypes.byref(desc), _ptr(xg), xg.stride(0), xg.stride(1), _ptr(yg), yg.stride(0), yg.stride(1), n, d,
        1 if same else 0, _ptr(out), n, B, _stream(),
    )
    check(rc, "gpk_kernel_diag")
    return out


def gemm_nt(A, Bm, C=None, *, alpha=1.0, beta=0.0, lower=False):
    """``C = beta * C + alpha * A @ Bm^T`` on padded ``[B, M, K]`` / ``[B, N, K]`` tensors (views with a unit inner
    stride are fine).  Returns ``C``."""
    _require_cuda(A, Bm, C)
    Bn, M, K = A.shape
    N = Bm.shape[1]
    if C is None:
        C = torch.empty(Bn, M, N, device=A.device, dtype=A.dtype)
        beta = 0.0
    for t in (A, Bm, C):
        if t.stride(2) != 1:
            raise ValueError("gemm_nt needs a unit inner stride")
    if Bn == 1:
        _emulation_for_gemm(A.device, A.dtype, M, N, K)
    rc = _fn("gpk_gemm_nt", A.dtype)(
        M, N, K, alpha, _ptr(A), A.stride(1), A.stride(0), _ptr(Bm), Bm.stride(1), Bm.stride(0), beta, _ptr(C),
        C.stride(1), C.stride(0), 1 if lower else 0, Bn, _stream(),
    )
    check(rc, "gpk_gemm_nt")
    return C


def _pad_copy(src, dst, rows, cols, rows_pad, cols_pad, diag_add, pad_identity):
    if src.stride(2) != 1:
        src = src.contiguous()
    rc = _fn("gpk_pad_copy", dst.dtype)(
        _ptr(src), src.stride(1), src.stride(0), rows, cols, _ptr(dst), dst.stride(1), dst.stride(0), rows_pad,
        cols_pad, float(diag_add), 1 if pad_identity else 0, dst.shape[0], _stream(),
    )
    check(rc, "gpk_pad_copy")


def symmetrize_(A, n):
    """Mirror the lower triangle of the leading ``n x n`` block into the upper one, in place."""
    rc = _fn("gpk_symmetrize", A.dtype)(_ptr(A), A.stride(1), A.stride(0), n, A.shape[0], _stream())
    check(rc, "gpk_symmetrize")
    return A


def transpose(src, rows, cols, out=None):
    """``out[B, cols, rows] = src[B, rows, cols]^T`` (leading block of possibly padded tensors)."""
    if src.stride(2) != 1:  # a transposed VIEW (e.g. a reversed cross-kernel): the kernels take a unit inner stride
        src = src.contiguous()
    if out is None:
        out = torch.empty(src.shape[0], cols, rows, device=src.device, dtype=src.dtype)
    rc = _fn("gpk_transpose", src.dtype)(
        _ptr(src), src.stride(1), src.stride(0), rows, cols, _ptr(out), out.stride(1), out.stride(0), src.shape[0],
        _stream(),
    )
    check(rc, "gpk_transpose")
    return out


def row_dot_sq(V, rows, n_cols, b=None, want_dot=True, want_sq=True):
    """Per-row ``<V[r, :n_cols], b>`` and ``|V[r, :n_cols]|^2`` of ``V[B, *, *]`` -> two ``[B, rows]`` tensors."""
    if V.stride(2) != 1:
        V = V.contiguous()
    Bn = V.shape[0]
    dot = torch.empty(Bn, rows, device=V.device, dtype=V.dtype) if (want_dot and b is not None) else None
    sq = torch.empty(Bn, rows, device=V.device, dtype=V.dtype) if want_sq else None
    if rows == 0:
        return dot, sq
    if b is not None:
        b = b.contiguous()
    rc = _fn("gpk_row_dot_sq", V.dtype)(
        _ptr(V), V.stride(1), V.stride(0), rows, n_cols, _ptr(b), (b.stride(0) if b is not None else 0), _ptr(dot),
        _ptr(sq), rows, Bn, _stream(),
    )
    check(rc, "gpk_row_dot_sq")
    return dot, sq


class Chol:
    """Lower Cholesky factor ``L`` of ``K + jitter I`` in padded workspace storage, with fused right-hand sides.

    Attributes: ``W [B, n_pad + extra, n_pad]``, ``n``, ``n_pad``, ``k`` (number of fused right-hand sides),
    ``logdet [B]`` (= ``2 sum log diag L``), ``info [B]`` (int32; first non-positive pivot, 0 = ok)."""

    def __init__(self, W, n, k, logdet, info):
        self.W, self.n, self.k, self.logdet, self.info = W, int(n), int(k), logdet, info
        self.n_pad = W.shape[2]
        self.batch = W.shape[0]

    @property
    def dtype(self):
        return self.W.dtype

    @property
    def device(self):
        return self.W.device

    def check(self):
        """Raise ``torch.linalg.LinAlgError`` if a pivot was non-positive (forces a host sync)."""
        bad = self.info.nonzero()
        if bad.numel():
            b = int(bad[0, 0])
            raise torch.linalg.LinAlgError(
                f"Cholesky: leading minor of order {int(self.info[b])} is not positive definite (batch {b})"
            )
        return self

    def L_padded(self):
        return self.W[:, : self.n_pad, :]

    def L(self):
        """Dense ``[B, n, n]`` lower-triangular factor (copy; strict upper triangle zeroed)."""
        return torch.tril(self.W[:, : self.n, : self.n])

    def L_lower_(self):
        """The padded factor ``[B, n_pad, n_pad]`` with its strict upper triangle zeroed IN PLACE (no copy; done once): the
        form products with ``L`` need (``L eps`` of sampling, ``L_z A L_z^T``).  Nothing else reads the upper triangle."""
        if not getattr(self, "_upper_zeroed", False):
            self.W[:, : self.n_pad, :].tril_()
            self._upper_zeroed = True
        return self.W[:, : self.n_pad, :]

    def rhs_half(self):
        """``(L^-1 rhs)^T`` for the fused right-hand sides: ``[B, k, n]`` (a view)."""
        return self.W[:, self.n_pad : self.n_pad + self.k, : self.n]

    def logpdf(self):
        """``-0.5 (logdet + n log 2 pi + |L^-1 rhs_c|^2)`` for every fused right-hand side -> ``[B, k]``."""
        if self.k == 0:
            raise ValueError("no right-hand side was fused into this factorisation")
        out = torch.empty(self.batch, self.k, device=self.device, dtype=self.dtype)
        rows = self.W[:, self.n_pad :, :]
        rc = _fn("gpk_logpdf_finish", self.dtype)(
            _ptr(rows), rows.stride(1), rows.stride(0), self.n, self.n_pad, self.k, _ptr(self.logdet), _ptr(out),
            self.batch, _stream(),
        )
        check(rc, "gpk_logpdf_finish")
        return out

    def new_rows(self, rows, zero=True):
        """A padded ``[B, round_up(rows), n_pad]`` buffer for :meth:`solve_rows_`."""
        f = torch.zeros if zero else torch.empty
        return f(self.batch, round_up(rows), self.n_pad, device=self.device, dtype=self.dtype)

    def solve_rows_(self, Bt):
        """In place ``Bt <- Bt L^-T`` on a padded ``[B, rows_pad, n_pad]`` buffer: row r becomes ``(L^-1 b_r)^T``."""
        _require_cuda(Bt)
        if Bt.shape[2] != self.n_pad or Bt.shape[1] % TILE or Bt.stride(2) != 1:
            raise ValueError("solve_rows_ needs a padded [B, rows_pad, n_pad] buffer")
        Lp = self.L_padded()
        if self.batch == 1:  # the recursive solve's largest GEMM: rows x n/2 x n/2
            h = round_up(self.n_pad // 2)
            _emulation_for_gemm(self.device, self.dtype, Bt.shape[1], self.n_pad - h + TILE, h)
        rc = _fn("gpk_trsm_right", self.dtype)(
            _ptr(Lp), Lp.stride(1), Lp.stride(0), self.n_pad, _ptr(Bt), Bt.stride(1), Bt.stride(0), Bt.shape[1],
            self.batch, _stream(),
        )
        check(rc, "gpk_trsm_right")
        return Bt

    def solve_rows_t_(self, Bt):
        """In place ``Bt <- Bt L^-1``: row r becomes ``(L^-T b_r)^T`` (backward substitution)."""
        _require_cuda(Bt)
        Lp = self.L_padded()
        rc = _fn("gpk_trsm_right_t", self.dtype)(
            _ptr(Lp), Lp.stride(1), Lp.stride(0), self.n_pad, _ptr(Bt), Bt.stride(1), Bt.stride(0), Bt.shape[1],
            self.batch, _stream(),
        )
        check(rc, "gpk_trsm_right_t")
        return Bt

    def half_solve(self, bt):
        """``bt [B, m, n]`` (rows = right-hand sides) -> ``(L^-1 b)^T [B, m, n]``."""
        m = bt.shape[1]
        buf = self.new_rows(m)
        buf[:, :m, : self.n] = bt
        self.solve_rows_(buf)
        return buf[:, :m, : self.n]

    def full_solve(self, bt):
        """``bt [B, m, n]`` -> ``(K^-1 b)^T [B, m, n]``."""
        m = bt.shape[1]
        buf = self.new_rows(m)
        buf[:, :m, : self.n] = bt
        self.solve_rows_(buf)
        self.solve_rows_t_(buf)
        return buf[:, :m, : self.n]


def _new_workspace(B, n, k, device, dtype, rhs_t):
    n_pad = round_up(max(n, 1))
    extra = round_up(k) if k > 0 else 0
    W = torch.empty(B, n_pad + extra, n_pad, device=device, dtype=dtype)
    if extra:
        W[:, n_pad:, :].zero_()
        W[:, n_pad : n_pad + k, :n] = rhs_t
    return W, n_pad, extra


def _potrf(W, n, n_pad, extra, k, well_conditioned=False):
    B = W.shape[0]
    logdet = torch.zeros(B, device=W.device, dtype=W.dtype)
    info = torch.zeros(B, device=W.device, dtype=torch.int32)
    from . import B as _Bns

    if W.dtype == torch.float64 and B == 1 and getattr(_Bns, "precision", "fp64") == "tf32x3" and n_pad > 512:
        # opt-in mixed precision: trailing updates on the tcgen05 tensor cores (3xTF32) from an fp32 panel copy
        ws = torch.empty((n_pad + extra) * 512, device=W.device, dtype=torch.float32)
        rc = _lib.load().gpk_potrf_f64_tf32x3(_ptr(W), W.stride(1), W.stride(0), n_pad, extra, _ptr(logdet), _ptr(info),
                                              B, _ptr(ws), ws.numel(), _stream())
        check(rc, "gpk_potrf_f64_tf32x3")
        return Chol(W, n, k, logdet, info)
    if W.dtype == torch.float64 and B == 1 and n_pad >= 2048:
        slices = _oz_slices(well_conditioned)
        _set_emulation(W.device, slices, _lib.load().gpk_potrf_oz_ws_bytes(n_pad, extra, slices) if slices else 0)
    rc = _fn("gpk_potrf", W.dtype)(_ptr(W), W.stride(1), W.stride(0), n_pad, extra, _ptr(logdet), _ptr(info), B,
                                   _stream())
    check(rc, "gpk_potrf")
    return Chol(W, n, k, logdet, info)


_PRODUCT_SLICES_AUTO = [8]


class product_slices:
    """``with ops.product_slices(7): ...`` -- inside the block, ``B.precision = "auto"`` emulates the large products and solves
    with that many int8 slices instead of 8.  For consumers with a looser accuracy target than the 1e-10 parity bar of
    log-pdfs and posteriors: the analytic backward pass (hyper-parameter gradients, checked at 1e-8) forms ``K^-1`` from one
    big solve and one big product, 1.3x faster with 7 slices."""

    def __init__(self, slices):
        self.slices = int(slices)

    def __enter__(self):
        self.prev = _PRODUCT_SLICES_AUTO[0]
        _PRODUCT_SLICES_AUTO[0] = self.slices

    def __exit__(self, *exc):
        _PRODUCT_SLICES_AUTO[0] = self.prev


def _oz_slices(well_conditioned=False):
    """``B.precision`` -> number of int8 slices of the emulated large fp64 updates (0: native fp64 tensor cores only).

    "auto" uses 8 slices (56-bit operands >= fp64's 53: the accuracy of the fp64 tensor-core kernel itself) everywhere EXCEPT
    the one place where 7 slices (49-bit operands, product error ~3e-14) are measured to stay three orders inside the 1e-10
    parity bar: the factorisation inside a stand-alone ``logpdf`` of a matrix that is well conditioned BY CONSTRUCTION (a known
    scalar noise of at least 1e-3 of the kernel's variance on the diagonal) -- the log-pdf is a well-conditioned functional of
    the factor (log-det and one quadratic form).  Measured (round 2, ``profiles/r02_conditioning_sweep.txt`` and the full-size
    oracle tests): at n = 16384, noise 0.1 the 7-slice log-pdf is 3e-14 from the CPU oracle, but the ELEMENTS of a posterior
    mean computed from a 7-slice factor and 7-slice solves are up to 2e-10 off (1.2e-10 of the largest element) -- so
    factorisations that serve a posterior (``Observations.K_x``) and every triangular solve / product use 8 slices; and the
    7-slice log-pdf error grows like 1.5e-14 / (noise / variance), reaching 1e-10 near 1e-4 -- hence the 1e-3 threshold."""
    from . import B as _Bns

    mode = getattr(_Bns, "precision", "auto")
    if mode == "auto":
        return 7 if well_conditioned else _PRODUCT_SLICES_AUTO[0]
    return {"int8x5": 5, "int8x6": 6, "int8x7": 7, "int8x8": 8}.get(mode, 0)


def _well_conditioned(flat, noise_scalar, noise_vec, jitter):
    """True when ``k(x, x) + noise`` is well conditioned by construction: scalar diagonal term >= 1e-3 of the kernel's
    variance scale (sum of |coefficients| of the bounded stationary terms; anything with a Linear factor is unbounded)."""
    if noise_vec is not None:
        return False
    scale = 0.0
    for coef, fs in flat.terms:
        if any(f[0] == "linear" for f in fs):
            return False
        if all(f[0] == "delta" for f in fs):
            continue  # a Delta term only adds to the diagonal
        scale += abs(coef)
    diag = float(noise_scalar) + float(jitter) + sum(c for c, fs in flat.terms if fs and all(f[0] == "delta" for f in fs) and c > 0)
    return diag >= 1e-3 * max(scale, 1e-300)


#: per-device scratch handed to the library for the int8-slice emulation: [tensor, slices registered]
_EMULATION = {}
_EMULATION_MAX_BYTES = 8 << 30


def _set_emulation(device, slices, need_bytes):
    """Make the library's fp64 emulation mode on ``device`` match ``B.precision`` and own a scratch buffer of at least
    ``need_bytes`` (grown on demand, capped: larger requests simply stay on the fp64 tensor cores)."""
    key = torch.device(device).index or 0
    buf, cur = _EMULATION.get(key, (None, 0))
    lib = _lib.load()
    if slices == 0:
        if cur:
            with torch.cuda.device(key):
                check(lib.gpk_set_f64_emulation(0, None, 0), "gpk_set_f64_emulation")
            _EMULATION[key] = (buf, 0)
        return
    need_bytes = min(int(need_bytes), _EMULATION_MAX_BYTES)
    if buf is None or buf.numel() < need_bytes:
        buf = _aligned_bytes(max(need_bytes, 64 << 20), torch.device("cuda", key))
        cur = 0
    if cur != slices:
        with torch.cuda.device(key):
            check(lib.gpk_set_f64_emulation(slices, _ptr(buf), buf.numel()), "gpk_set_f64_emulation")
    _EMULATION[key] = (buf, slices)


def _emulation_for_gemm(device, dtype, M, N, K):
    if dtype != torch.float64:
        return
    slices = _oz_slices(False)
    kc = K if K <= 65536 else -(-K // (-(-K // 65536)) // 128) * 128 + 128  # long reductions run in K chunks <= 65536
    need = _lib.load().gpk_f64_emulation_scratch_bytes(M, N, min(K, kc), slices) if slices and M * N * K >= 1.5e9 else 0
    _set_emulation(device, slices, need)


def _aligned_bytes(nbytes, device, align=1024):
    buf = torch.empty(int(nbytes) + align, device=device, dtype=torch.uint8)
    off = (-buf.data_ptr()) % align
    return buf[off : off + int(nbytes)]


def gemm_nt_oz(A, Bm, C=None, *, alpha=1.0, beta=0.0, lower=False, slices=6):
    """``C = beta C + alpha A B^T`` for fp64 row-major 2-D tensors through the int8 tensor-core emulation
    (``gpk_gemm_nt_f64_oz``).  ``A: [M, K]``, ``Bm: [N, K]``; M % 128 == 0, N % 64 == 0, K % 128 == 0."""
    _require_cuda(A, Bm, C)
    M, K = A.shape
    N = Bm.shape[0]
    if C is None:
        C = torch.zeros(M, N, device=A.device, dtype=A.dtype)
    lib = _lib.load()
    need = ((lib.gpk_oz_ws_bytes(M, K, slices) + 1023) // 1024) * 1024 + lib.gpk_oz_ws_bytes(N, K, slices)
    ws = _aligned_bytes(need, A.device)
    rc = lib.gpk_gemm_nt_f64_oz(M, N, K, float(alpha), _ptr(A), A.stride(0), _ptr(Bm), Bm.stride(0), float(beta), _ptr(C),
                                C.stride(0), int(lower), int(slices), _ptr(ws), ws.numel(), _stream())
    check(rc, "gpk_gemm_nt_f64_oz")
    return C


def chol_from_kernel(flat, xg, *, noise_scalar=0.0, noise_vec=None, jitter=0.0, rhs_t=None, full_precision=False):
    """Build ``k(x, x) + noise + jitter I`` straight into the padded lower workspace (K1), factorise it in place
    (K2) and carry ``rhs_t [B, k, n]`` through the factorisation (fused K3).  Returns a :class:`Chol`.
    ``full_precision``: the factor will serve element-wise quantities (posterior means / variances): never the 7-slice
    emulation (see :func:`_oz_slices`)."""
    _check_groups(xg, flat)
    _require_cuda(xg, noise_vec, rhs_t)
    B, n, d = xg.shape[1], xg.shape[2], xg.shape[3]
    k = 0 if rhs_t is None else rhs_t.shape[1]
    W, n_pad, extra = _new_workspace(B, n, k, xg.device, xg.dtype, rhs_t)
    _km_launch(flat, xg, xg, n, n, d, KM_LOWER | KM_SAME | KM_PAD_IDENTITY, noise_scalar, noise_vec, jitter, W,
               W.stride(1), W.stride(0), B)
    return _potrf(W, n, n_pad, extra, k, (not full_precision) and _well_conditioned(flat, noise_scalar, noise_vec, jitter))


def chol_from_dense(K, *, jitter=0.0, rhs_t=None):
    """Factorise a given dense SPD ``K [B, n, n]`` (+ ``jitter I``); only its lower triangle is used."""
    _require_cuda(K, rhs_t)
    if K.stride(2) != 1:
        K = K.contiguous()
    B, n = K.shape[0], K.shape[1]
    k = 0 if rhs_t is None else rhs_t.shape[1]
    W, n_pad, extra = _new_workspace(B, n, k, K.device, K.dtype, rhs_t)
    _pad_copy(K, W, n, n, n_pad, n_pad, jitter, True)
    return _potrf(W, n, n_pad, extra, k)


def kernel_rows_padded(flat, xsg, xg, chol):
    """``k(x*, x)`` as a zero-padded ``[B, m_pad, n_pad]`` buffer (rows = test points) ready for ``solve_rows_``."""
    _check_groups(xg, flat)
    B, m, d = xsg.shape[1], xsg.shape[2], xsg.shape[3]
    n = xg.shape[2]
    out = torch.empty(B, round_up(max(m, 1)), chol.n_pad, device=xg.device, dtype=xg.dtype)
    _km_launch(flat, xsg, xg, m, n, d, KM_PAD_ZERO, 0.0, None, 0.0, out, out.stride(1), out.stride(0), B)
    return out


def posterior_marginals(flat, xsg, xg, chol, half_y=None, want_sq=True, chunk=4096):
    """K3 in one call (``gpk_posterior_marginals``): ``(dot [m], sq [m])`` with ``V^T = k(x*, x) L^-T``, ``dot = V^T half_y``,
    ``sq = |v_i|^2``; one problem (no batch), test points streamed in chunks of ``chunk`` rows."""
    _check_groups(xg, flat)
    _require_cuda(xsg, xg, half_y)
    if chol.batch != 1 or xsg.shape[1] != 1 or xg.shape[1] != 1:
        raise ValueError("posterior_marginals handles a single problem")
    xsg, xg = xsg.contiguous(), xg.contiguous()
    m, d = xsg.shape[2], xsg.shape[3]
    dt, dev = chol.dtype, chol.device
    dot = torch.empty(m, dtype=dt, device=dev) if half_y is not None else None
    sq = torch.empty(m, dtype=dt, device=dev) if want_sq else None
    if m == 0:
        return dot, sq
    chunk = min(round_up(chunk), round_up(m))
    ws = torch.empty(chunk * chol.n_pad, dtype=dt, device=dev)
    if dt == torch.float64:  # the solve's largest product: rows x n/2 x n/2
        h = round_up(chol.n_pad // 2)
        slices = _oz_slices(False)
        need = _lib.load().gpk_f64_emulation_scratch_bytes(chunk, chol.n_pad - h + TILE, h, slices) if (
            slices and chunk * (chol.n_pad - h + TILE) * h >= 1.5e9) else 0
        _set_emulation(dev, slices, need)
    Lp = chol.L_padded()
    hy = None if half_y is None else half_y.contiguous()
    desc = flat.desc()
    rc = _fn("gpk_posterior_marginals", dt)(
        ctypes.byref(desc), _ptr(xsg), xsg.stride(0), m, _ptr(xg), xg.stride(0), xg.shape[2], d, _ptr(Lp), Lp.stride(1),
        chol.n_pad, _ptr(hy), _ptr(dot), _ptr(sq), chunk, _ptr(ws), ws.numel(), _stream(),
    )
    check(rc, "gpk_posterior_marginals")
    return dot, sq


SPARSE_METHOD = {"vfe": 0, "fitc": 1, "dtc": 2}


class SparseAccumulator:
    """Streamed ``AbstractPseudoObservations._compute`` (``stheno/model/observations.py:279-336``) for ONE problem (no batch
    dimension): the data are walked in chunks of ``chunk`` points through ``gpk_sparse_accumulate``; ``K_zx`` is never held.

    ``A [1, m_pad, m_pad]`` starts at the identity and receives ``W K_n^-1 W^T`` on its lower tiles, ``prod [m_pad]`` receives
    ``W K_n^-1 ybar``, ``scalars`` = (sum log(2 pi K_n), sum ybar^2 / K_n, trace part).  Device memory: two
    ``chunk x m_pad`` buffers + ``A``."""

    def __init__(self, flat, zg, ch_z, method, chunk=16384):
        _require_cuda(zg, ch_z.W)
        if ch_z.batch != 1 or zg.shape[1] != 1:
            raise ValueError("SparseAccumulator handles a single problem (batched sparse problems use the materialised path)")
        self.flat, self.zg, self.ch, self.method = flat, zg.contiguous(), ch_z, SPARSE_METHOD[method]
        self.m, self.m_pad, self.d = ch_z.n, ch_z.n_pad, zg.shape[3]
        self.chunk = int(chunk)
        dt, dev = ch_z.dtype, ch_z.device
        self.A = torch.zeros(1, self.m_pad, self.m_pad, dtype=dt, device=dev)
        self.A.diagonal(dim1=1, dim2=2).fill_(1.0)
        self.prod = torch.zeros(self.m_pad, dtype=dt, device=dev)
        self.scalars = torch.zeros(3, dtype=dt, device=dev)
        self.lib = _lib.load()
        self.ws = None

    def _workspace(self, c):
        need = int(self.lib.gpk_sparse_ws_elems(c, self.m_pad))
        if self.ws is None or self.ws.numel() < need:
            self.ws = torch.empty(need, dtype=self.ch.dtype, device=self.ch.device)
        if self.ch.dtype == torch.float64:
            # the emulation scratch has to hold the largest product of the solve and the K = c accumulation
            slices = _oz_slices(False)
            c_pad, h = round_up(c), round_up(self.m_pad // 2)
            need_b = 0
            if slices:
                f = self.lib.gpk_f64_emulation_scratch_bytes
                need_b = max(f(self.m_pad, self.m_pad, c_pad, slices), f(c_pad, self.m_pad - h + TILE, h, slices))
            _set_emulation(self.ch.device, slices, need_b)
        return self.ws

    def add(self, xg_chunk, kdiag, kn, ybar):
        """One chunk: ``xg_chunk [G, 1, c, d]`` pre-stretched points, ``kdiag / kn / ybar [c]`` (``kdiag`` None for DTC)."""
        _require_cuda(xg_chunk, kdiag, kn, ybar)
        xg_chunk = xg_chunk.contiguous()
        c = xg_chunk.shape[2]
        if c == 0:
            return
        ws = self._workspace(c)
        kd = None if kdiag is None else kdiag.contiguous()
        kn, ybar = kn.contiguous(), ybar.contiguous()
        Lp = self.ch.L_padded()
        desc = self.flat.desc()
        rc = _fn("gpk_sparse_accumulate", self.ch.dtype)(
            ctypes.byref(desc), _ptr(xg_chunk), xg_chunk.stride(0), c, _ptr(self.zg), self.zg.stride(0), self.m, self.d,
            _ptr(Lp), Lp.stride(1), self.m_pad, _ptr(kd), _ptr(kn), _ptr(ybar), self.method, _ptr(self.A), self.A.stride(1),
            _ptr(self.prod), _ptr(self.scalars), _ptr(ws), ws.numel(), _stream(),
        )
        check(rc, "gpk_sparse_accumulate")


def gemm_profile(enable):
    """Switch the in-situ event timing of the fp64 GEMM kernel on / off (clears the record)."""
    _lib.load().gpk_gemm_profile_enable(1 if enable else 0)


def gemm_profile_read(kind=0):
    """``(total_ms, algorithmic_flops, launches)`` of the profiled GEMM launches since ``gemm_profile(True)``; synchronises.
    ``kind``: 0 = fp64 DMMA trailing-update kernel, 1 = int8-slice emulation kernel (fp64-equivalent flops), -1 = both."""
    torch.cuda.synchronize()
    ms, fl, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
    check(_lib.load().gpk_gemm_profile_read_kind(int(kind), ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(n)),
          "gpk_gemm_profile_read_kind")
    return ms.value, fl.value, n.value


def probe_dmma_tflops():
    return float(_lib.load().gpk_probe_dmma_tflops())


LOG_2_PI = math.log(2 * math.pi)
