"""NumPy integer model of the int8-slice fp64 emulation in ``stheno_b200/csrc/gemm_oz.cu`` (test infrastructure).

It restates, operation for operation, what ``oz_slice_kernel`` and ``oz_gemm_kernel`` compute: the power-of-two row scaling,
the round-to-nearest 7-bit slicing (error-free: every step is exact in fp64), the exact integer slice products grouped by
``s + t``, the fp64 Horner recombination and the scaling.  Because every inexact step of the kernel is a single correctly
rounded fp64 operation, the kernel's output must equal this model BIT FOR BIT (``tests/test_emulation.py``)."""
import numpy as np


def slice_rows(X, S):
    """``X [rows, K]`` -> (``e [rows]``, list of ``S`` int64 arrays ``q_s [rows, K]``) with
    ``X = 2^e * sum_s q_s 2^-(6 + 7 s) + remainder``, ``|q_s| <= 64``, ``|remainder| <= 2^(e - 7 S)``."""
    X = np.asarray(X, np.float64)
    m = np.abs(X).max(axis=1)
    _, ex = np.frexp(m)  # m = f * 2^ex, f in [0.5, 1)  ->  ilogb(m) + 1 = ex
    e = np.where(m > 0, ex, 0).astype(np.int64)
    r = X * np.ldexp(1.0, -e)[:, None]  # exact
    qs, pw = [], 64.0
    for _ in range(S):
        t = np.rint(r * pw)  # round half to even, like the device rint()
        r = r - t / pw  # exact
        qs.append(t.astype(np.int64))
        pw *= 128.0
    return e, qs, r


def gemm(A, B, C0, alpha, beta, S):
    """``beta * C0 + alpha * A @ B.T`` the way the kernel forms it (A: [M, K], B: [N, K])."""
    ea, qa, _ = slice_rows(A, S)
    eb, qb, _ = slice_rows(B, S)
    acc = [sum(qa[s] @ qb[d - s].T for s in range(d + 1)) for d in range(S)]  # exact integers, diagonal d = s + t
    v = acc[0].astype(np.float64)
    for d in range(1, S):
        v = v * 128.0 + acc[d].astype(np.float64)  # v * 128 is exact: one rounding per step, like the device fma
    w_last = np.ldexp(1.0, -(12 + 7 * (S - 1)))
    rs = alpha * w_last * np.ldexp(1.0, ea)  # exact
    out = v * (rs[:, None] * np.ldexp(1.0, eb)[None, :])
    if beta == 0.0:
        return out
    base = C0 if beta == 1.0 else C0 * beta
    return base + out
