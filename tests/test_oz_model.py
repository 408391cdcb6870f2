"""CPU checks of the NumPy integer model of the fp64 emulation (``tests/_oz_model.py``): the slicing is error-free and the
modelled product has the accuracy the design claims.  (The GPU kernel is compared with this model bit for bit in
``tests/test_emulation.py``.)"""
import numpy as np
import pytest

from tests._oz_model import gemm, slice_rows


@pytest.mark.parametrize("S", [5, 6, 7, 8])
def test_slicing_is_error_free(S):
    rng = np.random.default_rng(S)
    X = rng.standard_normal((64, 256)) * np.exp(4 * rng.standard_normal((64, 1)))
    X[3] = 0.0  # an all-zero row (identity padding)
    X[5, 7] = 2.0 ** 10  # an exact power of two as the row maximum
    e, qs, rem = slice_rows(X, S)
    assert all(np.abs(q).max() <= 64 for q in qs)
    recon = sum(q.astype(np.float64) * 2.0 ** -(6 + 7 * s) for s, q in enumerate(qs)) * np.ldexp(1.0, e)[:, None]
    scale = np.ldexp(1.0, e)[:, None]
    # the first S slices carry 6 + 7 (S - 1) bits below the row's power-of-two scale; what is left is the remainder, exactly
    # (at 8 slices that is 55 bits, i.e. the reconstruction sum itself rounds at the last bit of the 53-bit mantissa)
    assert np.all(np.abs(X - recon) <= scale * 2.0 ** -(7 * S) * (1 + 1e-9) + np.abs(X) * 2.0 ** -52)
    assert np.all(np.abs(rem) <= 2.0 ** -(7 * S) * (1 + 1e-9))
    assert np.all(recon[3] == 0.0)


@pytest.mark.parametrize("S,tol", [(6, 2e-11), (7, 2e-13), (8, 5e-15)])
def test_model_product_accuracy(S, tol):
    rng = np.random.default_rng(10 + S)
    M, N, K = 96, 80, 512
    A = rng.standard_normal((M, K)) * np.exp(3 * rng.standard_normal((M, 1)))
    B = rng.standard_normal((N, K))
    ref = A @ B.T
    got = gemm(A, B, None, 1.0, 0.0, S)
    scale = np.abs(A).max(1)[:, None] * np.abs(B).max(1)[None, :] * np.sqrt(K)
    assert np.max(np.abs(got - ref) / scale) < tol


def test_model_is_exact_on_small_integers():
    rng = np.random.default_rng(0)
    A = rng.integers(-1000, 1000, (32, 128)).astype(np.float64)
    B = rng.integers(-1000, 1000, (48, 128)).astype(np.float64)
    assert np.array_equal(gemm(A, B, None, 1.0, 0.0, 7), A @ B.T)


def test_auto_slice_count_heuristic():
    """``B.precision = "auto"``: 7 slices only for factorisations that are well conditioned by construction."""
    from stheno_b200 import B, ops

    eq = ops.FlatKernel([(1.0, [("eq", 0)])], 1)
    eq_delta = ops.FlatKernel([(2.0, [("eq", 0)]), (0.1, [("delta", 0)])], 1)
    lin = ops.FlatKernel([(1.0, [("linear", 0)])], 1)
    assert ops._well_conditioned(eq, 0.1, None, 1e-12)
    assert ops._well_conditioned(eq_delta, 0.0, None, 1e-12)  # the Delta term is the noise
    assert not ops._well_conditioned(eq, 0.0, None, 1e-12)  # noise-free: only the jitter
    assert not ops._well_conditioned(eq, 1e-9, None, 0.0)
    assert not ops._well_conditioned(eq, 1e-4, None, 0.0)  # measured: the 7-slice log-pdf error reaches 1e-10 near 1e-4
    assert not ops._well_conditioned(eq, 0.1, object(), 0.0)  # per-point noise: smallest entry unknown on the host
    assert not ops._well_conditioned(lin, 0.1, None, 0.0)  # unbounded kernel
    before = B.precision
    try:
        B.precision = "auto"
        assert ops._oz_slices(True) == 7 and ops._oz_slices(False) == 8
        B.precision = "int8x6"
        assert ops._oz_slices(True) == 6 and ops._oz_slices(False) == 6
        B.precision = "fp64"
        assert ops._oz_slices(True) == 0
        B.precision = "tf32x3"
        assert ops._oz_slices(False) == 0
    finally:
        B.precision = before
