"""``GP.diff`` / ``GP.diff_approx``: the reference's own tests ``tests/model/test_model.py:510-529`` and
``tests/model/test_cases.py:95-113``, the README literal for the finite-difference step, and derivative kernels against the
closed form in the oracle."""
import numpy as np
import pytest
import torch

from oracle import gp_oracle as O


@pytest.fixture(params=["cpu", pytest.param("gpu", marks=pytest.mark.gpu)])
def S(request, monkeypatch):
    import stheno_b200 as s

    if request.param == "cpu":
        from tests import _cpu_backend

        _cpu_backend.install(monkeypatch)
    s.B.epsilon = 1e-12
    monkeypatch.setattr(s.Measure, "default", None)
    return s


def test_derivative_kernels_vs_closed_form(S):
    rng = np.random.default_rng(0)
    x, y = rng.standard_normal((9, 3)), rng.standard_normal((7, 3))
    k = S.EQ().stretch(1.4)
    inner = ("stretched", 1.4, ("eq",))
    for dims in ((1, 1), (0, 2), (2, None), (None, 0)):
        got = S.B.to_numpy(S.B.dense(k.diff(*dims)(x, y)))
        np.testing.assert_allclose(got, O.kernel_matrix(("diff", dims, inner), x, y), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(S.B.to_numpy(S.B.dense(k.diff(1)(x))), O.kernel_matrix(("diff", (1, 1), inner), x, x), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(k.diff(1).elwise(x, x)[:, 0], np.diag(O.kernel_matrix(("diff", (1, 1), inner), x, x)), rtol=1e-10, atol=1e-12)


def test_reference_derivative_test(S):
    # tests/model/test_model.py:510-529
    p = S.GP(lambda x: x ** 2, S.EQ())
    assert str(p.diff(1)) == "GP(d(1) <lambda>, d(1) EQ())"
    dp = p.diff()
    x = np.linspace(0, 1, 100)
    y = 2 * x
    x_check = np.linspace(0.2, 0.8, 100)
    post = p.measure | (p(x), y)
    np.testing.assert_allclose(post(dp)(x_check).mean, 2 * np.ones((100, 1)), atol=1e-4)
    post = p.measure | ((p(0.0), 0.0), (dp(x), y))
    np.testing.assert_allclose(post(p)(x_check).mean, x_check[:, None] ** 2, atol=1e-4)


def test_reference_approximate_derivative_test(S):
    # tests/model/test_cases.py:95-113
    p = S.GP(S.EQ().stretch(1.0))
    dp = p.diff_approx()
    x = np.linspace(0, 1, 100)
    y = 2 * x
    x_check = np.linspace(0.2, 0.8, 100)
    post = p.measure | (p(x), y)
    np.testing.assert_allclose(post(dp)(x_check).mean, 2 * np.ones((100, 1)), atol=1e-3)
    S.B.epsilon = 1e-10
    try:
        post = p.measure | ((p(0.0), 0.0), (dp(x), y))
        np.testing.assert_allclose(post(p)(x_check).mean, x_check[:, None] ** 2, atol=1e-3)
    finally:
        S.B.epsilon = 1e-12


def test_fdm_step_readme_literal():
    # README.md:292-293: GP(EQ()).diff_approx(deriv=1, order=2) shifts by 0.0001414213562373095 and scales by 50000000.0
    from stheno_b200.model.gp import _central_fdm

    grid, coefs, step = _central_fdm(2, 1)
    assert list(grid) == [-1.0, 1.0] and list(coefs) == [-0.5, 0.5]
    assert step == 0.0001414213562373095
    assert abs(1 / step ** 2 - 50000000.0) < 1e-3
