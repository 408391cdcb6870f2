"""Pin the CPU oracle on the reference's own known answers.

G1 README.md:48-85, G2 README.md:477-479, G3 README.md:482-497, G4 README.md:699-719,
T1 tests/test_random.py:185-192 (SciPy as independent oracle)."""
import numpy as np
import pytest
from scipy.stats import multivariate_normal

from oracle import gp_oracle as O


def test_G1_posterior_readme():
    x = np.linspace(0, 2, 10)
    y = x**2
    mean, var = O.posterior(("eq",), x, None, y, np.array([1.0, 2.0, 3.0]))
    # kappa(K) ~ 1e16 here: the third value moves by ~5e-7 with the LAPACK build (summation order),
    # while dropping the jitter moves it by 1.7e-2 (next test) -- so 2e-6 still pins eps.
    np.testing.assert_allclose(mean[:, 0], [1.00000068, 3.99999999, 8.4825932], rtol=0, atol=2e-6)
    # The variance at x*=3 is printed as 3.31283378e-03 (the other two entries are pure jitter noise).
    assert abs(var[2, 2] - 3.31283378e-03) < 5e-6
    assert abs(var[0, 0]) < 1e-10 and abs(var[1, 1]) < 1e-10


def test_G1_needs_the_jitter():
    x = np.linspace(0, 2, 10)
    mean0, _ = O.posterior(("eq",), x, None, x**2, np.array([3.0]), eps=0.0)
    # Without the 1e-12 jitter the third mean is ~8.4997, not 8.4826: the golden value pins eps.
    assert abs(mean0[0, 0] - 8.4825932) > 1e-3


def test_G2_eq_matrix():
    K = O.kernel_matrix(("eq",), np.array([0.0, 1.0, 2.0]))
    ref = np.array([[1.0, 0.607, 0.135], [0.607, 1.0, 0.607], [0.135, 0.607, 1.0]])
    np.testing.assert_allclose(K, ref, atol=5e-4)
    np.testing.assert_allclose(K[0, 1], np.exp(-0.5))
    np.testing.assert_allclose(K[0, 2], np.exp(-2.0))


def test_G3_logpdf_readme():
    x = np.array([0.0, 1.0, 2.0])
    y1 = np.array([-0.45172746, 0.46581948, 0.78929767])
    lp = O.fdd_logpdf(("eq",), x, None, y1)
    assert np.ndim(lp) == 0
    assert abs(lp - (-2.811609567720761)) < 5e-8  # y printed to 8 digits
    y2 = np.array([[-0.43771276, -2.36741858], [0.86080043, -1.22503079], [2.15779126, -0.75319405]])
    lp2 = O.fdd_logpdf(("eq",), x, None, y2)
    assert lp2.shape == (2,)
    np.testing.assert_allclose(lp2, [-4.82949038, -5.40084225], atol=5e-7)


def test_T1_logpdf_vs_scipy():
    rng = np.random.default_rng(0)
    a = rng.standard_normal((3, 3))
    var = a @ a.T + 0.5 * np.eye(3)
    mean = rng.standard_normal((3, 1))
    x = rng.standard_normal((3, 10))
    lp = O.normal_logpdf(mean, var, x)
    ref = multivariate_normal(mean[:, 0], var).logpdf(x.T)
    np.testing.assert_allclose(lp, ref, rtol=1e-6)
    assert O.normal_logpdf(mean, var, x[:, :1]).shape == ()


def test_missing_data_masking():
    # stheno/random.py:261-270 and tests/test_random.py:195-204
    rng = np.random.default_rng(1)
    a = rng.standard_normal((4, 4))
    var = a @ a.T + np.eye(4)
    y = rng.standard_normal(4)
    y_nan = y.copy()
    y_nan[1] = np.nan
    keep = np.array([0, 2, 3])
    ref = O.normal_logpdf(None, var[np.ix_(keep, keep)], y[keep])
    np.testing.assert_allclose(O.normal_logpdf(None, var, y_nan), ref)


def test_G4_vfe_close_to_exact():
    # README.md:699-719: n=2000, m=100, noise 1: ELBO - logpdf = -3.5e-10 (same magnitude here).
    rng = np.random.default_rng(4)
    x = np.linspace(0, 10, 2000)
    z = np.linspace(0, 10, 100)
    K = O.kernel_matrix(("eq",), x) + np.eye(2000)
    y = np.linalg.cholesky(K) @ rng.standard_normal(2000)
    exact = O.fdd_logpdf(("eq",), x, 1.0, y)
    for method, tol in (("vfe", 5e-9), ("fitc", 5e-9), ("dtc", 5e-9)):
        elbo = O.sparse_compute(("eq",), z, x, 1.0, y, method)["elbo"]
        assert abs(elbo - exact) < tol, (method, elbo - exact)
    assert O.sparse_compute(("eq",), z, x, 1.0, y, "vfe")["elbo"] <= exact + 1e-9


def test_T2_sparse_equals_exact_when_z_is_x():
    # tests/model/test_model.py:284-308
    rng = np.random.default_rng(5)
    x = np.linspace(0, 5, 10)
    xs = np.linspace(0, 5, 7) + 0.1
    y = rng.standard_normal(10)
    spec = ("eq",)
    pm, pv = O.posterior(spec, x, 0.1, y, xs)
    exact = O.fdd_logpdf(spec, x, 0.1, y)
    for method in ("vfe", "fitc", "dtc"):
        sm, sv = O.sparse_posterior(spec, x, x, 0.1, y, xs, method)
        np.testing.assert_allclose(sm, pm, atol=1e-4)
        np.testing.assert_allclose(sv, pv, atol=1e-4)
        np.testing.assert_allclose(O.sparse_compute(spec, x, x, 0.1, y, method)["elbo"], exact, atol=1e-4)


def test_kernel_algebra_and_stretch():
    rng = np.random.default_rng(6)
    x = rng.standard_normal((7, 3))
    y = rng.standard_normal((5, 3))
    d2 = ((x[:, None, :] - y[None, :, :]) ** 2).sum(-1)
    np.testing.assert_allclose(O.kernel_matrix(("eq",), x, y), np.exp(-0.5 * d2), atol=1e-14)
    r = np.sqrt(d2)
    np.testing.assert_allclose(O.kernel_matrix(("matern12",), x, y), np.exp(-r), atol=1e-13)
    np.testing.assert_allclose(O.kernel_matrix(("matern32",), x, y), (1 + np.sqrt(3) * r) * np.exp(-np.sqrt(3) * r), atol=1e-13)
    np.testing.assert_allclose(
        O.kernel_matrix(("matern52",), x, y), (1 + np.sqrt(5) * r + 5 * d2 / 3) * np.exp(-np.sqrt(5) * r), atol=1e-13
    )
    np.testing.assert_allclose(O.kernel_matrix(("linear",), x, y), x @ y.T)
    ell = np.array([0.5, 2.0, 3.0])
    d2s = (((x[:, None, :] - y[None, :, :]) / ell) ** 2).sum(-1)
    np.testing.assert_allclose(O.kernel_matrix(("stretched", ell, ("eq",)), x, y), np.exp(-0.5 * d2s), atol=1e-14)
    spec = ("sum", ("scaled", 2.0, ("stretched", 2.0, ("eq",))), ("product", ("matern32",), ("linear",)))
    ref = 2 * np.exp(-0.5 * d2 / 4) + (1 + np.sqrt(3) * r) * np.exp(-np.sqrt(3) * r) * (x @ y.T)
    np.testing.assert_allclose(O.kernel_matrix(spec, x, y), ref, atol=1e-13)
    # Delta: identity on the same object, exact-match indicator otherwise.
    np.testing.assert_allclose(O.kernel_matrix(("delta",), x), np.eye(7))
    np.testing.assert_allclose(O.kernel_matrix(("delta",), x, x[:3]), np.eye(7)[:, :3])
    # elwise agrees with the diagonal of pairwise
    for s in (("eq",), ("matern52",), spec):
        np.testing.assert_allclose(O.kernel_elwise(s, x, x + 0.1)[:, 0], np.diag(O.kernel_matrix(s, x, x + 0.1)), atol=1e-13)


def test_noise_as_process_equals_noise_argument():
    # tests/model/test_model.py:181-195 : k + s2*Delta  ==  (k, noise=s2)
    rng = np.random.default_rng(7)
    x = rng.standard_normal((20, 2))
    y = rng.standard_normal(20)
    a = O.fdd_logpdf(("sum", ("eq",), ("scaled", 0.3, ("delta",))), x, None, y)
    b = O.fdd_logpdf(("eq",), x, 0.3, y)
    np.testing.assert_allclose(a, b, rtol=1e-13)


def test_batched_logpdf_shape():
    # tests/model/test_cases.py:134-155
    rng = np.random.default_rng(8)
    x = rng.standard_normal((16, 10, 1))
    y = rng.standard_normal((16, 10, 1))
    K = O.kernel_matrix(("eq",), x) + 0.1 * np.eye(10)
    lp = O.normal_logpdf(None, K, y)
    assert lp.shape == (16,)
    np.testing.assert_allclose(lp[3], O.fdd_logpdf(("eq",), x[3], 0.1, y[3]))


def test_chunked_sparse_oracle_equals_plain():
    """The memory-bounded form used by the full-size C4 parity test is the same arithmetic as ``sparse_compute``."""
    rng = np.random.default_rng(41)
    n, m, d = 700, 40, 3
    x, z = rng.standard_normal((n, d)), rng.standard_normal((m, d))
    y = rng.standard_normal(n)
    spec = ("stretched", 2.0, ("matern52",))
    noise = 0.05 + rng.uniform(0, 0.1, n)
    for method in ("vfe", "fitc", "dtc"):
        a = O.sparse_compute(spec, z, x, noise, y, method)
        b = O.sparse_compute_chunked(spec, z, x, noise, y, method, chunk=128, workers=3 if method == "vfe" else 1)
        assert abs(a["elbo"] - b["elbo"]) < 1e-10 * abs(a["elbo"])
        np.testing.assert_allclose(b["mu"], a["mu"], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(b["A"], a["A"], rtol=1e-9, atol=1e-10)
