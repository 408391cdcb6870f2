"""``GP.shift / select / transform`` and the per-argument kernel maps behind them (SURVEY 8f rank 3;
``stheno/model/measure.py:272-345``, ``stheno/model/gp.py:190-216``) against the oracle.

Every test runs twice: on the CPU with the torch stand-in backend (host logic; ``-m "not gpu"``) and on the GPU through
the real CUDA kernels (``-m gpu``)."""
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


def approx(a, b, rtol=1e-9, atol=1e-9):
    import stheno_b200 as s

    np.testing.assert_allclose(np.asarray(s.B.to_numpy(a)), np.asarray(s.B.to_numpy(b)), rtol=rtol, atol=atol)


RNG = np.random.default_rng(12)
X = RNG.uniform(-2, 2, (40, 3))
Y = RNG.uniform(-2, 2, (25, 3))


def sq(t):  # works on torch tensors (the library) and numpy arrays (the oracle)
    return t**2


def test_kernel_maps_pairwise_and_elwise(S):
    base = S.EQ().stretch(1.3) + 0.5 * S.Linear()
    spec = ("sum", ("stretched", 1.3, ("eq",)), ("scaled", 0.5, ("linear",)))
    cases = [
        (base.shift(0.7), ("shifted", 0.7, spec)),
        (base.shift(np.array([0.1, -0.2, 0.3])), ("shifted", np.array([0.1, -0.2, 0.3]), spec)),
        (base.select((0, 2)), ("selected", (0, 2), spec)),  # one tuple = both inputs (two arguments = one per input)
        (base.select([1]), ("selected", (1,), spec)),
        (base.transform(sq), ("transformed", sq, spec)),
        (base.shift(0.7, 0), ("shifted2", (0.7, None), spec)),
        (base.shift(0.4, -0.3), ("shifted2", (0.4, -0.3), spec)),
        (base.select((0, 1), None), ("selected2", ((0, 1), None), spec)),
        (base.transform(sq, None), ("transformed2", (sq, None), spec)),
        (base.stretch(2.0, 1), ("stretched2", (2.0, None), spec)),
        (base.stretch(2.0, 0.5), ("stretched2", (2.0, 0.5), spec)),
    ]
    for k, ospec in cases:
        if ospec[0] == "selected2":  # k(x[:, :2], y) needs matching widths: compare on 2-column second arguments
            approx(k(X, Y[:, :2]), O.kernel_matrix(ospec, X, Y[:, :2]))
            continue
        approx(k(X, Y), O.kernel_matrix(ospec, X, Y))
        approx(S.B.dense(k(X)), O.kernel_matrix(ospec, X) if not ospec[0].endswith("2") else O.kernel_matrix(ospec, X, X))
        approx(k.elwise(X, X + 0.1), O.kernel_elwise(ospec, X, X + 0.1))


def test_maps_compose_and_stay_symmetric(S):
    k = S.Matern52().stretch(0.8).shift(1.0).select((2, 0))
    spec = ("selected", (2, 0), ("shifted", 1.0, ("stretched", 0.8, ("matern52",))))
    approx(S.B.dense(k(X)), O.kernel_matrix(spec, X))
    assert k.symmetric and k.reversed() is k
    k2 = S.EQ().shift(0.5, 0)
    assert not k2.symmetric
    approx(k2.reversed()(X, Y), O.kernel_matrix(("shifted2", (None, 0.5), ("eq",)), X, Y))
    assert isinstance(S.EQ().shift(0, 0), type(S.EQ()))  # identity maps drop out


def test_gp_shift_select_transform_logpdf_and_posterior(S):
    y = np.sin(X[:, 0]) + 0.1 * RNG.standard_normal(len(X))
    xs = RNG.uniform(-2, 2, (15, 3))
    kspec = ("stretched", 0.9, ("matern32",))
    for make, wrap in [
        (lambda f: f.shift(0.5), lambda s: ("shifted", 0.5, s)),
        (lambda f: f.select(0, 1), lambda s: ("selected", (0, 1), s)),
        (lambda f: f.transform(sq), lambda s: ("transformed", sq, s)),
    ]:
        f = make(S.GP(S.Matern32().stretch(0.9)))
        spec = wrap(kspec)
        approx(f(X, 0.2).logpdf(y), O.fdd_logpdf(spec, X, 0.2, y), rtol=1e-10, atol=0)
        post = f | (f(X, 0.2), y)
        mean, var = post(xs).marginals()
        mo, vo = O.posterior_marginals(spec, X, 0.2, y, xs)
        approx(mean, np.ravel(mo), atol=1e-8)
        approx(var, np.ravel(vo), atol=1e-8)


def test_shifted_gp_is_correlated_with_its_source(S):
    """The cross-kernel of ``f.shift(c)`` with ``f`` maps only one argument (``measure.py:286``): check the joint."""
    x = np.linspace(0, 3, 12)[:, None]
    c = 0.4
    m = S.Measure()
    f = S.GP(S.EQ(), measure=m)
    g = f.shift(c)
    y1, y2 = np.sin(x[:, 0]), np.sin(x[:, 0] - c)
    k = ("eq",)
    K = np.block([
        [O.kernel_matrix(k, x), O.kernel_matrix(("shifted2", (None, c), k), x, x)],
        [O.kernel_matrix(("shifted2", (c, None), k), x, x), O.kernel_matrix(("shifted", c, k), x)],
    ]) + 0.05 * np.eye(24)
    want = O.normal_logpdf(np.zeros((24, 1)), K, np.concatenate([y1, y2])[:, None])
    approx(m.logpdf((f(x, 0.05), y1), (g(x, 0.05), y2)), np.ravel(want)[0], rtol=1e-9, atol=0)
    # observing f pins down g = f(. - c)
    post = g | (f(x, 1e-6), y1)
    approx(post(x[4:8] + c).mean, y1[4:8, None], atol=2e-3)


def test_mean_follows_the_input_map(S):
    f = S.GP(sq, S.EQ())
    x = np.linspace(-1, 1, 7)
    approx(f.shift(0.5)(x).mean, (x[:, None] - 0.5) ** 2)
    approx(f.transform(lambda t: 2 * t)(x).mean, (2 * x[:, None]) ** 2)
    x2 = RNG.standard_normal((6, 2))
    f2 = S.GP(lambda t: t.sum(-1, keepdim=True), S.EQ())
    approx(f2.select(1)(x2).mean, x2[:, 1:2])


def test_display_of_mapped_kernels(S):
    assert str(S.EQ().shift(1.0)) == "EQ() shift 1"
    assert str(S.EQ().select((0, 2))) == "EQ() : [0, 2]"
    assert "transform" in str(S.EQ().transform(sq))


# ---- GP * function and the moment-matched GP * GP (stheno/model/measure.py:241-270) ---------------------------------
def gfun(t):
    return (torch.sin(t) if isinstance(t, torch.Tensor) else np.sin(t)) + 2.0


def test_gp_times_function(S):
    x = np.linspace(0, 3, 14)
    xs = np.linspace(0.2, 2.8, 5)
    m = S.Measure()
    f = S.GP(sq, S.EQ().stretch(0.7), measure=m)
    h = f * gfun
    h2 = gfun * f  # either order
    spec = ("stretched", 0.7, ("eq",))
    K, gx = O.kernel_matrix(spec, x), np.sin(x) + 2.0
    approx(S.B.dense(h(x).var), gx[:, None] * K * gx[None, :])
    approx(S.B.dense(h2(x).var), gx[:, None] * K * gx[None, :])
    approx(h(x).mean, (gx * x**2)[:, None])
    approx(h(x).var_diag if hasattr(h(x), "var_diag") else np.diag(S.B.to_numpy(S.B.dense(h(x).var))), gx**2 * np.diag(K))
    approx(S.B.dense(m.kernels[h, f](x)), gx[:, None] * K)  # only the left argument is scaled
    approx(S.B.dense(m.kernels[f, h](x)), K * gx[None, :])
    # log-pdf of the product, and the posterior of h after observing f
    y = RNG.standard_normal(len(x))
    want = O.normal_logpdf((gx * x**2)[:, None], gx[:, None] * K * gx[None, :] + 0.1 * np.eye(len(x)), y[:, None])
    approx(h(x, 0.1).logpdf(y), np.ravel(want)[0], rtol=1e-9, atol=0)
    yf = np.cos(x)
    post = h | (f(x, 0.05), yf)
    Ks = O.kernel_matrix(spec, xs, x)
    gs = np.sin(xs) + 2.0
    A = np.linalg.solve(K + 0.05 * np.eye(len(x)) + 1e-12 * np.eye(len(x)), yf - x**2)
    approx(post(xs).mean, (gs * (xs**2 + Ks @ A))[:, None], atol=1e-8)


def test_gp_times_gp_moment_matching(S):
    x = np.linspace(0, 2, 9)
    m = S.Measure()
    f1 = S.GP(sq, S.EQ(), measure=m)
    f2 = S.GP(lambda t: t + 1.0, S.Matern32().stretch(1.5), measure=m)
    p = f1 * f2
    K1 = O.kernel_matrix(("eq",), x)
    K2 = O.kernel_matrix(("stretched", 1.5, ("matern32",)), x)
    m1, m2 = x**2, x + 1.0
    approx(p(x).mean, (m1 * m2)[:, None])
    approx(S.B.dense(p(x).var), np.outer(m1, m1) * K2 + np.outer(m2, m2) * K1 + K1 * K2)
    # the product is correlated with its factors: cov(f1 f2, f1) = m2 k1
    approx(S.B.dense(m.kernels[p, f1](x)), m2[:, None] * K1)
    approx(S.B.dense(m.kernels[f2, p](x)), K2 * m1[None, :])


def test_periodic_kernel(S):
    x = np.linspace(0, 7, 30)
    y = np.sin(2 * np.pi * x / 2.5) + 0.05 * RNG.standard_normal(len(x))
    k = S.EQ().stretch(0.8).periodic(2.5)
    spec = ("periodic", 2.5, ("stretched", 0.8, ("eq",)))
    approx(S.B.dense(k(x)), O.kernel_matrix(spec, x))
    approx(k(x, x + 2.5), O.kernel_matrix(spec, x))  # period 2.5
    f = S.GP(k)
    approx(f(x, 0.1).logpdf(y), O.fdd_logpdf(spec, x, 0.1, y), rtol=1e-10, atol=0)
    assert str(S.EQ().periodic(2.0)) == "EQ() per 2"


def test_w2_between_normals(S):
    rng = np.random.default_rng(3)
    n = 6
    A1, A2 = rng.standard_normal((n, n)), rng.standard_normal((n, n))
    V1, V2 = A1 @ A1.T + 0.5 * np.eye(n), A2 @ A2.T + 0.5 * np.eye(n)
    m1, m2 = rng.standard_normal((n, 1)), rng.standard_normal((n, 1))
    d1, d2 = S.Normal(m1, V1), S.Normal(m2, V2)

    def root(a):
        lam, v = np.linalg.eigh(a)
        return (v * np.sqrt(np.maximum(lam, 0))) @ v.T

    r1 = root(V1)
    want = np.sqrt(np.sum((m1 - m2) ** 2) + np.trace(V1) + np.trace(V2) - 2 * np.trace(root(r1 @ V2 @ r1)))
    approx(d1.w2(d2), want, rtol=1e-8)
    approx(d1.w2(d1), 0.0, atol=1e-6)
    approx(d1.w2(d2), d2.w2(d1), rtol=1e-8)


# ---- the reference's own identity tests for these operations (behaviour of tests/model/test_model.py:429-507) ----------
def assert_equal_normals(S, d1, d2, atol=1e-6):
    approx(d1.mean, d2.mean, atol=atol, rtol=1e-6)
    approx(S.B.dense(d1.var), S.B.dense(d2.var), atol=atol, rtol=1e-6)


def test_reference_shifting_identities(S):
    p = S.GP(lambda t: t**2, S.Linear())
    assert str(p.shift(1)) == "GP(<lambda> shift 1, Linear() shift 1)"
    p_shifted = p.shift(5)
    x = np.linspace(0, 5, 10)
    y = np.asarray(S.B.to_numpy(p_shifted(x).sample())).reshape(-1)
    post = p.measure | (p_shifted(x, 1e-8), y)
    assert_equal_normals(S, post(p(x - 5)), post(p_shifted(x)))
    assert_equal_normals(S, post(p(x)), post(p_shifted(x + 5)))


def test_reference_input_transform_identities(S):
    p = S.GP(lambda t: t**2, S.Linear())
    assert str(p.transform(lambda t: t)) == "GP(<lambda> transform <lambda>, Linear() transform <lambda>)"
    root = lambda t: torch.sqrt(t) if isinstance(t, torch.Tensor) else np.sqrt(t)
    p_t = p.transform(root)
    x = np.linspace(0.1, 5, 10)
    y = np.asarray(S.B.to_numpy(p_t(x).sample())).reshape(-1)
    post = p.measure | (p_t(x, 1e-8), y)
    assert_equal_normals(S, post(p(np.sqrt(x))), post(p_t(x)))
    assert_equal_normals(S, post(p(x)), post(p_t(x * x)))


def test_reference_selection_identities(S):
    p = S.GP(lambda t: t**2, S.EQ())
    assert str(p.select(1)) == "GP(<lambda> : [1], EQ() : [1])"
    assert str(p.select(1, 2)) == "GP(<lambda> : [1, 2], EQ() : [1, 2])"
    p2 = p.select(0)  # a GP on 2-D inputs that only looks at the first column
    x = np.linspace(0, 5, 10)
    rng = np.random.default_rng(1)
    x21 = np.stack([x, rng.standard_normal(10)], axis=1)
    x22 = np.stack([x, rng.standard_normal(10)], axis=1)
    y = np.asarray(S.B.to_numpy(p2(x21).sample())).reshape(-1)
    post = p.measure | (p2(x21, 1e-8), y)
    approx(post(p(x)).mean, y[:, None], atol=1e-4)
    assert_equal_normals(S, post(p(x)), post(p2(x21)), atol=1e-5)
    post = p.measure | (p(x, 1e-8), y)
    approx(post(p2(x22)).mean, y[:, None], atol=1e-4)
    assert_equal_normals(S, post(p2(x21)), post(p(x)), atol=1e-5)


def test_reference_stretching_identities(S):
    """The cross-kernel of a stretched GP stretches one argument only (``measure.py:305``) -- now a per-argument map."""
    p = S.GP(lambda t: t**2, S.Linear())
    p_stretched = p.stretch(5)
    x = np.linspace(0, 5, 10)
    y = np.asarray(S.B.to_numpy(p_stretched(x).sample())).reshape(-1)
    post = p.measure | (p_stretched(x, 1e-8), y)
    assert_equal_normals(S, post(p(x / 5)), post(p_stretched(x)))
    assert_equal_normals(S, post(p(x)), post(p_stretched(x * 5)))


def test_reference_approximate_multiplication(S):
    """``tests/model/test_model.py:573-592``: the moment-matched product tracks the product of the sampled factors."""
    m = S.Measure()
    p1 = S.GP(20, S.EQ(), measure=m)
    p2 = S.GP(20, S.EQ(), measure=m)
    p_prod = p1 * p2
    x = np.linspace(0, 10, 50)
    s1, s2 = m.sample(p1(x), p2(x))
    s1, s2 = np.asarray(S.B.to_numpy(s1)), np.asarray(S.B.to_numpy(s2))
    post = m | ((p1(x), s1), (p2(x), s2))
    approx(post(p_prod)(x).mean, s1 * s2, rtol=5e-2, atol=0)
