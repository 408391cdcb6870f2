"""Regression tests for the round-1 advisor findings: behaviours the reference supports that failed here."""
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


def test_stretch_of_composite_and_posterior_means(S):
    """``StretchedMean`` over scaled / summed / posterior means (only ``dev`` is defined for those)."""
    x = np.linspace(0, 3, 25)
    f = S.GP(lambda t: t ** 2, S.EQ())
    g = S.GP(lambda t: torch.sin(t), S.Matern52(), measure=f.measure)
    np.testing.assert_allclose((2 * f).stretch(3.0)(x).mean[:, 0], 2 * (x / 3) ** 2, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose((f + g).stretch(3.0)(x).mean[:, 0], (x / 3) ** 2 + np.sin(x / 3), rtol=1e-12, atol=1e-12)
    xo = np.linspace(0, 2, 12)
    yo = np.cos(xo)
    post = f | (f(xo, 0.1), yo)
    m = post.stretch(2.0)(x).mean
    np.testing.assert_allclose(m, post(x / 2.0).mean, rtol=1e-10, atol=1e-10)


def test_more_than_eight_summed_gps(S):
    """Additive models with more components than one K1 descriptor holds fall back to per-child evaluation."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal((40, 2))
    y = rng.standard_normal(40)
    m = S.Measure()
    comps, spec = [], None
    for i in range(10):
        c, l = 0.5 + 0.1 * i, 0.7 + 0.2 * i
        comps.append(S.GP(c * S.EQ().stretch(l), measure=m))
        node = ("scaled", c, ("stretched", l, ("eq",)))
        spec = node if spec is None else ("sum", spec, node)
    f = comps[0]
    for g in comps[1:]:
        f = f + g
    want = float(O.fdd_logpdf(spec, x, 0.1, y))
    assert abs(float(f(x, 0.1).logpdf(y)) - want) < 1e-9 * abs(want)
    # a product that expands to 9 terms
    k3 = S.EQ() + S.Matern32() + S.Linear()
    k9 = k3 * (S.EQ().stretch(2.0) + S.Matern52() + S.Matern32().stretch(0.5))
    s3 = ("sum", ("sum", ("eq",), ("matern32",)), ("linear",))
    s9 = ("product", s3, ("sum", ("sum", ("stretched", 2.0, ("eq",)), ("matern52",)), ("stretched", 0.5, ("matern32",))))
    np.testing.assert_allclose(S.B.dense(k9(x)), O.kernel_matrix(s9, x), rtol=1e-9, atol=2e-7)
    np.testing.assert_allclose(S.B.dense(k9(x, x[:7])), O.kernel_matrix(s9, x, x[:7]), rtol=1e-9, atol=2e-7)
    np.testing.assert_allclose(k9.elwise(x)[:, 0], np.diag(O.kernel_matrix(s9, x)), rtol=1e-9, atol=2e-7)


def test_conditioning_a_pure_noise_process(S):
    """``e = GP(Delta()); e | (e(x, 0.1), y)``: K_x is Diagonal and still needs a factor object."""
    x = np.linspace(0, 1, 9)
    y = np.sin(x)
    e = S.GP(S.Delta())
    post = e | (e(x, 0.1), y)
    xs = np.linspace(2, 3, 4)  # away from the data: prior
    mean, var = post(xs).marginals()
    np.testing.assert_allclose(mean, 0.0, atol=1e-12)
    np.testing.assert_allclose(var, 1.0, atol=1e-12)


@pytest.mark.gpu
def test_scale_with_grad_keeps_its_graph():
    """``s * Linear()`` / ``s * Delta()`` with a tensor scale that requires grad (the structured shortcuts detached it).
    GPU only: the analytic backward is the K1-backward CUDA kernel."""
    import stheno_b200 as S

    S.B.epsilon = 1e-12
    S.Measure.default = None
    dev = "cuda"
    g = torch.Generator().manual_seed(0)
    x = torch.randn(30, 3, dtype=torch.float64, generator=g).to(dev)
    y = torch.randn(30, dtype=torch.float64, generator=g).to(dev)

    def ref(s, kind):
        K = s * (x @ x.T) if kind == "linear" else torch.exp(-0.5 * torch.cdist(x, x) ** 2) + s * torch.eye(30, dtype=x.dtype, device=x.device)
        if kind == "linear":
            K = K + 0.1 * torch.eye(30, dtype=x.dtype, device=x.device)
        K = K + 1e-12 * torch.eye(30, dtype=x.dtype, device=x.device)
        L = torch.linalg.cholesky(K)
        a = torch.linalg.solve_triangular(L, y[:, None], upper=False)
        return -0.5 * (2 * torch.log(torch.diagonal(L)).sum() + 30 * np.log(2 * np.pi) + (a * a).sum())

    for kind in ("linear", "delta"):
        s = torch.tensor(0.7, dtype=torch.float64, device=dev, requires_grad=True)
        if kind == "linear":
            lp = S.GP(s * S.Linear())(x, 0.1).logpdf(y)
        else:
            lp = S.GP(S.EQ() + s * S.Delta())(x).logpdf(y)
        assert lp.requires_grad
        lp.backward()
        s2 = torch.tensor(0.7, dtype=torch.float64, device=dev, requires_grad=True)
        want = ref(s2, kind)
        want.backward()
        assert abs(lp.item() - want.item()) < 1e-9 * abs(want.item())
        assert abs(s.grad.item() - s2.grad.item()) < 1e-7 * max(1.0, abs(s2.grad.item())), (kind, s.grad, s2.grad)


def test_strict_mode_raises_on_non_positive_definite(S):
    """``B.strict``: a failed factorisation raises like the reference's backend instead of returning NaN."""
    bad = np.array([[1.0, 2.0], [2.0, 1.0]])
    S.B.strict = True
    try:
        with pytest.raises(torch.linalg.LinAlgError):
            S.Normal(bad).logpdf(np.array([0.1, 0.2]))
    finally:
        S.B.strict = False
    assert np.isnan(S.Normal(bad).logpdf(np.array([0.1, 0.2])))


@pytest.mark.parametrize("method", ["vfe", "fitc", "dtc"])
def test_sparse_elbo_gradients_reach_every_parameter(S, method):
    """``elbo.backward()`` returned a partial gradient (noise only).  Now: gradients w.r.t. kernel variance,
    length scale, noise and inducing points, against central finite differences of the ORACLE's ELBO."""
    rng = np.random.default_rng(3)
    n, m, d = 60, 7, 2
    x, z, y = rng.standard_normal((n, d)), rng.standard_normal((m, d)), rng.standard_normal(n)
    cls = {"vfe": S.PseudoObs, "fitc": S.PseudoObsFITC, "dtc": S.PseudoObsDTC}[method]
    p0 = np.array([1.3, 0.9, 0.2])

    def oracle(p, zz):
        spec = ("scaled", p[0], ("stretched", p[1], ("matern52",)))
        return O.sparse_compute(spec, zz, x, p[2], y, method)["elbo"]

    var, ell, noise = (torch.tensor(v, dtype=torch.float64, requires_grad=True) for v in p0)
    zt = torch.tensor(z, requires_grad=True)
    f = S.GP(var * S.Matern52().stretch(ell))
    obs = cls(f(zt), f(torch.tensor(x), noise), torch.tensor(y))
    elbo = obs.elbo(f.measure)
    assert abs(float(elbo) - oracle(p0, z)) < 1e-9 * abs(float(elbo))
    elbo.backward()
    h = 1e-6
    for i, t in enumerate((var, ell, noise)):
        e = np.zeros(3)
        e[i] = h
        fd = (oracle(p0 + e, z) - oracle(p0 - e, z)) / (2 * h)
        assert t.grad is not None and abs(t.grad.item() - fd) < 1e-5 * max(1.0, abs(fd)), (method, i, t.grad, fd)
    zp, zm = z.copy(), z.copy()
    zp[2, 1] += h
    zm[2, 1] -= h
    fd = (oracle(p0, zp) - oracle(p0, zm)) / (2 * h)
    assert abs(zt.grad[2, 1].item() - fd) < 1e-5 * max(1.0, abs(fd))
    # posterior through the differentiable route agrees with the oracle as well
    mean = (f | obs)(x[:5]).mean
    mo, _ = O.sparse_posterior(("scaled", p0[0], ("stretched", p0[1], ("matern52",))), z, x, p0[2], y, x[:5], method)
    np.testing.assert_allclose(S.B.to_numpy(mean), mo, rtol=1e-7, atol=1e-8)


def test_woodbury_logpdf_gradients(S):
    """``GP(s * Linear())(x, noise).logpdf(y)`` keeps the O(n d^2) Woodbury route AND a complete graph."""
    rng = np.random.default_rng(8)
    n, d = 50, 3
    x, y = rng.standard_normal((n, d)), rng.standard_normal(n)

    def oracle(p):
        return float(O.fdd_logpdf(("scaled", p[0], ("linear",)), x, p[1], y))

    p0 = np.array([0.7, 0.3])
    s, noise = (torch.tensor(v, dtype=torch.float64, requires_grad=True) for v in p0)
    fdd = S.GP(s * S.Linear())(torch.tensor(x), noise)
    assert type(fdd.var).__name__ == "Woodbury"
    lp = fdd.logpdf(torch.tensor(y))
    assert abs(float(lp) - oracle(p0)) < 1e-9 * abs(oracle(p0))
    lp.backward()
    h = 1e-6
    for i, t in enumerate((s, noise)):
        e = np.zeros(2)
        e[i] = h
        fd = (oracle(p0 + e) - oracle(p0 - e)) / (2 * h)
        assert t.grad is not None and abs(t.grad.item() - fd) < 1e-5 * max(1.0, abs(fd)), (i, t.grad, fd)
