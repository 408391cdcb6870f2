"""Behaviour of the stheno-compatible model layer against the oracle and the reference's own identity tests
(tests/model/test_model.py, test_gp.py, test_fdd.py, test_cases.py, tests/test_random.py of the reference).

Every test runs twice: on the CPU with the torch stand-in backend (host logic; ``-m "not gpu"``) and on the GPU
through the real CUDA kernels (``-m gpu``)."""
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
    def conv(v):
        import stheno_b200 as s

        v = s.B.to_numpy(v)
        return np.asarray(v)

    np.testing.assert_allclose(conv(a), conv(b), rtol=rtol, atol=atol)


def test_readme_regression_G1(S):
    x = np.linspace(0, 2, 10)
    y = x**2
    f = S.GP(S.EQ())
    f_post = f | (f(x), y)
    pred = f_post(np.array([1.0, 2.0, 3.0]))
    mean = S.B.dense(pred.mean)
    assert isinstance(mean, np.ndarray) and mean.shape == (3, 1)
    np.testing.assert_allclose(mean[:, 0], [1.00000068, 3.99999999, 8.4825932], atol=5e-6)
    var = S.B.dense(pred.var)
    assert var.shape == (3, 3)
    assert abs(var[2, 2] - 3.31283378e-03) < 2e-5


def test_readme_logpdf_G3(S):
    x = np.array([0.0, 1.0, 2.0])
    f = S.GP(S.EQ())
    approx(S.B.dense(f(x).var), O.kernel_matrix(("eq",), x))
    approx(f(x).mean, np.zeros((3, 1)))
    y1 = np.array([-0.45172746, 0.46581948, 0.78929767])
    lp = f(x).logpdf(y1)
    assert np.ndim(lp) == 0
    assert abs(float(lp) + 2.811609567720761) < 5e-8
    y2 = np.array([[-0.43771276, -2.36741858], [0.86080043, -1.22503079], [2.15779126, -0.75319405]])
    approx(f(x).logpdf(y2), [-4.82949038, -5.40084225], atol=5e-7)


KERNEL_CASES = [
    (lambda S: S.EQ(), ("eq",)),
    (lambda S: 2.0 * S.EQ().stretch(1.5), ("scaled", 2.0, ("stretched", 1.5, ("eq",)))),
    (lambda S: S.Matern52().stretch(np.array([1.0, 2.0])) + 0.5 * S.Matern32(),
     ("sum", ("stretched", np.array([1.0, 2.0]), ("matern52",)), ("scaled", 0.5, ("matern32",)))),
    (lambda S: S.Matern32() * S.EQ().stretch(3.0) + S.Linear(),
     ("sum", ("product", ("matern32",), ("stretched", 3.0, ("eq",))), ("linear",))),
    # Matern12 only in d = 1: for d > 1 the reference's own diagonal carries ~6e-8 of rounding noise (see
    # tests/test_gpu_primitives.py), which caps reference-vs-anything agreement of logpdf at ~1e-9.
    (lambda S: S.Exp().stretch(0.7) + 0.1 * S.EQ(), ("sum", ("stretched", 0.7, ("matern12",)), ("scaled", 0.1, ("eq",)))),
]


@pytest.mark.parametrize("case", range(len(KERNEL_CASES)))
@pytest.mark.parametrize("noise_kind", ["scalar", "vector", "matrix", "none_plus_delta"])
def test_logpdf_and_posterior_vs_oracle(S, case, noise_kind):
    mk, spec = KERNEL_CASES[case]
    rng = np.random.default_rng(case)
    n, m, d = 40, 15, (1 if case == 4 else 2)
    x = rng.standard_normal((n, d))
    xs = rng.standard_normal((m, d))
    y = rng.standard_normal(n)
    k = mk(S)
    if noise_kind == "scalar":
        noise, onoise, ospec = 0.3, 0.3, spec
    elif noise_kind == "vector":
        noise = rng.uniform(0.2, 0.5, n)
        onoise, ospec = noise, spec
    elif noise_kind == "matrix":
        a = rng.standard_normal((n, n))
        noise = 0.01 * a @ a.T + 0.2 * np.eye(n)
        onoise, ospec = noise, spec
    else:
        k = k + 0.3 * S.Delta()
        noise, onoise, ospec = None, None, ("sum", spec, ("scaled", 0.3, ("delta",)))
    f = S.GP(k)
    approx(f(x, noise).logpdf(y), O.fdd_logpdf(ospec, x, onoise, y), rtol=1e-10, atol=0)
    post = f | (f(x, noise), y)
    if noise_kind == "none_plus_delta":
        # predicting the noisy process at NEW points: Delta contributes nothing across, 0.3 on the prior diagonal
        mean_ref, var_ref = O.posterior(ospec, x, None, y, xs)
    else:
        mean_ref, var_ref = O.posterior(ospec, x, onoise, y, xs)
    pred = post(xs)
    approx(pred.mean, mean_ref, rtol=1e-8, atol=1e-9)
    approx(S.B.dense(pred.var), var_ref, rtol=1e-7, atol=1e-8)
    mm, vv = post(xs).marginals()
    approx(mm, mean_ref[:, 0], rtol=1e-8, atol=1e-9)
    approx(vv, np.maximum(np.diag(var_ref), 0), rtol=1e-7, atol=1e-8)
    mean2, var2 = post(xs).mean_var
    approx(mean2, mean_ref, rtol=1e-8, atol=1e-9)
    approx(S.B.dense(var2), var_ref, rtol=1e-7, atol=1e-8)
    m3, lo, hi = post(xs).marginal_credible_bounds()
    approx(hi - lo, 2 * 1.96 * np.sqrt(np.maximum(np.diag(var_ref), 0)), rtol=1e-6, atol=1e-7)


def test_prior_var_and_noise(S):
    # tests/model/test_gp.py:84-92
    x = np.linspace(0, 1, 7)
    f = S.GP(S.EQ())
    approx(S.B.dense(f(x).var), O.kernel_matrix(("eq",), x))
    approx(S.B.dense(f(x, 1.0).var), O.kernel_matrix(("eq",), x) + np.eye(7))
    approx(f(x, 1.0).var_diag, np.full(7, 2.0))


def test_conditioning_syntaxes_agree(S):
    # tests/model/test_model.py:123-178
    rng = np.random.default_rng(3)
    x = rng.standard_normal((12, 1))
    y = rng.standard_normal(12)
    xs = np.linspace(-1, 1, 5)
    m = S.Measure()
    f = S.GP(S.EQ(), measure=m)
    e = S.GP(0.2 * S.Delta(), measure=m)
    ref = O.posterior(("eq",), x, 0.2, y, xs)[0]
    for post in (
        f | (f(x, 0.2), y),
        f | S.Obs(f(x, 0.2), y),
        f.condition(f(x, 0.2), y),
        (m | (f(x, 0.2), y))(f),
        m.condition(S.Obs(f(x, 0.2), y))(f),
        f | ((f + e)(x), y),  # noise as a process == noise as an argument (tests/model/test_model.py:181-195)
    ):
        approx(post(xs).mean, ref, rtol=1e-8, atol=1e-9)


def test_posterior_concentrates_and_reverts(S):
    # tests/model/test_gp.py:177-198
    x = np.linspace(0, 5, 10)
    y = np.sin(x)
    f = S.GP(S.EQ())
    post = f | (f(x), y)
    mean, var = post(x).marginals()
    approx(mean, y, atol=1e-4)
    assert np.max(np.abs(var)) < 1e-4
    far = post(np.array([100.0, 200.0]))
    approx(far.mean, np.zeros((2, 1)), atol=1e-8)
    approx(far.var_diag, np.ones(2), atol=1e-8)


def test_empty_observations_return_prior(S):
    f = S.GP(S.EQ())
    post = f | (f(np.zeros((0, 1))), np.zeros((0, 1)))
    xs = np.array([0.0, 1.0])
    approx(S.B.dense(post(xs).var), O.kernel_matrix(("eq",), xs))


def test_missing_data(S):
    # tests/model/test_model.py:231-238 and tests/test_random.py:195-204
    rng = np.random.default_rng(5)
    x = rng.standard_normal((10, 1))
    y = rng.standard_normal(10)
    y_nan = y.copy()
    y_nan[[2, 7]] = np.nan
    keep = ~np.isnan(y_nan)
    f = S.GP(S.EQ())
    approx(f(x, 0.1).logpdf(y_nan), O.fdd_logpdf(("eq",), x[keep], 0.1, y[keep]), rtol=1e-10)
    xs = np.array([0.3, 0.6])
    post = f | (f(x, 0.1), y_nan)
    approx(post(xs).mean, O.posterior(("eq",), x[keep], 0.1, y[keep], xs)[0], rtol=1e-8, atol=1e-9)


def test_shape_errors(S):
    f = S.GP(S.EQ())
    x = np.linspace(0, 1, 5)
    with pytest.raises(ValueError):
        f | (f(x), np.ones((5, 2)))
    u = S.GP(S.EQ(), measure=f.measure)
    with pytest.raises(RuntimeError):
        S.PseudoObs(f(x), f(x, np.eye(5)), np.ones(5)).elbo(f.measure)
    with pytest.raises(RuntimeError):
        S.GP().measure


def test_sum_scale_stretch_and_decomposition(S):
    rng = np.random.default_rng(7)
    x = rng.standard_normal((25, 1))
    xs = np.linspace(-2, 2, 9)
    y = rng.standard_normal(25)
    m = S.Measure()
    f1 = S.GP(S.EQ().stretch(2.0), measure=m)
    f2 = 0.5 * S.GP(S.Matern32(), measure=m)
    f = f1 + f2
    spec1 = ("stretched", 2.0, ("eq",))
    spec2 = ("scaled", 0.25, ("matern32",))
    spec = ("sum", spec1, spec2)
    approx(S.B.dense(f(x).var), O.kernel_matrix(spec, x))
    approx(f(x, 0.1).logpdf(y), O.fdd_logpdf(spec, x, 0.1, y), rtol=1e-10)
    # decomposition: condition on the sum, predict a component (README example 2)
    post = m | (f(x, 0.1), y)
    K = O.kernel_matrix(spec, x) + 0.1 * np.eye(25)
    L = O.chol_eps(K)
    K1s = O.kernel_matrix(spec1, x, xs)
    a = np.linalg.solve(L, K1s)
    b = np.linalg.solve(L, y[:, None])
    approx(post(f1)(xs).mean, a.T @ b, rtol=1e-8, atol=1e-9)
    approx(S.B.dense(post(f1)(xs).var), O.kernel_matrix(spec1, xs) - a.T @ a, rtol=1e-7, atol=1e-8)
    # stretch of a GP and a mean function
    g = S.GP(lambda t: t**2, S.EQ(), measure=m).stretch(3.0)
    approx(g(xs).mean, (xs[:, None] / 3.0) ** 2)
    approx(S.B.dense(g(xs).var), O.kernel_matrix(("stretched", 3.0, ("eq",)), xs))


def test_repeated_conditioning_and_chain_rule(S):
    # tests/model/test_model.py:211-228, 375-404
    rng = np.random.default_rng(8)
    x1, x2 = rng.standard_normal((8, 1)), rng.standard_normal((6, 1))
    y1, y2 = rng.standard_normal(8), rng.standard_normal(6)
    f = S.GP(S.EQ())
    joint = f.measure.logpdf((f(x1, 0.1), y1), (f(x2, 0.1), y2))
    ref = O.fdd_logpdf(("eq",), np.concatenate([x1, x2]), 0.1, np.concatenate([y1, y2]))
    approx(joint, ref, rtol=1e-10)
    post1 = f | (f(x1, 0.1), y1)
    chain = f(x1, 0.1).logpdf(y1) + post1(x2, 0.1).logpdf(y2)
    approx(chain, ref, rtol=1e-9)
    post12 = post1 | (post1(x2, 0.1), y2)
    xs = np.linspace(-1, 1, 4)
    ref_mean = O.posterior(("eq",), np.concatenate([x1, x2]), 0.1, np.concatenate([y1, y2]), xs)[0]
    approx(post12(xs).mean, ref_mean, rtol=1e-7, atol=1e-8)


@pytest.mark.parametrize("method", ["vfe", "fitc", "dtc"])
def test_sparse_vs_oracle(S, method):
    cls = {"vfe": S.PseudoObs, "fitc": S.PseudoObsFITC, "dtc": S.PseudoObsDTC}[method]
    rng = np.random.default_rng(9)
    n, mz = 150, 20
    x = np.sort(rng.uniform(0, 10, n))
    z = np.linspace(0, 10, mz)
    y = np.sin(x) + 0.3 * rng.standard_normal(n)
    xs = np.linspace(0, 10, 11)
    spec = ("stretched", 1.5, ("matern52",))
    f = S.GP(S.Matern52().stretch(1.5))
    obs = cls(f(z), f(x, 0.09), y)
    c = O.sparse_compute(spec, z, x, 0.09, y, method)
    approx(obs.elbo(f.measure), c["elbo"], rtol=1e-9)
    approx(f.measure.logpdf(obs), c["elbo"], rtol=1e-9)
    approx(obs.mu(f.measure), c["mu"], rtol=1e-6, atol=1e-7)
    approx(S.B.dense(obs.A(f.measure)), c["A"], rtol=1e-6, atol=1e-7)
    post = f | obs
    mean_ref, var_ref = O.sparse_posterior(spec, z, x, 0.09, y, xs, method)
    approx(post(xs).mean, mean_ref, rtol=1e-6, atol=1e-7)
    approx(S.B.dense(post(xs).var), var_ref, rtol=1e-5, atol=1e-6)


def test_sparse_equals_exact_when_z_is_x(S):
    # tests/model/test_model.py:284-308
    rng = np.random.default_rng(10)
    x = np.linspace(0, 5, 10)
    y = rng.standard_normal(10)
    xs = np.linspace(0, 5, 7) + 0.1
    f = S.GP(S.EQ())
    exact = f(x, 0.1).logpdf(y)
    post = f | (f(x, 0.1), y)
    for cls in (S.PseudoObs, S.PseudoObsFITC, S.PseudoObsDTC):
        obs = cls(f(x), f(x, 0.1), y)
        approx(obs.elbo(f.measure), exact, atol=1e-4, rtol=0)
        sp = f | obs
        approx(sp(xs).mean, post(xs).mean, atol=1e-4)
        approx(S.B.dense(sp(xs).var), S.B.dense(post(xs).var), atol=1e-4)
        assert obs.elbo(f.measure) is not None and id(obs.K_z(f.measure)) == id(obs.K_z(f.measure))


def test_batched(S):
    # tests/model/test_cases.py:134-155
    rng = np.random.default_rng(11)
    x = rng.standard_normal((16, 10, 1))
    y = rng.standard_normal((16, 10, 1))
    f = S.GP(S.EQ())
    lp = f(x, 0.1).logpdf(y)
    assert lp.shape == (16,)
    for b in (0, 7, 15):
        approx(lp[b], O.fdd_logpdf(("eq",), x[b], 0.1, y[b]), rtol=1e-10)
    post = f | (f(x, 0.1), y)
    xs = rng.standard_normal((16, 4, 1))
    mean = post(xs).mean
    assert mean.shape == (16, 4, 1)
    approx(mean[3], O.posterior(("eq",), x[3], 0.1, y[3], xs[3])[0], rtol=1e-8, atol=1e-9)
    s = f(x, 0.1).sample(3)
    assert s.shape == (16, 10, 3)


def test_multi_output_blocks_and_joint_logpdf(S):
    # tests/mo/test_kernel.py:30-102, tests/model/test_observations.py:8-41
    rng = np.random.default_rng(12)
    m = S.Measure()
    u1 = S.GP(S.EQ(), measure=m)
    u2 = S.GP(S.Matern32().stretch(2.0), measure=m)
    f1 = 1.0 * u1 + 0.5 * u2
    f2 = -0.7 * u1 + 2.0 * u2
    x1, x2 = rng.standard_normal((7, 1)), rng.standard_normal((5, 1))
    s1, s2 = ("eq",), ("stretched", 2.0, ("matern32",))
    def kk(a, b, xa, xb=None):
        return a[0] * b[0] * O.kernel_matrix(s1, xa, xb) + a[1] * b[1] * O.kernel_matrix(s2, xa, xb)
    h1, h2 = (1.0, 0.5), (-0.7, 2.0)
    K = np.block([[kk(h1, h1, x1), kk(h1, h2, x1, x2)], [kk(h2, h1, x2, x1), kk(h2, h2, x2)]])
    p = S.cross(f1, f2)
    joint = p((f1(x1), f2(x2)))
    approx(S.B.dense(joint.var), K, rtol=1e-10, atol=1e-12)
    y1, y2 = rng.standard_normal(7), rng.standard_normal(5)
    noise = 0.2 * np.eye(12)
    ref = O.normal_logpdf(None, K + noise, np.concatenate([y1, y2]))
    approx(m.logpdf((f1(x1, 0.2), y1), (f2(x2, 0.2), y2)), ref, rtol=1e-10)
    # condition on both outputs, predict a latent process
    post = m | ((f1(x1, 0.2), y1), (f2(x2, 0.2), y2))
    xs = np.linspace(-1, 1, 4)
    Kc = np.vstack([h1[0] * O.kernel_matrix(s1, x1, xs), h2[0] * O.kernel_matrix(s1, x2, xs)])
    L = O.chol_eps(K + noise)
    a = np.linalg.solve(L, Kc)
    b = np.linalg.solve(L, np.concatenate([y1, y2])[:, None])
    approx(post(u1)(xs).mean, a.T @ b, rtol=1e-8, atol=1e-9)
    approx(S.B.dense(post(u1)(xs).var), O.kernel_matrix(s1, xs) - a.T @ a, rtol=1e-7, atol=1e-8)
    # same-input multi-output evaluation: p(x) stacks the outputs
    pv = S.B.dense(p(x1).var)
    Kx = np.block([[kk(h1, h1, x1), kk(h1, h2, x1, x1)], [kk(h2, h1, x1, x1), kk(h2, h2, x1)]])
    approx(pv, Kx, rtol=1e-10, atol=1e-12)


def test_sampling_and_arithmetic(S):
    x = np.linspace(0, 3, 30)
    f = S.GP(lambda t: 2 * t, S.EQ())
    fd = f(x, 0.01)
    g = torch.Generator(device=S._util._device_fn()).manual_seed(0)
    state, s1 = fd.sample(g, 4000)
    assert s1.shape == (30, 4000)
    emp_mean = np.asarray(s1).mean(1)
    approx(emp_mean, 2 * x, atol=0.08, rtol=0)
    emp_cov = np.cov(np.asarray(s1))
    approx(emp_cov, O.kernel_matrix(("eq",), x) + 0.01 * np.eye(30), atol=0.12, rtol=0)
    assert fd.sample().shape == (30, 1) and fd.sample(2, noise=0.1).shape == (30, 2)
    a, b = f.measure.sample(f(x[:5]), f(x[5:8]))
    assert a.shape == (5, 1) and b.shape == (3, 1)
    n1 = S.Normal(np.ones((3, 1)), np.eye(3)) + S.Normal(np.eye(3))
    approx(S.B.dense(n1.var), 2 * np.eye(3))
    approx(n1.mean, np.ones((3, 1)))
    approx(S.Normal(np.eye(3)).entropy(), 0.5 * 3 * (np.log(2 * np.pi) + 1), rtol=1e-9)
    p, q = S.Normal(np.zeros((2, 1)), np.eye(2)), S.Normal(np.ones((2, 1)), 2 * np.eye(2))
    kl_ref = 0.5 * (2 * 0.5 + 2 * 0.5 - 2 + 2 * np.log(2))
    approx(p.kl(q), kl_ref, rtol=1e-9)


def test_torch_inputs_stay_torch(S):
    dev = S._util._device_fn()
    x = torch.linspace(0, 1, 6, dtype=torch.float64, device=dev)
    y = torch.sin(x)
    f = S.GP(S.EQ())
    lp = f(x, 0.1).logpdf(y)
    assert isinstance(lp, torch.Tensor) and lp.device.type == dev.type and lp.dim() == 0
    post = f | (f(x, 0.1), y)
    m, v = post(x).marginals()
    assert isinstance(m, torch.Tensor) and m.shape == (6,)
    x32 = x.to(torch.float32)
    S.B.epsilon = 1e-6
    lp32 = f(x32, 0.1).logpdf(y.to(torch.float32))
    assert lp32.dtype == torch.float32
    assert abs(float(lp32) - float(lp)) < 1e-3 * abs(float(lp)) + 1e-3


def test_named_gps_and_default_measure(S):
    with S.Measure() as m:
        f = S.GP(S.EQ(), name="f")
        g = S.GP(S.EQ(), name="g")
    assert f.measure is m and g.measure is m and m["f"] is f and m[g] == "g"
    with pytest.raises(RuntimeError):
        m.name(g, "f")
    h = S.GP(S.EQ())
    assert h.measure is not m
    with pytest.raises(AssertionError):
        f + h


def test_normal_with_explicit_variance(S):
    # tests/test_random.py:185-192 : logpdf against SciPy's formula via the oracle, 10 right-hand sides
    rng = np.random.default_rng(20)
    a = rng.standard_normal((30, 30))
    var = a @ a.T + 0.5 * np.eye(30)
    mean = rng.standard_normal((30, 1))
    x = rng.standard_normal((30, 10))
    dist = S.Normal(mean, var)
    approx(dist.logpdf(x), O.normal_logpdf(mean, var, x), rtol=1e-9)
    assert np.ndim(dist.logpdf(x[:, 0])) == 0
    assert dist.dim == 30 and dist.dtype == torch.float64
    approx(dist.m2, var + mean @ mean.T)
    # Diagonal variance keeps its structure
    dn = S.Normal(S.Diagonal(S._util.to_dev(np.full(30, 2.0))))
    approx(dn.logpdf(x[:, :3]), O.normal_logpdf(None, 2.0 * np.eye(30), x[:, :3]), rtol=1e-12)


def test_float32_and_batched_posterior(S):
    rng = np.random.default_rng(21)
    S.B.epsilon = 1e-6
    x = rng.standard_normal((3, 60, 2)).astype(np.float32)
    y = rng.standard_normal((3, 60, 1)).astype(np.float32)
    xs = rng.standard_normal((3, 9, 2)).astype(np.float32)
    noise = rng.uniform(0.2, 0.4, (3, 60)).astype(np.float32)
    f = S.GP(S.Matern52().stretch(1.2))
    post = f | (f(x, noise), y)
    mean, var = post(xs).marginals()
    assert mean.shape == (3, 9) and mean.dtype == np.float32
    for b in range(3):
        mref, vref = O.posterior_marginals(("stretched", 1.2, ("matern52",)), x[b].astype(np.float64),
                                           noise[b].astype(np.float64), y[b].astype(np.float64),
                                           xs[b].astype(np.float64), eps=1e-6)
        approx(mean[b], mref, rtol=1e-3, atol=1e-4)
        approx(var[b], vref, rtol=1e-3, atol=1e-4)


def test_kernel_beyond_descriptor_limits_falls_back(S):
    # more product terms than ONE K1 descriptor holds -> evaluated per child and combined (the reference has no limit)
    k = S.EQ()
    spec = ("eq",)
    for i in range(9):
        k = k + (i + 2.0) * S.Matern32().stretch(float(i + 2))
        spec = ("sum", spec, ("scaled", i + 2.0, ("stretched", float(i + 2), ("matern32",))))
    x = np.linspace(0, 1, 5)
    approx(S.GP(k)(x).var.mat, O.kernel_matrix(spec, x))


def test_combine_and_multi_fdd_observations(S):
    # tests/model/test_observations.py:8-41 : joint of independent FDDs = concat mean + block-diag var
    rng = np.random.default_rng(30)
    m = S.Measure()
    f1 = S.GP(lambda t: t, S.EQ(), measure=m)
    f2 = S.GP(2.0 * S.Matern32(), measure=m)
    x1, x2 = rng.standard_normal((4, 1)), rng.standard_normal((3, 1))
    joint = S.combine(f1(x1, 0.1), f2(x2, np.array([0.2, 0.3, 0.4])))
    K = np.zeros((7, 7))
    K[:4, :4] = O.kernel_matrix(("eq",), x1) + 0.1 * np.eye(4)
    K[4:, 4:] = 2.0 * O.kernel_matrix(("matern32",), x2) + np.diag([0.2, 0.3, 0.4])
    approx(S.B.dense(joint.var), K, rtol=1e-10, atol=1e-12)
    approx(joint.mean, np.concatenate([x1, np.zeros((3, 1))]))
    # sparse observations with inducing points in two processes
    y1, y2 = rng.standard_normal(4), rng.standard_normal(3)
    obs = S.PseudoObs((f1(x1), f2(x2)), (f1(x1, 0.1), y1), (f2(x2, 0.2), y2))
    exact = m.logpdf((f1(x1, 0.1), y1), (f2(x2, 0.2), y2))
    approx(obs.elbo(m), exact, atol=1e-5, rtol=0)


def test_fdd_take_and_mask_errors(S):
    f = S.GP(S.EQ())
    x = np.linspace(0, 1, 6)
    fd = f(x, np.arange(1.0, 7.0))
    mask = np.array([True, False, True, True, False, True])
    sub = fd.take(mask)
    approx(S.B.dense(sub.var), O.kernel_matrix(("eq",), x[mask]) + np.diag(np.arange(1.0, 7.0)[mask]))
    with pytest.raises(AssertionError):
        fd.take(np.array([0, 2]))


def test_normal_arithmetic_and_lazy_contracts(S):
    # tests/test_random.py:97-158 : laziness contracts of mean_var / marginals
    calls = {"mean": 0, "var": 0, "mv": 0, "mvd": 0}
    dev = S._util._device_fn()

    def mean():
        calls["mean"] += 1
        return torch.ones(3, 1, dtype=torch.float64, device=dev)

    def var():
        calls["var"] += 1
        return torch.eye(3, dtype=torch.float64, device=dev)

    def mv():
        calls["mv"] += 1
        return mean(), var()

    def mvd():
        calls["mvd"] += 1
        return torch.ones(3, 1, dtype=torch.float64, device=dev), torch.ones(3, dtype=torch.float64, device=dev)

    d = S.Normal(mean, var, mean_var=mv, mean_var_diag=mvd)
    m, v = d.mean_var
    assert calls["mv"] == 1
    d2 = S.Normal(mean, var, mean_var=mv, mean_var_diag=mvd)
    mm, vv = d2.marginals()
    assert calls["mvd"] == 1 and mm.shape == (3,) and vv.shape == (3,)
    scaled = S.Normal(np.ones((2, 1)), np.eye(2)) * 3.0
    approx(S.B.dense(scaled.var), 9 * np.eye(2))
    a = np.array([[1.0, 2.0]])
    lm = S.Normal(np.ones((2, 1)), np.eye(2)).lmatmul(a)
    approx(S.B.dense(lm.var), a @ a.T)
    approx(lm.mean, a @ np.ones((2, 1)))


def test_linear_kernel_stays_low_rank(S):
    # Linear() + diagonal noise -> Woodbury: logpdf / posterior in O(n d^2), same numbers as dense
    rng = np.random.default_rng(40)
    n, d = 200, 3
    x = rng.standard_normal((n, d))
    y = rng.standard_normal(n)
    f = S.GP(2.0 * S.Linear())
    fd = f(x, 0.3)
    assert isinstance(fd.var, S.matrix.Woodbury) and isinstance(f(x).var, S.matrix.LowRank)
    spec = ("scaled", 2.0, ("linear",))
    approx(fd.logpdf(y), O.fdd_logpdf(spec, x, 0.3, y), rtol=1e-9)
    approx(S.B.dense(fd.var), O.kernel_matrix(spec, x) + 0.3 * np.eye(n), rtol=1e-12, atol=1e-12)
    approx(fd.var_diag, np.diag(O.kernel_matrix(spec, x)) + 0.3, rtol=1e-12)
    noise_vec = rng.uniform(0.2, 0.5, n)
    approx(f(x, noise_vec).logpdf(y), O.fdd_logpdf(spec, x, noise_vec, y), rtol=1e-9)
    xs = rng.standard_normal((7, d))
    post = f | (f(x, 0.3), y)
    mref, vref = O.posterior(spec, x, 0.3, y, xs)
    approx(post(xs).mean, mref, rtol=1e-7, atol=1e-8)
    approx(S.B.dense(post(xs).var), vref, rtol=1e-6, atol=1e-8)
    # a sum with a dense kernel falls back to the dense path and still agrees
    g = S.GP(S.Linear() + S.EQ())
    approx(g(x, 0.3).logpdf(y), O.fdd_logpdf(("sum", ("linear",), ("eq",)), x, 0.3, y), rtol=1e-9)


def test_rq_kernel(S):
    # mlkernels RQ(alpha) (README.md:1076-1088): matrix, elwise, logpdf and posterior against the oracle, alone and composed
    rng = np.random.default_rng(77)
    x, xs = rng.standard_normal((30, 2)), rng.standard_normal((6, 2))
    y = rng.standard_normal(30)
    k = 1.5 * S.RQ(0.7).stretch(1.3) + S.EQ() * S.RQ(2.0)
    spec = ("sum", ("scaled", 1.5, ("stretched", 1.3, ("rq", 0.7))), ("product", ("eq",), ("rq", 2.0)))
    approx(S.B.dense(k(x, xs)), O.kernel_matrix(spec, x, xs), rtol=1e-11, atol=1e-13)
    approx(k.elwise(x), O.kernel_elwise(spec, x), rtol=1e-11, atol=1e-13)
    f = S.GP(k)
    approx(f(x, 0.1).logpdf(y), O.fdd_logpdf(spec, x, 0.1, y), rtol=1e-10, atol=0)
    approx((f | (f(x, 0.1), y))(xs).mean, O.posterior(spec, x, 0.1, y, xs)[0], rtol=1e-8, atol=1e-9)
    assert str(S.RQ(0.5)) == "RQ(0.5)"


def test_block_joint_with_unequal_blocks_and_mixed_noise(S):
    # B.block assembly kept symbolic (matrix.BlockDense): unequal block sizes, scalar / vector / no noise per block, a
    # non-symbolic (conditioned) block next to symbolic ones -- logpdf, dense matrix and a sample against the oracle
    rng = np.random.default_rng(91)
    x1, x2, x3 = rng.standard_normal((37, 2)), rng.standard_normal((140, 2)), rng.standard_normal((5, 2))
    m = S.Measure()
    f = S.GP(S.EQ().stretch(1.2), measure=m)
    g = S.GP(0.7 * S.Matern52(), measure=m)
    h = f + 2.0 * g
    nv = rng.uniform(0.1, 0.3, 140)
    fdds = (f(x1, 0.2), h(x2, nv), g(x3))
    y = [rng.standard_normal(37), rng.standard_normal(140), rng.standard_normal(5)]
    kf, kg = ("stretched", 1.2, ("eq",)), ("scaled", 0.7, ("matern52",))
    kh = ("sum", kf, ("scaled", 4.0, kg))
    khg = ("scaled", 2.0, kg)
    zero = ("zero",)
    specs = [[kf, kf, zero], [kf, kh, khg], [zero, khg, kg]]
    K = O.mo_block_kernel(specs, [x1, x2, x3])
    K[:37, :37] += 0.2 * np.eye(37)
    K[37:177, 37:177] += np.diag(nv)
    want = float(O.normal_logpdf(None, K, np.concatenate(y)))
    got = m.logpdf(*[(fd, yi) for fd, yi in zip(fdds, y)])
    approx(got, want, rtol=1e-10, atol=0)
    from stheno_b200.model.observations import combine

    joint = combine(*fdds)
    approx(S.B.dense(joint.var), K, rtol=1e-10, atol=1e-12)
