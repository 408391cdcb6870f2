"""Streamed sparse accumulation (``gpk_sparse_accumulate``): several ragged chunks give the oracle's
ELBO / mu / A for VFE, FITC and DTC, and the posterior; batched problems keep the materialised route."""
import numpy as np
import pytest

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


@pytest.mark.parametrize("method", ["vfe", "fitc", "dtc"])
@pytest.mark.parametrize("chunk", [96, 250, 100000])
def test_streamed_chunks_match_oracle(S, monkeypatch, method, chunk):
    monkeypatch.setattr(S.B, "sparse_chunk", chunk)
    rng = np.random.default_rng(12)
    n, m, d = 700, 37, 3
    x, z, y = rng.standard_normal((n, d)), rng.standard_normal((m, d)), rng.standard_normal(n)
    noise = 0.05 + rng.uniform(0, 0.1, n)
    spec = ("sum", ("scaled", 1.2, ("stretched", 1.7, ("matern52",))), ("scaled", 0.3, ("eq",)))
    k = 1.2 * S.Matern52().stretch(1.7) + 0.3 * S.EQ()
    cls = {"vfe": S.PseudoObs, "fitc": S.PseudoObsFITC, "dtc": S.PseudoObsDTC}[method]
    f = S.GP(k)
    obs = cls(f(z), f(x, noise), y)
    want = O.sparse_compute(spec, z, x, noise, y, method)
    assert abs(float(obs.elbo(f.measure)) - want["elbo"]) < 1e-10 * abs(want["elbo"])
    np.testing.assert_allclose(S.B.to_numpy(obs.mu(f.measure)), want["mu"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(S.B.to_numpy(S.B.dense(obs.A(f.measure))), want["A"], rtol=1e-8, atol=1e-9)
    xs = rng.standard_normal((11, d))
    mean, var = O.sparse_posterior(spec, z, x, noise, y, xs, method)
    post = f | obs
    np.testing.assert_allclose(S.B.to_numpy(post(xs).mean), mean, rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(S.B.to_numpy(S.B.dense(post(xs).var)), var, rtol=1e-6, atol=1e-8)


def test_batched_sparse_keeps_the_materialised_route(S):
    rng = np.random.default_rng(13)
    Bn, n, m = 3, 90, 8
    x, z = rng.standard_normal((Bn, n, 2)), rng.standard_normal((Bn, m, 2))
    y = rng.standard_normal((Bn, n, 1))
    f = S.GP(S.EQ().stretch(1.3))
    e = S.PseudoObs(f(z), f(x, 0.2), y).elbo(f.measure)
    want = [O.sparse_compute(("stretched", 1.3, ("eq",)), z[b], x[b], 0.2, y[b], "vfe")["elbo"] for b in range(Bn)]
    np.testing.assert_allclose(S.B.to_numpy(e), want, rtol=1e-9)
