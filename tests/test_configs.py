"""BASELINE.json configs: the committed golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the
oracle) at reduced size, and size-independent properties at the full sizes."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture
def S():
    import stheno_b200 as s

    s.B.epsilon = 1e-12
    s.Measure.default = None
    return s


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


def test_config1_golden(S):
    g = np.load(os.path.join(GOLD, "config1.npz"))
    f = S.GP(S.EQ())
    assert rel(f(g["x"], 0.1).logpdf(g["y"]), g["logpdf"]) < 1e-10
    post = f | (f(g["x"], 0.1), g["y"])
    pred = post(g["xs"])
    mean, var = pred.marginals()
    np.testing.assert_allclose(mean, g["mean"][:, 0], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(var, g["var_diag"], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(S.B.dense(pred.var)[:20, :20], g["var_block"], rtol=1e-6, atol=1e-9)


def test_config2_golden_reduced(S):
    g = np.load(os.path.join(GOLD, "config2_n1536.npz"))
    f = S.GP(S.EQ().stretch(2.0) + 0.1 * S.Delta())
    assert rel(f(g["x"]).logpdf(g["y"]), g["logpdf"]) < 1e-10
    f2 = S.GP(S.EQ().stretch(2.0))
    assert rel(f2(g["x"], 0.1).logpdf(g["y"]), g["logpdf"]) < 1e-10  # noise as process == noise as argument
    post = f2 | (f2(g["x"], 0.1), g["y"])
    mean, var = post(g["xs"]).marginals()
    np.testing.assert_allclose(mean, g["mean"][:, 0], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(var, g["var_diag"], rtol=1e-7, atol=1e-9)


def test_config2_full_size_properties(S):
    """n = 16384, d = 8, fp64: (i) against torch-fp64 potrf/trsm on the same matrix, (ii) the chain rule
    logpdf(y) = logpdf(y_a) + logpdf(y_b | y_a) (tests/model/test_model.py:375-404 at full size)."""
    rng = np.random.default_rng(2)
    n, d = 16384, 8
    x = torch.as_tensor(rng.standard_normal((n, d)), device="cuda")
    y = torch.as_tensor(rng.standard_normal(n), device="cuda")
    k = S.EQ().stretch(2.0)
    f = S.GP(k)
    fd = f(x, 0.1)
    lp = fd.logpdf(y)
    K = S.B.dense(f(x, 0.1 + 1e-12).var)
    L = torch.linalg.cholesky(K)
    a = torch.linalg.solve_triangular(L, y[:, None], upper=False)
    ref = -0.5 * (2 * torch.log(L.diagonal()).sum() + n * np.log(2 * np.pi) + (a * a).sum())
    assert abs((lp - ref).item() / ref.item()) < 1e-11
    del K, L, a
    h = n // 2
    post = f | (f(x[:h], 0.1), y[:h])
    chain = f(x[:h], 0.1).logpdf(y[:h]) + post(x[h:], 0.1).logpdf(y[h:])
    assert abs((chain - lp).item() / lp.item()) < 1e-9


def test_config3_golden_reduced_fp32(S):
    g = np.load(os.path.join(GOLD, "config3_B6_n256.npz"))
    S.B.epsilon = 1e-6
    f = S.GP(S.EQ())
    lp = f(g["x"], 0.1).logpdf(g["y"])
    assert lp.shape == (6,) and lp.dtype == np.float32
    assert rel(lp, g["logpdf"]) < 1e-4
    from stheno_b200.dist import sharded_logpdf

    tot = sharded_logpdf(lambda xl: S.GP(S.EQ())(xl, 0.1), g["x"], g["y"], reduce="sum")
    assert abs(float(tot) - g["logpdf"].sum()) < 1e-4 * abs(g["logpdf"].sum())


def test_config3_full_size_batch_properties(S):
    """B = 64 (one GPU's share of the 512), n = 2048, fp32: per-problem values equal single-problem evaluations."""
    S.B.epsilon = 1e-6
    gen = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(64, 2048, 8, device="cuda", generator=gen)
    y = torch.randn(64, 2048, 1, device="cuda", generator=gen)
    f = S.GP(S.EQ())
    lp = f(x, 0.1).logpdf(y)
    assert lp.shape == (64,)
    for b in (0, 31, 63):
        single = f(x[b], 0.1).logpdf(y[b])
        assert abs((lp[b] - single).item()) < 1e-4 * abs(single.item())
        ref = S.GP(S.EQ())(x[b].double(), 0.1).logpdf(y[b].double())
        assert abs(lp[b].item() - ref.item()) < 1e-4 * abs(ref.item())


def test_config4_golden_reduced(S):
    g = np.load(os.path.join(GOLD, "config4_n2048_m96.npz"))
    f = S.GP(S.Matern52().stretch(2.0))
    for method, cls in (("vfe", S.PseudoObs), ("fitc", S.PseudoObsFITC), ("dtc", S.PseudoObsDTC)):
        obs = cls(f(g["z"]), f(g["x"], 0.1), g["y"])
        assert rel(obs.elbo(f.measure), g[f"elbo_{method}"]) < 1e-10
        np.testing.assert_allclose(S.B.to_numpy(obs.mu(f.measure)), g[f"mu_{method}"], rtol=1e-6, atol=1e-8)


def test_config4_larger_sparse_bound(S):
    """n = 32768, m = 1024: the VFE ELBO is a lower bound that tightens when z grows (size-independent property)."""
    rng = np.random.default_rng(4)
    n, d = 32768, 8
    x = torch.as_tensor(rng.standard_normal((n, d)), device="cuda")
    y = torch.as_tensor(rng.standard_normal(n), device="cuda")
    z = torch.as_tensor(rng.standard_normal((1024, d)), device="cuda")
    f = S.GP(S.Matern52().stretch(2.0))
    e_small = S.PseudoObs(f(z[:256]), f(x, 0.1), y).elbo(f.measure)
    e_big = S.PseudoObs(f(z), f(x, 0.1), y).elbo(f.measure)
    assert torch.isfinite(e_small) and torch.isfinite(e_big)
    assert e_big >= e_small - 1e-6 * abs(e_small)


def test_config5_golden_reduced(S):
    g = np.load(os.path.join(GOLD, "config5_p4_n96.npz"))
    x, H, ells, y = g["x"], g["H"], g["ells"], g["y"]
    m = S.Measure()
    us = [S.GP(S.EQ().stretch(float(l)), measure=m) for l in ells]
    fs = []
    for i in range(4):
        fi = float(H[i, 0]) * us[0] + float(H[i, 1]) * us[1]
        fs.append(fi)
    n = len(x)
    pairs = [(fs[i](x, 0.5), y[i * n:(i + 1) * n]) for i in range(4)]
    assert rel(m.logpdf(*pairs), g["logpdf"]) < 1e-10


def test_opt_in_tf32x3_trailing_update(S):
    """BASELINE north_star: "tf32/bf16 where the user opts in".  B.precision = "tf32x3" moves the K >= 128 trailing updates of
    the fp64 Cholesky to the tcgen05 tensor cores (fp32 panel copy, 3xTF32 products).  It is an accuracy/speed trade:
    fp32-level agreement, not the 1e-10 parity bar -- hence opt-in."""
    rng = np.random.default_rng(7)
    n, d = 3000, 8
    x = torch.as_tensor(rng.standard_normal((n, d)), device="cuda")
    y = torch.as_tensor(rng.standard_normal(n), device="cuda")
    f = S.GP(S.EQ().stretch(2.0))
    before = S.B.precision
    try:
        S.B.precision = "fp64"
        ref = f(x, 0.1).logpdf(y)
        S.B.precision = "tf32x3"
        fast = f(x, 0.1).logpdf(y)
        S.B.precision = "fp64"
        again = f(x, 0.1).logpdf(y)
    finally:
        S.B.precision = before
    rel = abs((fast - ref).item() / ref.item())
    assert rel < 1e-4, rel
    assert rel > 0  # it really took the other path
    assert again.item() == ref.item()
