"""Generate the golden fixtures for the BASELINE.json configs from the CPU oracle (run in the build container).

The reference itself cannot be imported (un-vendored deps, no network), so the fixtures come from the oracle that is
pinned on the reference's README literals (tests/test_oracle_golden.py).  Inputs are seeded as in SURVEY.md 8d; sizes of
configs 2-5 are reduced so that the fixtures stay small (full sizes are covered by property tests).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import gp_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def config1():
    rng = np.random.default_rng(1)
    x = np.linspace(0, 10, 1000)
    K = O.kernel_matrix(("eq",), x) + 0.1 * np.eye(1000)
    y = np.linalg.cholesky(K + 1e-12 * np.eye(1000)) @ rng.standard_normal(1000)
    xs = np.linspace(0, 10, 200) + 0.005
    lp = O.fdd_logpdf(("eq",), x, 0.1, y)
    mean, var = O.posterior(("eq",), x, 0.1, y, xs)
    np.savez_compressed(os.path.join(HERE, "config1.npz"), x=x, y=y, xs=xs, logpdf=lp, mean=mean, var_diag=np.diag(var),
                        var_block=var[:20, :20])


def config2_small():
    rng = np.random.default_rng(2)
    n, d = 1536, 8
    x = rng.standard_normal((n, d))
    y = rng.standard_normal(n)
    xs = rng.standard_normal((64, d))
    spec = ("sum", ("stretched", 2.0, ("eq",)), ("scaled", 0.1, ("delta",)))
    lp = O.fdd_logpdf(spec, x, None, y)
    mean, var = O.posterior(("stretched", 2.0, ("eq",)), x, 0.1, y, xs)
    np.savez_compressed(os.path.join(HERE, "config2_n1536.npz"), x=x, y=y, xs=xs, logpdf=lp, mean=mean, var_diag=np.diag(var))


def config3_small():
    rng = np.random.default_rng(3)
    B, n, d = 6, 256, 8
    x = rng.standard_normal((B, n, d)).astype(np.float32)
    y = rng.standard_normal((B, n, 1)).astype(np.float32)
    lp = np.array([O.fdd_logpdf(("eq",), x[b].astype(np.float64), 0.1, y[b].astype(np.float64), eps=1e-6) for b in range(B)])
    np.savez_compressed(os.path.join(HERE, "config3_B6_n256.npz"), x=x, y=y, logpdf=lp)


def config4_small():
    rng = np.random.default_rng(4)
    n, m, d = 2048, 96, 8
    x = rng.standard_normal((n, d))
    z = rng.standard_normal((m, d))
    y = rng.standard_normal(n)
    spec = ("stretched", 2.0, ("matern52",))
    out = {}
    for method in ("vfe", "fitc", "dtc"):
        c = O.sparse_compute(spec, z, x, 0.1, y, method)
        out[f"elbo_{method}"] = c["elbo"]
        out[f"mu_{method}"] = c["mu"]
    np.savez_compressed(os.path.join(HERE, "config4_n2048_m96.npz"), x=x, z=z, y=y, **out)


def config5_small():
    rng = np.random.default_rng(5)
    p, m, n = 4, 2, 96
    x = np.linspace(0, 10, n)
    H = rng.standard_normal((p, m))
    ells = np.array([1.0, 2.5])
    Ks = [O.kernel_matrix(("stretched", ells[j], ("eq",)), x) for j in range(m)]
    K = np.block([[sum(H[i, j] * H[k, j] * Ks[j] for j in range(m)) for k in range(p)] for i in range(p)])
    y = rng.standard_normal(p * n)
    lp = O.normal_logpdf(None, K + 0.5 * np.eye(p * n), y)
    np.savez_compressed(os.path.join(HERE, "config5_p4_n96.npz"), x=x, H=H, ells=ells, y=y, logpdf=lp)


if __name__ == "__main__":
    config1()
    config2_small()
    config3_small()
    config4_small()
    config5_small()
    print(sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))
