"""How does the emulated trailing update behave on numerically singular covariances (noise-free EQ, jitter 1e-12)?"""
import numpy as np, torch
import stheno_b200.torch as S
from stheno_b200 import ops
torch.manual_seed(0)
for n, d, ell, eps in [(4096, 1, 1.0, 1e-12), (4096, 1, 0.3, 1e-12), (4096, 2, 1.0, 1e-12), (3000, 1, 1.0, 1e-10), (8192, 1, 0.05, 1e-12), (4096, 1, 1.0, 1e-8)]:
    S.B.epsilon = eps
    x = torch.rand(n, d, dtype=torch.float64, device="cuda") * 10
    y = torch.randn(n, dtype=torch.float64, device="cuda")
    f = S.GP(S.EQ().stretch(ell))
    out = {}
    for prec in ("fp64", "int8x7", "int8x8", "auto"):
        S.B.precision = prec
        lp = f(x).logpdf(y)
        out[prec] = float(lp)
    S.B.precision = "auto"
    r = lambda a, b: abs(a - b) / abs(b) if np.isfinite(a) and np.isfinite(b) else float("nan")
    print(f"n={n} d={d} ell={ell} eps={eps:g}: fp64={out['fp64']:.6e} int8x7={out['int8x7']:.6e} int8x8={out['int8x8']:.6e} "
          f"auto={out['auto']:.6e} rel(x7)={r(out['int8x7'], out['fp64']):.2e} rel(x8)={r(out['int8x8'], out['fp64']):.2e}", flush=True)
S.B.epsilon = 1e-12
