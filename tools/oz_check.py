"""Dev check of the int8-slice (Ozaki) fp64 emulation: accuracy vs torch fp64 matmul, throughput vs the DMMA kernel, and
the emulated Cholesky on the headline problem."""
import sys, time
import numpy as np
import torch
from stheno_b200 import ops, B as Bns
import stheno_b200.torch as S

dev = "cuda"
torch.manual_seed(0)


def ev(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


what = sys.argv[1] if len(sys.argv) > 1 else "all"
if what in ("all", "acc"):
    for (M, N, K) in [(128, 64, 128), (256, 192, 512), (1024, 1024, 512), (512, 256, 2048)]:
        A = torch.randn(M, K, device=dev, dtype=torch.float64) * torch.exp(2 * torch.randn(M, 1, device=dev, dtype=torch.float64))
        Bm = torch.randn(N, K, device=dev, dtype=torch.float64)
        C0 = torch.randn(M, N, device=dev, dtype=torch.float64)
        ref = 0.5 * C0 - 1.5 * A @ Bm.T
        scale = (A.abs().amax(1, keepdim=True) * Bm.abs().amax(1)[None, :]) * np.sqrt(K)
        for sl in (5, 6, 7, 8):
            C = ops.gemm_nt_oz(A, Bm, C0.clone(), alpha=-1.5, beta=0.5, slices=sl)
            err = ((C - ref).abs() / scale).max().item()
            rel = ((C - ref).norm() / ref.norm()).item()
            print(f"M{M} N{N} K{K} S={sl}: max|err|/(rowmax colmax sqrtK)={err:.3e}  fro-rel={rel:.3e}", flush=True)
        if N % 128 == 0:
            Cd = ops.gemm_nt(A[None], Bm[None])[0]
            print("   DMMA kernel fro-rel vs torch:", ((Cd - A @ Bm.T).norm() / (A @ Bm.T).norm()).item())

if what in ("all", "perf"):
    n, K = 16384, 512
    P = torch.randn(n, K, device=dev, dtype=torch.float64)
    C = torch.zeros(n, n, device=dev, dtype=torch.float64)
    fl = n * n * K  # lower triangle: n^2/2 * K * 2
    t = ev(lambda: ops.gemm_nt(P[None], P[None], C[None], alpha=-1.0, beta=1.0, lower=True))
    print(f"DMMA syrk n={n} K={K}: {t:.3f} ms  {fl / t / 1e9:.1f} TF/s")
    for sl in (6, 7):
        t = ev(lambda: ops.gemm_nt_oz(P, P, C, alpha=-1.0, beta=1.0, lower=True, slices=sl))
        print(f"int8x{sl} syrk (incl. slicing): {t:.3f} ms  {fl / t / 1e9:.1f} TF/s-equivalent")
    for n2 in (4096, 8192):
        P2 = P[:n2]; C2 = C[:n2, :n2]
        t0 = ev(lambda: ops.gemm_nt(P2[None], P2[None], C2[None], alpha=-1.0, beta=1.0, lower=True))
        t1 = ev(lambda: ops.gemm_nt_oz(P2, P2, C2, alpha=-1.0, beta=1.0, lower=True, slices=6))
        print(f"n={n2}: DMMA {t0:.3f} ms, int8x6 {t1:.3f} ms")

if what in ("all", "chol"):
    n, d = 16384, 8
    g = torch.Generator().manual_seed(1)
    x = torch.rand(n, d, generator=g, dtype=torch.float64).to(dev)
    y = torch.randn(n, generator=g, dtype=torch.float64).to(dev)
    f = S.GP(S.EQ().stretch(1.5))
    res = {}
    for prec in ("fp64", "int8x6", "int8x7", "tf32x3"):
        Bns.precision = prec
        lp = float(f(x, 0.1).logpdf(y))
        t = ev(lambda: f(x, 0.1).logpdf(y), reps=3)
        res[prec] = lp
        print(f"{prec}: logpdf={lp:.12f} rel diff vs fp64={abs(lp - res['fp64']) / abs(res['fp64']):.3e}  {t:.2f} ms", flush=True)
    Bns.precision = "fp64"
