import sys, torch, numpy as np
sys.path.insert(0, ".")
import stheno_b200 as S
def ev(f, reps=3):
    f(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); out = f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts), out
for n in (4096, 16384):
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(n, 8, device="cuda", dtype=torch.float64, generator=g); y = torch.randn(n, device="cuda", dtype=torch.float64, generator=g)
    f = S.GP(S.EQ().stretch(2.0))
    S.B.precision = "fp64"; t0, r0 = ev(lambda: f(x, 0.1).logpdf(y))
    S.B.precision = "tf32x3"; t1, r1 = ev(lambda: f(x, 0.1).logpdf(y))
    S.B.precision = "fp64"
    print(f"n={n}: fp64 {t0:.2f} ms  tf32x3 {t1:.2f} ms  speedup {t0/t1:.2f}x  rel diff {abs((r1-r0).item()/r0.item()):.3e}")
