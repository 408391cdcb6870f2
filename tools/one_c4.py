"""One streamed sparse ELBO (BASELINE configs[3]: n = 262144, m = 4096, Matern52) or one posterior-marginals call -- launch lists."""
import sys
import torch
sys.path.insert(0, ".")
import stheno_b200 as S
what = sys.argv[1] if len(sys.argv) > 1 else "c4"
g = torch.Generator(device="cuda").manual_seed(4)
if what == "c4":
    n, m, d = 262144, 4096, 8
    x = torch.randn(n, d, device="cuda", dtype=torch.float64, generator=g); y = torch.randn(n, device="cuda", dtype=torch.float64, generator=g)
    z = torch.randn(m, d, device="cuda", dtype=torch.float64, generator=g)
    f = S.GP(S.Matern52().stretch(2.0))
    print(float(S.PseudoObs(f(z), f(x, 0.1), y).elbo(f.measure)))
else:
    n, m, d = 16384, 4096, 8
    x = torch.randn(n, d, device="cuda", dtype=torch.float64, generator=g); y = torch.randn(n, device="cuda", dtype=torch.float64, generator=g)
    xs = torch.randn(m, d, device="cuda", dtype=torch.float64, generator=g)
    f = S.GP(S.EQ().stretch(2.0))
    post = f | (f(x, 0.1), y)
    mean, var = post(xs).marginals()
    print(float(mean.sum()), float(var.sum()))
torch.cuda.synchronize()
