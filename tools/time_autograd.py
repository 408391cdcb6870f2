import sys, time, torch
sys.path.insert(0, ".")
import stheno_b200 as S
from stheno_b200 import ops
def t(): torch.cuda.synchronize(); return time.perf_counter()
for n, d in [(700, 8), (4096, 8)]:
    x = torch.randn(n, d, device="cuda", dtype=torch.float64); y = torch.randn(n, device="cuda", dtype=torch.float64)
    for rep in range(2):
        var = torch.tensor(1.3, device="cuda", dtype=torch.float64, requires_grad=True)
        scale = torch.tensor(0.8, device="cuda", dtype=torch.float64, requires_grad=True)
        noise = torch.tensor(0.15, device="cuda", dtype=torch.float64, requires_grad=True)
        t0 = t(); lp = S.GP(var * S.EQ().stretch(scale))(x, noise).logpdf(y); t1 = t(); lp.backward(); t2 = t()
        print(n, rep, "fwd %.4f s  bwd %.4f s" % (t1 - t0, t2 - t1), var.grad.item(), flush=True)
