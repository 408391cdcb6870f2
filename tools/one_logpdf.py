"""One (or a few) fused logpdf evaluations at size n (dev tool for ncu launch lists)."""
import sys
import torch
sys.path.insert(0, ".")
from stheno_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
flat = ops.FlatKernel([(1.0, [("eq", 0)])], 1)
g = torch.Generator(device="cuda").manual_seed(n)
x = torch.randn(1, 1, n, 8, device="cuda", dtype=torch.float64, generator=g) / 2.0
y = torch.randn(1, 1, n, device="cuda", dtype=torch.float64, generator=g)
for _ in range(reps):
    lp = ops.chol_from_kernel(flat, x, noise_scalar=0.1, jitter=1e-12, rhs_t=y).logpdf()
torch.cuda.synchronize()
print(lp.item())
