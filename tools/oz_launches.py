import sys, torch
import stheno_b200.torch as S
from stheno_b200 import B as Bns
Bns.precision = sys.argv[1] if len(sys.argv) > 1 else "int8x7"
n, d = 16384, 8
g = torch.Generator().manual_seed(1)
x = torch.rand(n, d, generator=g, dtype=torch.float64).cuda()
y = torch.randn(n, generator=g, dtype=torch.float64).cuda()
f = S.GP(S.EQ().stretch(1.5))
for _ in range(2):
    lp = f(x, 0.1).logpdf(y)
torch.cuda.synchronize()
print(float(lp))
