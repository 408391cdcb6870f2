"""One batched fp32 logpdf (BASELINE configs[2] share: B x 2048, d = 8) -- dev tool for ncu launch lists."""
import sys
import torch
sys.path.insert(0, ".")
import stheno_b200 as S
Bn = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
S.B.epsilon = 1e-6
g = torch.Generator(device="cuda").manual_seed(3)
x = torch.randn(Bn, 2048, 8, device="cuda", generator=g)
y = torch.randn(Bn, 2048, 1, device="cuda", generator=g)
for _ in range(reps):
    lp = S.GP(S.EQ())(x, 0.1).logpdf(y)
torch.cuda.synchronize()
print(float(lp.sum()))
