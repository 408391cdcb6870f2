import sys, torch
from stheno_b200 import ops
n, K = int(sys.argv[1]) if len(sys.argv) > 1 else 8192, 512
sl = int(sys.argv[2]) if len(sys.argv) > 2 else 6
P = torch.randn(n, K, device="cuda", dtype=torch.float64)
C = torch.zeros(n, n, device="cuda", dtype=torch.float64)
for _ in range(2):
    ops.gemm_nt_oz(P, P, C, alpha=-1.0, beta=1.0, lower=True, slices=sl)
torch.cuda.synchronize()
