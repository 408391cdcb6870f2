import sys, torch
sys.path.insert(0, ".")
from stheno_b200 import ops
M = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
K = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
A = torch.randn(1, M, K, device="cuda", dtype=torch.float64); C = torch.zeros(1, M, M, device="cuda", dtype=torch.float64)
for _ in range(2):
    ops.gemm_nt(A, A, C, alpha=-1.0, beta=1.0, lower=True)
torch.cuda.synchronize()
