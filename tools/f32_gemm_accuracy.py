import sys, torch
sys.path.insert(0, ".")
from stheno_b200 import ops
for K in (128, 512, 1024, 4096):
    g = torch.Generator(device="cuda").manual_seed(K)
    A = torch.randn(1, 1024, K, device="cuda", generator=g); B = torch.randn(1, 512, K, device="cuda", generator=g)
    ref = A.double() @ B.double().transpose(1, 2)
    out = ops.gemm_nt(A, B)
    scale = (A.double().abs() @ B.double().abs().transpose(1, 2)).max().item()
    err = (out.double() - ref)
    t32 = (A @ B.transpose(1, 2)).double() - ref
    print(f"K={K}: max|err|/scale={err.abs().max().item()/scale:.3e} mean err/scale={err.mean().item()/scale:.3e} rms/scale={err.pow(2).mean().sqrt().item()/scale:.3e} | torch fp32 matmul max/scale={t32.abs().max().item()/scale:.3e}")
