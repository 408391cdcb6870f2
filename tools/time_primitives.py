"""Timing of libgpk primitives on the GPU box (dev tool; prints JSON-ish lines)."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from stheno_b200 import ops

def ev(f, reps=3, warm=1):
    for _ in range(warm): f()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
res = {}
res["dmma_peak_tflops"] = ops.probe_dmma_tflops()
for (M, N, K) in [(16384, 16384, 1024), (16384, 16384, 128), (15360, 896, 128), (4096, 4096, 4096), (8192, 8192, 512)]:
    A = torch.randn(1, M, K, device="cuda", dtype=torch.float64); B = torch.randn(1, N, K, device="cuda", dtype=torch.float64)
    C = torch.zeros(1, M, N, device="cuda", dtype=torch.float64)
    t = ev(lambda: ops.gemm_nt(A, B, C, alpha=-1.0, beta=1.0))
    res[f"gemm_f64_{M}x{N}x{K}_tflops"] = 2.0 * M * N * K / t / 1e9
    if M == N:
        t = ev(lambda: ops.gemm_nt(A, B, C, alpha=-1.0, beta=1.0, lower=True))
        res[f"syrk_f64_{M}x{N}x{K}_tflops"] = 1.0 * M * (N + 128) * K / t / 1e9
    del A, B, C
A = torch.randn(1, 8192, 2048, device="cuda"); B = torch.randn(1, 8192, 2048, device="cuda"); C = torch.zeros(1, 8192, 8192, device="cuda")
t = ev(lambda: ops.gemm_nt(A, B, C, alpha=-1.0, beta=1.0)); res["gemm_f32_8192x8192x2048_tflops"] = 2.0 * 8192 * 8192 * 2048 / t / 1e9
del A, B, C
flat = ops.FlatKernel([(1.0, [("eq", 0)])], 1)
for n in (2048, 4096, 8192, 16384):
    g = torch.Generator(device="cuda").manual_seed(n)
    x = torch.randn(1, 1, n, 8, device="cuda", dtype=torch.float64, generator=g) / 2.0
    y = torch.randn(1, 1, n, device="cuda", dtype=torch.float64, generator=g)
    res[f"k1_full_{n}_ms"] = ev(lambda: ops.kernel_matrix(flat, x, noise_scalar=0.1))
    def full():
        ch = ops.chol_from_kernel(flat, x, noise_scalar=0.1, jitter=1e-12, rhs_t=y)
        return ch.logpdf()
    t = ev(full, reps=3)
    res[f"logpdf_{n}_ms"] = t
    res[f"logpdf_{n}_tflops"] = n**3 / 3 / t / 1e9
    lp = full()
    K = ops.kernel_matrix(flat, x, noise_scalar=0.1, jitter=1e-12)[0]
    L = torch.linalg.cholesky(K); a = torch.linalg.solve_triangular(L, y[0].T, upper=False)
    ref = -0.5 * (2 * torch.log(L.diagonal()).sum() + n * np.log(2 * np.pi) + (a * a).sum())
    res[f"logpdf_{n}_relerr_vs_torch"] = abs((lp[0, 0] - ref).item() / ref.item())
    res[f"logpdf_{n}_launches"] = ops.launch_count(reset=True)
    del K, L
# batched fp32
flat = ops.FlatKernel([(1.0, [("eq", 0)])], 1)
x = torch.randn(1, 64, 2048, 8, device="cuda") ; y = torch.randn(64, 1, 2048, device="cuda")
def fullb():
    return ops.chol_from_kernel(flat, x, noise_scalar=0.1, jitter=1e-6, rhs_t=y).logpdf()
t = ev(fullb); res["batched_f32_64x2048_ms"] = t; res["batched_f32_64x2048_tflops"] = 64 * 2048**3 / 3 / t / 1e9
print(json.dumps(res, indent=1))
import os; os.makedirs("gpurun_out", exist_ok=True); json.dump(res, open("gpurun_out/time_primitives.json", "w"), indent=1)
