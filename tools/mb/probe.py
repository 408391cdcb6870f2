"""Hardware probe (not product code): fp64 library baselines on the B200 box."""
import json, os, time, subprocess, sys
import torch
out = {}
dev = torch.device("cuda:0")
out["gpu"] = torch.cuda.get_device_name(0)
out["cpu_count"] = os.cpu_count()
try:
    out["cpu_model"] = [l for l in open("/proc/cpuinfo") if "model name" in l][0].split(":")[1].strip()
except Exception as e:
    out["cpu_model"] = str(e)
def ev(f, reps=3, warm=1):
    for _ in range(warm): f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sorted(ts)[len(ts)//2]
# fp64 GEMM
for n in (4096, 8192):
    a = torch.randn(n, n, device=dev, dtype=torch.float64); b = torch.randn(n, n, device=dev, dtype=torch.float64)
    mn, md = ev(lambda: a @ b, reps=5)
    out[f"dgemm_{n}_tflops_best"] = 2 * n**3 / mn / 1e9
    out[f"dgemm_{n}_tflops_median"] = 2 * n**3 / md / 1e9
    del a, b
# syrk-like K=1024: C (16384x16384) -= A A^T
n = 16384
for K in (128, 256, 512, 1024):
    A = torch.randn(n, K, device=dev, dtype=torch.float64); C = torch.randn(n, n, device=dev, dtype=torch.float64)
    mn, md = ev(lambda: torch.addmm(C, A, A.T, beta=1.0, alpha=-1.0, out=C), reps=3)
    out[f"dgemm_rankK{K}_n16384_tflops"] = 2 * n * n * K / mn / 1e9
    del A, C
# sustained dgemm 3 s
a = torch.randn(8192, 8192, device=dev, dtype=torch.float64); b = torch.randn(8192, 8192, device=dev, dtype=torch.float64)
torch.cuda.synchronize(); t0 = time.time(); cnt = 0
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
while time.time() - t0 < 3.0:
    for _ in range(5): a @ b
    cnt += 5; torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
out["dgemm_8192_tflops_sustained"] = 2 * 8192**3 * cnt / e0.elapsed_time(e1) / 1e9
try:
    out["clocks_after_sustained"] = subprocess.check_output("nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv,noheader", shell=True).decode().strip()
except Exception as e:
    out["clocks_after_sustained"] = str(e)
del a, b
# potrf
for n in (2048, 4096, 8192, 16384):
    x = torch.randn(n, 8, device=dev, dtype=torch.float64)
    def build():
        n2 = (x * x).sum(-1)
        d2 = n2[:, None] + n2[None, :] - 2 * x @ x.T
        K = torch.exp(-0.5 * d2 / 4.0)
        K.diagonal().add_(0.1 + 1e-12)
        return K
    mn, md = ev(build, reps=3)
    out[f"eager_kbuild_{n}_ms"] = mn
    K = build()
    mn, md = ev(lambda: torch.linalg.cholesky(K), reps=3)
    out[f"cusolver_potrf_{n}_ms"] = mn
    out[f"cusolver_potrf_{n}_tflops"] = n**3 / 3 / mn / 1e9
    L = torch.linalg.cholesky(K)
    y = torch.randn(n, 1, device=dev, dtype=torch.float64)
    mn, md = ev(lambda: torch.linalg.solve_triangular(L, y, upper=False), reps=3)
    out[f"trsv_{n}_ms"] = mn
    if n <= 16384:
        m = 4096
        Y = torch.randn(n, m, device=dev, dtype=torch.float64)
        mn, md = ev(lambda: torch.linalg.solve_triangular(L, Y, upper=False), reps=3)
        out[f"trsm_{n}x{m}_ms"] = mn
        out[f"trsm_{n}x{m}_tflops"] = n * n * m / mn / 1e9
        del Y
    def full():
        K = build(); L = torch.linalg.cholesky(K)
        a = torch.linalg.solve_triangular(L, y, upper=False)
        return -0.5 * (2 * torch.log(L.diagonal()).sum() + n * 1.8378770664093453 + (a * a).sum())
    mn, md = ev(full, reps=3)
    out[f"eager_logpdf_{n}_ms"] = mn
    del K, L
# batched fp32
B_, n = 64, 2048
Kb = torch.randn(B_, n, 8, device=dev)
Kb = torch.exp(-0.5 * torch.cdist(Kb, Kb) ** 2); Kb.diagonal(dim1=-2, dim2=-1).add_(0.1)
mn, md = ev(lambda: torch.linalg.cholesky(Kb), reps=3)
out["batched_potrf_f32_64x2048_ms"] = mn
out["batched_potrf_f32_64x2048_tflops"] = B_ * n**3 / 3 / mn / 1e9
del Kb
# fp32 / tf32 gemm for reference
a = torch.randn(8192, 8192, device=dev); b = torch.randn(8192, 8192, device=dev)
torch.backends.cuda.matmul.allow_tf32 = False
mn, md = ev(lambda: a @ b, reps=5); out["sgemm_8192_tflops"] = 2 * 8192**3 / mn / 1e9
torch.backends.cuda.matmul.allow_tf32 = True
mn, md = ev(lambda: a @ b, reps=5); out["tf32gemm_8192_tflops"] = 2 * 8192**3 / mn / 1e9
ai = torch.randint(-128, 127, (8192, 8192), device=dev, dtype=torch.int8); bi = torch.randint(-128, 127, (8192, 8192), device=dev, dtype=torch.int8)
try:
    mn, md = ev(lambda: torch._int_mm(ai, bi.T), reps=5); out["int8gemm_8192_tops"] = 2 * 8192**3 / mn / 1e9
except Exception as e:
    out["int8gemm_8192_tops"] = str(e)[:200]
# CPU numpy
import numpy as np
nn = 4096
xa = np.random.randn(nn, nn); t0 = time.time(); xa @ xa; out["numpy_dgemm_4096_gflops"] = 2 * nn**3 / (time.time() - t0) / 1e9
S = xa @ xa.T + nn * np.eye(nn); t0 = time.time(); np.linalg.cholesky(S); out["numpy_potrf_4096_s"] = time.time() - t0
try:
    import threadpoolctl; out["blas"] = [ (d.get("internal_api"), d.get("num_threads")) for d in threadpoolctl.threadpool_info()]
except Exception as e:
    out["blas"] = str(e)
out["dmma_bin"] = subprocess.run(["./tools/mb/dmma.bin"], capture_output=True, text=True).stdout
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
print(json.dumps(out, indent=1))
