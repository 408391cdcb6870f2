"""Run the BASELINE.json configs at (near) full size on one GPU and print timings (dev / evidence tool)."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import stheno_b200 as S
from stheno_b200 import ops

def ev(f, reps=2, warm=1):
    for _ in range(warm): out = f()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); out = f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts), out
res = {}
which = sys.argv[1:] or ["c1", "c2", "c3", "c4", "c5"]
dev = "cuda"
if "c1" in which:
    S.B.epsilon = 1e-12
    x = torch.linspace(0, 10, 1000, dtype=torch.float64, device=dev); y = torch.sin(x)
    xs = x + 0.005
    f = S.GP(S.EQ())
    t, lp = ev(lambda: f(x, 0.1).logpdf(y), reps=5)
    res["c1_logpdf_ms"] = t
    def post():
        p = f | (f(x, 0.1), y); return p(xs).marginals()
    res["c1_posterior_marginals_ms"], _ = ev(post, reps=5)
if "c2" in which:
    S.B.epsilon = 1e-12
    g = torch.Generator(device=dev).manual_seed(2)
    n, d, m = 16384, 8, 4096
    x = torch.randn(n, d, device=dev, dtype=torch.float64, generator=g); y = torch.randn(n, device=dev, dtype=torch.float64, generator=g)
    xs = torch.randn(m, d, device=dev, dtype=torch.float64, generator=g)
    f = S.GP(S.EQ().stretch(2.0))
    t, lp = ev(lambda: f(x, 0.1).logpdf(y), reps=3)
    res["c2_logpdf_ms"] = t; res["c2_logpdf_per_s"] = 1e3 / t
    obs = S.Obs(f(x, 0.1), y)
    post = f | obs
    post(xs[:8]).marginals(); torch.cuda.synchronize()   # factorisation cached on obs
    t, _ = ev(lambda: post(xs).marginals(), reps=3)
    res["c2_posterior_marginals_m4096_ms"] = t
    res["c2_posterior_solve_gflops"] = (n * n * m + 4 * n * m) / t / 1e6
    t, _ = ev(lambda: post(xs).mean_var, reps=2)
    res["c2_posterior_full_cov_m4096_ms"] = t
    res["c2_posterior_full_gflops"] = (n * n * m + n * m * m + 2 * n * m) / t / 1e6
    del obs, post
if "c3" in which:
    S.B.epsilon = 1e-6
    g = torch.Generator(device=dev).manual_seed(3)
    for B in (64, 512):
        x = torch.randn(B, 2048, 8, device=dev, generator=g); y = torch.randn(B, 2048, 1, device=dev, generator=g)
        f = S.GP(S.EQ())
        t, lp = ev(lambda: f(x, 0.1).logpdf(y), reps=2)
        res[f"c3_B{B}_n2048_f32_ms"] = t; res[f"c3_B{B}_tflops"] = B * 2048**3 / 3 / t / 1e9
        del x, y
if "c4" in which:
    S.B.epsilon = 1e-12
    g = torch.Generator(device=dev).manual_seed(4)
    n, m, d = 262144, 4096, 8
    x = torch.randn(n, d, device=dev, dtype=torch.float64, generator=g); y = torch.randn(n, device=dev, dtype=torch.float64, generator=g)
    z = torch.randn(m, d, device=dev, dtype=torch.float64, generator=g)
    f = S.GP(S.Matern52().stretch(2.0))
    def elbo():
        return S.PseudoObs(f(z), f(x, 0.1), y).elbo(f.measure)
    torch.cuda.reset_peak_memory_stats()
    t, e = ev(elbo, reps=2)
    res["c4_peak_mem_gb"] = torch.cuda.max_memory_allocated() / 1e9
    res["c4_elbo_ms"] = t; res["c4_elbo"] = float(e); res["c4_tflops"] = (2.0 * m * m * n + 2 * m**3 / 3) / t / 1e9
    del x, y, z
    torch.cuda.empty_cache()
if "c5" in which:
    S.B.epsilon = 1e-12
    rng = np.random.default_rng(5)
    p, ml, n = 4, 2, 8192
    x = torch.linspace(0, 10, n, dtype=torch.float64, device=dev)
    H = torch.tensor(rng.standard_normal((p, ml)), device=dev, requires_grad=True)
    ells = torch.tensor([1.0, 2.5], dtype=torch.float64, device=dev, requires_grad=True)
    noise = torch.tensor(0.5, dtype=torch.float64, device=dev, requires_grad=True)
    y = torch.tensor(rng.standard_normal(p * n), device=dev)
    # ILMM as ONE flattened kernel over the stacked inputs is not expressible; evaluate it the way the reference does:
    # the joint of the p outputs through the measure (block assembly), forward only here.
    def fwd():
        with torch.no_grad():
            m = S.Measure()
            us = [S.GP(S.EQ().stretch(float(ells[j])), measure=m) for j in range(ml)]
            fs = [sum(float(H[i, j]) * us[j] for j in range(ml)) for i in range(p)]
            return m.logpdf(*[(fs[i](x, 0.5), y[i * n:(i + 1) * n]) for i in range(p)])
    t, lp = ev(fwd, reps=1)
    res["c5_forward_N32768_ms"] = t; res["c5_logpdf"] = float(lp)
    # the real thing: loss + gradients w.r.t. H, length scales and noise through the 4-output joint (N = 32768)
    def loss_grad():
        for v in (H, ells, noise): v.grad = None
        m = S.Measure()
        us = [S.GP(S.EQ().stretch(ells[j]), measure=m) for j in range(ml)]
        fs = [sum(H[i, j] * us[j] for j in range(ml)) for i in range(p)]
        l = -m.logpdf(*[(fs[i](x, noise), y[i * n:(i + 1) * n]) for i in range(p)])
        l.backward(); return l
    t, l = ev(loss_grad, reps=1)
    res["c5_loss_and_grad_N32768_ms"] = t; res["c5_loss"] = float(l)
    res["c5_grad_H_norm"] = float(H.grad.norm()); res["c5_peak_mem_gb"] = torch.cuda.max_memory_allocated() / 1e9
print(json.dumps(res, indent=1))
import os; os.makedirs("gpurun_out", exist_ok=True); json.dump(res, open("gpurun_out/r02_configs.json", "w"), indent=1)
