#!/bin/bash
set -u
cat > /tmp/c4.py <<'P'
import sys, torch
sys.path.insert(0, ".")
import stheno_b200 as S
g = torch.Generator(device="cuda").manual_seed(4)
n, m, d = 262144, 4096, 8
x = torch.randn(n, d, device="cuda", dtype=torch.float64, generator=g); y = torch.randn(n, device="cuda", dtype=torch.float64, generator=g)
z = torch.randn(m, d, device="cuda", dtype=torch.float64, generator=g)
f = S.GP(S.Matern52().stretch(2.0))
for chunk in (8192, 16384, 32768, 65536):
    S.B.sparse_chunk = chunk
    def elbo(): return S.PseudoObs(f(z), f(x, 0.1), y).elbo(f.measure)
    e = elbo(); torch.cuda.synchronize(); torch.cuda.reset_peak_memory_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); e = elbo(); e = elbo(); e1.record(); torch.cuda.synchronize()
    print("chunk", chunk, "ms", round(e0.elapsed_time(e1) / 2, 2), "peak GB", round(torch.cuda.max_memory_allocated() / 1e9, 2), repr(float(e)))
P
timeout 300 python /tmp/c4.py
