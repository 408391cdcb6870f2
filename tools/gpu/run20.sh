#!/bin/bash
set -u
cat > /tmp/t16k.py <<'P'
import os, sys, torch
sys.path.insert(0, ".")
import stheno_b200 as S
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(16384, 8, device="cuda", dtype=torch.float64, generator=g); y = torch.randn(16384, device="cuda", dtype=torch.float64, generator=g)
k = S.EQ().stretch(2.0) + 0.1 * S.Delta()
for prec in ("auto", "int8x8"):
    S.B.precision = prec
    for _ in range(3): lp = S.GP(k)(x).logpdf(y)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): lp = S.GP(k)(x).logpdf(y)
    e1.record(); torch.cuda.synchronize()
    print(prec, "logpdf ms", round(e0.elapsed_time(e1) / 20, 3), repr(float(lp)))
P
timeout 150 python /tmp/t16k.py
echo "== tests"
timeout 400 python -m pytest tests/test_emulation.py tests/test_configs.py tests/test_gpu_primitives.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
timeout 300 python -m pytest tests/test_full_size_parity.py -m gpu -q -x -p no:cacheprovider -k "c2" 2>&1 | tail -2
