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
for cfg in "GPK_OZ_BAND=16 GPK_OZ_TPC_CAP=2" "GPK_OZ_BAND=8 GPK_OZ_TPC_CAP=2" "GPK_OZ_BAND=32 GPK_OZ_TPC_CAP=2" "GPK_OZ_BAND=16 GPK_OZ_TPC_CAP=4" "GPK_OZ_BAND=24 GPK_OZ_TPC_CAP=3" "GPK_OZ_BAND=12 GPK_OZ_TPC_CAP=1"; do
echo "== $cfg"; env $cfg timeout 120 python /tmp/t16k.py
done
