#!/bin/bash
set -u
for cfg in "GPK_TC_STAGES1=2" "GPK_TC_STAGES1=3"; do
env $cfg timeout 150 python - "$cfg" <<'P'
import sys, torch
sys.path.insert(0, ".")
import stheno_b200 as S
S.B.epsilon = 1e-6
g = torch.Generator(device="cuda").manual_seed(3)
for Bn in (64, 512):
    x = torch.randn(Bn, 2048, 8, device="cuda", generator=g); y = torch.randn(Bn, 2048, 1, device="cuda", generator=g)
    for _ in range(3): lp = S.GP(S.EQ())(x, 0.1).logpdf(y)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): lp = S.GP(S.EQ())(x, 0.1).logpdf(y)
    e1.record(); torch.cuda.synchronize()
    print(sys.argv[1], "B", Bn, "ms", round(e0.elapsed_time(e1) / 5, 3), float(lp.sum()))
P
done
echo "== tests with the one-stage kernel for batched problems"
GPK_TC_STAGES1=3 timeout 200 python -m pytest tests/test_configs.py tests/test_gpu_primitives.py -m gpu -q -x -p no:cacheprovider -k "config3 or gemm or fp32 or batch" 2>&1 | tail -3
