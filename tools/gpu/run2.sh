#!/bin/bash
# round-2 GPU run 2: recursive leaf + K1 fast path A/B, then the GPU suite and a bench line
set -u
mkdir -p gpurun_out
echo "== leaf A/B"; timeout 600 python tools/time_leaf.py 2>&1 | tee gpurun_out/r02_leaf_ab.jsonl | cut -c1-1800
echo "== K1 A/B"; timeout 300 python tools/time_k1.py 2>&1 | tee gpurun_out/r02_k1_ab.jsonl | cut -c1-1800
echo "== pytest gpu (all, incl. the full-size oracle tests)"
timeout 1800 python -m pytest tests -m gpu -q -s -p no:cacheprovider --durations=8 2>&1 | grep -v "^conditioning sweep" | tail -40 | tee gpurun_out/r02b_pytest_gpu.log
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r02b_bench_n1.err
python - <<'P'
import json
b=json.load(open('gpurun_out/r02b_bench_n1.json'))
print({k:b[k] for k in ('value','ms_per_step')}, b['e2e']['value'], b['native_fp64']['ms_per_step'], b['emulated_8_slices']['ms_per_step'], b['roofline']['frac'], b['posterior_solve'], b['sharded_c3'].get('ms_per_step'))
P
echo "== launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02b_launches_logpdf16384.csv python tools/one_logpdf.py 16384 1 > gpurun_out/r02b_launches.log 2>&1; echo "ncu rc=$?"
echo "== C3 launch list (64 x 2048 fp32)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02b_launches_c3_64x2048.csv python tools/one_c3.py 64 1 > gpurun_out/r02b_launches_c3.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/r02b_launches_c3_64x2048.csv | head -16
python tools/launch_summary.py gpurun_out/r02b_launches_logpdf16384.csv | head -12
echo "== NB_OUTER experiment"
for nb in 768 1024; do GPK_NB_OUTER=$nb timeout 200 python - <<'P'
import os, torch, sys
sys.path.insert(0, ".")
import stheno_b200 as S
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(16384, 8, device="cuda", dtype=torch.float64, generator=g); y = torch.randn(16384, device="cuda", dtype=torch.float64, generator=g)
k = S.EQ().stretch(2.0) + 0.1 * S.Delta()
for _ in range(3): lp = S.GP(k)(x).logpdf(y)
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): lp = S.GP(k)(x).logpdf(y)
e1.record(); torch.cuda.synchronize()
print("NB_OUTER", os.environ["GPK_NB_OUTER"], "logpdf ms", e0.elapsed_time(e1) / 10, float(lp))
P
done
echo "== C3 knobs (64 x 2048 fp32, ms per batched logpdf)"
for cfg in "GPK_NB_OUTER=512" "GPK_NB_OUTER=256" "GPK_NB_OUTER=1024" "GPK_NO_LOOKAHEAD=1" "GPK_NB_OUTER=256 GPK_NO_LOOKAHEAD=1"; do
env $cfg timeout 200 python - "$cfg" <<'P'
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
    print(sys.argv[1], "B", Bn, "ms", e0.elapsed_time(e1) / 5, float(lp.sum()))
P
done
