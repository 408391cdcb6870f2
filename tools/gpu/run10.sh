#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== A/B pairs (banded tile order)"
timeout 200 python /tmp/t16k.py 2>/dev/null || true
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
    for _ in range(10): lp = S.GP(k)(x).logpdf(y)
    e1.record(); torch.cuda.synchronize()
    print("pairs" if not os.environ.get("GPK_NO_PAIRS") else "single", prec, "logpdf ms", round(e0.elapsed_time(e1) / 10, 3), repr(float(lp)))
P
timeout 200 python /tmp/t16k.py; GPK_NO_PAIRS=1 timeout 200 python /tmp/t16k.py
echo "== tests"
timeout 900 python -m pytest tests/test_emulation.py tests/test_configs.py tests/test_gpu_primitives.py tests/test_autograd.py tests/test_sparse_streamed.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
timeout 600 python -m pytest tests/test_full_size_parity.py -m gpu -q -x -p no:cacheprovider -k "c2 or keeps" 2>&1 | tail -3
echo "== ncu: K=1024 launch traffic"
timeout 300 ncu --set full --clock-control none -k regex:'oz_gemm_kernel' -s 6 -c 1 -o gpurun_out/r02i_oz_gemm_k1024 -f python tools/one_logpdf.py 16384 1 > gpurun_out/r02i_ncu_oz.log 2>&1; echo "ncu rc=$?"
GPK_NO_PAIRS=1 timeout 300 ncu --set full --clock-control none -k regex:'oz_gemm_kernel' -s 8 -c 1 -o gpurun_out/r02i_oz_gemm_k512 -f python tools/one_logpdf.py 16384 1 > gpurun_out/r02i_ncu_oz512.log 2>&1; echo "ncu rc=$?"
echo "== bench"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-c3 > gpurun_out/r02i_bench_n1.json 2> gpurun_out/r02i_bench_n1.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02i_bench_n1.err
python - <<'P'
import json
b=json.load(open('gpurun_out/r02i_bench_n1.json')); r=b['roofline']
print({k:b[k] for k in ('value','ms_per_step')}, b['e2e']['value'], r['frac'], r['kernel_ms_per_step'], r['launches_per_step'], b['emulated_8_slices']['ms_per_step'], b['posterior_solve']['marginals']['ms'])
P
