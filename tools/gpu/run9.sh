#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== bench (full line)"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02h_bench_n1.json 2> gpurun_out/r02h_bench_n1.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r02h_bench_n1.err
python - <<'P'
import json
b=json.load(open('gpurun_out/r02h_bench_n1.json'))
r=b['roofline']
print({k:b[k] for k in ('value','ms_per_step')}, b['e2e']['value'], r['frac'], r['frac_vs_sustained'], r['kernel_ms_per_step'], r['launches_per_step'], r['fp64_equivalent_tflops'], b['emulated_8_slices']['ms_per_step'], b['native_fp64']['ms_per_step'], b['posterior_solve']['marginals']['ms'], b['parity_vs_oracle_rel'])
P
echo "== launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02h_launches_logpdf16384.csv python tools/one_logpdf.py 16384 1 > gpurun_out/r02h_launches.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/r02h_launches_logpdf16384.csv 2>/dev/null | head -9
echo "== ncu full: oz gemm K=1024 launch"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'oz_gemm_kernel' -s 6 -c 1 -o gpurun_out/r02h_oz_gemm_k1024 -f python tools/one_logpdf.py 16384 1 > gpurun_out/r02h_ncu_oz.log 2>&1; echo "ncu rc=$?"
echo "== configs c4 c5"
timeout 600 python tools/run_configs.py c4 c5 2>&1 | grep -v Warn | tail -14
