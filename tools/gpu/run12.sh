#!/bin/bash
# final evidence run of round 2: whole GPU suite, smoke, config timings, full bench line, launch list
set -u
mkdir -p gpurun_out
echo "== pytest gpu (all)"
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider --durations=3 2>&1 | tail -8 | tee gpurun_out/r02n_pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
echo "== configs"
timeout 900 python tools/run_configs.py 2>&1 | grep -v Warn | tail -26
echo "== bench"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02n_bench_n1.json 2> gpurun_out/r02n_bench_n1.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02n_bench_n1.err
python - <<'P'
import json
b=json.load(open('gpurun_out/r02n_bench_n1.json')); r=b['roofline']
print({k:b[k] for k in ('value','ms_per_step')}, b['e2e']['value'], r['frac'], r['frac_vs_sustained'], r['kernel_ms_per_step'], r['fp64_equivalent_tflops'], r['traffic'], b['emulated_8_slices']['ms_per_step'], b['native_fp64']['ms_per_step'], b['gpu_library_baseline']['ms_per_step'], b['posterior_solve']['marginals']['ms'], b['posterior_solve']['full_covariance']['ms'], b['sharded_c3']['ms_per_step'], b['cpu_baseline']['value'], b['parity_vs_oracle_rel'])
P
echo "== launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02n_launches_logpdf16384.csv python tools/one_logpdf.py 16384 1 > gpurun_out/r02n_launches.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/r02n_launches_logpdf16384.csv 2>/dev/null | head -9
