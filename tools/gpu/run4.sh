#!/bin/bash
# round-2 GPU run 4: leaf phases + A/B, the whole GPU suite, config timings, the bench line with its CPU baseline, launch list, ncu
set -u
mkdir -p gpurun_out
echo "== leaf phases"; timeout 120 python tools/time_leaf_phases.py 2>&1 | tail -2 | tee gpurun_out/r02d_leaf_phases.txt
echo "== leaf A/B"; timeout 600 python tools/time_leaf.py 2>&1 | tee gpurun_out/r02d_leaf_ab.jsonl | python -c "
import sys, json
for line in sys.stdin:
    try: d = json.loads(line)
    except Exception: print(line[:1500]); continue
    print({k: v for k, v in d.items() if not isinstance(v, dict)})
"
echo "== pytest gpu (all)"
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=5 2>&1 | tail -14 | tee gpurun_out/r02d_pytest_gpu.log
echo "== configs"
timeout 900 python tools/run_configs.py 2>&1 | tail -40
echo "== bench"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02d_bench_n1.json 2> gpurun_out/r02d_bench_n1.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r02d_bench_n1.err
python - <<'P'
import json
b=json.load(open('gpurun_out/r02d_bench_n1.json'))
print({k:b[k] for k in ('value','ms_per_step')}, b['e2e']['value'], b['native_fp64']['ms_per_step'], b['emulated_8_slices']['ms_per_step'], b['roofline']['frac'], b['posterior_solve']['marginals'], b['posterior_solve']['full_covariance'], b['sharded_c3'].get('ms_per_step'), b['cpu_baseline'], b['parity_vs_oracle_rel'], b['gpu_library_baseline']['ms_per_step'])
P
echo "== launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02d_launches_logpdf16384.csv python tools/one_logpdf.py 16384 1 > gpurun_out/r02d_launches.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/r02d_launches_logpdf16384.csv 2>/dev/null | head -9
echo "== ncu full: recursive leaf"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'potrf_leaf_rec_kernel|diag_syrk_kernel' -c 3 -o gpurun_out/r02d_leaf -f python tools/one_logpdf.py 16384 1 > gpurun_out/r02d_ncu_leaf.log 2>&1; echo "ncu rc=$?"
