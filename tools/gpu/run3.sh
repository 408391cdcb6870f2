#!/bin/bash
# round-2 GPU run 3: square-root-free leaf A/B, GPU suite (minus the two slow full-size oracle tests), bench, ncu of the new kernels
set -u
mkdir -p gpurun_out
echo "== leaf A/B"; timeout 600 python tools/time_leaf.py 2>&1 | tee gpurun_out/r02c_leaf_ab.jsonl | python -c "
import sys, json
for line in sys.stdin:
    try: d = json.loads(line)
    except Exception: print(line[:1500]); continue
    print({k: v for k, v in d.items() if not isinstance(v, dict)})
"
echo "== pytest gpu"
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -k "not c4_full_size and not c5_full_size" 2>&1 | tail -15 | tee gpurun_out/r02c_pytest_gpu.log
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r02c_bench_n1.err
python - <<'P'
import json
b=json.load(open('gpurun_out/r02c_bench_n1.json'))
print({k:b[k] for k in ('value','ms_per_step')}, b['e2e']['value'], b['native_fp64']['ms_per_step'], b['emulated_8_slices']['ms_per_step'], b['roofline']['frac'], b['posterior_solve']['marginals'], b['posterior_solve']['full_covariance'], b['sharded_c3'].get('ms_per_step'))
P
echo "== launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02c_launches_logpdf16384.csv python tools/one_logpdf.py 16384 1 > gpurun_out/r02c_launches.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/r02c_launches_logpdf16384.csv | head -9
echo "== ncu full: recursive leaf, fast K1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'kernel_matrix_fast_kernel|potrf_leaf_rec_kernel' -c 4 -o gpurun_out/r02c_new_kernels -f python tools/one_logpdf.py 16384 1 > gpurun_out/r02c_ncu_new.log 2>&1; echo "ncu rc=$?"
