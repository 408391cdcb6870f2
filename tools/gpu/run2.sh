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
