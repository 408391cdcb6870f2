#!/bin/bash
# final validation of the tree: whole GPU suite, smoke, bench line
set -u
mkdir -p gpurun_out
echo "== pytest gpu (all)"
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider --durations=4 2>&1 | tail -10 | tee gpurun_out/r02f_pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
echo "== bench"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r02f_bench_n1.err
python - <<'P'
import json
b=json.load(open('gpurun_out/r02f_bench_n1.json'))
print({k:b[k] for k in ('value','ms_per_step')}, b['e2e']['value'], b['roofline']['frac'], b['roofline']['traffic'], b['parity_vs_oracle_rel'], b['sharded_c3'].get('ms_per_step'))
P
