#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== K1 A/B (strip kernel)"; timeout 300 python tools/time_k1.py 2>&1 | tee gpurun_out/r02e_k1_ab.jsonl | cut -c1-1500
echo "== primitives + configs tests"; timeout 900 python -m pytest tests/test_gpu_primitives.py tests/test_configs.py tests/test_model.py tests/test_sparse_streamed.py -m gpu -q -p no:cacheprovider 2>&1 | tail -5
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== c5 / c4 timings"; timeout 600 python tools/run_configs.py c2 c4 c5 2>&1 | grep -v Warning | tail -22
