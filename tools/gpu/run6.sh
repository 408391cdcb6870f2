#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pass 1 (default kernels), per-test timeout 100 s"
timeout 1000 python -m pytest tests/test_gpu_primitives.py tests/test_configs.py tests/test_model.py tests/test_sparse_streamed.py -m gpu -q -x -p no:cacheprovider --timeout=100 2>&1 | tail -30 | cut -c1-300
echo "== pass 2 (GPK_K1_GENERIC=1)"
GPK_K1_GENERIC=1 timeout 1000 python -m pytest tests/test_gpu_primitives.py tests/test_configs.py tests/test_model.py tests/test_sparse_streamed.py -m gpu -q -x -p no:cacheprovider --timeout=100 2>&1 | tail -6 | cut -c1-300
