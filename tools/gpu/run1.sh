#!/bin/bash
# round-2 GPU run 1: full GPU test suite (incl. the new full-size oracle parity tests), bench line, launch list, ncu captures
set -u
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_smi.txt 2>&1
nproc > gpurun_out/r02_host.txt; free -g >> gpurun_out/r02_host.txt
echo "== pytest (all gpu tests except the full-size file)"; 
timeout 900 python -m pytest tests -m gpu -q --ignore=tests/test_full_size_parity.py -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/r02_pytest_gpu.log
echo "== full-size parity"
timeout 1500 python -m pytest tests/test_full_size_parity.py -m gpu -q -s --durations=10 -p no:cacheprovider 2>&1 | tail -60 | tee gpurun_out/r02_pytest_full_size.log
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r02_bench_n1.err
echo "== launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_logpdf16384.csv python tools/one_logpdf.py 16384 1 > gpurun_out/r02_launches.log 2>&1; echo "ncu rc=$?"
echo "== ncu full: K1, leaf, trsm, slice, diag syrk"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'kernel_matrix_kernel|potrf_leaf_kernel|trsm_leaf_tc_kernel|oz_slice_kernel|diag_syrk_kernel' -c 10 -o gpurun_out/r02_small_kernels -f python tools/one_logpdf.py 16384 1 > gpurun_out/r02_ncu_small.log 2>&1; echo "ncu rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'oz_gemm_kernel' -s 8 -c 1 -o gpurun_out/r02_oz_gemm -f python tools/one_logpdf.py 16384 1 > gpurun_out/r02_ncu_oz.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out | tail -20
