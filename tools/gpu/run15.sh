#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none -k regex:'gemm_nt_f64_v2_kernel|diag_syrk_kernel|logpdf_finish_kernel' -s 2 -c 5 -o gpurun_out/r02k_chol_small -f python tools/one_logpdf.py 16384 1 > gpurun_out/r02k_ncu1.log 2>&1; echo "ncu rc=$?"
timeout 400 ncu --set full --clock-control none -k regex:'transpose_scaled_kernel|row_dot_acc_kernel|sparse_rows_kernel|row_dot_sq_kernel|kernel_matrix_fast_kernel' -c 5 -o gpurun_out/r02k_sparse -f python tools/one_c4.py c4 > gpurun_out/r02k_ncu2.log 2>&1; echo "ncu rc=$?"
