#!/bin/bash
# ncu captures only (permute W4, gemm W4, svd round kernel). Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:svd_round -s 60 -c 3 -o gpurun_out/svd_round -f python tools/profile_kernels.py svd > gpurun_out/ncu_svd.log 2>&1
tail -3 gpurun_out/ncu_svd.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:permute_kernel -c 3 -o gpurun_out/permute_w4 -f python tools/profile_kernels.py permute > gpurun_out/ncu_permute.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_grouped -c 1 -o gpurun_out/gemm_w4 -f python tools/profile_kernels.py gemm > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out
