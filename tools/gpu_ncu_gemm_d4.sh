#!/bin/bash
mkdir -p gpurun_out
XB_CHECK_NCU_LAST=1 XB_SVD_NOGRAPH=1 timeout 500 ncu --profile-from-start off -k regex:gemm_grouped --set full --import-source on --clock-control none -s 24 -c 10 -o gpurun_out/gemm_d4_iter -f python tools/rsvd_check.py 4 64 10 > gpurun_out/ncu_gemm_d4_iter.log 2>&1
tail -2 gpurun_out/ncu_gemm_d4_iter.log
