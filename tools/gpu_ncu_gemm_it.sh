#!/bin/bash
mkdir -p gpurun_out
XB_SVD_NOGRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_grouped_kernel -s 700 -c 6 -o gpurun_out/gemm_iter_d6 -f python tools/rsvd_check.py 6 144 4 > gpurun_out/ncu_gemm_iter.log 2>&1
ls -la gpurun_out/gemm_iter_d6.ncu-rep
