#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/rsvd_check.py 4 64 10 > gpurun_out/sk_new_d4.log 2>&1; grep "^step [6-9]\|tSVD of\|U S Vd\|rror" gpurun_out/sk_new_d4.log
XB_GEMM_SPLITK_KERNEL=1 timeout 600 python tools/rsvd_check.py 4 64 10 > gpurun_out/sk_old_d4.log 2>&1; grep "^step [8-9]" gpurun_out/sk_old_d4.log
timeout 900 python tools/rsvd_check.py 6 144 6 > gpurun_out/sk_new_d6.log 2>&1; grep "^step [4-5]\|tSVD of\|U S Vd\|rror" gpurun_out/sk_new_d6.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
