#!/bin/bash
mkdir -p gpurun_out
for ty in 4 8 16; do
XB_CHOL_TY=$ty XB_SVD_TRUNC=1 XB_SVD_PROF=1 timeout 600 python tools/rsvd_check.py 4 64 8 > gpurun_out/rsvd_d4_ty$ty.log 2>&1
echo "TY=$ty"; grep "^step" gpurun_out/rsvd_d4_ty$ty.log | tail -2; grep "\[rsvd\]" gpurun_out/rsvd_d4_ty$ty.log | tail -2
done
timeout 900 python -m pytest tests -m gpu -x -q -k "not bench_shapes" > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
