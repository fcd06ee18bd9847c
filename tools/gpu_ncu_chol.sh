#!/bin/bash
mkdir -p gpurun_out
XB_SVD_NOGRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:rsvd_chol -s 60 -c 1 -o gpurun_out/rsvd_chol -f python tools/rsvd_check.py 4 64 4 > gpurun_out/ncu_chol.log 2>&1
ncu -i gpurun_out/rsvd_chol.ncu-rep --page details --csv > gpurun_out/rsvd_chol_details.csv 2>/dev/null
ncu -i gpurun_out/rsvd_chol.ncu-rep --page raw --csv > gpurun_out/rsvd_chol_raw.csv 2>/dev/null
ls -la gpurun_out | head
timeout 900 python -m pytest tests/test_gpu_bench_shapes.py -x -q -k "ctm_d4" > gpurun_out/pytest_shapes.log 2>&1; tail -30 gpurun_out/pytest_shapes.log
