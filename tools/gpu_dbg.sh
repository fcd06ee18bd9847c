#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_bench_shapes.py -x -q > gpurun_out/pytest_shapes.log 2>&1; tail -30 gpurun_out/pytest_shapes.log
XB_SVD_TRUNC=1 XB_SVD_PROF=1 timeout 600 python tools/rsvd_check.py 4 64 12 > gpurun_out/rsvd_d4_prof.log 2>&1
grep "^step" gpurun_out/rsvd_d4_prof.log | tail -4; grep "\[rsvd\]" gpurun_out/rsvd_d4_prof.log | tail -12
XB_SVD_TRUNC=1 XB_SVD_PROF=1 timeout 900 python tools/rsvd_check.py 6 144 8 > gpurun_out/rsvd_d6_t1.log 2>&1
grep "^step" gpurun_out/rsvd_d6_t1.log | tail -4; grep "\[rsvd\]" gpurun_out/rsvd_d6_t1.log | tail -10
timeout 900 python -m pytest tests -m gpu -x -q -k "not bench_shapes" > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
