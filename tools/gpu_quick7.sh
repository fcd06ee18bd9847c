#!/bin/bash
mkdir -p gpurun_out
XB_SVD_PROF=1 timeout 600 python tools/rsvd_check.py 4 64 12 > gpurun_out/q7_d4.log 2>&1; grep "^step [7-9]\|^step 1[01]\|rror" gpurun_out/q7_d4.log; grep "^\[rsvd\] K" gpurun_out/q7_d4.log | tail -3 | cut -c1-250
XB_SVD_PROF=1 timeout 600 python tools/rsvd_check.py 6 144 8 > gpurun_out/q7_d6.log 2>&1; grep "^step [4-7]\|rror" gpurun_out/q7_d6.log; grep "^\[rsvd\] K" gpurun_out/q7_d6.log | tail -6 | cut -c1-250
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
