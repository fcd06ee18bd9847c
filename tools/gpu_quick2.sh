#!/bin/bash
mkdir -p gpurun_out
XB_SVD_PROF=1 timeout 600 python tools/rsvd_check.py 4 64 12 > gpurun_out/q_d4.log 2>&1; grep "^step [7-9]\|^step 1[01]\|U S Vd\|rror" gpurun_out/q_d4.log; grep "^\[rsvd\] K" gpurun_out/q_d4.log | tail -4 | cut -c1-230
XB_SVD_PROF=1 timeout 900 python tools/rsvd_check.py 6 144 7 > gpurun_out/q_d6.log 2>&1; grep "^step [4-6]\|U S Vd\|rror" gpurun_out/q_d6.log; grep "^\[rsvd\] K" gpurun_out/q_d6.log | tail -3 | cut -c1-230
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
