#!/bin/bash
mkdir -p gpurun_out
XB_SVD_PROF=1 timeout 600 python tools/rsvd_check.py 4 64 12 > gpurun_out/cheap_d4.log 2>&1; grep "^step [6-9]\|^step 1\|tSVD of\|U S Vd\|rror" gpurun_out/cheap_d4.log; grep "^\[rsvd\] K" gpurun_out/cheap_d4.log | tail -8 | cut -c1-260
XB_SVD_PROF=1 timeout 900 python tools/rsvd_check.py 6 144 8 > gpurun_out/cheap_d6.log 2>&1; grep "^step [3-7]\|tSVD of\|U S Vd\|rror" gpurun_out/cheap_d6.log;  grep "^\[rsvd\] K" gpurun_out/cheap_d6.log | tail -6 | cut -c1-260
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
