#!/bin/bash
mkdir -p gpurun_out
XB_SVD_NOPRECOND=1 XB_SVD_PROF=1 timeout 600 python tools/rsvd_check.py 4 64 10 > gpurun_out/q6_d4.log 2>&1; grep "^step [7-9]\|rror" gpurun_out/q6_d4.log; grep "^\[rsvd\] K" gpurun_out/q6_d4.log | tail -5 | cut -c1-250
XB_SVD_NOPRECOND=1 XB_SVD_PROF=1 timeout 600 python tools/rsvd_check.py 6 144 6 > gpurun_out/q6_d6.log 2>&1; grep "^step [4-5]\|rror" gpurun_out/q6_d6.log; grep "^\[rsvd\] K" gpurun_out/q6_d6.log | tail -3 | cut -c1-250
