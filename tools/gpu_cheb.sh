#!/bin/bash
mkdir -p gpurun_out
XB_SVD_PROF=1 timeout 600 python tools/rsvd_check.py 4 64 14 > gpurun_out/cheb_d4.log 2>&1; grep "^step [7-9]\|^step 1[0-3]\|tSVD of\|U S Vd\|rror" gpurun_out/cheb_d4.log; grep "^\[rsvd\] K" gpurun_out/cheb_d4.log | tail -6 | cut -c1-250
XB_SVD_PROF=1 timeout 900 python tools/rsvd_check.py 6 144 9 > gpurun_out/cheb_d6.log 2>&1; grep "^step [4-8]\|tSVD of\|U S Vd\|rror" gpurun_out/cheb_d6.log;  grep "^\[rsvd\] K" gpurun_out/cheb_d6.log | tail -6 | cut -c1-250
