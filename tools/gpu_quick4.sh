#!/bin/bash
mkdir -p gpurun_out
XB_SVD_PROF=1 timeout 900 python tools/rsvd_check.py 6 144 7 > gpurun_out/q4_d6.log 2>&1; grep "^step [4-6]\|tSVD of\|U S Vd\|rror" gpurun_out/q4_d6.log; grep "^\[rsvd\] K" gpurun_out/q4_d6.log | tail -3 | cut -c1-250
timeout 600 python tools/rsvd_check.py 4 64 10 > gpurun_out/q4_d4.log 2>&1; grep "^step [7-9]\|rror" gpurun_out/q4_d4.log
timeout 900 python -m pytest tests -m gpu -x -q -k "tsvd or shapes or ctm or multiply or contract" > gpurun_out/pytest_q4.log 2>&1; tail -3 gpurun_out/pytest_q4.log
