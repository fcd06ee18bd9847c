#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/rsvd_check.py 4 64 12 > gpurun_out/q9_d4.log 2>&1; grep "^step [7-9]\|^step 10\|U S Vd\|rror" gpurun_out/q9_d4.log
timeout 900 python tools/rsvd_check.py 6 144 7 > gpurun_out/q9_d6.log 2>&1; grep "^step [5-6]\|U S Vd\|rror" gpurun_out/q9_d6.log
timeout 900 python -m pytest tests -m gpu -x -q -k "tsvd or shapes or ctm or decompose or mps" > gpurun_out/pytest_q9.log 2>&1; tail -2 gpurun_out/pytest_q9.log
