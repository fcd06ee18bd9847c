#!/bin/bash
# quick single-GPU check: CTM step times at D=4 / D=6 (truncated tSVD path) + the tSVD / CTM / shape tests
mkdir -p gpurun_out
timeout 600 python tools/rsvd_check.py 4 64 12 > gpurun_out/q_d4.log 2>&1; grep "^step [7-9]\|^step 1[01]\|tSVD of\|U S Vd\|rror" gpurun_out/q_d4.log
timeout 900 python tools/rsvd_check.py 6 144 7 > gpurun_out/q_d6.log 2>&1; grep "^step [4-6]\|tSVD of\|U S Vd\|rror" gpurun_out/q_d6.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
