#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/rsvd_check.py 4 64 24 > gpurun_out/q_d4.log 2>&1; grep "^step" gpurun_out/q_d4.log | cut -c1-22 | tr '\n' ' '; echo; grep "U S Vd\|rror" gpurun_out/q_d4.log
timeout 900 python tools/rsvd_check.py 6 144 9 > gpurun_out/q_d6.log 2>&1; grep "^step" gpurun_out/q_d6.log | cut -c1-22 | tr '\n' ' '; echo; grep "^step [78]" gpurun_out/q_d6.log | cut -c1-140; grep "U S Vd\|rror" gpurun_out/q_d6.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
