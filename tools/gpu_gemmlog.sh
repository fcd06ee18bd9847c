#!/bin/bash
mkdir -p gpurun_out
XB_GEMM_LOG=1 timeout 600 python tools/rsvd_check.py 6 144 4 > gpurun_out/gemmlog_d6.log 2>&1
XB_GEMM_LOG=1 timeout 600 python tools/rsvd_check.py 4 64 7 > gpurun_out/gemmlog_d4.log 2>&1
XB_SVD_NOGRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 7000 -c 2500 --csv --log-file gpurun_out/launches_r02b_d4.csv python tools/rsvd_check.py 4 64 8 > gpurun_out/ncu_r02b_d4.log 2>&1
python tools/launch_summary.py gpurun_out/launches_r02b_d4.csv > gpurun_out/launch_summary_r02b_d4.txt; head -14 gpurun_out/launch_summary_r02b_d4.txt
