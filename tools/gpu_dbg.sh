#!/bin/bash
mkdir -p gpurun_out
XB_SVD_NOGRAPH=1 timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 5200 -c 2200 --csv --log-file gpurun_out/launches_rsvd_d6.csv python tools/rsvd_check.py 6 144 4 > gpurun_out/ncu_rsvd_d6.log 2>&1
python tools/launch_summary.py gpurun_out/launches_rsvd_d6.csv | head -24
