#!/bin/bash
# launch lists (ncu, one metric) of ONE steady-state CTM step at D=4 chi=64 and D=6 chi=144: the last step of tools/rsvd_check.py between cuProfilerStart / Stop
mkdir -p gpurun_out
for cfg in "4 64 10 d4" "6 144 8 d6"; do
  set -- $cfg
  XB_CHECK_NCU_LAST=1 XB_SVD_NOGRAPH=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_last_$4.csv python tools/rsvd_check.py $1 $2 $3 > gpurun_out/ncu_last_$4.log 2>&1
  python tools/launch_summary.py gpurun_out/launches_last_$4.csv > gpurun_out/launch_summary_last_$4.txt; head -14 gpurun_out/launch_summary_last_$4.txt; rm -f gpurun_out/launches_last_$4.csv
done
