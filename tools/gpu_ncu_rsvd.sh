#!/bin/bash
# launch list (per-kernel durations) of the CTM steps with the truncated tSVD: D=4 steps 4..5 and one D=6 step
mkdir -p gpurun_out
XB_SVD_NOGRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP4:-2500} -c ${CNT4:-2500} --csv --log-file gpurun_out/launches_rsvd_d4.csv python tools/rsvd_check.py 4 64 6 > gpurun_out/ncu_rsvd_d4.log 2>&1
python tools/launch_summary.py gpurun_out/launches_rsvd_d4.csv | head -30
if [ "${D6:-1}" = "1" ]; then
XB_SVD_NOGRAPH=1 timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP6:-6000} -c ${CNT6:-2500} --csv --log-file gpurun_out/launches_rsvd_d6.csv python tools/rsvd_check.py 6 144 4 > gpurun_out/ncu_rsvd_d6.log 2>&1
python tools/launch_summary.py gpurun_out/launches_rsvd_d6.csv | head -30
fi
