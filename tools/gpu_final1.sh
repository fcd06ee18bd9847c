#!/bin/bash
# final single-GPU pass of round 2: tests, smoke, bench (both arms), ncu captures of the round-2 kernels
bash tools/gpu_round2.sh
XB_SVD_NOGRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:rsvd_chol_blocked -s 80 -c 1 -o gpurun_out/chol_blocked_d4 -f python tools/rsvd_check.py 4 64 6 > gpurun_out/ncu_chol_blocked.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:permute_kernel -c 3 -o gpurun_out/permute_w3_r02b -f python tools/perm_w3.py > gpurun_out/ncu_permute_w3b.log 2>&1
XB_SVD_NOGRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 7000 -c 2500 --csv --log-file gpurun_out/launches_final_d4.csv python tools/rsvd_check.py 4 64 8 > gpurun_out/ncu_final_d4.log 2>&1
python tools/launch_summary.py gpurun_out/launches_final_d4.csv > gpurun_out/launch_summary_final_d4.txt
ls -la gpurun_out/*.ncu-rep | tail -3
