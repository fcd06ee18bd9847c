#!/bin/bash
# round-2 profiling pass: launch lists of the D=4 and D=6 CTM steps, full captures of the dominant kernels
mkdir -p gpurun_out
XB_SVD_NOGRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 3000 --csv --log-file gpurun_out/launches_rsvd_d4.csv python tools/rsvd_check.py 4 64 8 > gpurun_out/ncu_rsvd_d4.log 2>&1
python tools/launch_summary.py gpurun_out/launches_rsvd_d4.csv > gpurun_out/launch_summary_d4.txt; head -30 gpurun_out/launch_summary_d4.txt
XB_SVD_NOGRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:rsvd_chol -s 400 -c 1 -o gpurun_out/rsvd_chol_d6 -f python tools/rsvd_check.py 6 144 3 > gpurun_out/ncu_chol_d6.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:permute_kernel -c 2 -o gpurun_out/permute_w3 -f python tools/profile_kernels.py su2 > gpurun_out/ncu_permute_w3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_grouped -c 1 -o gpurun_out/gemm_w3 -f python tools/profile_kernels.py su2 > gpurun_out/ncu_gemm_w3.log 2>&1
ls -la gpurun_out/*.ncu-rep
