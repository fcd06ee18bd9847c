#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:permute_lut -c 3 -o gpurun_out/permute_lut_w3 -f python tools/perm_w3.py > gpurun_out/ncu_permute_lut.log 2>&1
XB_PERMUTE_LUT=1 timeout 600 python tools/rsvd_check.py 4 64 10 > gpurun_out/perm_d4_lut1.log 2>&1; echo "LUT forced:"; grep "^step [8-9]\|rror" gpurun_out/perm_d4_lut1.log
XB_PERMUTE_LUT=1 timeout 600 python tools/rsvd_check.py 6 144 6 > gpurun_out/perm_d6_lut1.log 2>&1; grep "^step [4-5]\|rror" gpurun_out/perm_d6_lut1.log
