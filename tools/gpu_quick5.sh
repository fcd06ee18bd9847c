#!/bin/bash
mkdir -p gpurun_out
for w in 9 12 16; do
XB_GEMM_WAVE_SPLIT=$w timeout 900 python tools/rsvd_check.py 6 144 7 > gpurun_out/q5_d6_$w.log 2>&1; echo "wave target $w"; grep "^step [5-6]\|rror" gpurun_out/q5_d6_$w.log
done
