#!/bin/bash
mkdir -p gpurun_out
for o in 16 48 64; do
XB_SVD_TRUNC_OVERSAMPLING=$o XB_SVD_PROF=1 timeout 600 python tools/rsvd_check.py 4 64 12 > gpurun_out/q8_d4_$o.log 2>&1; echo "oversampling $o"; grep "^step [89]\|^step 10\|rror" gpurun_out/q8_d4_$o.log; grep "^\[rsvd\] K" gpurun_out/q8_d4_$o.log | tail -2 | cut -c1-250
done
