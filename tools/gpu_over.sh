#!/bin/bash
mkdir -p gpurun_out
for o in 16 32; do
XB_SVD_TRUNC_OVERSAMPLING=$o timeout 600 python tools/rsvd_check.py 4 64 14 > gpurun_out/over_d4_$o.log 2>&1; echo "oversampling $o"; grep "^step [6-9]\|^step 1[0-3]\|rror" gpurun_out/over_d4_$o.log | cut -c1-100
done
XB_SVD_TRUNC_OVERSAMPLING=16 timeout 900 python -m pytest tests -m gpu -x -q -k "tsvd or ctm or shapes" 2>&1 | tail -2
