#!/bin/bash
mkdir -p gpurun_out
XB_PLAN_NOSORT=1 python tools/perm_w3.py 2>&1 | tail -3
python tools/perm_w3.py 2>&1 | tail -3
timeout 600 python tools/rsvd_check.py 6 144 6 > gpurun_out/perm_d6.log 2>&1; grep "^step [4-5]\|rror" gpurun_out/perm_d6.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "permute or contract" 2>&1 | tail -2
