#!/bin/bash
mkdir -p gpurun_out
XB_PERMUTE_LUT=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_io.py -x -q > gpurun_out/pytest_perm_lut.log 2>&1; tail -4 gpurun_out/pytest_perm_lut.log
python tools/perm_w3.py 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "permute or contract or fill" > gpurun_out/pytest_perm.log 2>&1; tail -2 gpurun_out/pytest_perm.log
timeout 600 python tools/rsvd_check.py 6 144 6 > gpurun_out/perm_d6.log 2>&1; grep "^step [4-5]\|rror" gpurun_out/perm_d6.log
timeout 600 python tools/rsvd_check.py 4 64 10 > gpurun_out/perm_d4.log 2>&1; grep "^step [8-9]\|rror" gpurun_out/perm_d4.log
