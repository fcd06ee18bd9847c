#!/bin/bash
mkdir -p gpurun_out
XB_CHOL_OLD=1 timeout 600 python tools/rsvd_check.py 4 64 10 > gpurun_out/chol_old_d4.log 2>&1; grep "^step [6-9]\|tSVD of\|sigma head\|U S Vd" gpurun_out/chol_old_d4.log
timeout 600 python tools/rsvd_check.py 4 64 10 > gpurun_out/chol_new_d4.log 2>&1; grep "^step [6-9]\|tSVD of\|sigma head\|U S Vd\|rror" gpurun_out/chol_new_d4.log
timeout 900 python tools/rsvd_check.py 6 144 6 > gpurun_out/chol_new_d6.log 2>&1; grep "^step\|tSVD of\|sigma head\|U S Vd\|rror" gpurun_out/chol_new_d6.log
timeout 1200 python -m pytest tests -m gpu -x -q -k "tsvd or decompose or ctm or mps or shapes" > gpurun_out/pytest_svd.log 2>&1; tail -5 gpurun_out/pytest_svd.log
