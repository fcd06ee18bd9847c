#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/dbg_ad7.py > gpurun_out/dbg_ad7.log 2>&1; grep "corners chi\|random chi 7" gpurun_out/dbg_ad7.log | head -12
timeout 900 python -m pytest tests/test_gpu_ad.py tests/test_integration_shim.py -x -q > gpurun_out/pytest_ad.log 2>&1; tail -12 gpurun_out/pytest_ad.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "tsvd_backward or diag or decompose" > gpurun_out/pytest_p.log 2>&1; tail -4 gpurun_out/pytest_p.log
