#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_io.py -x -q > gpurun_out/pytest_io.log 2>&1; tail -15 gpurun_out/pytest_io.log
bash tools/gpu_round2.sh
