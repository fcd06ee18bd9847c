#!/bin/bash
# quick GPU check: parity tests + headline bench (no extras) ; outputs under gpurun_out/
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 900 python bench.py $BENCH_ARGS > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/bench_quick.json; tail -5 gpurun_out/bench_quick.err
