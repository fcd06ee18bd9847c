#!/bin/bash
# One gpurun call: parity tests, smoke, bench, launch list, ncu captures. Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"
# launch list of the same command (per-launch times are cold-cache and serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-extras --cpu-steps 0 > gpurun_out/bench_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:permute_bulk -c 3 -o gpurun_out/permute_w4 -f python tools/profile_kernels.py permute > gpurun_out/ncu_permute.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_grouped -c 1 -o gpurun_out/gemm_w4 -f python tools/profile_kernels.py gemm > gpurun_out/ncu_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:svd_round_resident -s 100 -c 2 -o gpurun_out/svd_round -f python tools/profile_kernels.py svd > gpurun_out/ncu_svd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qr_ -s 20 -c 2 -o gpurun_out/svd_qr -f python tools/profile_kernels.py svd > gpurun_out/ncu_qr.log 2>&1
ls -la gpurun_out
