#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 300 python - <<'PY' > gpurun_out/svd_timing.log 2>&1
import sys, time
sys.path.insert(0, ".")
import xped_b200 as xb
ctx = xb.Context([("SU2", 0, 0)])
a, e = ctx.basis([(1, 2), (2, 2)]), ctx.basis([(1, 8), (2, 14), (3, 16), (4, 12), (5, 8), (6, 4), (7, 2)])
Q4 = ctx.tensor([e, a, a], [e, a, a]).setRandom(1)
Q1 = ctx.tensor([e, a, a], [e, a, a]).setRandom(2)
M = Q4 @ Q1
for rep in range(3):
    l0 = ctx.stats()["launches"]
    ctx.sync(); ctx.timer_start(); t0 = time.perf_counter()
    U, S, Vd = M.tSVD(64, 1e-10, False)
    ms = ctx.timer_stop(); wall = (time.perf_counter() - t0) * 1e3
    n = ctx.stats()["launches"] - l0
    print(f"tSVD: {ms:.2f} ms device-events, {wall:.2f} ms wall, {n} launches, {1e3*ms/n:.1f} us/launch")
PY
cat gpurun_out/svd_timing.log
timeout 900 python bench.py $BENCH_ARGS > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_quick.err
