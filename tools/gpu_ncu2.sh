#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:svd_round -s 60 -c 2 -o gpurun_out/svd_round -f python tools/profile_kernels.py svd > gpurun_out/ncu_svd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:permute_kernel -c 3 -o gpurun_out/permute_w4 -f python tools/profile_kernels.py permute > gpurun_out/ncu_permute.log 2>&1
# un-profiled timing of the svd rounds: nsys-free, plain run with the library's own event profile
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
