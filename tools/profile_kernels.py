#!/usr/bin/env python
"""Small driver for ncu captures of the two roofline kernels on bandwidth/compute-bound sizes (run under gpurun):

    ncu --set full --clock-control none --import-source on -k regex:permute_bulk -c 3 -o gpurun_out/permute python tools/profile_kernels.py permute
    ncu --set full --clock-control none --import-source on -k regex:gemm_grouped  -c 1 -o gpurun_out/gemm    python tools/profile_kernels.py gemm
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xped_b200 as xb  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "permute"
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 24
if what in ("permute", "gemm"):
    ctx = xb.Context([("U1", 0, 0)])
    B = ctx.basis([(q, dim) for q in range(-4, 5)])
    T = ctx.tensor([B, B], [B, B]).setRandom(0)
    if what == "permute":
        for shift, p in [(0, (0, 1, 3, 2)), (0, (1, 0, 2, 3)), (1, (3, 2, 1, 0))]:   # copy mode, merged-load transpose, full reversal
            r = T.permute(shift, p)
            ctx.sync()
            del r
    else:
        r = T @ T
        ctx.sync()
elif what == "su2":
    ctx = xb.Context([("SU2", 0, 0)])
    a, e = ctx.basis([(1, 3), (2, 3)]), ctx.basis([(1, 16), (2, 32), (3, 36), (4, 30), (5, 18), (6, 8), (7, 4)])
    Q = ctx.tensor([e, a, a], [e, a, a]).setRandom(1)
    for shift, p in [(-1, (0, 2, 3, 5, 1, 4)), (1, (0, 3, 5, 1, 2, 4))]:
        r = Q.permute(shift, p)
        ctx.sync()
    r = Q @ Q
    ctx.sync()
elif what == "svd":
    ctx = xb.Context([("SU2", 0, 0)])
    a, e = ctx.basis([(1, 2), (2, 2)]), ctx.basis([(1, 8), (2, 14), (3, 16), (4, 12), (5, 8), (6, 4), (7, 2)])
    Q4 = ctx.tensor([e, a, a], [e, a, a]).setRandom(1)
    Q1 = ctx.tensor([e, a, a], [e, a, a]).setRandom(2)
    M = Q4 @ Q1
    U, S, Vd = M.tSVD(64, 1e-10, False)
    ctx.sync()
    print("tSVD done", ctx.stats())
ctx.close()
