"""GPU experiment driver (not part of the product): tSVD phase timing inside a D=4 chi=64 CTM step, W4 permute per-case table."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xped_b200 as xb

what = sys.argv[1:] or ["svd", "perm"]
if "svd" in what:
    ctx = xb.Context([("SU2", 0, 0)])
    a, phys = ctx.basis([(1, 2), (2, 2)]), ctx.basis([(2, 1)])
    A = ctx.tensor([a, a], [a, a, phys]).setRandom(0)
    ctm = xb.CTM(A, 64)
    for _ in range(6):
        ctm.grow_all()
    ctx.sync()
    for rep in range(2):
        l0 = ctx.stats()["launches"]
        ctx.timer_start(); t0 = time.perf_counter()
        ctm.grow_all()
        ms = ctx.timer_stop()
        print(f"CTM step: {ms:.2f} ms device, {(time.perf_counter()-t0)*1e3:.2f} ms wall, launches {ctx.stats()['launches']-l0}", flush=True)
    ctx.close()
if "perm" in what:
    ctx = xb.Context([("U1", 0, 0)])
    B = ctx.basis([(q, 24) for q in range(-4, 5)])
    T = ctx.tensor([B, B], [B, B]).setRandom(0)
    perms = [(1, 0, 2, 3), (0, 1, 3, 2), (2, 3, 0, 1), (0, 2, 1, 3), (3, 2, 1, 0)]
    for p in perms:
        for shift in (0, 1, -1, 2, -2):
            T.bench_permute(shift, p, reps=2)
            ctx.flush_l2()
            ms, nbytes = T.bench_permute(shift, p, reps=5)
            print(f"perm p={p} shift={shift:+d}: {ms:.3f} ms {nbytes/ms/1e6:.0f} GB/s", flush=True)
    ctx.close()
if "w3" in what:
    ctx3 = xb.Context([("SU2", 0, 0)])
    a, e = ctx3.basis([(1, 3), (2, 3)]), ctx3.basis([(1, 16), (2, 32), (3, 36), (4, 30), (5, 18), (6, 8), (7, 4)])
    Q4 = ctx3.tensor([e, a, a], [e, a, a]).setRandom(1)
    Q1 = ctx3.tensor([e, a, a], [e, a, a]).setRandom(2)
    M = Q4 @ Q1
    for rep in range(2):
        ctx3.sync(); ctx3.timer_start()
        U, S, Vd = M.tSVD(144, 1e-10, False)
        print(f"tSVD W3: {ctx3.timer_stop():.1f} ms", flush=True)
    ctx3.close()
