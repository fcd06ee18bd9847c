"""W3 (SU(2) D=6 chi=144 corner) permutes: GB/s of the three CTM-type cases + the permutes inside a D=6 CTM step"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xped_b200 as xb
ctx = xb.Context([("SU2", 0, 0)])
a, e = ctx.basis([(1, 3), (2, 3)]), ctx.basis([(1, 16), (2, 32), (3, 36), (4, 30), (5, 18), (6, 8), (7, 4)])
Q = ctx.tensor([e, a, a], [e, a, a]).setRandom(1)
for shift, p in [(-1, (0, 2, 3, 5, 1, 4)), (1, (0, 3, 5, 1, 2, 4)), (0, (1, 0, 2, 4, 3, 5))]:
    Q.bench_permute(shift, p, reps=2)
    ctx.flush_l2()
    ms, nbytes = Q.bench_permute(shift, p, reps=5)
    print(f"W3 permute shift {shift} p {p}: {ms:.4f} ms  {nbytes / ms / 1e6:.0f} GB/s", flush=True)
