import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import xped_b200 as xb
from test_gpu_ad import heisenberg_setup
from helpers import SPECS
ctx = xb.Context(SPECS["SU2"])
for D, chi, steps in [(2, 6, 1), (2, 6, 2), (2, 3, 2), (2, 20, 2), (3, 8, 1), (3, 8, 2), (3, 30, 1)]:
    A, H = heisenberg_setup(ctx, D)
    e0, g = xb.ctm_energy_grad(A, chi, steps, H, H)
    gd = g.download(); base = A.download(); h = 1e-6
    errs = []
    for k in range(min(base.size, 12)):
        vals = []
        for s in (+1, -1):
            d = base.copy(); d[k] += s * h; A.upload(d)
            vals.append(xb.ctm_energy_grad(A, chi, steps, H, H, want_grad=False)[0])
        A.upload(base)
        fd = (vals[0] - vals[1]) / (2 * h)
        errs.append((k, gd[k], fd, abs(gd[k] - fd)))
    c = xb.CTM(A, chi)
    for _ in range(steps): c.grow_all()
    print(f"D={D} chi={chi} steps={steps} E={e0:.10f} size={base.size} C1 structure {c.get('C1').structure()[1:3]} max abs err {max(e[3] for e in errs):.2e} (|g| max {np.abs(gd).max():.3f})")
    for e in errs[:6]: print("    k=%d tape % .8e fd % .8e diff %.2e" % e)
