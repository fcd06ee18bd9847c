import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import xped_b200 as xb
from helpers import SPECS
from test_gpu_ad import heisenberg_setup, generic_env_energy
ctx = xb.Context(SPECS["SU2"])
for D, chis in ((3, (5, 6, 7, 8, 9, 12, 20)), (2, (4, 5, 6, 8))):
    A, H = heisenberg_setup(ctx, D)
    for chi in chis:
        for seed in (40, 60):
            e0, g = generic_env_energy(ctx, A, H, chi, seed, tape=True)
            gd = g.download(); base = A.download(); h = 1e-6
            worst = 0
            for k in (1, 7, 23 % base.size, 41 % base.size):
                vals = []
                for s in (+1, -1):
                    d = base.copy(); d[k] += s * h; A.upload(d)
                    vals.append(generic_env_energy(ctx, A, H, chi, seed, tape=False)[0])
                A.upload(base)
                fd = (vals[0] - vals[1]) / (2 * h)
                worst = max(worst, abs(fd - gd[k]) / max(1.0, np.abs(gd).max()))
            # spectrum at the cut of the first move
            ctm = xb.CTM(A, chi)
            for k, name in enumerate(["C1", "C2", "C3", "C4", "T1", "T2", "T3", "T4"]):
                ctm.set(name, ctm.get(name).setRandom(seed + k))
            M = ctm.contractCorner(3) @ ctm.contractCorner(0)
            U, S, Vd = M.tSVD(10**6, 0.0, False)
            s = np.sort(S.download()[S.download() > 0])[::-1]
            print(f"D={D} chi={chi} seed={seed}: E={e0:+.6e} |g|max={np.abs(gd).max():.3e} worst rel err {worst:.2e}; nsv {len(s)}, sigma[chi-2:chi+2]/s0 = {s[max(chi-2,0):chi+2] / s[0]}")
