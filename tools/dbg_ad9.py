import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import xped_b200 as xb
from helpers import SPECS
from test_gpu_ad import heisenberg_setup
ctx = xb.Context(SPECS["SU2"])
ctx.set_int("svd.truncated", 0)
A3, H = heisenberg_setup(ctx, 3)
ctm = xb.CTM(A3, 7)
for k, name in enumerate(["C1", "C2", "C3", "C4", "T1", "T2", "T3", "T4"]):
    ctm.set(name, ctm.get(name).setRandom(40 + k))
Q4, Q1 = ctm.contractCorner(3), ctm.contractCorner(0)
M = Q4 @ Q1
def blocks(t):
    st = t.structure(); d = t.download()
    return {tuple(q): d[o:o + r * c].reshape((r, c), order="F") for q, r, c, o in zip(*st)}
runs = []
for i in range(3):
    U, S, Vd = M.tSVD(10**6, 0.0, False)
    runs.append((blocks(U), blocks(S), blocks(Vd)))
for i in (1, 2):
    for q in runs[0][0]:
        du = np.abs(runs[i][0][q] - runs[0][0][q]).max(); dv = np.abs(runs[i][2][q] - runs[0][2][q]).max()
        # sign pattern
        sg = np.sign(np.sum(runs[i][0][q] * runs[0][0][q], axis=0))
        print(f"full run {i} vs 0, sector {q}: max |dU| {du:.2e} max |dVd| {dv:.2e} columns with flipped sign {int((sg < 0).sum())}")
Ut, St, Vt = M.tSVD(7, 1e-10, False)
bu = blocks(Ut)
for q in bu:
    k = bu[q].shape[1]
    d = np.abs(bu[q] - runs[0][0][q][:, :k]).max()
    sg = np.sign(np.sum(bu[q] * runs[0][0][q][:, :k], axis=0))
    print(f"truncated(7) vs full, sector {q}: kept {k}, max |dU| {d:.2e}, flipped {int((sg < 0).sum())}")
