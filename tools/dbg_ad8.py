import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import xped_b200 as xb
from helpers import SPECS
from test_gpu_ad import heisenberg_setup
ctx = xb.Context(SPECS["SU2"])
A3, H = heisenberg_setup(ctx, 3)
ctm = xb.CTM(A3, 7)
for k, name in enumerate(["C1", "C2", "C3", "C4", "T1", "T2", "T3", "T4"]):
    ctm.set(name, ctm.get(name).setRandom(40 + k))
Q4, Q1 = ctm.contractCorner(3), ctm.contractCorner(0)
M = Q4 @ Q1
st = M.structure(); d = M.download()
for q, r, c, o in zip(*st):
    blk = d[o:o + r * c].reshape((r, c), order="F")
    s = np.linalg.svd(blk, compute_uv=False)
    gaps = np.abs(np.diff(s)) / s[:-1] if len(s) > 1 else []
    print("sector", q, "size", r, "top", s[:8], "min rel gap among top 12:", (gaps[:12].min() if len(gaps) else None), "smallest", s[-3:])
for nm, T in (("Q4", Q4), ("Q1", Q1)):
    st = T.structure(); d = T.download()
    for q, r, c, o in zip(*st):
        blk = d[o:o + r * c].reshape((r, c), order="F")
        s = np.linalg.svd(blk, compute_uv=False)
        print(nm, "sector", q, "rank", int((s > 1e-10 * s[0]).sum()), "of", r, "asym", np.abs(blk - blk.T).max() / np.abs(blk).max())
