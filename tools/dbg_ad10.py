import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import xped_b200 as xb
import xped_oracle as xo
from helpers import SPECS
from test_gpu_ad import heisenberg_setup
from bench import oracle_tensor_like
ctx = xb.Context(SPECS["SU2"]); osym = xo.Symmetry.SU2()
ctx.set_int("svd.truncated", 0)
A3, H = heisenberg_setup(ctx, 3)
ctm = xb.CTM(A3, 7)
for k, name in enumerate(["C1", "C2", "C3", "C4", "T1", "T2", "T3", "T4"]):
    ctm.set(name, ctm.get(name).setRandom(40 + k))
Q4, Q1 = ctm.contractCorner(3), ctm.contractCorner(0)
G = Q4.clone().setRandom(77)
for label, X, Y in (("corners", Q4, Q1), ("random on corner bases", Q4.clone().setRandom(8), Q1.clone().setRandom(5))):
    M = X @ Y
    oM = oracle_tensor_like(xo, osym, M); oG = oracle_tensor_like(xo, osym, G)
    for keep in (7, 10**6):
        gU, gS, gV = M.tSVD(keep, 0.0, False)
        oU, oS, oV = oM.tSVD(keep, 0.0, False)
        # L = tr(U U^T G)  ->  dU = (G + G^T)-weighted: use the tape-free closed form dU = W U with W = degeneracy-weighted (G^T + G)
        g_dU = (G.adjoint() + G) @ gU
        o_dU = (oG.adjoint() + oG) @ oU
        gg = M.tSVD_backward(gU, gS, gV, g_dU, None, None).download()
        og = oM.tSVD_backward(o_dU, None, None).data
        print(f"{label} keep {keep}: |grad gpu| {np.linalg.norm(gg):.3e} |grad oracle| {np.linalg.norm(og):.3e} rel diff {np.linalg.norm(gg - og) / max(np.linalg.norm(og), 1e-300):.2e}")
        # orthogonality of the device factors
        one = gU.multiply(gU, True, False)
        print("    max |U^T U - 1| =", np.abs(one.download() - one.clone().setIdentity().download()).max(), " max |V V^T - 1| =",
              np.abs(gV.multiply(gV, False, True).download() - gV.multiply(gV, False, True).clone().setIdentity().download()).max())
