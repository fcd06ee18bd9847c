import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import xped_b200 as xb
import xped_oracle as xo
from helpers import SPECS, obasis
from xped_oracle.rng import uniform_fill
ctx = xb.Context(SPECS["SU2"]); osym = xo.Symmetry.from_spec(SPECS["SU2"])
a = [((1,), 2), ((2,), 1)]; e = [((1,), 2), ((2,), 3), ((3,), 2), ((4,), 1)]
def mk(seed):
    g = ctx.tensor([ctx.basis(e), ctx.basis(a), ctx.basis(a)], [ctx.basis(e), ctx.basis(a), ctx.basis(a)]).setRandom(seed)
    o = xo.Tensor(osym, [obasis(osym, e), obasis(osym, a), obasis(osym, a)], [obasis(osym, e), obasis(osym, a), obasis(osym, a)])
    o.data = uniform_fill(seed, o.plainSize())
    return g, o
g1, o1 = mk(1); g2, o2 = mk(2); gB, oB = mk(3)
print("blocks", g1.structure()[1])
for label, g, o in (("random full rank", g1, o1), ("product (same shape)", g1 @ g2, o1 @ o2)):
    if label.startswith("product"):
        o.data = g.download().copy()
    for keep in (5, 9, 30):
        gU, gS, gV = g.tSVD(keep, 1e-10, False)
        oU, oS, oV = o.tSVD(keep, 1e-10, False)
        sg = np.sort(gS.download()[gS.download() > 0])[::-1]
        full = np.sort(np.concatenate([np.linalg.svd(o.block(i), compute_uv=False) for i in range(len(o.sector))]))[::-1]
        g_dU, o_dU = gB.multiply(gV, False, True) @ gS, (oB @ oV.adjoint()) @ oS
        g_dS, o_dS = gU.multiply(gB, True, False).multiply(gV, False, True), (oU.adjoint() @ oB) @ oV.adjoint()
        g_dV, o_dV = gS @ gU.multiply(gB, True, False), oS @ (oU.adjoint() @ oB)
        for nm, (a1, a2, a3), (b1, b2, b3) in (("all", (g_dU, g_dS, g_dV), (o_dU, o_dS, o_dV)), ("dU", (g_dU, None, None), (o_dU, None, None)), ("dS", (None, g_dS, None), (None, o_dS, None)),
                                               ("dV", (None, None, g_dV), (None, None, o_dV))):
            gg = g.tSVD_backward(gU, gS, gV, a1, a2, a3).download(); og = o.tSVD_backward(b1, b2, b3).data
            print(f"{label} keep {keep} {nm}: rel diff {np.linalg.norm(gg - og) / np.linalg.norm(og):.2e}  (sigma min/max of the input {full[-1] / full[0]:.1e})")
