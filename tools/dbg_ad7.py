import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import xped_b200 as xb
from helpers import SPECS
from test_gpu_ad import heisenberg_setup
ctx = xb.Context(SPECS["SU2"])
def fd(label, T, f, ks=(1, 50, 300, 777)):
    v, g = f(True)
    base = T.download(); h = 1e-6 * max(1.0, np.abs(base).max()); out = []
    for k in ks:
        k = k % base.size
        vals = []
        for s in (1, -1):
            d = base.copy(); d[k] += s * h; T.upload(d); vals.append(f(False)[0])
        T.upload(base)
        out.append((g[k], (vals[0] - vals[1]) / (2 * h)))
    print(label, " ".join(f"[{a:+.6e} {b:+.6e}]" for a, b in out), flush=True)
def suite(tag, X, Y, G):
    M0 = X @ Y
    U, S, Vd = M0.tSVD(10**6, 0.0, False)
    s = np.sort(S.download()[S.download() > 0])[::-1]
    print(tag, "blocks", M0.structure()[1], "sigma max", s[0], "sigma min", s[-1], "count", len(s), "of", sum(min(r, c) for r, c in zip(M0.structure()[1], M0.structure()[2])))
    for chi in (3, 7):
        for nm, loss in (("U S Vd", lambda U, S, Vd: ((U @ S) @ Vd) @ G), ("U U^T", lambda U, S, Vd: U.multiply(U, False, True) @ G), ("Vd^T Vd", lambda U, S, Vd: Vd.multiply(Vd, True, False) @ G)):
            def f(tape, chi=chi, loss=loss):
                if tape: ctx.tape_begin()
                M = X @ Y
                U, S, Vd = M.tSVD(chi, 1e-10, False)
                L = loss(U, S, Vd)
                val = L.trace(); g = None
                if tape:
                    ctx.tape_seed_trace(L, 1.0); ctx.tape_backward(); g = ctx.tape_grad(X).download(); ctx.tape_end()
                return val, g
            fd(f"  {tag} chi {chi} {nm}:", X, f)
a = ctx.basis([((1,), 2), ((2,), 1)]); e = ctx.basis([((1,), 2), ((2,), 3), ((3,), 2), ((4,), 1)])
X = ctx.tensor([e, a, a], [e, a, a]).setRandom(1); Y = ctx.tensor([e, a, a], [e, a, a]).setRandom(2); G = ctx.tensor([e, a, a], [e, a, a]).setRandom(3)
suite("random", X, Y, G)
A3, H = heisenberg_setup(ctx, 3)
ctm = xb.CTM(A3, 7)
for k, name in enumerate(["C1", "C2", "C3", "C4", "T1", "T2", "T3", "T4"]):
    ctm.set(name, ctm.get(name).setRandom(40 + k))
Q4, Q1 = ctm.contractCorner(3), ctm.contractCorner(0)
suite("corners", Q4, Q1, Q4.clone().setRandom(77))
suite("corners, Q1 random", Q4, Q1.clone().setRandom(5), Q4.clone().setRandom(77))
suite("random on the corner bases", Q4.clone().setRandom(8), Q1.clone().setRandom(5), Q4.clone().setRandom(77))
