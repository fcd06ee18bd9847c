import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import xped_b200 as xb
from helpers import SPECS
from test_gpu_ad import heisenberg_setup
ctx = xb.Context(SPECS["SU2"])
A3, H = heisenberg_setup(ctx, 3)
NAMES = ["C1", "C2", "C3", "C4", "T1", "T2", "T3", "T4"]
ctm = xb.CTM(A3, 7)
for k, name in enumerate(NAMES):
    ctm.set(name, ctm.get(name).setRandom(40 + k))
Q4, Q1 = ctm.contractCorner(3), ctm.contractCorner(0)
print("Q structure rows", Q4.structure()[1], "cols", Q4.structure()[2])
G = Q4.clone().setRandom(77)
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
for chi in (3, 7, 12, 1000):
    def f(tape, chi=chi):
        if tape: ctx.tape_begin()
        P1, P2 = xb.decompose(Q4, Q1, chi)
        L = (P2 @ P1) @ G
        val = L.trace(); g = None
        if tape:
            ctx.tape_seed_trace(L, 1.0); ctx.tape_backward(); g = ctx.tape_grad(Q4).download(); ctx.tape_end()
        return val, g
    fd(f"decompose(Q4,Q1,{chi}) wrt Q4", Q4, f)
    def f2(tape, chi=chi):
        if tape: ctx.tape_begin()
        M = Q4 @ Q1
        U, S, Vd = M.tSVD(chi, 1e-10, False)
        L = ((U @ S) @ Vd) @ G
        val = L.trace(); g = None
        if tape:
            ctx.tape_seed_trace(L, 1.0); ctx.tape_backward(); g = ctx.tape_grad(Q4).download(); ctx.tape_end()
        return val, g
    fd(f"U S Vd (chi {chi}) wrt Q4", Q4, f2)
    def f3(tape, chi=chi):
        if tape: ctx.tape_begin()
        M = Q4 @ Q1
        U, S, Vd = M.tSVD(chi, 1e-10, False)
        L = (U.multiply(U, False, True)) @ G
        val = L.trace(); g = None
        if tape:
            ctx.tape_seed_trace(L, 1.0); ctx.tape_backward(); g = ctx.tape_grad(Q4).download(); ctx.tape_end()
        return val, g
    fd(f"U U^T (chi {chi}) wrt Q4", Q4, f3)
    def f4(tape, chi=chi):
        if tape: ctx.tape_begin()
        M = Q4 @ Q1
        U, S, Vd = M.tSVD(chi, 1e-10, False)
        L = S @ S
        val = L.trace(); g = None
        if tape:
            ctx.tape_seed_trace(L, 1.0); ctx.tape_backward(); g = ctx.tape_grad(Q4).download(); ctx.tape_end()
        return val, g
    fd(f"tr S^2 (chi {chi}) wrt Q4", Q4, f4)
