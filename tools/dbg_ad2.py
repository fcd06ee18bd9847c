import os, sys, itertools
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import xped_b200 as xb
from helpers import SPECS
ctx = xb.Context(SPECS["SU2"])
a = ctx.basis([((1,), 1), ((2,), 1)]); e = ctx.basis([((1,), 2), ((2,), 2), ((3,), 1)]); p = ctx.basis([((2,), 1)])
def inner(x, y): return float(np.dot(x.download(), y.download()))
worst = 0
cases = [
 ([e, a, a], [e, a, a], [(-1, (0, 2, 3, 5, 1, 4)), (1, (0, 3, 5, 1, 2, 4)), (0, (1, 0, 2, 4, 3, 5)), (-2,(1,2,3,0,4,5)), (2,(0,2,4,1,3,5))]),
 ([e], [e, a, a], [(-2, (1, 2, 3, 0)), (-2, (0, 2, 3, 1)), (0, (0, 3, 1, 2)), (1,(1,0,2,3))]),
 ([e, a, a, a], [a, a, p], [(0, (0, 2, 4, 5, 1, 3, 6)), (0, (0, 3, 4, 5, 1, 2, 6)), (-1,(0,2,3,5,4,1,6)), (1,(1,3,4,5,0,2,6))]),
 ([a, a], [a, a, p], [(0, (0, 3, 1, 2, 4)), (0, (1, 2, 0, 3, 4)), (0,(3,4,2,0,1)), (1,(2,0,1,3,4)), (-1,(4,3,2,1,0))]),
]
for dom, cod, perms in cases:
    T = ctx.tensor(dom, cod).setRandom(1)
    for shift, pi in perms:
        try:
            P = T.permute(shift, pi)
        except Exception as ex:
            print("skip", shift, pi, ex); continue
        W = P.clone().setRandom(2)
        lhs = inner(P, W)
        rhs = inner(T, W.permute_adj(shift, pi))
        err = abs(lhs - rhs) / max(abs(lhs), 1e-300)
        worst = max(worst, err)
        print(f"rank ({len(dom)},{len(cod)}) shift {shift} p {pi}: <P T, W> = {lhs:.12e}, <T, P^T W> = {rhs:.12e}, rel diff {err:.1e}")
print("worst", worst)
# multiply with adjoint flags: gradient through the tape vs finite differences (W is a constant created OUTSIDE the tape)
A = ctx.tensor([e, a], [e, a]).setRandom(3); B = ctx.tensor([e, a], [e, a]).setRandom(4); W = ctx.tensor([e, a], [e, a]).setRandom(5)
for aa, bb in itertools.product((False, True), repeat=2):
    ctx.tape_begin()
    C = A.multiply(B, aa, bb)
    L = C.multiply(W, False, True)
    ctx.tape_seed_trace(L, 1.0); ctx.tape_backward()
    gA = ctx.tape_grad(A).download(); gB = ctx.tape_grad(B).download(); ctx.tape_end()
    for T, g, nm in ((A, gA, "A"), (B, gB, "B")):
        base = T.download(); h = 1e-6; k = 3
        vals = []
        for s in (1, -1):
            d = base.copy(); d[k] += s * h; T.upload(d)
            vals.append(A.multiply(B, aa, bb).multiply(W, False, True).trace())
        T.upload(base)
        print("multiply adj flags", aa, bb, nm, "tape", g[k], "fd", (vals[0] - vals[1]) / (2 * h))
# decompose (multiply, tSVD, diag_map op 2, products with adjoint operands)
G = ctx.tensor([e, a, a], [e, a, a]).setRandom(9)
for chi in (3, 5, 40):
    T1 = ctx.tensor([e, a, a], [e, a, a]).setRandom(6); T2 = ctx.tensor([e, a, a], [e, a, a]).setRandom(7)
    def f(tape):
        if tape: ctx.tape_begin()
        P1, P2 = xb.decompose(T1, T2, chi)
        L = (P2 @ P1) @ G
        val = L.trace()
        g = None
        if tape:
            ctx.tape_seed_trace(L, 1.0); ctx.tape_backward(); g = (ctx.tape_grad(T1).download(), ctx.tape_grad(T2).download()); ctx.tape_end()
        return val, g
    v, g = f(True)
    for T, gg, nm in ((T1, g[0], "T1"), (T2, g[1], "T2")):
        base = T.download(); h = 1e-6
        for k in (0, 7, 19):
            vals = []
            for s in (1, -1):
                d = base.copy(); d[k] += s * h; T.upload(d); vals.append(f(False)[0])
            T.upload(base)
            print("decompose chi", chi, nm, "k", k, "tape", gg[k], "fd", (vals[0] - vals[1]) / (2 * h))
# a single CTM move from a generic environment, loss = <new C1, W1> + <new T4, W4> + <new C4, W3>
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_ad import heisenberg_setup
A2, H = heisenberg_setup(ctx, 2)
for direction in (0, 1, 2, 3):
    def g_env(tape, chi=5):
        if tape: ctx.tape_begin()
        ctm = xb.CTM(A2, chi)
        for k, name in enumerate(["C1", "C2", "C3", "C4", "T1", "T2", "T3", "T4"]):
            ctm.set(name, ctm.get(name).setRandom(40 + k))
        ctm.move(direction)
        val, roots = 0.0, []
        for name in ["C1", "C2", "C3", "C4", "T1", "T2", "T3", "T4"]:
            t = ctm.get(name)
            L = t.multiply(t, False, True)
            val += L.trace(); roots.append(L)
        g = None
        if tape:
            for r in roots: ctx.tape_seed_trace(r, 1.0)
            ctx.tape_backward(); g = ctx.tape_grad(A2).download(); ctx.tape_end()
        return val, g
    v, g = g_env(True)
    base = A2.download(); h = 1e-6
    for k in (0, 5, 11):
        vals = []
        for s in (1, -1):
            d = base.copy(); d[k] += s * h; A2.upload(d); vals.append(g_env(False)[0])
        A2.upload(base)
        print("move", direction, "k", k, "tape", g[k], "fd", (vals[0] - vals[1]) / (2 * h))
