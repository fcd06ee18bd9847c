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
def run(fn_after, tape, chi=7):
    if tape: ctx.tape_begin()
    try:
        ctm = xb.CTM(A3, chi)
        for k, name in enumerate(NAMES):
            ctm.set(name, ctm.get(name).setRandom(40 + k))
        roots, val = fn_after(ctm)
        g = None
        if tape:
            for r in roots: ctx.tape_seed_trace(r, 1.0)
            ctx.tape_backward(); g = ctx.tape_grad(A3).download()
        return val, g
    finally:
        if tape: ctx.tape_end()
def check(label, fn_after, chi=7):
    v, g = run(fn_after, True, chi)
    base = A3.download(); h = 1e-6; out = []
    for k in (1, 7, 23, 41):
        vals = []
        for s in (1, -1):
            d = base.copy(); d[k] += s * h; A3.upload(d); vals.append(run(fn_after, False, chi)[0])
        A3.upload(base)
        out.append((g[k], (vals[0] - vals[1]) / (2 * h)))
    print(label, " ".join(f"[{a:+.6e} {b:+.6e}]" for a, b in out), flush=True)
def sq(ts):
    roots, val = [], 0.0
    for t in ts:
        L = t.multiply(t, False, True); roots.append(L); val += L.trace()
    return roots, val
for d in range(4):
    def f(ctm, d=d):
        ctm.move(d); return sq([ctm.get(n) for n in NAMES])
    check(f"move {d}", f)
for c in range(4):
    def f(ctm, c=c):
        return sq([ctm.contractCorner(c)])
    check(f"corner {c}", f)
def frho(ctm):
    roots, val = [], 0.0
    for rho in (ctm.computeRDM_h(), ctm.computeRDM_v()):
        L = rho.twist(0).twist(1) @ H; roots.append(L); val += L.trace()
    return roots, val
check("rdm energy (no move)", frho)
def f2(ctm):
    ctm.move(0); ctm.move(1); return sq([ctm.get(n) for n in NAMES])
check("moves 0,1", f2)
def f3(ctm):
    ctm.grow_all(); return sq([ctm.get(n) for n in NAMES])
check("grow_all", f3)
def f4(ctm):
    ctm.grow_all(); return frho(ctm)
check("grow_all+energy", f4)
