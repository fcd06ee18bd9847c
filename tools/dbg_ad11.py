import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
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
G = Q4.clone().setRandom(77)
def blocks(t):
    st = t.structure(); d = t.download()
    return {tuple(q): d[o:o + r * c].reshape((r, c), order="F") for q, r, c, o in zip(*st)}
M = Q4 @ Q1
U, S, Vd = M.tSVD(10**6, 0.0, False)
dU = (G.adjoint() + G) @ U
gg = blocks(M.tSVD_backward(U, S, Vd, dU, None, None))
bU, bS, bV, bdU, bM = blocks(U), blocks(S), blocks(Vd), blocks(dU), blocks(M)
for q in bU:
    u, s, vt, du = bU[q], np.diag(bS[q]), bV[q], bdU[q]
    n = len(s)
    d = s[None, :] ** 2 - s[:, None] ** 2
    F = np.where(np.abs(d) < 1e-12, d / (d * d + 1e-12), 1 / np.where(d == 0, 1, d)); np.fill_diagonal(F, 0)
    G1 = u.T @ du
    J = F * G1
    term1 = u @ ((J + J.T) * s[None, :]) @ vt
    term2 = (np.eye(n) - u @ u.T) @ du @ np.diag(1 / s) @ vt
    ref = term1 + term2
    print(f"sector {q} n={n}: |G1 - G1^T|/|G1| = {np.abs(G1 - G1.T).max() / np.abs(G1).max():.1e}; numpy formula |dA| = {np.abs(ref).max():.2e}; device |dA| = {np.abs(gg[q]).max():.2e}; "
          f"|A - U S Vt|/|A| = {np.abs((u * s) @ vt - bM[q]).max() / np.abs(bM[q]).max():.1e}; sigma sorted desc: {bool(np.all(np.diff(s) <= 0))}; min s {s.min():.3e}")
print("---- probes")
rng = np.random.default_rng(0)
# (a) dS only
dS = S.clone()
gS = blocks(M.tSVD_backward(U, S, Vd, None, dS, None))
for q in bU:
    s = np.diag(bS[q]); ref = (bU[q] * s) @ bV[q]
    print(f"(a) dS = S: sector {q}: device vs U dS Vt rel diff {np.abs(gS[q] - ref).max() / np.abs(ref).max():.1e}")
# (c) random dU
dUr = U.clone().setRandom(5)
gR = blocks(M.tSVD_backward(U, S, Vd, dUr, None, None)); bdUr = blocks(dUr)
for q in bU:
    u, s, vt, du = bU[q], np.diag(bS[q]), bV[q], bdUr[q]
    n = len(s)
    d = s[None, :] ** 2 - s[:, None] ** 2
    F = np.where(np.abs(d) < 1e-12, d / (d * d + 1e-12), 1 / np.where(d == 0, 1, d)); np.fill_diagonal(F, 0)
    G1 = u.T @ du; J = F * G1
    ref = u @ ((J + J.T) * s[None, :]) @ vt + (np.eye(n) - u @ u.T) @ du @ np.diag(1 / s) @ vt
    print(f"(c) random dU: sector {q}: device vs numpy rel diff {np.abs(gR[q] - ref).max() / np.abs(ref).max():.1e}  (|ref| {np.abs(ref).max():.2e}, max |F| {np.abs(F).max():.2e}, min |d| offdiag {np.abs(d + np.eye(n) * 1e300).min():.2e})")
print("---- device GEMMs on this data")
G1d = blocks(U.multiply(dUr, True, False))
P = blocks(U @ S)
R = blocks((U @ S) @ Vd)
for q in bU:
    u, s, vt, du = bU[q], np.diag(bS[q]), bV[q], bdUr[q]
    print(f"sector {q}: U^T dU device vs numpy {np.abs(G1d[q] - u.T @ du).max():.2e}; U S {np.abs(P[q] - u * s).max():.2e}; U S Vd {np.abs(R[q] - (u * s) @ vt).max():.2e}")
print("---- dumped intermediates of the LAST tSVD_backward call (random dU)")
gR = blocks(M.tSVD_backward(U, S, Vd, dUr, None, None))
st = S.structure()
def load(name):
    d = np.fromfile(f"gpurun_out/dbg_{name}.bin")
    return {tuple(q): d[o:o + r * c].reshape((r, c), order="F") for q, r, c, o in zip(*st)}
Mc, G1f, G2f, FSf, dFSf = load("Mcore"), load("G1"), load("G2"), load("FS"), load("dFS")
for q in bU:
    u, s, vt, du = bU[q], np.diag(bS[q]), bV[q], bdUr[q]
    n = len(s)
    d = s[None, :] ** 2 - s[:, None] ** 2
    F = np.where(np.abs(d) < 1e-12, d / (d * d + 1e-12), 1 / np.where(d == 0, 1, d)); np.fill_diagonal(F, 0)
    G1 = u.T @ du; J = F * G1
    ref = (J + J.T) * s[None, :] - G1 / s[None, :]
    bad = np.argwhere(np.abs(Mc[q] - ref) > 1e-9 * np.abs(ref).max())
    print(f"sector {q}: G1 dump vs numpy {np.abs(G1f[q] - G1).max():.1e}; |G2| {np.abs(G2f[q]).max():.1e}; |dFS| {np.abs(dFSf[q]).max():.1e}; FS diag ok {np.abs(np.diag(FSf[q]) - s).max():.1e}; offdiag FS {np.abs(FSf[q] - np.diag(np.diag(FSf[q]))).max():.1e}; Mcore diff {np.abs(Mc[q] - ref).max():.2e} at {len(bad)} entries, e.g. {bad[:4].tolist()}")
