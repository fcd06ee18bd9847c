"""Truncated tSVD vs the full Jacobi path on the CTM projector matrices (D=4 chi=64 and, optionally, D=6 chi=144): timing and agreement.
Usage: python tools/rsvd_check.py D chi steps   (XB_SVD_TRUNC=0/1 selects the path; XB_SVD_PROF=1 prints the iteration log)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xped_b200 as xb

D, CHI, STEPS = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ctx = xb.Context([("SU2", 0, 0)])
aux = {4: [(1, 2), (2, 2)], 6: [(1, 3), (2, 3)], 3: [(1, 1), (2, 1)], 2: [(1, 1), (2, 1)]}[D] if D != 2 else [(2, 1)]
a, phys = ctx.basis(aux), ctx.basis([(2, 1)])
A = ctx.tensor([a, a], [a, a, phys]).setRandom(0)
ctm = xb.CTM(A, CHI)
_cuda = None
if os.environ.get("XB_CHECK_NCU_LAST"):   # ncu --profile-from-start off: capture the last step only
    import ctypes
    _cuda = ctypes.CDLL("libcuda.so.1")
for step in range(STEPS):
    if _cuda is not None and step == STEPS - 1:
        ctx.sync(); _cuda.cuProfilerStart()
    ctx.sync(); ctx.profile(True); ctx.profile_read()
    ctx.timer_start()
    ctm.grow_all()
    ms = ctx.timer_stop()
    prof = ctx.profile_read(); ctx.profile(False)
    print(f"step {step}: {ms:.2f} ms; prof {dict((k, (round(v[0], 2), v[1])) for k, v in prof.items())}; truncated calls / fallbacks / iterations so far "
          f"{ctx.get_int('svd.truncated_calls')} / {ctx.get_int('svd.truncated_fallbacks')} / {ctx.get_int('svd.truncated_iterations')}", flush=True)
    if _cuda is not None and step == STEPS - 1:
        ctx.sync(); _cuda.cuProfilerStop()
# gauge-invariant fingerprint of the environment: singular values of C1 and T1 (per sector)
Q = ctm.contractCorner(0)
Q4 = ctm.contractCorner(3)
M = Q4 @ Q
t0 = time.perf_counter()
U, S, Vd, info = M.tSVD(CHI, 1e-10, False, return_info=True)
ctx.sync()
print(f"tSVD of Q4*Q1: {(time.perf_counter() - t0) * 1e3:.2f} ms wall; info {info}; S structure {S.structure()[1]}")
s = np.sort(S.download())[::-1]
s = s[s > 0]
print("sigma head", s[:6], "tail", s[-4:])
rec = (U @ S) @ Vd
print("||U S Vd||_F", rec.norm(), " ||M||_F", M.norm())
np.save(f"gpurun_out/rsvd_sigma_D{D}_trunc{os.environ.get('XB_SVD_TRUNC', '1')}.npy", S.download())
c1 = ctm.get("C1"); 
print("C1 maxnorm", c1.maxNorm(), "sqnorm", c1.squaredNorm())
