"""D=6, chi=144 SU(2) CTMRG step on the device (BASELINE.json north_star target workload): timing only."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xped_b200 as xb
ctx = xb.Context([("SU2", 0, 0)])
a, phys = ctx.basis([(1, 3), (2, 3)]), ctx.basis([(2, 1)])
A = ctx.tensor([a, a], [a, a, phys]).setRandom(0)
ctm = xb.CTM(A, 144)
for step in range(6):
    ctx.sync(); ctx.profile(True); ctx.profile_read()
    ctx.timer_start(); t0 = time.perf_counter()
    ctm.grow_all()
    ms = ctx.timer_stop()
    prof = ctx.profile_read(); ctx.profile(False)
    c1 = ctm.get("C1")
    print(f"step {step}: {ms:.1f} ms device ({(time.perf_counter()-t0)*1e3:.1f} wall), C1 {c1.structure()[1]}, prof {{k:(round(v[0],1),v[1]) for k,v in prof.items()}}".replace("{k:(round(v[0],1),v[1]) for k,v in prof.items()}", str({k:(round(v[0],1),v[1]) for k,v in prof.items()})), flush=True)
