import os, sys
sys.path.insert(0, "/root/repo")
import xped_b200 as xb
ctx = xb.Context([("SU2", 0, 0)])
a, phys = ctx.basis([(1, 2), (2, 2)]), ctx.basis([(2, 1)])
A = ctx.tensor([a, a], [a, a, phys]).setRandom(0)
ctm = xb.CTM(A, 64)
for _ in range(6): ctm.grow_all()
ctx.sync()
print("==== step", file=sys.stderr, flush=True)
ctx.profile(True); ctx.profile_read()
ctm.move(0)
print(ctx.profile_read(), file=sys.stderr)
