"""D=4 chi=64 CTM, device vs oracle step by step (corner spectra); XB_SVD_TRUNC selects the tSVD path."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import xped_b200 as xb
import xped_oracle as xo
from test_gpu_bench_shapes import oracle_su2_ctm, sv_spectra
D, CHI, STEPS = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ctx = xb.Context([("SU2", 0, 0)])
sym, phys, aux, oA, octm = oracle_su2_ctm(D, CHI)
entries = [(tuple(q), d) for q, _, d in aux.data]
a, p = ctx.basis(entries), ctx.basis([((2,), 1)])
A = ctx.tensor([a, a], [a, a, p]).setRandom(0)
ctm = xb.CTM(A, CHI)
for step in range(STEPS):
    # projector spectra of this step's first move, before the move
    gM = ctm.contractCorner(3) @ ctm.contractCorner(0)
    oM = octm.contractCorner(3) @ octm.contractCorner(0)
    gs, os_ = sv_spectra(gM.structure(), gM.download()), sv_spectra(oM.structure(), oM.data)
    alls = np.sort(np.concatenate(list(os_.values())))[::-1]
    top = alls[0]
    dm = max(np.abs(gs[q] - os_[q]).max() for q in os_) / top
    gap = (alls[CHI - 1] - alls[CHI]) / alls[CHI - 1] if len(alls) > CHI else None
    print(f"step {step}: spectrum of M = Q4*Q1 differs by {dm:.2e}; relative gap at the cut {gap}; sigma around the cut {alls[CHI-3:CHI+3] / top}")
    ctm.grow_all(); octm.grow_all()
    for name in ["C1", "C2", "C3", "C4", "T1"]:
        g, o = ctm.get(name), octm.tensors()[name]
        ok = g.structure() == o.structure()
        if ok:
            gs, os_ = sv_spectra(g.structure(), g.download()), sv_spectra(o.structure(), o.data)
            top = max(v[0] for v in os_.values())
            w = max(np.abs(gs[q] - os_[q]).max() for q in os_) / top
        print(f"   after step {step}: {name} structure equal {ok}, spectra differ by {w if ok else None}")
