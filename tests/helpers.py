"""Shared fixtures: build the same bases / tensors in the oracle (CPU) and the product (GPU, via the C ABI)."""
import numpy as np

import xped_oracle as xo
from xped_oracle.rng import uniform_fill

SPECS = {
    "U1": [("U1", 0, 0)],
    "SU2": [("SU2", 0, 0)],
    "U0": [("U0", 0, 0)],
    "Z3": [("ZN", 3, 0)],
    "fU1": [("U1", 0, 1)],
    "U1xU1": [("U1", 0, 1), ("U1", 0, 0)],
    "fU1xSU2": [("U1", 0, 1), ("SU2", 0, 0)],
}


def rand_entries(spec, size, rng, max_sector=4):
    """[(q, dim)] like Qbasis::setRandom (src/Core/Qbasis.cpp:93-103) with small charges; insertion order kept."""
    sym = xo.Symmetry.from_spec(spec)
    out, full = [], 0
    guard = 0
    while full < size and guard < 1000:
        guard += 1
        q = []
        for kind, N, _ in spec:
            if kind == "SU2":
                q.append(int(rng.integers(1, 5)))
            elif kind == "U1":
                q.append(int(rng.integers(-2, 3)))
            elif kind == "ZN":
                q.append(int(rng.integers(0, N)))
        q = tuple(q)
        d = int(rng.integers(1, min(max_sector, size - full) + 1))
        if full + sym.degeneracy(q) * d <= size:
            out.append((q, d))
            full += sym.degeneracy(q) * d
    return out


def obasis(osym, entries):
    return xo.Qbasis.from_list(osym, entries)


def pair(ctx, osym, dom, cod, seed):
    """Same tensor in both worlds, filled with the shared counter-based generator."""
    gt = ctx.tensor([ctx.basis(e) for e in dom], [ctx.basis(e) for e in cod]).setRandom(seed)
    ot = xo.Tensor(osym, [obasis(osym, e) for e in dom], [obasis(osym, e) for e in cod])
    ot.data = uniform_fill(seed, ot.plainSize())
    return gt, ot


def rel_err(g, o):
    """relative Frobenius error of the raw storage images"""
    a, b = g.download(), o.data
    assert a.shape == b.shape
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def assert_same(g, o, tol=1e-12, what=""):
    assert g.structure() == o.structure(), f"{what}: block structure differs"
    e = rel_err(g, o)
    assert e < tol, f"{what}: relative Frobenius error {e:.3e} >= {tol}"


def mps_chain_bases(osym, spec, L, loc_entries, chi_cap):
    """Auxiliary bases of an open chain: aux[0] = vacuum, aux[l+1] = (aux[l] x loc) with every sector dimension capped at
    `chi_cap` (a small stand-in for Mps::gen_maxBasis, src/MPS/Mps.cpp:68-179); returned as [(q, dim)] lists."""
    loc = obasis(osym, loc_entries)
    aux = [[(osym.qvacuum(), 1)]]
    for _ in range(L):
        comb = obasis(osym, aux[-1]).combine(loc).forgetHistory()
        aux.append([(q, min(d, chi_cap)) for q, _, d in comb.data])
    aux[-1] = [(osym.qvacuum(), 1)]   # target quantum number Qtot = vacuum, one state (Mps constructor: Qtot)
    return aux
