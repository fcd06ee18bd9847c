"""Known answers of the reference's own tests/physics.cpp through teigh: the model Hamiltonians are rebuilt densely (tests/physics_models.py),
cut into their charge sectors and stored as Tensor<double,1,1,Symmetry>; `Es.block(k)(0,0)` of `teigh()` must be the golden energy the
reference asserts, for the oracle (CPU) and for xb_teigh (GPU).  Pins teigh + the sorted sector order ("eigh: tests/physics.cpp
12 golden energies")."""
import numpy as np
import pytest

import xped_oracle as xo
from helpers import SPECS
from physics_models import GOLDEN, hubbard_chain, sector_blocks, spin_chain

CASES = {
    "spin_u1": ("U1", spin_chain),
    "hubbard_u1xu1": ("U1xU1", hubbard_chain),
    "spin_u0": ("U0", spin_chain),
}


def build_oracle(case):
    name, model = CASES[case]
    osym = xo.Symmetry.from_spec(SPECS[name])
    H, charges = model()
    if name == "U0":
        charges = [() for _ in charges]
    blocks = sector_blocks(H, charges)
    qs = sorted(blocks)
    entries = [(q, blocks[q].shape[0]) for q in qs]
    B = xo.Qbasis.from_list(osym, entries)
    t = xo.Tensor(osym, [B], [B])
    assert [tuple(q) for q in t.sector] == [tuple(q) for q in qs]      # Qbasis::sort order == the order Es.block(k) counts in
    for i, q in enumerate(t.sector):
        t.block(i)[...] = blocks[tuple(q)]
    return name, entries, t


@pytest.mark.parametrize("case", sorted(CASES))
def test_oracle_teigh_reproduces_reference_energies(case):
    _, _, t = build_oracle(case)
    D, V = t.teigh()
    e, k = GOLDEN[case]
    assert abs(D.block(k)[0, 0] - e) < 1e-11
    for i in range(len(t.sector)):     # V D V^T = H per block
        assert np.abs(V.block(i) @ D.block(i) @ V.block(i).T - t.block(i)).max() < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(CASES))
def test_gpu_teigh_reproduces_reference_energies(case):
    import xped_b200 as xb
    name, entries, t = build_oracle(case)
    ctx = xb.Context(SPECS[name])
    try:
        B = ctx.basis(entries)
        g = ctx.tensor([B], [B])
        assert g.structure() == t.structure()
        g.upload(t.data)
        gD, gV = g.teigh()
        oD, _ = t.teigh()
        e, k = GOLDEN[case]
        st = gD.structure()
        d = gD.download()
        n = st[1][k]
        blk = d[st[3][k]:st[3][k] + n * n].reshape((n, n), order="F")
        assert abs(blk[0, 0] - e) < 1e-11
        assert gD.structure() == oD.structure()
        assert np.linalg.norm(d - oD.data) / np.linalg.norm(oD.data) < 1e-12
        rec = (gV @ gD).multiply(gV, False, True)
        assert np.linalg.norm(rec.download() - t.data) / np.linalg.norm(t.data) < 1e-12
    finally:
        ctx.close()
