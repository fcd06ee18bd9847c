"""computeRDM_h / computeRDM_v and the two-site average (src/PEPS/CTM.cpp:528-874 TRank == 2, src/PEPS/LinearAlgebra.cpp:57-60;
SURVEY.md section 8f rank 1).  The reference holds no test values for them ("parity unpinned"): the oracle is pinned by
invariants (unit trace, hermiticity, bounds of the Heisenberg bond energy, symmetry of a rotation-invariant site tensor), the
device path is compared with the oracle (the RDM is gauge invariant on the chi legs, so the comparison is element-wise)."""
import numpy as np
import pytest

import xped_oracle as xo
from helpers import SPECS, obasis, pair
from xped_oracle.ctm import CTM as OCTM
from xped_oracle.ctm import avg_one_site as oavg1
from xped_oracle.ctm import avg_two_site as oavg
from xped_oracle.ctm import rdm_one_site as ordm1
from xped_oracle.ctm import heisenberg_bond_SU2, init_basis
from xped_oracle.rng import uniform_fill


def _octm(name, D, chi, seed):
    osym = xo.Symmetry.from_spec(SPECS[name])
    phys_e = [((2,), 1)] if name == "SU2" else [((1,), 1), ((-1,), 1)]
    phys = obasis(osym, phys_e)
    aux = init_basis(osym, D, phys)
    A = xo.Tensor(osym, [aux, aux], [aux, aux, phys])
    A.data = uniform_fill(seed, A.plainSize())
    return osym, phys, phys_e, [(q, d) for q, _, d in aux.data], OCTM(A, chi)


@pytest.mark.parametrize("name", ["SU2", "U1"])
def test_oracle_rdm_invariants(name):
    osym, phys, _, _, c = _octm(name, 3, 9, 5)
    for _ in range(3):
        c.grow_all()
    Id = xo.Tensor.Identity(osym, [phys, phys], [phys, phys])
    for rho in (c.computeRDM_h(), c.computeRDM_v()):
        assert rho.rank == 2 and rho.corank == 2
        assert abs(oavg(rho, Id) - 1.0) < 1e-12                                   # unit trace (CTM.cpp:660-667)
        herm = np.linalg.norm((rho - rho.adjoint()).data) / np.linalg.norm((rho + rho.adjoint()).data)
        assert herm < 1e-3, herm                                                   # rho_h_hermitian_check: exact only for a converged environment
        rho1 = ordm1(rho)
        Id1 = xo.Tensor.Identity(osym, [phys], [phys])
        assert rho1.rank == 1 and rho1.corank == 1 and abs(oavg1(rho1, Id1) - 1.0) < 1e-12
    if name == "SU2":
        H = heisenberg_bond_SU2(osym, phys)
        for rho in (c.computeRDM_h(), c.computeRDM_v()):
            assert -0.75 - 1e-9 <= oavg(rho, H) <= 0.25 + 1e-9                     # spectrum of S.S on two spins 1/2


@pytest.mark.gpu
@pytest.mark.parametrize("name,D,chi", [("SU2", 3, 10), ("U1", 2, 8)])
def test_gpu_rdm_and_energy_match_the_oracle(name, D, chi):
    import xped_b200 as xb
    ctx = xb.Context(SPECS[name])
    try:
        osym, phys, phys_e, aux_e, o = _octm(name, D, chi, 51)
        gA, oA = pair(ctx, osym, [aux_e, aux_e], [aux_e, aux_e, phys_e], 51)
        assert np.array_equal(oA.data, o.A.data)
        g = xb.CTM(gA, chi)
        gO, oO = pair(ctx, osym, [phys_e, phys_e], [phys_e, phys_e], 77)           # a random two-site observable
        gO1, oO1 = pair(ctx, osym, [phys_e], [phys_e], 78)                           # a random one-site observable
        if name == "SU2":
            oH = heisenberg_bond_SU2(osym, phys)
            gH = ctx.tensor([ctx.basis(phys_e)] * 2, [ctx.basis(phys_e)] * 2)
            assert gH.structure() == oH.structure()
            gH.upload(oH.data)
        for step in range(3):
            for (grho, orho, what) in ((g.computeRDM_h(), o.computeRDM_h(), "rho_h"), (g.computeRDM_v(), o.computeRDM_v(), "rho_v")):
                assert grho.structure() == orho.structure(), what
                err = np.linalg.norm(grho.download() - orho.data) / np.linalg.norm(orho.data)
                assert err < 1e-9, f"step {step} {what}: {err:.2e}"
                assert abs(xb.avg_two_site(grho, gO) - oavg(orho, oO)) < 1e-9
                grho1, orho1 = xb.rdm_one_site(grho), ordm1(orho)
                assert grho1.structure() == orho1.structure()
                assert np.linalg.norm(grho1.download() - orho1.data) < 1e-9 * np.linalg.norm(orho1.data)
                assert abs(xb.avg_one_site(grho1, gO1) - oavg1(orho1, oO1)) < 1e-9
                if name == "SU2":
                    assert abs(xb.avg_two_site(grho, gH) - oavg(orho, oH)) < 1e-9
            g.grow_all()
            o.grow_all()
    finally:
        ctx.close()


# ---------------------------------------------------------------------------------------------------------------- known answer: spin-2 AKLT state
# Site tensor with aux legs = spin 1/2 (D = 2) and a spin-2 physical leg: Hom(1/2 x 1/2, 1/2 x 1/2 x 2) is ONE-dimensional (one sector, q = 3, one tree pair),
# and its only element is the AKLT map (four virtual spins 1/2 symmetrised into spin 2, singlets on the bonds).  Two neighbouring sites share a singlet of two
# of their eight virtual spins, so their total spin is at most 3 -- for ANY environment: <P_4> = 0 exactly, the AKLT Hamiltonian H = sum_bonds P_4 has energy 0.
# A wrong leg pairing in computeRDM_h / computeRDM_v, in the CTM moves that build the environment, or in avg() breaks this.
def _aklt(osym):
    aux = obasis(osym, [((2,), 1)])
    phys = obasis(osym, [((5,), 1)])
    A = xo.Tensor(osym, [aux, aux], [aux, aux, phys])
    assert A.plainSize() == 1, "the invariant space is one-dimensional"
    A.data = np.ones(1)
    return aux, phys, A


def _spin_projector(osym, phys, q_total, data_of=None):
    P = xo.Tensor(osym, [phys, phys], [phys, phys])
    secs, rows, cols, offs = P.structure()
    data = np.zeros(P.plainSize())
    for q, r, c, o in zip(secs, rows, cols, offs):
        assert r == 1 and c == 1
        if q == (q_total,):
            data[o] = 1.0
    P.data = data
    return P


def test_oracle_aklt_bond_has_no_spin4_component():
    osym = xo.Symmetry.from_spec(SPECS["SU2"])
    aux, phys, A = _aklt(osym)
    c = OCTM(A, 12)
    for _ in range(4):
        c.grow_all()
    for rho in (c.computeRDM_h(), c.computeRDM_v()):
        w = {S: oavg(rho, _spin_projector(osym, phys, 2 * S + 1)) for S in range(5)}
        assert abs(sum(w.values()) - 1.0) < 1e-12                       # the projectors resolve the identity
        assert all(v > -1e-12 for v in w.values())
        assert abs(w[4]) < 1e-12, w                                     # E_AKLT = <P_4> = 0
        assert w[0] > 1e-3 and w[3] > 1e-3                              # (the other sectors are populated)


@pytest.mark.gpu
def test_gpu_aklt_bond_has_no_spin4_component():
    import xped_b200 as xb
    ctx = xb.Context(SPECS["SU2"])
    try:
        osym = xo.Symmetry.from_spec(SPECS["SU2"])
        _, ophys, _ = _aklt(osym)
        a, p = ctx.basis([((2,), 1)]), ctx.basis([((5,), 1)])
        A = ctx.tensor([a, a], [a, a, p])
        assert A.plainSize() == 1
        A.upload(np.ones(1))
        g = xb.CTM(A, 12)
        for _ in range(4):
            g.grow_all()
        for rho in (g.computeRDM_h(), g.computeRDM_v()):
            w = {}
            for S in range(5):
                oP = _spin_projector(osym, ophys, 2 * S + 1)
                gP = ctx.tensor([p, p], [p, p])
                assert gP.structure() == oP.structure()
                gP.upload(oP.data)
                w[S] = xb.avg_two_site(rho, gP)
            assert abs(sum(w.values()) - 1.0) < 1e-12
            assert abs(w[4]) < 1e-12, w
            assert w[0] > 1e-3 and w[3] > 1e-3
    finally:
        ctx.close()
