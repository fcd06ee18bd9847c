"""Mps::leftSweepStep / rightSweepStep / dot (src/MPS/Mps.cpp:181-229, src/MPS/MpsAlgebra.cpp:14-51).
CPU part: the oracle restatement obeys the invariants that define a sweep (norm conserved, site tensors become isometries).
GPU part: the device sweep against the oracle -- truncation weights, entropies and bond spectra (gauge invariant)."""
import numpy as np
import pytest

import xped_oracle as xo
from helpers import SPECS, mps_chain_bases, obasis, pair
from xped_oracle.mps import dot as odot
from xped_oracle.mps import left_sweep_step, right_sweep_step
from xped_oracle.rng import uniform_fill

LOC = {"SU2": [((2,), 1)], "U1": [((1,), 1), ((-1,), 1)], "U0": [((), 2)]}


def oracle_chain(osym, name, L, chi, seed0):
    aux = mps_chain_bases(osym, SPECS[name], L, LOC[name], chi)
    Ac = []
    for l in range(L):
        t = xo.Tensor(osym, [obasis(osym, aux[l]), obasis(osym, LOC[name])], [obasis(osym, aux[l + 1])])
        t.data = uniform_fill(seed0 + l, t.plainSize())
        Ac.append(t)
    return aux, Ac


@pytest.mark.parametrize("name", ["SU2", "U1", "U0"])
def test_oracle_sweeps_conserve_the_state(name):
    osym = xo.Symmetry.from_spec(SPECS[name])
    L = 6
    _, Ac = oracle_chain(osym, name, L, 6, 100)
    n0 = odot(Ac, Ac, "RIGHT")
    assert abs(odot(Ac, Ac, "LEFT") - n0) <= 1e-12 * n0
    for l in range(L - 1):
        right_sweep_step(Ac, l)
    assert abs(odot(Ac, Ac, "RIGHT") - n0) <= 1e-11 * n0   # the norm travels with S * Vdag into the next site
    for l in range(L - 1):                                    # left-isometries: A^dagger A = 1 on the right bond
        one = Ac[l].adjoint() @ Ac[l]
        eye = one.copy()
        eye.setIdentity()
        assert np.abs(one.data - eye.data).max() < 1e-12
    right_sweep_step(Ac, L - 1)                               # the last step drops S * Vdag (Mps.cpp:226): the state is normalised
    assert abs(odot(Ac, Ac, "RIGHT") - 1.0) <= 1e-11
    for l in range(L - 1, 0, -1):
        left_sweep_step(Ac, l)
    assert abs(odot(Ac, Ac, "LEFT") - 1.0) <= 1e-11
    assert abs(odot(Ac, Ac, "RIGHT") - 1.0) <= 1e-11


@pytest.mark.gpu
@pytest.mark.parametrize("name,chi", [("SU2", 8), ("U1", 6), ("U0", 7)])
def test_gpu_sweeps_match_the_oracle(name, chi):
    import xped_b200 as xb
    ctx = xb.Context(SPECS[name])
    try:
        osym = xo.Symmetry.from_spec(SPECS[name])
        L = 6
        aux, oAc = oracle_chain(osym, name, L, chi, 200)
        gAc = []
        for l in range(L):
            g, o = pair(ctx, osym, [aux[l], LOC[name]], [aux[l + 1]], 200 + l)
            assert np.array_equal(o.data, oAc[l].data)
            gAc.append(g)
        psi = xb.Mps(gAc, max_Nsv=5, eps_svd=1e-10)           # truncating sweep: max_Nsv below the bond dimensions
        n0 = odot(oAc, oAc, "RIGHT")
        assert abs(xb.dot(psi, psi, "RIGHT") - n0) <= 1e-12 * n0
        assert abs(xb.dot(psi, psi, "LEFT") - n0) <= 1e-12 * n0
        for sweep, order, gstep, ostep in (("right", range(L), psi.rightSweepStep, right_sweep_step),
                                           ("left", range(L - 1, -1, -1), psi.leftSweepStep, left_sweep_step)):
            for l in order:
                gstep(l)
                info = ostep(oAc, l, 5, 1e-10)
                assert psi.Ac[l].structure() == oAc[l].structure(), f"{sweep} sweep, site {l}"
                scale = max(info["truncWeight"], 1e-30, n0 * n0)
                assert abs(psi.truncWeight[l] - info["truncWeight"]) <= 1e-10 * scale, f"{sweep} sweep, site {l}: truncWeight"
                if l != L - 1:
                    assert abs(psi.S[l] - info["entropy"]) <= 1e-9 * max(abs(info["entropy"]), n0 * n0), f"{sweep} sweep, site {l}: entropy"
            ng, no = xb.dot(psi, psi, "RIGHT"), odot(oAc, oAc, "RIGHT")
            assert abs(ng - no) <= 1e-10 * no, f"norm after the {sweep} sweep"
    finally:
        ctx.close()


@pytest.mark.parametrize("name,L,Qtarget,Minit,Qinit", [("SU2", 8, (1,), 12, 3), ("SU2", 10, (3,), 200, 10), ("U1", 8, (0,), 10, 3), ("U1", 7, (1,), 6, 2)])
def test_aux_basis_generator_matches_the_oracle(name, L, Qtarget, Minit, Qinit):
    """Mps::gen_maxBasis / gen_auxBasis (src/MPS/Mps.cpp:56-127): the product's host logic (planning-only context, no GPU) against
    the oracle restatement, plus the defining properties (ends, reachability from both sides, caps)."""
    from xped_b200.sharded import PlanningContext
    from xped_b200.mps import gen_aux_bases
    from xped_oracle.mps import gen_aux_bases as ogen
    ctx = PlanningContext(SPECS[name])
    osym = xo.Symmetry.from_spec(SPECS[name])
    got = gen_aux_bases(ctx, L, LOC[name], Qtarget, Minit, Qinit)
    ref = ogen(osym, L, obasis(osym, LOC[name]), Qtarget, Minit, Qinit)
    assert got == ref
    assert got[0] == [(osym.qvacuum(), 1)] and got[L] == [(tuple(Qtarget), 1)]
    for l in range(L + 1):
        assert len(got[l]) <= max(Qinit, 1) and sum(d for _, d in got[l]) <= max(Minit, len(got[l]))
    ctx.close()


def test_aux_basis_known_answer_su2_singlet():
    """Spin-1/2 chain of four sites with total spin 0: the bond sectors are forced (q = 2S+1): 1 | 2 | 1,3 | 2 | 1."""
    from xped_b200.sharded import PlanningContext
    from xped_b200.mps import gen_aux_bases
    ctx = PlanningContext(SPECS["SU2"])
    got = gen_aux_bases(ctx, 4, LOC["SU2"], (1,), 100, 10)
    assert got == [[((1,), 1)], [((2,), 1)], [((1,), 1), ((3,), 1)], [((2,), 1)], [((1,), 1)]]
    ctx.close()
