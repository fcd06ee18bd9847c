"""`contract` checked WITHOUT a GPU, end to end through the host logic of the device path: `xb_contract_plan_emulate` derives the leg permutations like
Context::contract (get_permutations, include/Xped/Util/Constfct.hpp:55-109), builds the three permute plans and the grouped-product lists, and walks all of them
on the host the way the kernels address device memory.  Compared with the oracle's Tensor::contract (include/Xped/Core/Tensor.hpp:328-351):
 * label patterns of tests/tensor_tests.hpp on four symmetries,
 * EVERY contraction of a CTM run -- the oracle's CTM driver (CTM.cpp:158-214, 585-603, 753-779, 1358-1738 label arrays) is executed with Tensor.contract
   wrapped, so that each of its calls (initialisation, four corners per move, projector products, renormalisation) is also lowered and walked on the host."""
import numpy as np
import pytest

import xped_oracle as xo
from xped_oracle.ctm import CTM as OCTM, init_basis
from xped_oracle.rng import uniform_fill
from helpers import SPECS, obasis, rand_entries
from xped_b200.sharded import PlanningContext, contract_plan_emulate


def product_basis(ctx, basis):
    """the oracle basis in the product: a conjugated basis (legs that a shifted permute moved between domain and codomain carry Qbasis::conj, i.e. dual
    leaves in the fusion trees and Frobenius-Schur signs in the turns) is rebuilt from its un-conjugated sectors and conjugated again"""
    if basis.CONJ:
        return ctx.basis([(tuple(basis.sym.conj(q)), d) for q, _, d in basis.data]).conj()
    return ctx.basis([(tuple(q), d) for q, _, d in basis.data])


def emulate(ctx, a, la, b, lb, r, n_c, table_form=True):
    g = lambda bases: [product_basis(ctx, x) for x in bases]
    return contract_plan_emulate(ctx, g(a.uncoupled_domain), g(a.uncoupled_codomain), la, g(b.uncoupled_domain), g(b.uncoupled_codomain), lb, r, a.data, b.data,
                                 n_c, table_form)


def assert_close(out, ref, what):
    assert out.shape == ref.data.shape, what
    if out.size == 0:
        return
    assert not np.isnan(out).any(), what
    scale = max(float(np.abs(ref.data).max()), 1.0)
    err = float(np.abs(out - ref.data).max()) if out.size else 0.0
    assert err <= 1e-12 * scale, f"{what}: max abs error {err:.3e} (scale {scale:.3e})"


@pytest.mark.parametrize("name", ["U1", "SU2", "fU1", "fU1xSU2"])
def test_label_patterns(name):
    ctx, osym = PlanningContext(SPECS[name]), xo.Symmetry.from_spec(SPECS[name])
    rng = np.random.default_rng(7)
    B, Cb, D = rand_entries(SPECS[name], 6, rng), rand_entries(SPECS[name], 5, rng), rand_entries(SPECS[name], 4, rng)
    mk = lambda dom, cod, seed: _filled(xo.Tensor(osym, [obasis(osym, e) for e in dom], [obasis(osym, e) for e in cod]), seed)
    t, s = mk([B, Cb], [Cb, D], 1), mk([Cb, D], [D, B], 2)   # the tensors and label patterns of tests/test_gpu_parity.py::test_multiply_adjoint_contract
    for la, lb, r in [([-1, -2, 2, 1], [2, 1, -3, -4], 1), ([1, -1, 2, -2], [2, -3, -4, 1], 1), ([-3, -1, -4, 1], [-2, 1, -5, -6], 4), ([-1, -2, 1, 2], [1, 2, -3, -4], 2)]:
        ref = t.contract(la, s, lb, r)
        for tf in (True, False):
            assert_close(emulate(ctx, t, la, s, lb, r, ref.plainSize(), tf), ref, f"{name} contract<{la},{lb},{r}> table_form={tf}")
    ctx.close()


def _filled(t, seed):
    t.data = uniform_fill(seed, t.plainSize())
    return t


@pytest.mark.parametrize("D,chi", [(2, 5), (3, 7)])
def test_every_contraction_of_a_ctm_run(D, chi, monkeypatch):
    ctx, osym = PlanningContext(SPECS["SU2"]), xo.Symmetry.SU2()
    phys = xo.Qbasis.from_list(osym, [((2,), 1)])
    aux = init_basis(osym, D, phys)
    A = _filled(xo.Tensor(osym, [aux, aux], [aux, aux, phys]), 0)
    calls = {"n": 0, "patterns": set()}
    orig = xo.Tensor.contract

    def checked(self, la, other, lb, r):
        ref = orig(self, la, other, lb, r)
        if ref.plainSize() > 0:
            assert_close(emulate(ctx, self, list(la), other, list(lb), r, ref.plainSize()), ref, f"CTM contract<{list(la)},{list(lb)},{r}>")
            calls["n"] += 1
            calls["patterns"].add((tuple(la), tuple(lb), r))
        return ref

    monkeypatch.setattr(xo.Tensor, "contract", checked)
    ctm = OCTM(A, chi)      # CTM_INIT::FROM_A: the init contractions
    ctm.grow_all()          # four moves: corners, projectors, renormalisation
    ctm.grow_all()          # second step: environment bases have grown to chi
    assert calls["n"] >= 100 and len(calls["patterns"]) >= 30, (calls["n"], len(calls["patterns"]))
    ctx.close()
