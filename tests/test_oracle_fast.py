"""The optimised execution of the oracle used as the CPU baseline of bench.py (oracle/xped_oracle/fast.py + oracle/csrc/xped_oracle_kernels.c: cached
recoupling plans applied by a compiled OpenMP kernel) must reproduce the plain oracle: permute on every symmetry (fan-in > 1 and fermionic signs
included), operator*, contract, and a whole CTM step."""
import itertools

import numpy as np
import pytest

import xped_oracle as xo
from helpers import SPECS, obasis, rand_entries
from xped_oracle import fast
from xped_oracle.rng import uniform_fill


def rand_tensor(osym, spec, rng, seed, r=2, c=2):
    ents = [rand_entries(spec, 7, rng, 3) for _ in range(r + c)]
    t = xo.Tensor(osym, [obasis(osym, e) for e in ents[:r]], [obasis(osym, e) for e in ents[r:]])
    t.data = uniform_fill(seed, t.plainSize())
    return t


@pytest.mark.parametrize("name", ["SU2", "U1", "Z3", "fU1", "U1xU1", "fU1xSU2", "U0"])
def test_fast_permute_and_product_equal_plain_oracle(name):
    osym = xo.Symmetry.from_spec(SPECS[name])
    rng = np.random.default_rng(11)
    t = rand_tensor(osym, SPECS[name], rng, 5)
    cases = [(s, p) for s in (-2, -1, 0, 1, 2) for p in list(itertools.permutations(range(4)))[::5]]
    plain = [t.permute(s, p) for s, p in cases]
    prod_plain = t @ t.permute(0, (0, 1, 2, 3)) if t.codomain.dim() and t.uncoupled_codomain == t.uncoupled_domain else None
    fast.install()
    try:
        for (s, p), ref in zip(cases, plain):
            for _ in range(2):      # second call: cached plan
                got = t.permute(s, p)
                assert got.structure() == ref.structure()
                assert np.allclose(got.data, ref.data, rtol=0, atol=1e-14 * max(1.0, np.abs(ref.data).max())), (name, s, p)
        a, b = t.permute(0, (0, 1, 2, 3)), t.adjoint()
        got = a @ b
    finally:
        fast.uninstall()
    ref = a @ b
    assert got.structure() == ref.structure() and np.array_equal(got.data, ref.data)
    del prod_plain


def test_fast_ctm_step_equals_plain_oracle():
    from xped_oracle.ctm import CTM, init_basis
    sym = xo.Symmetry.SU2()
    phys = xo.Qbasis.from_list(sym, [((2,), 1)])
    aux = init_basis(sym, 3, phys)
    A = xo.Tensor(sym, [aux, aux], [aux, aux, phys])
    A.data = uniform_fill(0, A.plainSize())
    ref = CTM(A, 12)
    ref.grow_all()
    fast.install()
    try:
        got = CTM(A, 12)
        got.grow_all()
    finally:
        fast.uninstall()
    for name, o in ref.tensors().items():
        g = got.tensors()[name]
        assert g.structure() == o.structure()
        assert np.linalg.norm(g.data - o.data) <= 1e-10 * np.linalg.norm(o.data), name
