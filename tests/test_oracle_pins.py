"""Pins the CPU oracle against the reference's own tests (no GPU, no product code).

Restates tests/tree.cpp + tests/tree_tests.hpp:2-144 (recoupling moves are self-inverse),
tests/tensor.cpp:75-225 + tests/tensor_tests.hpp:2-199,306-333 (permute == dense shuffle through
explicit Clebsch-Gordan tensors; norms) and the analytic SU(2) known-answer checks
tests/tensor.cpp:230-383 of /root/reference.
"""
import itertools
import math

import numpy as np
import pytest

from xped_oracle import FusionTree, Qbasis, Symmetry, Tensor, decompose, treepair_permute, turn, wigner_3j_2j, wigner_6j_2j


def rand_basis(sym, size, rng, max_sector=5):
    """Qbasis::setRandom (Qbasis.cpp:93-103) with a seeded generator and small charges."""
    b = Qbasis(sym)
    while b.fullDim() < size:
        q = []
        for f in sym.active:
            q.append({"SU2": lambda: int(rng.integers(1, 5)), "U1": lambda: int(rng.integers(-3, 4)),
                      "ZN": lambda: int(rng.integers(0, getattr(f, "N", 2)))}[f.name]())
        q = tuple(q)
        d = int(rng.integers(1, min(max_sector, size - b.fullDim()) + 1))
        if b.fullDim() + sym.degeneracy(q) * d <= size:
            b.push_back(q, d)
    return b


SYMS = {
    "SU2": (Symmetry.SU2, 7),
    "U1": (Symmetry.U1, 6),
    "U0": (Symmetry.U0, 3),
    "fU1": (lambda: Symmetry.U1(fermionic=True), 5),
    "Z3": (lambda: Symmetry.ZN(3), 5),
    "fU1xSU2": (lambda: Symmetry.from_spec([("U1", 0, 1), ("SU2", 0, 0)]), 7),
}


def test_wigner_symbols_match_sympy():
    sympy = pytest.importorskip("sympy")
    from sympy import Rational
    from sympy.physics.wigner import wigner_3j, wigner_6j
    rng = np.random.default_rng(5)
    for _ in range(150):
        a, b, c, d, e, f = (int(x) for x in rng.integers(0, 9, 6))
        try:
            ref = float(wigner_6j(*[Rational(x, 2) for x in (a, b, c, d, e, f)]))
        except ValueError:
            ref = 0.0
        assert abs(ref - wigner_6j_2j(a, b, c, d, e, f)) < 1e-14
    for _ in range(150):
        a, b, c = (int(x) for x in rng.integers(0, 7, 3))
        ma = int(rng.choice(range(-a, a + 1, 2))); mb = int(rng.choice(range(-b, b + 1, 2))); mc = -(ma + mb)
        try:
            ref = float(wigner_3j(*[Rational(x, 2) for x in (a, b, c, ma, mb, mc)]))
        except ValueError:
            ref = 0.0
        assert abs(ref - wigner_3j_2j(a, b, c, ma, mb, mc)) < 1e-14


def test_decompose_reproduces_permutation():
    for n in range(1, 6):
        for p in itertools.permutations(range(n)):
            a = list(range(n))
            for pos in decompose(p):
                a[pos], a[pos + 1] = a[pos + 1], a[pos]
            assert a == list(p)


def _apply(maps, fn):
    out = {}
    for t, c in maps.items():
        for t2, c2 in fn(t).items():
            out[t2] = out.get(t2, 0.0) + c * c2
    return {k: v for k, v in out.items() if abs(v) > 1e-12}


@pytest.mark.parametrize("name", ["SU2", "U1", "U0", "fU1"])
def test_tree_moves_are_self_inverse(name):
    """tree_tests.hpp:2-144."""
    mk, size = SYMS[name]
    sym = mk()
    rng = np.random.default_rng(11)
    bases = [rand_basis(sym, size, rng) for _ in range(4)]
    dom = bases[0].combine(bases[1]).combine(bases[2]).combine(bases[3])
    trees = [t for q in dom.qs() for t in dom.tree(q)]
    for tree in trees[:: max(1, len(trees) // 25)]:
        for pos in range(3):                                   # swap o swap = 1
            back = _apply(tree.swap(sym, pos), lambda t: t.swap(sym, pos))
            assert list(back) == [tree] and abs(back[tree] - 1) < 1e-12
        for p in itertools.permutations(range(4)):            # permute(p) o permute(p^-1) = 1
            pinv = [p.index(i) for i in range(4)]
            back = _apply(tree.permute(sym, p), lambda t: t.permute(sym, pinv))
            assert list(back) == [tree] and abs(back[tree] - 1) < 1e-12
    d2, c2 = bases[0].combine(bases[1]), bases[2].combine(bases[3])
    pairs = [(a, b) for q in d2.qs() if c2.IS_PRESENT(q) for a in d2.tree(q) for b in c2.tree(q)]
    for a, b in pairs[:: max(1, len(pairs) // 15)]:
        for s in (-2, -1, 1, 2):                               # turn<s> o turn<-s> = 1
            out = {}
            for (x, y), c in turn(sym, s, a, b).items():
                for k, c2_ in turn(sym, -s, x, y).items():
                    out[k] = out.get(k, 0.0) + c * c2_
            assert list(out) == [(a, b)] and abs(out[(a, b)] - 1) < 1e-12
            for p in [(0, 1, 2, 3), (1, 0, 3, 2), (2, 0, 3, 1), (3, 2, 1, 0)]:   # treepair::permute<s>(p) o permute<-s>(p^-1) = 1
                pinv = [p.index(i) for i in range(4)]
                out = {}
                for (x, y), c in treepair_permute(sym, s, a, b, p).items():
                    for k, c2_ in treepair_permute(sym, -s, x, y, pinv).items():
                        out[k] = out.get(k, 0.0) + c * c2_
                out = {k: v for k, v in out.items() if abs(v) > 1e-12}
                assert list(out) == [(a, b)] and abs(out[(a, b)] - 1) < 1e-12


@pytest.mark.parametrize("name", list(SYMS))
def test_permute_equals_dense_shuffle(name):
    """tensor_tests.hpp:2-48: plainTensor(permute(T)) == shuffle(plainTensor(T)), all 24 x 5 (the reference samples 8 x 5)."""
    mk, size = SYMS[name]
    sym = mk()
    rng = np.random.default_rng(3)
    B, C = rand_basis(sym, size, rng), rand_basis(sym, size, rng)
    t = Tensor(sym, [B, C], [B, C]).setRandom(rng)
    tp = t.plainTensor()
    assert abs(t.squaredNorm() - np.sum(tp ** 2)) < 1e-10 * max(1.0, np.sum(tp ** 2))     # tensor_tests.hpp:306-333
    perms = list(itertools.permutations(range(4)))
    if sym.fermionic:
        pytest.skip("fermionic dense pin not asserted by the reference either (no fermionic case in tests/tensor.cpp)")
    for shift in (-2, -1, 0, 1, 2):
        for p in perms:
            r = t.permute(shift, p)
            assert np.abs(r.plainTensor() - tp.transpose(p)).max() < 1e-12


@pytest.mark.parametrize("name", ["SU2", "U1", "U0"])
def test_permute_within_domain(name):
    """tensor_tests.hpp:112-199: Tensor<4,0>, intra-domain permutations."""
    mk, size = SYMS[name]
    sym = mk()
    rng = np.random.default_rng(4)
    B = rand_basis(sym, min(size, 5), rng)
    t = Tensor(sym, [B, B, B, B], []).setRandom(rng)
    if t.plainSize() == 0:
        pytest.skip("no vacuum sector drawn")
    tp = t.plainTensor()
    for p in itertools.permutations(range(4)):
        assert np.abs(t.permute(0, p).plainTensor() - tp.transpose(p)).max() < 1e-12


@pytest.mark.parametrize("name", ["SU2", "U1", "fU1", "fU1xSU2"])
def test_permute_adj_is_transpose(name):
    """permute_adj (Tensor.cpp:258-333) is the transpose of permute on the raw block data (what ADTensor.hpp:171-185 needs)."""
    mk, size = SYMS[name]
    sym = mk()
    rng = np.random.default_rng(6)
    B, C = rand_basis(sym, size, rng), rand_basis(sym, size, rng)
    t = Tensor(sym, [B, C], [B, C]).setRandom(rng)
    for shift in (-2, -1, 0, 1, 2):
        for p in list(itertools.permutations(range(4)))[::3]:
            r = t.permute(shift, p)
            s = r.like().setRandom(rng)
            back = s.permute_adj(shift, p)
            assert back.structure() == t.structure()
            assert abs(np.dot(r.data, s.data) - np.dot(t.data, back.data)) < 1e-11


@pytest.mark.parametrize("twoS1", [1, 2, 3, 4, 7, 10])
def test_su2_spin_operator_known_answers(twoS1):
    """tests/tensor.cpp:230-383 (analytic)."""
    sym = Symmetry.SU2()
    S1 = 0.5 * twoS1
    n = twoS1 + 1
    qb = lambda q: Qbasis.from_list(sym, [(q, 1)])
    B1, C, one, three, five = qb(n), qb(3), qb(1), qb(3), qb(5)
    pv = [np.zeros((n, n)) for _ in range(3)]
    for m_ in range(1, n):
        m = S1 - m_
        pv[0][m_ - 1, m_] = math.sqrt(S1 * (S1 + 1) - m * (m + 1))
    pv[0] *= -math.sqrt(0.5)
    for m_ in range(n):
        pv[1][m_, m_] = S1 - m_
    for m_ in range(n - 1):
        m = S1 - m_
        pv[2][m_ + 1, m_] = math.sqrt(S1 * (S1 + 1) - m * (m - 1))
    pv[2] *= math.sqrt(0.5)

    s1 = Tensor(sym, [B1, C], [B1]).setConstant(math.sqrt(S1 * (S1 + 1.0)))
    tplain = s1.adjoint().plainTensor()
    for k in range(3):
        assert np.abs(tplain[:, :, k] - pv[k]).max() < 1e-12

    def coupled(mid):
        cpl = Tensor(sym, [C, mid], [C]).setConstant(1.0)
        return (s1.adjoint().permute(-1, (0, 1, 2)) @ cpl.permute(+1, (0, 2, 1))).permute(0, (0, 3, 1, 2)) @ s1

    check1 = coupled(one).plainTensor()
    assert np.abs(check1[:, 0, :] - S1 * (S1 + 1.0) * np.eye(n)).max() < 1e-12
    check3 = coupled(three).adjoint().plainTensor()
    for k in range(3):
        assert np.abs(check3[:, :, k] - math.sqrt(0.5) * pv[k]).max() < 1e-12
    check5 = coupled(five).adjoint().plainTensor()
    expect = [-math.sqrt(0.6) * pv[0] @ pv[0], -math.sqrt(0.3) * (pv[0] @ pv[1] + pv[1] @ pv[0]),
              -math.sqrt(0.1) * (pv[0] @ pv[2] + pv[2] @ pv[0] + 2 * pv[1] @ pv[1]),
              -math.sqrt(0.3) * (pv[2] @ pv[1] + pv[1] @ pv[2]), -math.sqrt(0.6) * pv[2] @ pv[2]]
    for k in range(5):
        assert np.abs(check5[:, :, k] - expect[k]).max() < 1e-12

    cpl1 = Tensor(sym, [C, one], [C]).setConstant(1.0)
    for twoS2 in (1, 2, 5):
        S2 = 0.5 * twoS2
        B2 = qb(twoS2 + 1)
        s2 = Tensor(sym, [B2, C], [B2]).setConstant(math.sqrt(S2 * (S2 + 1.0)))
        outer = (((s1.adjoint().permute(-1, (0, 1, 2)) @ cpl1.permute(+1, (0, 2, 1))).permute(-1, (0, 1, 3, 2)))
                 @ s2.permute(+1, (1, 2, 0))).permute(0, (0, 4, 2, 1, 3))
        for i, q in enumerate(outer.sector):
            assert outer.block(i).size == 1
            Stot = 0.5 * (q[0] - 1.0)
            assert abs(outer.block(i)[0, 0] - 0.5 * (Stot * (Stot + 1) - S1 * (S1 + 1.0) - S2 * (S2 + 1.0))) < 1e-12


@pytest.mark.parametrize("name", ["SU2", "U1", "U0", "fU1xSU2"])
def test_mps_contractions_equal_permute_formulation(name):
    """tests/mps_tests.hpp:19-33: contract_R / contract_L == the permute + operator* formulation."""
    from xped_oracle.mps import contract_L, contract_R
    make, size = SYMS[name]
    sym = make()
    rng = np.random.default_rng(23)
    inb, loc = rand_basis(sym, size + 4, rng), rand_basis(sym, 4, rng, 2)
    out = inb.combine(loc).forgetHistory()
    T1, T2 = Tensor(sym, [inb, loc], [out]), Tensor(sym, [inb, loc], [out])
    T1.setRandom(rng)
    T2.setRandom(rng)
    Bright = Tensor(sym, [out], [out])
    Bright.setRandom(rng)
    Bcheck = (T2 @ Bright).permute(1, (0, 1, 2)) @ T1.permute(1, (0, 1, 2)).adjoint()
    Bn = contract_R(Bright, T1, T2)
    assert Bn.structure() == Bcheck.structure()
    assert np.linalg.norm(Bn.data - Bcheck.data) <= 1e-12 * max(np.linalg.norm(Bcheck.data), 1.0)
    Bleft = Tensor(sym, [inb], [inb])
    Bleft.setRandom(rng)
    Bcheck = (T1.permute(1, (0, 1, 2)).adjoint() @ Bleft).permute(1, (1, 2, 0)) @ T2
    Bn = contract_L(Bleft, T1, T2)
    assert Bn.structure() == Bcheck.structure()
    assert np.linalg.norm(Bn.data - Bcheck.data) <= 1e-12 * max(np.linalg.norm(Bcheck.data), 1.0)


@pytest.mark.parametrize("name", ["SU2", "U1"])
def test_tsvd_backward_matches_finite_differences(name):
    """tests/ad.cpp:421-590 shape: the SVD reverse rule against central finite differences of a gauge-invariant
    function of the truncated factors, L(A) = <B, U_k S_k Vdag_k>_weighted (trace with sector degeneracies)."""
    make, size = SYMS[name]
    sym = make()
    rng = np.random.default_rng(5)
    b1, b2 = rand_basis(sym, size + 3, rng), rand_basis(sym, size, rng)
    out = b1.combine(b2).forgetHistory()
    A = Tensor(sym, [b1, b2], [out])                       # tall blocks: exercises the (1 - U U^dagger) correction
    A.setRandom(rng)
    B = A.like()
    B.setRandom(rng)
    keep = max(2, sum(min(A.block(i).shape) for i in range(len(A.sector))) // 2)

    def loss(T):
        U, S, Vd = T.tSVD(keep, 0.0, False)
        return ((U @ S) @ Vd @ B.adjoint()).trace()

    U, S, Vd = A.tSVD(keep, 0.0, False)
    # dL/dX of L = sum_q deg(q) tr(B_q^T U S Vd): dU = deg B Vd^T S, dS = deg diag(U^T B Vd^T), dVd = deg S U^T B
    dU, dS, dVd = U.like(), S.like(), Vd.like()
    for j, q in enumerate(S.sector):
        i = A.dict[q]
        d = sym.degeneracy(q)
        u, sm, vd, b = U.block(j), S.block(j), Vd.block(j), B.block(i)
        dU.block(j)[...] = d * b @ vd.T @ sm
        dS.block(j)[...] = d * np.diag(np.diag(u.T @ b @ vd.T))
        dVd.block(j)[...] = d * sm @ u.T @ b
    grad = A.tSVD_backward(dU, dS, dVd)
    eps = 1e-6
    idx = rng.choice(A.plainSize(), size=min(25, A.plainSize()), replace=False)
    for e in idx:
        Ap, Am = A.copy(), A.copy()
        Ap.data[e] += eps
        Am.data[e] -= eps
        fd = (loss(Ap) - loss(Am)) / (2 * eps)
        assert abs(fd - grad.data[e]) < 1e-6 * max(1.0, abs(fd)), (e, fd, grad.data[e])
