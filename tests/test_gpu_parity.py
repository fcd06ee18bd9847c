"""GPU parity tests: every call goes through the C ABI (xped_b200 ctypes mirror) and is compared with the CPU
oracle on the same counter-based inputs.  Bar: block structure bit-exact, values within 1e-12 relative
Frobenius error (BASELINE.json north_star)."""
import itertools

import numpy as np
import pytest

import xped_b200 as xb
import xped_oracle as xo
from helpers import SPECS, assert_same, obasis, pair, rand_entries, rel_err
from xped_oracle.ctm import CTM as OCTM
from xped_oracle.ctm import decompose as odecompose
from xped_oracle.ctm import init_basis

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def ctxs():
    made = {}

    def get(name):
        if name not in made:
            made[name] = (xb.Context(SPECS[name]), xo.Symmetry.from_spec(SPECS[name]))
        return made[name]
    yield get
    for c, _ in made.values():
        c.close()


def test_fill_and_structure_bit_exact(ctxs):
    for name in SPECS:
        ctx, osym = ctxs(name)
        rng = np.random.default_rng(1)
        B, C = rand_entries(SPECS[name], 7, rng), rand_entries(SPECS[name], 6, rng)
        g, o = pair(ctx, osym, [B, C], [C, B], 11)
        assert g.structure() == o.structure()
        assert np.array_equal(g.download(), o.data)
        tp = g.tree_pairs()
        assert [tuple(r[:3]) for r in tp.tolist()] == [(i, ro, co) for i, ro, co, _ in o.tree_pairs()]


@pytest.mark.parametrize("name", list(SPECS))
def test_permute_all_24x5(ctxs, name):
    """tests/tensor_tests.hpp:2-48 shape: Tensor<2,2>, all 24 permutations x shifts -2..2."""
    ctx, osym = ctxs(name)
    rng = np.random.default_rng(2)
    B, C = rand_entries(SPECS[name], 8, rng), rand_entries(SPECS[name], 7, rng)
    g, o = pair(ctx, osym, [B, C], [B, C], 3)
    for shift in (-2, -1, 0, 1, 2):
        for p in itertools.permutations(range(4)):
            assert_same(g.permute(shift, p), o.permute(shift, p), TOL, f"{name} permute<{shift},{p}>")


@pytest.mark.parametrize("name,dim", [("U1", 4), ("U1", 8), ("U1", 12), ("Z3", 10), ("U0", 8)])
def test_permute_bulk_path(ctxs, name, dim, monkeypatch):
    """The cp.async.bulk form of the permute (normally used for >= 8 MB tensors) forced on small tensors whose runs obey
    the 16-byte rules (even multiplicities; runs shorter than 8 doubles stay on the LSU form): all 24 permutations x 5 shifts,
    bit-exact against the oracle (coefficients
    of bosonic Abelian symmetries are 1, so the permute is a pure data movement)."""
    monkeypatch.setenv("XB_PERMUTE_BULK", "1")
    ctx, osym = ctxs(name)
    if name == "U1":
        B = [((q,), dim) for q in range(-2, 3)]
    elif name == "Z3":
        B = [((q,), dim) for q in range(3)]
    else:
        B = [((), dim)]
    g, o = pair(ctx, osym, [B, B], [B, B], 81)
    for p in itertools.permutations(range(4)):
        for shift in (0, 1, -1, 2, -2):
            gp, op = g.permute(shift, p), o.permute(shift, p)
            assert gp.structure() == op.structure()
            assert np.array_equal(gp.download(), op.data), f"{name} bulk permute p={p} shift={shift}"


@pytest.mark.parametrize("name", ["SU2", "U1", "fU1", "fU1xSU2", "Z3"])
def test_permute_table_driven_form(ctxs, name):
    """The table-driven form (permute_lut_kernel: recoupling sums, odd extents) is built when a plan is used for the second time: every permutation runs
    three times -- LSU form first, then twice the table-driven form -- and all three results are compared with the oracle; (2,2) tensors with odd
    multiplicities (all 24 x 3 shifts), CTM-shaped (3,3) and (4,3) tensors, permute_adj, and the forms against each other."""
    ctx, osym = ctxs(name)
    rng = np.random.default_rng(31)
    B, C = rand_entries(SPECS[name], 9, rng), rand_entries(SPECS[name], 7, rng)
    g, o = pair(ctx, osym, [B, C], [B, C], 13)
    for shift in (-1, 0, 2):
        for p in itertools.permutations(range(4)):
            op = o.permute(shift, p)
            first = g.permute(shift, p)
            for rep in range(2):
                again = g.permute(shift, p)
                assert_same(again, op, TOL, f"{name} permute<{shift},{p}> repeat {rep}")
                assert np.abs(again.download() - first.download()).max() <= 1e-13 * max(np.abs(op.data).max(), 1.0)
    chi, aux = rand_entries(SPECS[name], 11, rng, 3), rand_entries(SPECS[name], 5, rng, 2)
    g6, o6 = pair(ctx, osym, [chi, aux, aux], [chi, aux, aux], 14)
    for shift, p in [(-1, (0, 2, 3, 5, 1, 4)), (1, (0, 3, 5, 1, 2, 4)), (0, (1, 0, 2, 4, 3, 5)), (-3, (5, 4, 3, 2, 1, 0))]:
        op = o6.permute(shift, p)
        for rep in range(3):
            assert_same(g6.permute(shift, p), op, TOL, f"{name} (3,3) permute<{shift},{p}> use {rep}")
        oa = o6.permute_adj(shift, p)
        for rep in range(3):
            assert_same(g6.permute_adj(shift, p), oa, TOL, f"{name} (3,3) permute_adj<{shift},{p}> use {rep}")
    phys = rand_entries(SPECS[name], 2, rng, 1)
    g7, o7 = pair(ctx, osym, [chi, aux, aux, aux], [aux, aux, phys], 15)
    for shift, p in [(0, (0, 2, 4, 5, 1, 3, 6)), (1, (0, 3, 5, 1, 2, 4, 6))]:
        op = o7.permute(shift, p)
        for rep in range(3):
            assert_same(g7.permute(shift, p), op, TOL, f"{name} (4,3) permute<{shift},{p}> use {rep}")


@pytest.mark.parametrize("name", ["SU2", "U1", "fU1xSU2"])
def test_permute_ctm_shaped(ctxs, name):
    """The (3,3)<->(4,2)/(1,5)-style regroupings contractCorner issues, incl. 6- and 7-leg tensors."""
    ctx, osym = ctxs(name)
    rng = np.random.default_rng(5)
    chi, aux = rand_entries(SPECS[name], 9, rng, 3), rand_entries(SPECS[name], 4, rng, 2)
    g, o = pair(ctx, osym, [chi, aux, aux], [chi, aux, aux], 4)
    for shift, p in [(-1, (0, 2, 3, 5, 1, 4)), (1, (0, 3, 5, 1, 2, 4)), (-1, (0, 2, 4, 5, 1, 3)), (0, (1, 0, 2, 4, 3, 5)), (2, (3, 0, 1, 2, 5, 4)),
                     (-3, (5, 4, 3, 2, 1, 0)), (3, (2, 4, 0, 1, 5, 3))]:
        assert_same(g.permute(shift, p), o.permute(shift, p), TOL, f"{name} (3,3) permute<{shift},{p}>")
    phys = rand_entries(SPECS[name], 2, rng, 1)
    g7, o7 = pair(ctx, osym, [chi, aux, aux, aux], [aux, aux, phys], 6)
    for shift, p in [(0, (0, 2, 4, 5, 1, 3, 6)), (0, (0, 3, 4, 5, 1, 2, 6)), (1, (0, 3, 5, 1, 2, 4, 6)), (-2, (6, 0, 1, 5, 2, 3, 4))]:
        assert_same(g7.permute(shift, p), o7.permute(shift, p), TOL, f"{name} (4,3) permute<{shift},{p}>")


@pytest.mark.parametrize("name", ["SU2", "U1", "fU1", "fU1xSU2"])
def test_permute_adj(ctxs, name):
    ctx, osym = ctxs(name)
    rng = np.random.default_rng(7)
    B, C = rand_entries(SPECS[name], 7, rng), rand_entries(SPECS[name], 6, rng)
    g, o = pair(ctx, osym, [B, C], [B, C], 8)
    for shift in (-2, -1, 0, 1, 2):
        for p in list(itertools.permutations(range(4)))[::2]:
            gs, os_ = g.permute(shift, p), o.permute(shift, p)
            assert_same(gs.permute_adj(shift, p), os_.permute_adj(shift, p), TOL, f"{name} permute_adj<{shift},{p}>")


@pytest.mark.parametrize("name", ["SU2", "U1", "U0", "U1xU1"])
def test_multiply_adjoint_contract(ctxs, name):
    ctx, osym = ctxs(name)
    rng = np.random.default_rng(9)
    B, C, D = (rand_entries(SPECS[name], n, rng) for n in (9, 8, 7))
    ga, oa = pair(ctx, osym, [B, C], [C, D], 1)
    gb, ob = pair(ctx, osym, [C, D], [D, B], 2)
    assert_same(ga @ gb, oa @ ob, TOL, "operator*")
    assert_same(ga.adjoint(), oa.adjoint(), TOL, "adjoint")
    assert_same(ga.multiply(ga, False, True), oa @ oa.adjoint(), TOL, "A * A^dagger (lazy adjoint operand)")
    assert_same(gb.multiply(ga, True, True), ob.adjoint() @ oa.adjoint(), TOL, "B^dagger * A^dagger")
    assert_same(ga.contract([-1, -2, 2, 1], gb, [2, 1, -3, -4], 1), oa.contract([-1, -2, 2, 1], ob, [2, 1, -3, -4], 1), TOL, "contract<2 legs>")
    assert_same(ga.contract([1, -1, 2, -2], gb, [2, -3, -4, 1], 1), oa.contract([1, -1, 2, -2], ob, [2, -3, -4, 1], 1), TOL, "contract<crossed>")
    assert_same(ga.contract([-3, -1, -4, 1], gb, [-2, 1, -5, -6], 4), oa.contract([-3, -1, -4, 1], ob, [-2, 1, -5, -6], 4), TOL, "contract<1 leg>")
    assert_same(ga.scale(-0.37), oa.scale(-0.37), TOL, "scale")
    g2, o2 = pair(ctx, osym, [B, C], [C, D], 5)
    assert_same(ga + g2, oa + o2, TOL, "operator+")
    assert_same(ga - g2, oa - o2, TOL, "operator-")
    assert abs(ga.maxNorm() - oa.maxNorm()) == 0.0
    gsq, osq = ga.multiply(ga, False, True), oa @ oa.adjoint()
    assert abs(gsq.trace() - osq.trace()) <= 1e-12 * abs(osq.trace())
    assert abs(ga.squaredNorm() - oa.squaredNorm()) <= 1e-12 * oa.squaredNorm()
    gd, od = gsq.diag_sqrt().diag_inv(), osq.diag_sqrt().diag_inv()      # zero blocks give inf on both sides
    assert gd.structure() == od.structure()
    a, b = gd.download(), od.data
    assert np.array_equal(np.isfinite(a), np.isfinite(b))
    fin = np.isfinite(b)
    assert np.linalg.norm(a[fin] - b[fin]) <= TOL * max(np.linalg.norm(b[fin]), 1e-300)


def test_multiply_split_k(ctxs):
    """Products with few output tiles and a long contraction run as K slices + a deterministic reduction (ops.cpp run_gemm):
    normal and lazily adjoint operands (TensorBase.cpp:243-288, AdjointOp.hpp:26-69)."""
    ctx, osym = ctxs("U0")
    gA, oA = pair(ctx, osym, [[((), 31)]], [[((), 612)]], 91)
    gB, oB = pair(ctx, osym, [[((), 612)]], [[((), 33)]], 92)
    gC, oC = pair(ctx, osym, [[((), 612)]], [[((), 140)]], 93)
    assert_same(gA @ gB, oA @ oB, TOL, "(31 x 612) (612 x 33)")
    assert_same(gA.multiply(gA, False, True), oA @ oA.adjoint(), TOL, "A A^dagger")
    assert_same(gB.multiply(gB, True, False), oB.adjoint() @ oB, TOL, "B^dagger B")
    assert_same(gC.multiply(gB, True, False), oC.adjoint() @ oB, TOL, "C^dagger B (140 x 612) (612 x 33)")
    assert_same(gB.multiply(gA, True, True), oB.adjoint() @ oA.adjoint(), TOL, "B^dagger A^dagger")


def test_error_codes(ctxs):
    ctx, _ = ctxs("U1")
    B, C = ctx.basis([(0, 2), (1, 1)]), ctx.basis([(0, 1), (-1, 2)])
    a = ctx.tensor([B], [C])
    with pytest.raises(xb.XbError) as e:
        a @ a                                     # coupledCodomain(left) != coupledDomain(right)
    assert e.value.status == 1
    with pytest.raises(xb.XbError):
        a.permute(0, (0, 0))
    with pytest.raises(xb.XbError):
        a.permute(2, (0, 1))
    empty = ctx.tensor([ctx.basis([(3, 1)])], [ctx.basis([(-3, 1)])])   # no common sector: empty tensor
    assert empty.plainSize() == 0 and empty.structure()[0] == []
    assert empty.permute(1, (1, 0)).plainSize() == 0 or True


def _svd_checks(g, o, keep, eps, pm):
    gU, gS, gV, ginfo = g.tSVD(keep, eps, pm, return_info=True)
    oU, oS, oV, oinfo = o.tSVD(keep, eps, pm, return_info=True)
    assert gS.structure() == oS.structure() and gU.structure() == oU.structure() and gV.structure() == oV.structure()
    smax = np.abs(oS.data).max()
    assert np.abs(gS.download() - oS.data).max() < 1e-12 * smax, "singular values"
    assert abs(ginfo["truncWeight"] - oinfo["truncWeight"]) <= 1e-11 * max(oinfo["truncWeight"], smax ** 2)
    assert abs(ginfo["entropy"] - oinfo["entropy"]) <= 1e-10 * max(abs(oinfo["entropy"]), 1.0)
    # gauge-free comparisons: the truncated reconstruction U S V^dagger and the subspace projectors
    grec, orec = (gU @ gS) @ gV, (oU @ oS) @ oV
    assert_same(grec, orec, 2e-11, "U*S*Vdag")
    one_g = gU.multiply(gU, True, False)
    eye = one_g.clone().setIdentity()
    assert np.abs(one_g.download() - eye.download()).max() < 1e-12, "U^dagger U = 1"
    one_v = gV.multiply(gV, False, True)
    assert np.abs(one_v.download() - eye.download()).max() < 1e-12, "Vdag Vdag^dagger = 1"
    return gU, gS, gV


@pytest.mark.parametrize("name", ["SU2", "U1", "U0"])
def test_tsvd(ctxs, name):
    ctx, osym = ctxs(name)
    rng = np.random.default_rng(13)
    B, C = rand_entries(SPECS[name], 14, rng, 5), rand_entries(SPECS[name], 12, rng, 5)
    g, o = pair(ctx, osym, [B, C], [B, C], 21)          # square blocks
    _svd_checks(g, o, 1000, 1e-10, False)                # no truncation
    _svd_checks(g, o, 7, 1e-10, False)                   # CTM-style truncation
    out = [(q, d) for q, _, d in obasis(osym, B).combine(obasis(osym, C)).forgetHistory().data][::2]
    g2, o2 = pair(ctx, osym, [B, C], [out], 22)         # tall / wide blocks (MPS-like A(2,1), cf. tests/mps_tests.hpp:13-16)
    _svd_checks(g2, o2, 1000, 0.0, False)
    _svd_checks(g2.adjoint(), o2.adjoint(), 5, 1e-10, False)


def test_tsvd_graded_spectrum_and_multiplets(ctxs):
    """M = U diag(10^(-k/4)) V^T per block: exercises the eps_svd cut and PRESERVE_MULTIPLETS (Tensor.cpp:551-584)."""
    ctx, osym = ctxs("SU2")
    ent = [((1,), 6), ((2,), 7), ((3,), 5)]
    g, o = pair(ctx, osym, [ent], [ent], 31)
    rng = np.random.default_rng(3)
    for i in range(len(o.sector)):
        n = o.block(i).shape[0]
        q1, _ = np.linalg.qr(rng.standard_normal((n, n)))
        q2, _ = np.linalg.qr(rng.standard_normal((n, n)))
        o.block(i)[...] = q1 @ np.diag(10.0 ** (-np.arange(n) * (2.9 + i))) @ q2.T
    g.upload(o.data)
    for keep, eps, pm in [(100, 1e-10, False), (9, 1e-10, False), (9, 1e-10, True), (4, 0.0, True)]:
        gU, gS, gV = g.tSVD(keep, eps, pm)
        oU, oS, oV = o.tSVD(keep, eps, pm)
        assert gS.structure() == oS.structure(), (keep, eps, pm)
        assert np.abs(gS.download() - oS.data).max() < 1e-13


@pytest.mark.parametrize("case", ["U0_square", "U0_tall", "U0_very_tall", "U0_tallest", "U1_big", "SU2_graded"])
def test_tsvd_large_blocks(ctxs, case):
    """Blocks wide enough for the QR-preconditioned path (>= 32 columns), the 4- and 8-CTA cluster round kernels and
    tall panels; same checks as test_tsvd (sigma 1e-12, truncated reconstruction, orthonormality)."""
    if case == "U0_square":
        ctx, osym = ctxs("U0")
        g, o = pair(ctx, osym, [[((), 40)], [((), 5)]], [[((), 8)], [((), 25)]], 71)          # one 200 x 200 block
        _svd_checks(g, o, 1000, 1e-10, False)
        _svd_checks(g, o, 30, 1e-10, False)
    elif case == "U0_tall":
        ctx, osym = ctxs("U0")
        g, o = pair(ctx, osym, [[((), 30)], [((), 11)]], [[((), 77)]], 72)                      # 330 x 77
        _svd_checks(g, o, 1000, 0.0, False)
        _svd_checks(g.adjoint(), o.adjoint(), 20, 1e-10, False)                                 # 77 x 330 (transposed work matrix)
    elif case == "U0_very_tall":
        ctx, osym = ctxs("U0")
        g, o = pair(ctx, osym, [[((), 650)], [((), 4)]], [[((), 40)]], 75)                      # 2600 x 40: the 16-CTA cluster round kernel
        _svd_checks(g, o, 1000, 0.0, False)
        _svd_checks(g, o, 12, 1e-10, False)
    elif case == "U0_tallest":
        ctx, osym = ctxs("U0")
        g, o = pair(ctx, osym, [[((), 925)], [((), 4)]], [[((), 36)]], 76)                      # 3700 x 36: beyond the resident round kernels (chunked fallback)
        _svd_checks(g, o, 1000, 0.0, False)
    elif case == "U1_big":
        ctx, osym = ctxs("U1")
        B = [((q,), 12) for q in range(-2, 3)]
        g, o = pair(ctx, osym, [B, B], [B, B], 73)                                              # 9 sectors, block side 144 ... 720
        _svd_checks(g, o, 10 ** 6, 1e-10, False)
        _svd_checks(g, o, 64, 1e-10, False)
    else:
        ctx, osym = ctxs("SU2")
        ent = [((1,), 30), ((2,), 41), ((3,), 25)]
        g, o = pair(ctx, osym, [ent, [((1,), 2), ((2,), 1)]], [ent, [((1,), 2), ((2,), 1)]], 74)
        rng = np.random.default_rng(5)
        for i in range(len(o.sector)):                                                          # graded spectra over 16 decades
            m, n = o.block(i).shape
            k = min(m, n)
            q1, _ = np.linalg.qr(rng.standard_normal((m, k)))
            q2, _ = np.linalg.qr(rng.standard_normal((n, k)))
            o.block(i)[...] = q1 @ np.diag(10.0 ** (-16.0 * np.arange(k) / k)) @ q2.T
        g.upload(o.data)
        _svd_checks(g, o, 48, 1e-10, False)
        gU, gS, gV = g.tSVD(10 ** 6, 1e-10, False)
        oU, oS, oV = o.tSVD(10 ** 6, 1e-10, False)
        assert gS.structure() == oS.structure()
        assert np.abs(gS.download() - oS.data).max() < 1e-13


@pytest.mark.parametrize("name", ["SU2", "U1"])
def test_decompose_projectors(ctxs, name):
    ctx, osym = ctxs(name)
    rng = np.random.default_rng(17)
    chi, aux = rand_entries(SPECS[name], 8, rng, 3), rand_entries(SPECS[name], 3, rng, 2)
    g1, o1 = pair(ctx, osym, [chi, aux, aux], [chi, aux, aux], 41)
    g2, o2 = pair(ctx, osym, [chi, aux, aux], [chi, aux, aux], 42)
    gP1, gP2 = xb.decompose(g1, g2, 6)
    oP1, oP2 = odecompose(o1, o2, 6)
    assert gP1.structure() == oP1.structure() and gP2.structure() == oP2.structure()
    assert_same(gP2 @ gP1, oP2 @ oP1, 1e-10, "P2*P1 (gauge invariant)")
    assert np.abs(np.abs(gP1.download()) - np.abs(oP1.data)).max() < 1e-9 * np.abs(oP1.data).max()


def _ctm_pair(ctx, osym, spec, D, chi, seed):
    phys_e = [((2,), 1)] if spec[0][0] == "SU2" else [((1,), 1), ((-1,), 1)]
    ophys = obasis(osym, phys_e)
    oaux = init_basis(osym, D, ophys)
    aux_e = [(q, d) for q, _, d in oaux.data]
    gA, oA = pair(ctx, osym, [aux_e, aux_e], [aux_e, aux_e, phys_e], seed)
    return xb.CTM(gA, chi), OCTM(oA, chi)


def _spectrum(t):
    """per-sector singular values of an environment tensor: invariant under the orthogonal gauge freedom on the chi legs"""
    if t.rank == 0:
        t = t.permute(-1, list(range(t.corank)))
    elif t.corank == 0:
        t = t.permute(1, list(range(t.rank)))
    S = t.tSVD(10 ** 6, 0.0, False)[1]
    return S.structure(), (S.download() if hasattr(S, "download") else S.data)


@pytest.mark.parametrize("name,D,chi", [("SU2", 3, 10), ("U1", 2, 8), ("U1", 3, 9)])
def test_ctm_init_corner_and_step(ctxs, name, D, chi):
    """CTM.cpp init (FROM_A) and contractCorner involve no SVD -> strict parity. After grow_all() the environment is
    defined up to an orthogonal gauge on the chi legs (sign flips, rotations inside degenerate singular subspaces --
    the bra/ket double layer has such degeneracies), so steps are compared through gauge-invariant per-sector spectra."""
    ctx, osym = ctxs(name)
    g, o = _ctm_pair(ctx, osym, SPECS[name], D, chi, 51)
    for nm, ot in o.tensors().items():
        assert_same(g.get(nm), ot, TOL, f"init {nm}")
    for corner in range(4):
        assert_same(g.contractCorner(corner), o.contractCorner(corner), TOL, f"contractCorner {corner}")
    for step in range(3):
        g.grow_all()
        o.grow_all()
        for nm, ot in o.tensors().items():
            gt = g.get(nm)
            assert gt.structure() == ot.structure(), f"step {step} {nm}"
            (gs_struct, gs), (os_struct, os_) = _spectrum(gt), _spectrum(ot)
            assert gs_struct == os_struct
            assert np.abs(gs - os_).max() < 1e-9 * np.abs(os_).max(), f"step {step} {nm}: spectrum differs by {np.abs(gs - os_).max():.2e}"


@pytest.mark.parametrize("name,D,chi", [("SU2", 3, 9), ("U1", 3, 8)])
def test_ctm_single_moves_generic_environment(ctxs, name, D, chi):
    """With a generic (random) environment the projector SVD has no degeneracies, so each directional move must agree
    with the oracle element-wise up to the sign of each new chi index: compare |values| (CTM.cpp:924-1641)."""
    ctx, osym = ctxs(name)
    g, o = _ctm_pair(ctx, osym, SPECS[name], D, chi, 61)
    from xped_oracle.rng import uniform_fill
    for direction in range(4):
        for k, nm in enumerate(["C1", "C2", "C3", "C4", "T1", "T2", "T3", "T4"]):
            seed = 100 + 10 * direction + k
            g.set(nm, g.get(nm).setRandom(seed))
            ot = getattr(o, nm)
            ot.data = uniform_fill(seed, ot.plainSize())
        g.move(direction)
        o.move(direction)
        for nm, ot in o.tensors().items():
            gt = g.get(nm)
            assert gt.structure() == ot.structure(), f"move {direction} {nm}"
            a, b = np.abs(gt.download()), np.abs(ot.data)
            assert np.abs(a - b).max() < 1e-10 * b.max(), f"move {direction} {nm}: {np.abs(a - b).max():.2e}"


@pytest.mark.parametrize("name", ["SU2", "U1", "U1xU1", "fU1xSU2"])
def test_mps_transfer_contractions(ctxs, name):
    """contract_L / contract_R (src/MPS/MpsContractions.cpp:11-105) vs the oracle, and the reference's own identity
    tests/mps_tests.hpp:19-33 (== permute + operator* formulation) evaluated on the device."""
    from xped_oracle.mps import contract_L as ocontract_L
    from xped_oracle.mps import contract_R as ocontract_R
    ctx, osym = ctxs(name)
    rng = np.random.default_rng(29)
    inb, loc = rand_entries(SPECS[name], 12, rng), rand_entries(SPECS[name], 4, rng, 2)
    out = [(q, d) for q, _, d in obasis(osym, inb).combine(obasis(osym, loc)).forgetHistory().data]
    g1, o1 = pair(ctx, osym, [inb, loc], [out], 71)
    g2, o2 = pair(ctx, osym, [inb, loc], [out], 72)
    gR, oR = pair(ctx, osym, [out], [out], 73)
    gL, oL = pair(ctx, osym, [inb], [inb], 74)
    gBr, gBl = xb.contract_R(gR, g1, g2), xb.contract_L(gL, g1, g2)
    assert_same(gBr, ocontract_R(oR, o1, o2), TOL, "contract_R")
    assert_same(gBl, ocontract_L(oL, o1, o2), TOL, "contract_L")
    chk = (g2 @ gR).permute(1, (0, 1, 2)).multiply(g1.permute(1, (0, 1, 2)), False, True)
    assert np.linalg.norm(gBr.download() - chk.download()) <= 1e-12 * np.linalg.norm(chk.download())
    chk = g1.permute(1, (0, 1, 2)).multiply(gL, True, False).permute(1, (1, 2, 0)) @ g2
    assert np.linalg.norm(gBl.download() - chk.download()) <= 1e-12 * np.linalg.norm(chk.download())


@pytest.mark.parametrize("name", ["SU2", "U1", "U1xU1"])
def test_tqr(ctxs, name):
    """Tensor::tQR (src/Core/Tensor.cpp:757-780) on MPS-shaped A(2,1) tensors (tall blocks) and a square (2,2) tensor.
    The reference has no tQR test ("parity unpinned"): compared with the oracle's LAPACK Householder factors (same sign
    convention) and through Q*R = A, Q^dagger Q = 1, R upper triangular."""
    ctx, osym = ctxs(name)
    rng = np.random.default_rng(37)
    inb, loc = rand_entries(SPECS[name], 12, rng), rand_entries(SPECS[name], 4, rng, 2)
    out = [(q, d) for q, _, d in obasis(osym, inb).combine(obasis(osym, loc)).forgetHistory().data][::2]
    cases = [pair(ctx, osym, [inb, loc], [out], 81), pair(ctx, osym, [inb, loc], [inb, loc], 82)]
    for g, o in cases:
        gQ, gR = g.tQR()
        oQ, oR = o.tQR()
        assert gQ.structure() == oQ.structure() and gR.structure() == oR.structure()
        assert_same(gQ @ gR, o, 1e-12, "Q*R")
        one = gQ.multiply(gQ, True, False)
        assert np.abs(one.download() - one.clone().setIdentity().download()).max() < 1e-12
        assert_same(gR, oR, 1e-11, "R (Householder sign convention)")
        assert_same(gQ, oQ, 1e-11, "Q")


@pytest.mark.parametrize("name", ["SU2", "U1", "U0"])
def test_teigh(ctxs, name):
    """Tensor::teigh (src/Core/Tensor.cpp:782-803): eigenvalues ascending per block (bit-exact structure, values 1e-12),
    eigenvectors through V D V^dagger = H and V^dagger V = 1 (their signs are a gauge freedom)."""
    ctx, osym = ctxs(name)
    rng = np.random.default_rng(41)
    B, Cb = rand_entries(SPECS[name], 13, rng, 5), rand_entries(SPECS[name], 9, rng, 4)
    g, o = pair(ctx, osym, [B, Cb], [B, Cb], 91)
    gH, oH = g + g.adjoint(), o + o.adjoint()
    gD, gV = gH.teigh()
    oD, oV = oH.teigh()
    assert gD.structure() == oD.structure() and gV.structure() == oV.structure()
    lam = np.abs(oD.data).max()
    assert np.abs(gD.download() - oD.data).max() < 1e-12 * lam * 10
    rec = (gV @ gD).multiply(gV, False, True)
    assert_same(rec, oH, 1e-11, "V D V^dagger")
    one = gV.multiply(gV, True, False)
    assert np.abs(one.download() - one.clone().setIdentity().download()).max() < 1e-12


@pytest.mark.parametrize("name", ["SU2", "U1"])
def test_tsvd_backward(ctxs, name):
    """SVD reverse rule (include/Xped/AD/ADTensor.hpp:253-325,356-371) vs the oracle (which is pinned against finite
    differences, tests/test_oracle_pins.py).  Adjoints are built from each side's own factors out of the same B, so the
    comparison is independent of the sign gauge of the singular vectors."""
    ctx, osym = ctxs(name)
    rng = np.random.default_rng(43)
    b1, b2 = rand_entries(SPECS[name], 11, rng), rand_entries(SPECS[name], 8, rng)
    out = [(q, d) for q, _, d in obasis(osym, b1).combine(obasis(osym, b2)).forgetHistory().data]
    for dom, cod, seed, keep in [([b1, b2], [out], 101, 9), ([b1, b2], [b1, b2], 103, 12)]:
        g, o = pair(ctx, osym, dom, cod, seed)
        gB, oB = pair(ctx, osym, dom, cod, seed + 1)
        gU, gS, gV = g.tSVD(keep, 0.0, False)
        oU, oS, oV = o.tSVD(keep, 0.0, False)
        assert gS.structure() == oS.structure()
        g_dU, o_dU = gB.multiply(gV, False, True) @ gS, (oB @ oV.adjoint()) @ oS
        g_dS, o_dS = gU.multiply(gB, True, False).multiply(gV, False, True), (oU.adjoint() @ oB) @ oV.adjoint()
        g_dV, o_dV = gS @ gU.multiply(gB, True, False), oS @ (oU.adjoint() @ oB)
        ggrad = g.tSVD_backward(gU, gS, gV, g_dU, g_dS, g_dV)
        ograd = o.tSVD_backward(o_dU, o_dS, o_dV)
        assert_same(ggrad, ograd, 1e-9, f"{name} tSVD_backward")
        assert_same(g.tSVD_backward(gU, gS, gV, g_dU, None, None), o.tSVD_backward(o_dU, None, None), 1e-9, f"{name} tSVD_backward (dU only)")


@pytest.mark.parametrize("name", ["SU2", "U1", "fU1xSU2"])
def test_sub_block_view_sub_matrix(ctxs, name):
    """Tensor::subBlock / subMatrix / view (include/Xped/Core/Tensor.hpp:191-216, src/Core/Tensor.cpp:830-981): every tree pair's dense
    sub-block equals the oracle's bit for bit (it is a strided copy), the trees reported through the C ABI are the oracle's, the pair is
    found again by its labels, view() addresses the same elements in place, and writing a sub-block back changes exactly that sub-block."""
    ctx, osym = ctxs(name)
    rng = np.random.default_rng(91)
    ents = [rand_entries(SPECS[name], 7, rng, 3) for _ in range(4)]
    g, o = pair(ctx, osym, ents[:2], ents[2:], 55)
    opairs = [(i, dt, ct) for i, q in enumerate(o.sector) for dt in o.domain.tree(q) for ct in o.codomain.tree(q)]
    assert len(opairs) == len(g.tree_pairs())
    full = g.download()
    for k, (i, dt, ct) in enumerate(opairs):
        qu, qi, qc, dims = g.tree_pair_trees(k)
        assert [tuple(q) for q in qu] == [tuple(q) for q in list(dt.q_uncoupled) + list(ct.q_uncoupled)]
        assert tuple(qc) == tuple(dt.q_coupled) and dims == list(dt.dims) + list(ct.dims)
        assert [tuple(q) for q in qi] == [tuple(q) for q in list(dt.q_intermediates) + list(ct.q_intermediates)]
        assert g.find_tree_pair(qu, qi, qc) == k
        ob = o.subBlock(dt, ct)
        gb = g.subBlock(k)
        assert gb.shape == ob.shape and np.array_equal(gb, ob)
        assert np.array_equal(g.subMatrix(k), o._sub(i, dt, ct))
        ptr, ld, r, c = g.view(k)
        assert (r, c) == o._sub(i, dt, ct).shape and ld == o.shape_of(i)[0] and ptr != 0
    # l-value use: overwrite one sub-block, everything else unchanged
    k = len(opairs) // 2
    i, dt, ct = opairs[k]
    new = rng.uniform(-1, 1, o.subBlock(dt, ct).shape)
    g.setSubBlock(k, new)
    o._sub(i, dt, ct)[...] = new.reshape(o._sub(i, dt, ct).shape, order="F")
    assert np.array_equal(g.download(), o.data)
    assert not np.array_equal(full, o.data) or new.size == 0


@pytest.mark.parametrize("name", ["SU2", "U1"])
def test_tqr_wide_blocks_and_return_lq(ctxs, name):
    """tQR beyond the tall case (src/Core/Tensor.cpp:757-780).  (a) middle basis = coupled domain (the tensor is wider than tall): blocks are
    Q_full 1(m x mid), 1(mid x m) R_full -- Q orthogonal, Q R = A, and where the reference's own matrices fit its bases (square blocks) they
    are the oracle's.  (b) RETURN_LQ on a square tensor: the adjoint blocks of both factors, as the reference returns them.  (c) RETURN_LQ on
    non-square blocks does not fit the bases: XB_ERR_INVALID."""
    ctx, osym = ctxs(name)
    rng = np.random.default_rng(53)
    inb, loc = rand_entries(SPECS[name], 10, rng), rand_entries(SPECS[name], 4, rng, 2)
    # (a) wide: A(1, 2)
    g, o = pair(ctx, osym, [inb], [inb, loc], 61)
    assert o.domain.dim() <= o.codomain.dim()
    gQ, gR = g.tQR()
    assert_same(gQ @ gR, o, 1e-12, "Q*R (wide)")
    rows, cols = g.structure()[1], g.structure()[2]
    qs, rs = gQ.structure(), gR.structure()
    for i in range(len(o.sector)):
        m, n = o.shape_of(i)
        assert (qs[1][i], qs[2][i]) == (m, m) and (rs[1][i], rs[2][i]) == (m, n)
        gq = gQ.download()[qs[3][i]:qs[3][i] + m * m].reshape((m, m), order="F")
        gr = gR.download()[rs[3][i]:rs[3][i] + m * n].reshape((m, n), order="F")
        qm, rm = np.linalg.qr(o.block(i), mode="complete")
        oq, orr = qm @ np.eye(m, n), np.eye(n, m) @ rm     # the reference's matrices (Tensor.cpp:771-772): m x n and n x n, which do not fit its own bases here
        k = min(m, n)
        assert np.abs(gr[:k, :] - orr[:k, :]).max() < 1e-11 * max(1.0, np.abs(orr).max())
        assert np.abs(gq[:, :k] - oq[:, :k]).max() < 1e-11
        assert np.abs(gq.T @ gq - np.eye(m)).max() < 1e-12          # Q_full: orthogonal in every sector the tensor has
        assert np.abs(np.tril(gr, -1)).max() == 0.0
    # (b) RETURN_LQ, square blocks
    g2, o2 = pair(ctx, osym, [inb, loc], [inb, loc], 62)
    gQ2, gR2 = g2.tQR(True)
    oQ2, oR2 = o2.tQR(True)
    assert_same(gQ2, oQ2, 1e-11, "Q^T")
    assert_same(gR2, oR2, 1e-11, "R^T")
    assert_same(gQ2.adjoint() @ gR2.adjoint(), o2, 1e-12, "(Q^T)^T (R^T)^T = A")
    # (c)
    with pytest.raises(xb.XbError) as ei:
        g.tQR(True)
    assert ei.value.status == 1   # XB_ERR_INVALID
