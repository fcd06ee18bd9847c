"""Parity at the shapes bench.py times (BASELINE.json configs[1..4]); everything goes through the C ABI and is compared with the CPU oracle
(numpy / LAPACK) on the same counter-based inputs.  Sizes: W3 = SU(2) D=6 chi=144 corners (blocks up to 2394^2), CTM at D=4 chi=64 and D=6 chi=144,
W4 = the 1.3 GB U(1) tensor, configs[2] = U(1)xU(1) m ~ 2000 MPS transfer contraction.  Minutes of CPU time in total: marked `slow` as well."""
import numpy as np
import pytest

import xped_b200 as xb
import xped_oracle as xo
from helpers import SPECS, obasis
from xped_oracle.ctm import CTM as OCTM
from xped_oracle.ctm import avg_two_site, heisenberg_bond_SU2, init_basis
from xped_oracle.rng import uniform_fill

pytestmark = [pytest.mark.gpu, pytest.mark.slow]

AUX6 = [((1,), 3), ((2,), 3)]
ENV6 = [((1,), 16), ((2,), 32), ((3,), 36), ((4,), 30), ((5,), 18), ((6,), 8), ((7,), 4)]


@pytest.fixture(scope="module")
def su2():
    ctx = xb.Context(SPECS["SU2"])
    yield ctx, xo.Symmetry.from_spec(SPECS["SU2"])
    ctx.close()


def blocks_of(t_struct, data):
    qs, rows, cols, offs = t_struct
    return {tuple(q): data[o:o + r * c].reshape((r, c), order="F") for q, r, c, o in zip(qs, rows, cols, offs)}


def test_w3_product_and_tsvd(su2):
    """configs[3]: M = Q4*Q1 of two random D=6 chi=144 corners (9 sectors, blocks 36..2394), then
    (a) the FULL decomposition (all 10800 values, the QR-preconditioned Jacobi kernels at that size): sigma vs LAPACK 1e-12, U S V^+ = M 1e-12, orthonormal factors;
    (b) tSVD(144, 1e-10): flat spectrum, so the truncated algorithm has to give up and hand over to the full path -- same kept values as the oracle,
        kept triplets are singular triplets of M, truncWeight / entropy;
    (c) the same shapes with a graded spectrum sigma_k = 10^(-k/40): the truncated algorithm converges; sigma vs the exact values, the
        truncated reconstruction vs the exact one."""
    ctx, osym = su2
    a, e = ctx.basis(AUX6), ctx.basis(ENV6)
    Q4 = ctx.tensor([e, a, a], [e, a, a]).setRandom(1)
    Q1 = ctx.tensor([e, a, a], [e, a, a]).setRandom(2)
    M = Q4 @ Q1
    st = M.structure()
    assert sorted(st[1]) == [36, 144, 414, 882, 1188, 1530, 2070, 2142, 2394] and st[1] == st[2]
    oa, oe = obasis(osym, AUX6), obasis(osym, ENV6)
    o4 = xo.Tensor(osym, [oe, oa, oa], [oe, oa, oa]); o4.data = uniform_fill(1, o4.plainSize())
    o1 = xo.Tensor(osym, [oe, oa, oa], [oe, oa, oa]); o1.data = uniform_fill(2, o1.plainSize())
    oM = o4 @ o1
    assert M.structure() == oM.structure()
    gm = M.download()
    assert np.linalg.norm(gm - oM.data) / np.linalg.norm(oM.data) < 1e-12          # grouped DGEMM at the W3 shape
    oM.data = gm.copy()                                                            # same input for both decompositions
    mblocks = blocks_of(st, gm)
    osig = {tuple(q): np.linalg.svd(oM.block(i), compute_uv=False) for i, q in enumerate(oM.sector)}
    nsv = sum(len(v) for v in osig.values())
    # (a) full decomposition
    U, S, Vd, info = M.tSVD(nsv, 0.0, False, return_info=True)
    sblocks = blocks_of(S.structure(), S.download())
    ublocks, vblocks = blocks_of(U.structure(), U.download()), blocks_of(Vd.structure(), Vd.download())
    smax = max(v[0] for v in osig.values())
    for q, ref in osig.items():
        sg = np.diag(sblocks[q])
        assert np.abs(sg - ref).max() < 1e-12 * smax, f"sigma, sector {q}"
        u, v = ublocks[q], vblocks[q]
        assert np.abs(u.T @ u - np.eye(u.shape[1])).max() < 1e-12 and np.abs(v @ v.T - np.eye(v.shape[0])).max() < 1e-12
        assert np.linalg.norm((u * sg) @ v - mblocks[q]) < 1e-12 * np.linalg.norm(gm)
    assert abs(info["truncWeight"]) < 1e-12 * np.linalg.norm(gm) ** 2
    # (b) chi = 144 of a flat spectrum
    calls0, fb0 = ctx.get_int("svd.truncated_calls"), ctx.get_int("svd.truncated_fallbacks")
    U, S, Vd, info = M.tSVD(144, 1e-10, False, return_info=True)
    oU, oS, oVd, oinfo = oM.tSVD(144, 1e-10, False, return_info=True)
    assert ctx.get_int("svd.truncated_calls") == calls0 + 1 and ctx.get_int("svd.truncated_fallbacks") == fb0 + 1
    assert S.structure() == oS.structure()
    assert np.linalg.norm(S.download() - oS.data) / np.linalg.norm(oS.data) < 1e-12
    assert abs(info["truncWeight"] - oinfo["truncWeight"]) < 1e-11 * oinfo["truncWeight"]
    assert abs(info["entropy"] - oinfo["entropy"]) < 1e-11 * abs(oinfo["entropy"])
    sblocks = blocks_of(S.structure(), S.download())
    ublocks, vblocks = blocks_of(U.structure(), U.download()), blocks_of(Vd.structure(), Vd.download())
    for q, sb in sblocks.items():
        sg, u, v, m = np.diag(sb), ublocks[q], vblocks[q], mblocks[q]
        assert np.linalg.norm(m @ v.T - u * sg) < 1e-12 * smax * np.sqrt(len(sg)) and np.linalg.norm(m.T @ u - v.T * sg) < 1e-12 * smax * np.sqrt(len(sg))
        assert np.abs(u.T @ u - np.eye(len(sg))).max() < 1e-12 and np.abs(v @ v.T - np.eye(len(sg))).max() < 1e-12
    # (c) graded spectrum on the same shapes
    rng = np.random.default_rng(3)
    exact = {}
    data = np.empty_like(gm)
    for q, r, c, o in zip(*st):
        n = r
        qu, _ = np.linalg.qr(rng.standard_normal((n, n)))
        qv, _ = np.linalg.qr(rng.standard_normal((n, n)))
        sg = 10.0 ** (-np.arange(n) / 40.0)
        exact[tuple(q)] = (qu, sg, qv)
        data[o:o + n * n] = ((qu * sg) @ qv.T).reshape(-1, order="F")
    G = M.clone()
    G.upload(data)
    calls0, fb0 = ctx.get_int("svd.truncated_calls"), ctx.get_int("svd.truncated_fallbacks")
    U, S, Vd, info = G.tSVD(144, 1e-10, False, return_info=True)
    assert ctx.get_int("svd.truncated_calls") == calls0 + 1 and ctx.get_int("svd.truncated_fallbacks") == fb0     # converged on the truncated path
    sblocks = blocks_of(S.structure(), S.download())
    ublocks, vblocks = blocks_of(U.structure(), U.download()), blocks_of(Vd.structure(), Vd.download())
    assert sorted(sb.shape[0] for sb in sblocks.values()) == [16] * 9
    tw_exact = 0.0
    for q, (qu, sg, qv) in exact.items():
        k = 16
        assert np.abs(np.diag(sblocks[q]) - sg[:k]).max() < 1e-12
        rec = (ublocks[q] * np.diag(sblocks[q])) @ vblocks[q]
        ref = (qu[:, :k] * sg[:k]) @ qv[:, :k].T
        assert np.linalg.norm(rec - ref) < 1e-11 * np.linalg.norm(ref)
        tw_exact += osym.degeneracy(q) * np.sum(sg[k:] ** 2)
    assert abs(info["truncWeight"] - tw_exact) < 1e-10 * tw_exact


def oracle_su2_ctm(D, chi, seed=0):
    sym = xo.Symmetry.SU2()
    phys = xo.Qbasis.from_list(sym, [((2,), 1)])
    aux = init_basis(sym, D, phys)
    A = xo.Tensor(sym, [aux, aux], [aux, aux, phys])
    A.data = uniform_fill(seed, A.plainSize())
    return sym, phys, aux, A, OCTM(A, chi)


def sv_spectra(struct, data):
    out = {}
    for q, blk in blocks_of(struct, data).items():
        out[q] = np.linalg.svd(blk, compute_uv=False)
    return out


def test_ctm_d4_chi64_against_oracle(su2):
    """configs[1]: one grow_all() (four moves: 4 tSVD of 1664 values, 64 contractions) of the SU(2) D=4 chi=64 CTM, device vs oracle.
    The environment is a GENERIC one on the bases the FROM_A run has after two steps (the same counter-based random C1..C4, T1..T4 on both sides):
    the FROM_A trajectory itself cuts exactly degenerate pairs of singular values in half at chi = 64 (step 0, right move: sigma_64 = sigma_65 in sector q=2),
    where the reference's result depends on LAPACK's arbitrary choice inside the pair and no two implementations agree beyond ~1e-4.
    The chi legs carry a sign gauge from the singular vectors, so the comparison uses gauge-invariant data: the singular spectra of the new corners and
    edge tensors, and the two-site density matrices rho_h, rho_v (physical legs only) with the bond energy."""
    ctx, osym = su2
    sym, phys, aux, oA, octm = oracle_su2_ctm(4, 64)
    entries = [(tuple(q), d) for q, _, d in aux.data]
    a, p = ctx.basis(entries), ctx.basis([((2,), 1)])
    A = ctx.tensor([a, a], [a, a, p]).setRandom(0)
    assert np.array_equal(A.download(), oA.data)
    ctm = xb.CTM(A, 64)
    for _ in range(2):
        ctm.grow_all()
        octm.grow_all()
    names = ["C1", "C2", "C3", "C4", "T1", "T2", "T3", "T4"]
    for k, name in enumerate(names):
        g, o = ctm.get(name), octm.tensors()[name]
        assert g.structure() == o.structure()
        o.data = uniform_fill(100 + k, o.plainSize())
        g.upload(o.data)
        ctm.set(name, g)
    ctm.grow_all()
    octm.grow_all()
    worst = 0.0
    for name in names:
        g, o = ctm.get(name), octm.tensors()[name]
        assert g.structure() == o.structure()
        gs, os_ = sv_spectra(g.structure(), g.download()), sv_spectra(o.structure(), o.data)
        top = max(v[0] for v in os_.values())
        for q in os_:
            worst = max(worst, np.abs(gs[q] - os_[q]).max() / top)
    assert worst < 1e-11, f"environment spectra differ by {worst:.2e}"
    bond = heisenberg_bond_SU2(sym, phys)
    gbond = ctx.tensor([p, p], [p, p])
    assert gbond.structure() == bond.structure()
    gbond.upload(bond.data)
    for which, (grho, orho) in {"h": (ctm.computeRDM_h(), octm.computeRDM_h()), "v": (ctm.computeRDM_v(), octm.computeRDM_v())}.items():
        assert grho.structure() == orho.structure()
        err = np.linalg.norm(grho.download() - orho.data) / np.linalg.norm(orho.data)
        assert err < 1e-10, f"rho_{which}: {err:.2e}"
        assert abs(xb.avg_two_site(grho, gbond) - avg_two_site(orho, bond)) < 1e-10


def test_ctm_d6_chi144_truncated_vs_full_on_device(su2):
    """The north-star workload size (D=6, chi=144: projector matrices with blocks up to ~3060^2, which run the chunked Jacobi round kernel on the full path and
    the subspace iteration on the truncated path): both algorithms on the same M = Q4*Q1 of a D=6 environment -- kept singular values 1e-12,
    the gauge-invariant projector product P2*P1 1e-10, and each result against M itself (singular triplets, orthonormality)."""
    ctx, osym = su2
    a, p = ctx.basis(AUX6[:2]), ctx.basis([((2,), 1)])
    A = ctx.tensor([a, a], [a, a, p]).setRandom(0)
    ctm = xb.CTM(A, 144)
    for _ in range(2):
        ctm.grow_all()
    Q1, Q4 = ctm.contractCorner(0), ctm.contractCorner(3)
    M = Q4 @ Q1
    assert max(M.structure()[1]) > 2432          # beyond the resident round kernel's range
    res = {}
    for mode in (1, 0):
        ctx.set_int("svd.truncated", mode)
        fb0 = ctx.get_int("svd.truncated_fallbacks")
        U, S, Vd = M.tSVD(144, 1e-10, False)
        if mode == 1:
            assert ctx.get_int("svd.truncated_fallbacks") == fb0, "the truncated path did not converge on a CTM spectrum"
        P1, P2 = xb.decompose(Q4, Q1, 144)
        res[mode] = (U, S, Vd, P2 @ P1)
    ctx.set_int("svd.truncated", -1)
    (U1, S1, V1, PP1), (U0, S0, V0, PP0) = res[1], res[0]
    assert S1.structure() == S0.structure()
    assert np.linalg.norm(S1.download() - S0.download()) / np.linalg.norm(S0.download()) < 1e-12
    assert np.linalg.norm(PP1.download() - PP0.download()) / np.linalg.norm(PP0.download()) < 1e-10
    nM = M.norm()
    for U, S, Vd in ((U1, S1, V1), (U0, S0, V0)):
        r1 = (M.multiply(Vd, False, True) - (U @ S)).norm() / nM
        one = U.multiply(U, True, False)
        assert r1 < 1e-12
        assert np.abs(one.download() - one.clone().setIdentity().download()).max() < 1e-12


def test_w4_permutes_and_product():
    """configs[4] on one GPU: the 1.3 GB U(1) Tensor<2,2> (leg q = -4..4 x 24).  Three of the benchmark's permutes bit-exact against the oracle
    (fan-in 1: pure data movement, the cp.async.bulk path), and T*T on four sectors (the largest one included) against numpy per-sector GEMMs."""
    ctx, osym = xb.Context(SPECS["U1"]), xo.Symmetry.from_spec(SPECS["U1"])
    try:
        ent = [((q,), 24) for q in range(-4, 5)]
        B, oB = ctx.basis(ent), obasis(osym, ent)
        T = ctx.tensor([B, B], [B, B]).setRandom(0)
        oT = xo.Tensor(osym, [oB, oB], [oB, oB])
        oT.data = uniform_fill(0, oT.plainSize())
        assert T.plainSize() == 162238464
        for shift, perm in [(0, (1, 0, 2, 3)), (1, (2, 3, 0, 1)), (-2, (3, 2, 1, 0))]:
            g, o = T.permute(shift, perm), oT.permute(shift, perm)
            assert g.structure() == o.structure()
            assert np.array_equal(g.download(), o.data), f"permute<{shift}, {perm}> is not bit-exact"
            del g, o
        P = T @ T
        st = P.structure()
        pd = P.download()
        for q in [(0,), (1,), (-3,), (8,)]:
            i = st[0].index(q)
            n = st[1][i]
            blk = pd[st[3][i]:st[3][i] + n * n].reshape((n, n), order="F")
            ob = oT.block(oT.dict[q])
            ref = ob @ ob
            assert np.linalg.norm(blk - ref) / np.linalg.norm(ref) < 1e-12, f"T*T sector {q}"
    finally:
        ctx.close()


def test_hubbard_transfer_contraction_m2000():
    """configs[2]: U(1)xU(1) (fermion number, 2 Sz) MPS site tensor A(2,1) with a bond basis of m ~ 2000 states in ~88 sectors (Gaussian over (N, 2Sz), N + 2Sz
    even) and the four-state Hubbard site; contract_L / contract_R against the oracle."""
    name = "U1xU1"
    ctx, osym = xb.Context(SPECS[name]), xo.Symmetry.from_spec(SPECS[name])
    try:
        from xped_b200.workloads import hubbard_bond_basis
        bond = hubbard_bond_basis(2000)
        loc = [((0, 0), 1), ((1, 1), 1), ((1, -1), 1), ((2, 0), 1)]
        assert 1800 <= sum(d for _, d in bond) <= 2200 and 80 <= len(bond) <= 95
        ob, ol = obasis(osym, bond), obasis(osym, loc)
        out = [(tuple(q), d) for q, _, d in ob.combine(ol).forgetHistory().data]
        out = [(q, min(d, dict(bond).get(q, 0))) for q, d in out if dict(bond).get(q, 0) > 0]
        gA = ctx.tensor([ctx.basis(bond), ctx.basis(loc)], [ctx.basis(out)]).setRandom(5)
        oA = xo.Tensor(osym, [ob, ol], [obasis(osym, out)])
        oA.data = uniform_fill(5, oA.plainSize())
        assert gA.structure() == oA.structure()
        gL = ctx.tensor([ctx.basis(bond)], [ctx.basis(bond)]).setRandom(6)
        oL = xo.Tensor(osym, [ob], [ob]); oL.data = uniform_fill(6, oL.plainSize())
        gR = ctx.tensor([ctx.basis(out)], [ctx.basis(out)]).setRandom(7)
        oR = xo.Tensor(osym, [obasis(osym, out)], [obasis(osym, out)]); oR.data = uniform_fill(7, oR.plainSize())
        from xped_oracle.mps import contract_L as o_contract_L, contract_R as o_contract_R
        gl, ol_ = xb.contract_L(gL, gA, gA), o_contract_L(oL, oA, oA)
        assert gl.structure() == ol_.structure()
        assert np.linalg.norm(gl.download() - ol_.data) / np.linalg.norm(ol_.data) < 1e-12
        gr, or_ = xb.contract_R(gR, gA, gA), o_contract_R(oR, oA, oA)
        assert gr.structure() == or_.structure()
        assert np.linalg.norm(gr.download() - or_.data) / np.linalg.norm(or_.data) < 1e-12
    finally:
        ctx.close()
