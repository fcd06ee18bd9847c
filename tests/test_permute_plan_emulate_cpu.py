"""The permute plan builder checked WITHOUT a GPU: a planning-only context keeps the host copies of a plan's tables, `xb_permute_plan_emulate` walks them over
host arrays the way the kernels address device memory (form 0: tile list of the LSU form, form 1: table-driven form; every offset bounds-checked, every
destination element written exactly once), and the result is compared with the oracle's Tensor::permute (src/Core/Tensor.cpp:83-214, 258-333).
What this covers: recoupling coefficients, sub-block offsets and strides of every source, tile origins / clipping, the offset / staging tables of the
table-driven form and its cut of a sub-block into full tiles + remainder regions.  What it does not cover: the kernels themselves (tests/test_gpu_parity.py)."""
import itertools

import numpy as np
import pytest

import xped_oracle as xo
from xped_oracle.rng import uniform_fill
from helpers import SPECS, obasis, rand_entries
from xped_b200.sharded import PlanningContext, layout_structure, permute_plan_emulate, permute_plan_stats

TOL = 1e-13


def oracle_tensor(osym, dom, cod, seed):
    t = xo.Tensor(osym, [obasis(osym, e) for e in dom], [obasis(osym, e) for e in cod])
    t.data = uniform_fill(seed, t.plainSize())
    return t


def check(ctx, osym, dom, cod, shift, p, seed, forms=(0, 1), adjoint=False, rank_after=None):
    o = oracle_tensor(osym, dom, cod, seed)
    ref = o.permute_adj(shift, p) if adjoint else o.permute(shift, p)
    gdom, gcod = [ctx.basis(e) for e in dom], [ctx.basis(e) for e in cod]
    stats = permute_plan_stats(ctx, gdom, gcod, shift, p) if not adjoint else None
    scale = max(float(np.abs(ref.data).max()), 1.0)
    for form in forms:
        try:
            out = permute_plan_emulate(ctx, gdom, gcod, shift, p, o.data, ref.plainSize(), form, adjoint)
        except Exception as e:   # plans with the bulk form (even Abelian runs) carry no table-driven form
            if form == 1 and "no table-driven form" in str(e):
                continue
            raise
        assert not np.isnan(out).any()
        err = float(np.abs(out - ref.data).max())
        assert err <= TOL * scale, f"form {form} permute<{shift},{p}> adjoint={adjoint}: max abs error {err:.3e}"
        if stats is not None and stats["max_fan_in"] <= 1 and not stats["identity"]:
            # fan-in 1: one product per element (coefficient x value) on both sides -- bit-exact for coefficients +-1 (bosonic / fermionic Abelian)
            if all(kind != "SU2" for kind, _, _ in ctx.spec):
                assert np.array_equal(out, ref.data), f"form {form} permute<{shift},{p}>: Abelian plan is not bit-exact"


@pytest.fixture(params=list(SPECS))
def world(request):
    name = request.param
    ctx = PlanningContext(SPECS[name])
    ctx.spec = SPECS[name]
    yield name, ctx, xo.Symmetry.from_spec(SPECS[name])
    ctx.close()


def test_all_24_permutations_of_a_2_2_tensor(world):
    """tests/tensor_tests.hpp:2-48 shape: Tensor<2,2> with odd multiplicities, all 24 permutations x shifts -2..2, both forms of the plan"""
    name, ctx, osym = world
    rng = np.random.default_rng(5)
    B, C = rand_entries(SPECS[name], 8, rng), rand_entries(SPECS[name], 7, rng)
    for shift in (-2, -1, 0, 1, 2):
        for p in itertools.permutations(range(4)):
            check(ctx, osym, [B, C], [B, C], shift, p, 3)


def test_ctm_shaped_plans(world):
    """corner-shaped (3,3) tensors and the (4,3) intermediate of a CTM move: the plans a CTM step actually builds (SU(2): recoupling sums with fan-in > 1),
    extents without useful divisors so that the table-driven form needs its remainder regions"""
    name, ctx, osym = world
    rng = np.random.default_rng(17)
    chi, aux = rand_entries(SPECS[name], 13, rng, 3), rand_entries(SPECS[name], 5, rng, 2)
    for shift, p in [(-1, (0, 2, 3, 5, 1, 4)), (1, (0, 3, 5, 1, 2, 4)), (0, (1, 0, 2, 4, 3, 5)), (-3, (5, 4, 3, 2, 1, 0)), (2, (3, 1, 5, 0, 2, 4))]:
        check(ctx, osym, [chi, aux, aux], [chi, aux, aux], shift, p, 14)
    for shift, p in [(0, (0, 2, 1, 3, 5, 4, 6)), (-2, (6, 0, 5, 1, 4, 2, 3)), (1, (1, 0, 3, 2, 4, 6, 5))]:
        check(ctx, osym, [chi, aux, aux, aux], [chi, aux, aux], shift, p, 15)


def test_permute_adj_plans(world):
    """the adjoint plans of the reverse pass (Tensor::permute_adj, src/Core/Tensor.cpp:258-333)"""
    name, ctx, osym = world
    rng = np.random.default_rng(23)
    B, C = rand_entries(SPECS[name], 7, rng), rand_entries(SPECS[name], 6, rng)
    for shift, p in [(1, (2, 0, 3, 1)), (-1, (0, 2, 3, 1)), (0, (1, 0, 3, 2)), (2, (3, 2, 1, 0))]:
        check(ctx, osym, [B, C], [C, B], shift, p, 29, adjoint=True)


def test_large_tiles_and_many_tiles():
    """an SU(2) tensor large enough that sub-blocks span several 2048-element tiles per region (body of full tiles + remainders along two legs)"""
    ctx = PlanningContext(SPECS["SU2"])
    ctx.spec = SPECS["SU2"]
    osym = xo.Symmetry.SU2()
    e, a = [((1,), 7), ((2,), 13), ((3,), 9)], [((1,), 3), ((2,), 2)]
    for shift, p in [(-1, (0, 2, 3, 5, 1, 4)), (1, (0, 3, 5, 1, 2, 4)), (0, (1, 0, 2, 4, 3, 5))]:
        check(ctx, osym, [e, a, a], [e, a, a], shift, p, 41)
    ctx.close()


def test_emulation_is_refused_by_nothing_but_planning_contexts():
    """the entry point is test support: it needs the host tables that only a planning-only context keeps"""
    ctx = PlanningContext(SPECS["U1"])
    B = ctx.basis([((q,), 3) for q in (-1, 0, 1)])
    st = layout_structure(ctx, [B, B], [B, B])
    n = sum(r * c for r, c in zip(st[1], st[2]))
    with pytest.raises(Exception):
        permute_plan_emulate(ctx, [B, B], [B, B], 0, (1, 0, 2, 3), np.zeros(n + 1), n, 0)   # wrong array size
    with pytest.raises(Exception):
        permute_plan_emulate(ctx, [B, B], [B, B], 0, (1, 0, 2, 3), np.zeros(n), n, 7)       # unknown form
    ctx.close()


def test_w3_corner_plans_both_forms_agree_at_full_size():
    """configs[3] shape: the SU(2) D=6 chi=144 corner Q(3,3) (19.3 M elements, 442 sub-blocks, recoupling sums with fan-in up to 5).  The tile list of the LSU
    form and the tables of the table-driven form (11.5 k tiles, ~10 MB of tables, full tiles + remainder regions) are walked over the same random data: both
    visit every destination element exactly once, stay inside the arrays, and -- accumulating over the sources in the same order -- agree bit for bit."""
    ctx = PlanningContext(SPECS["SU2"])
    a, e = ctx.basis([((1,), 3), ((2,), 3)]), ctx.basis([((1,), 16), ((2,), 32), ((3,), 36), ((4,), 30), ((5,), 18), ((6,), 8), ((7,), 4)])
    st = layout_structure(ctx, [e, a, a], [e, a, a])
    n = int(sum(r * c for r, c in zip(st[1], st[2])))
    assert n == 19327896
    src = uniform_fill(9, n)
    for shift, p in [(-1, (0, 2, 3, 5, 1, 4)), (1, (0, 3, 5, 1, 2, 4)), (0, (1, 0, 2, 4, 3, 5))]:
        lsu = permute_plan_emulate(ctx, [e, a, a], [e, a, a], shift, p, src, n, 0)
        lut = permute_plan_emulate(ctx, [e, a, a], [e, a, a], shift, p, src, n, 1)
        assert not np.isnan(lsu).any() and np.array_equal(lsu, lut), (shift, p)
    ctx.close()


@pytest.mark.parametrize("name,dim", [("U1", 4), ("U1", 8), ("U1", 12), ("Z3", 10), ("U0", 8), ("U0", 48), ("fU1", 8)])
def test_bulk_form_all_24_permutations(name, dim):
    """the cp.async.bulk form (the kernel behind the headline HBM roofline): every bulk operation of every tile is replayed on the host -- source runs into the
    staging buffer, transposing pass, destination runs out of it -- with the hardware's demands on a bulk copy checked (16-byte aligned addresses and sizes,
    inside the staging buffer) and the byte count announced to the mbarrier compared with what the loads deliver.  Bit-exact against the oracle for bosonic
    symmetries; fermionic plans carry -1 coefficients through the transposing pass."""
    ctx, osym = PlanningContext(SPECS[name]), xo.Symmetry.from_spec(SPECS[name])
    ctx.spec = SPECS[name]
    if name in ("U1", "fU1"):
        B = [((q,), dim) for q in range(-2, 3)]
    elif name == "Z3":
        B = [((q,), dim) for q in range(3)]
    else:
        B = [((), dim)]
    o = oracle_tensor(osym, [B, B], [B, B], 81)
    gB = ctx.basis(B)
    used = 0
    for p in itertools.permutations(range(4)):
        for shift in (0, 1, -1, 2, -2):
            st = permute_plan_stats(ctx, [gB, gB], [gB, gB], shift, p)
            if st["identity"] or st["bulk_tiles"] == 0:
                continue
            ref = o.permute(shift, p)
            out = permute_plan_emulate(ctx, [gB, gB], [gB, gB], shift, p, o.data, ref.plainSize(), 2)
            assert np.array_equal(out, ref.data), f"{name} dim {dim} bulk permute<{shift},{p}>"
            used += 1
    assert used > (0 if dim < 8 else 60), used   # runs shorter than 8 doubles are not worth a bulk operation: dim 4 has few or no bulk plans
    ctx.close()


@pytest.mark.parametrize("qmax,dim", [(4, 8), (1, 24)])
def test_bulk_form_w4_shaped_tensor(qmax, dim):
    """configs[4] at reduced sizes: U(1) Tensor<2,2>, leg q = -4..4 x 8 (17 sectors, 2.0 M elements) and q = -1..1 x 24 (the W4 multiplicity: the tile extents
    the 1.30 GB tensor gets), the five permutes x three shifts of the W4 bench set"""
    ctx, osym = PlanningContext(SPECS["U1"]), xo.Symmetry.U1()
    B = [((q,), dim) for q in range(-qmax, qmax + 1)]
    o = oracle_tensor(osym, [B, B], [B, B], 7)
    from xped_oracle import fast
    fast.install()
    try:
        gB = ctx.basis(B)
        for p in [(1, 0, 2, 3), (0, 1, 3, 2), (2, 3, 0, 1), (0, 2, 1, 3), (3, 2, 1, 0)]:
            for shift in (0, 1, -2):
                st = permute_plan_stats(ctx, [gB, gB], [gB, gB], shift, p)
                assert st["bulk_tiles"] > 0
                ref = o.permute(shift, p)
                out = permute_plan_emulate(ctx, [gB, gB], [gB, gB], shift, p, o.data, ref.plainSize(), 2)
                assert np.array_equal(out, ref.data), (p, shift)
    finally:
        fast.uninstall()
    ctx.close()
