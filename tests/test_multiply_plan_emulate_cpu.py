"""The host side of the grouped product checked WITHOUT a GPU: `xb_multiply_plan_emulate` builds the problem list of op(A) * op(B) and the tile / split-K lists
exactly as the device path does (multiply_problems + build_gemm_lists, xped_b200/csrc/ops.cpp) and walks them over host arrays the way gemm_grouped_kernel and
splitk_reduce_kernel address device memory (clipped tiles, K slices into scratch, slice sums), bounds-checked and write-once-checked.  Compared with the oracle's
operator* (src/Core/TensorBase.cpp:243-288).  Covers: sector matching, offsets / leading dimensions, adjoint operands, tile coverage of ragged blocks, the
128-row / 64-row tile choice (small launches), split-K slicing (chunks rounded to 8) and its scratch layout.  Not covered: the kernels (tests/test_gpu_parity.py)."""
import numpy as np
import pytest

import xped_oracle as xo
from xped_oracle.rng import uniform_fill
from helpers import SPECS, obasis, rand_entries
from xped_b200.sharded import PlanningContext, multiply_plan_emulate


def otensor(osym, dom, cod, seed):
    t = xo.Tensor(osym, [obasis(osym, e) for e in dom], [obasis(osym, e) for e in cod])
    t.data = uniform_fill(seed, t.plainSize())
    return t


def check(ctx, osym, a_dom, a_cod, adj_a, b_dom, b_cod, adj_b, sm_count=0, tol=1e-12):
    oa, ob = otensor(osym, a_dom, a_cod, 5), otensor(osym, b_dom, b_cod, 6)
    ref = (oa.adjoint() if adj_a else oa) @ (ob.adjoint() if adj_b else ob)
    g = lambda es: [ctx.basis(e) for e in es]
    out, counts = multiply_plan_emulate(ctx, g(a_dom), g(a_cod), adj_a, g(b_dom), g(b_cod), adj_b, oa.data, ob.data, ref.plainSize(), sm_count)
    assert not np.isnan(out).any()
    assert ref.plainSize() > 0
    scale = max(float(np.abs(ref.data).max()), 1.0)
    assert float(np.abs(out - ref.data).max()) <= tol * scale
    return counts


@pytest.mark.parametrize("name", list(SPECS))
def test_small_products_all_adjoint_combinations(name):
    """tests/tensor_tests.hpp multiplication shapes: T(2,2) * T(2,2), with lazily adjoint operands on either side (AdjointOp.hpp:26-69)"""
    ctx, osym = PlanningContext(SPECS[name]), xo.Symmetry.from_spec(SPECS[name])
    rng = np.random.default_rng(3)
    B, Cb, D = rand_entries(SPECS[name], 8, rng), rand_entries(SPECS[name], 7, rng), rand_entries(SPECS[name], 6, rng)
    check(ctx, osym, [B, Cb], [Cb, D], False, [Cb, D], [D, B], False)
    check(ctx, osym, [Cb, D], [B, Cb], True, [Cb, D], [D, B], False)
    check(ctx, osym, [B, Cb], [Cb, D], False, [D, B], [Cb, D], True)
    check(ctx, osym, [Cb, D], [B, Cb], True, [D, B], [Cb, D], True)
    ctx.close()


def test_tile_choice_and_split_k():
    """blocks large enough for every branch of the list builder: 128-row tiles (a launch with more than two CTAs per SM: few SMs), 64-row tiles for small launches
    (148 SMs), split K (k >= 256 with <= 32 tiles: slices of a multiple of 8, ragged last slice), the wave rule (k >= 768), ragged tile edges"""
    ctx, osym = PlanningContext(SPECS["U1"]), xo.Symmetry.U1()
    X = [((-1,), 131), ((0,), 97), ((1,), 45)]      # rows: ragged against 128 / 64
    K = [((-1,), 300), ((0,), 781), ((1,), 259)]    # contraction: 300 / 781 / 259
    Y = [((-1,), 70), ((0,), 150), ((1,), 33)]
    one = [((0,), 1)]
    a_dom, a_cod, b_dom, b_cod = [X, one], [K, one], [K, one], [Y, one]
    seen = {"tiles128": 0, "tiles64": 0, "jobs": 0}
    for sm in (148, 16, 2):
        for adj_a, adj_b in [(False, False), (True, False), (False, True)]:
            ad, ac = (a_cod, a_dom) if adj_a else (a_dom, a_cod)
            bd, bc = (b_cod, b_dom) if adj_b else (b_dom, b_cod)
            c = check(ctx, osym, ad, ac, adj_a, bd, bc, adj_b, sm_count=sm, tol=1e-11)
            for k in seen:
                seen[k] += c[k]
    assert seen["tiles128"] > 0 and seen["tiles64"] > 0 and seen["jobs"] > 0, seen
    # the same product on 148 SMs is a small launch (64-row tiles only) and cuts K; on 2 SMs it keeps 128-row tiles for the tall blocks
    c148 = check(ctx, osym, a_dom, a_cod, False, b_dom, b_cod, False, sm_count=148, tol=1e-11)
    c2 = check(ctx, osym, a_dom, a_cod, False, b_dom, b_cod, False, sm_count=2, tol=1e-11)
    assert c148["tiles128"] == 0 and c148["jobs"] > 0
    assert c2["tiles128"] > 0
    ctx.close()


def test_su2_corner_product():
    """Q4 * Q1 of two CTM-shaped corners (SU(2), small chi): nine-sector product with blocks of very different sizes"""
    ctx, osym = PlanningContext(SPECS["SU2"]), xo.Symmetry.SU2()
    e, a = [((1,), 3), ((2,), 5), ((3,), 4), ((4,), 2)], [((1,), 2), ((2,), 2)]
    check(ctx, osym, [e, a, a], [e, a, a], False, [e, a, a], [e, a, a], False)
    check(ctx, osym, [e, a, a], [e, a, a], True, [e, a, a], [e, a, a], False)
    ctx.close()
