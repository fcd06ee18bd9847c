"""Host logic of the permute plan builder in a planning-only context (no GPU): which plans get the cp.async.bulk form
(fan-in 1, 16-byte aligned runs of even length on both sides), algorithmic bytes, fan-in of SU(2) plans."""
import itertools

import pytest

from helpers import SPECS
from xped_b200.sharded import PlanningContext, layout_structure, permute_plan_stats


def _ctx(name):
    return PlanningContext(SPECS[name])


def test_even_abelian_tensor_gets_the_bulk_form():
    ctx = _ctx("U1")
    B = ctx.basis([(q, 8) for q in range(-2, 3)])   # runs of >= 8 doubles: shorter ones are not worth a bulk operation
    st = layout_structure(ctx, [B, B], [B, B])
    nelem = sum(r * c for r, c in zip(st[1], st[2]))
    for p in itertools.permutations(range(4)):
        for shift in (0, 1, -2):
            s = permute_plan_stats(ctx, [B, B], [B, B], shift, p)
            if s["identity"]:
                assert p == (0, 1, 2, 3) and shift == 0
                continue
            assert s["max_fan_in"] == 1
            assert s["bytes"] == 16 * nelem                      # 8 * (N_src + N_dst), SURVEY.md section 8d
            assert s["bulk_tiles"] > 0, (p, shift)
            assert s["tiles"] > 0                                # the LSU form is always built (sharded gather, small tensors)
    ctx.close()


@pytest.mark.parametrize("name,entries", [("U1", [(-1, 3), (0, 2), (1, 3)]), ("U0", [((), 5)])])
def test_odd_run_lengths_fall_back_to_the_lsu_form(name, entries):
    ctx = _ctx(name)
    B = ctx.basis(entries)
    s = permute_plan_stats(ctx, [B, B], [B, B], 0, (1, 0, 2, 3))
    assert s["bulk_tiles"] == 0 and s["tiles"] > 0
    ctx.close()


def test_su2_plans_sum_over_recoupled_trees():
    ctx = _ctx("SU2")
    B = ctx.basis([(1, 2), (2, 2), (3, 2)])
    s = permute_plan_stats(ctx, [B, B], [B, B], 0, (1, 0, 2, 3))
    assert s["max_fan_in"] == 1                                  # a swap inside the domain is a phase (coeff_swap), no sum
    s = permute_plan_stats(ctx, [B, B], [B, B], 1, (2, 0, 3, 1))
    assert s["max_fan_in"] > 1 and s["bulk_tiles"] == 0          # recoupling sums: LSU form with fan-in
    ctx.close()


def test_su2_plans_get_the_table_driven_form():
    """plans with recoupling sums (no bulk form) carry the table-driven form: full tiles + remainder regions cover every destination element once, the
    tables stay small next to the tensor; Abelian plans with the bulk form do not build it"""
    import ctypes as C
    from xped_b200.sharded import _handles, _ints
    from xped_b200.capi import check

    def census(ctx, dom, cod, shift, p):
        out = (C.c_int64 * 6)()
        check(ctx.lib.xb_permute_plan_census(ctx.h, _handles(dom), len(dom), _handles(cod), len(cod), shift, _ints(p), len(p), out))
        return dict(dst=out[0], reads=out[1], tasks=out[2], srcs=out[3], lut_tiles=out[4], lut_bytes=out[5])
    ctx = _ctx("SU2")
    a, e = ctx.basis([(1, 3), (2, 3)]), ctx.basis([(1, 17), (2, 31), (3, 37), (4, 29), (5, 18), (6, 9), (7, 3)])   # multiplicities without useful divisors
    for shift, p in [(-1, (0, 2, 3, 5, 1, 4)), (1, (0, 3, 5, 1, 2, 4)), (0, (1, 0, 2, 4, 3, 5))]:
        c = census(ctx, [e, a, a], [e, a, a], shift, p)
        assert c["lut_tiles"] > 0
        assert c["dst"] / c["lut_tiles"] > 1000          # tiles stay near the 2048-element budget in spite of the odd extents
        assert c["lut_bytes"] < 0.1 * 8 * c["dst"]       # deduplicated tables
    ctx.close()
    ctx = _ctx("U1")
    B = ctx.basis([(q, 8) for q in range(-2, 3)])
    c = census(ctx, [B, B], [B, B], 0, (1, 0, 2, 3))
    assert c["lut_tiles"] == 0                           # bulk form available
    ctx.close()
