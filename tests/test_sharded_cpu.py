"""CPU tests of the N > 1 path: owner maps, exchange plans, alltoallv counts and offsets.

The exchange plan is pure host logic of the C library (planning-only context, no GPU).  Its three work lists are executed
here with numpy (test infrastructure, standing in for the pack / gather kernels), the exchange itself runs either in memory
(all ranks simulated in one process, world = 1..4) or over a real world_size-2 `gloo` process group, and the result is compared
with the oracle's permute of the unsharded tensor.
"""
import os
import socket

import numpy as np
import pytest

import xped_oracle as xo
from xped_oracle.rng import uniform_fill

import xped_b200 as xb
from xped_b200.sharded import PermuteExchangePlan, PlanningContext, default_owner, layout_structure, owner_lpt

from helpers import SPECS, obasis

L = 16


def np_pack(plan, local_in):
    send = np.zeros(sum(plan.send_counts))
    for rec in plan.records(0):
        n = int(rec[3])
        dims, sst = rec[4:4 + n], rec[4 + L:4 + L + n]
        src = np.lib.stride_tricks.as_strided(local_in[rec[2]:], shape=tuple(dims), strides=tuple(8 * sst))
        cnt = int(np.prod(dims))
        send[rec[1]:rec[1] + cnt] = src.reshape(-1, order="F")
    return send


def np_gather(plan, local_in, recv, out_size):
    out = np.full(out_size, np.nan)
    tasks, srcs = plan.records(1), plan.records(2)
    for t in tasks:
        n = int(t[1])
        dims, dst = tuple(t[4:4 + n]), tuple(8 * t[4 + L:4 + L + n])
        acc = np.zeros(dims)
        for s in srcs[t[2]:t[2] + t[3]]:
            buf = recv if s[0] else local_in
            coeff = np.frombuffer(np.int64(s[2]).tobytes(), dtype=np.float64)[0]
            acc += coeff * np.lib.stride_tricks.as_strided(buf[s[1]:], shape=dims, strides=tuple(8 * s[4:4 + n]))
        view = np.lib.stride_tricks.as_strided(out[t[0]:], shape=dims, strides=dst)
        view[...] = acc
    return out


def shard_of(data, st, owner, rank):
    _, rows, cols, offs = st
    return np.concatenate([data[offs[i]:offs[i] + rows[i] * cols[i]] for i in range(len(owner)) if owner[i] == rank] + [np.zeros(0)])


def make_case(name, spec, dom_e, cod_e, seed):
    ctx = PlanningContext(spec)
    osym = xo.Symmetry.from_spec(spec)
    dom, cod = [ctx.basis(e) for e in dom_e], [ctx.basis(e) for e in cod_e]
    ot = xo.Tensor(osym, [obasis(osym, e) for e in dom_e], [obasis(osym, e) for e in cod_e])
    ot.data = uniform_fill(seed, ot.plainSize())
    st = layout_structure(ctx, dom, cod)
    assert st == ot.structure(), "canonical structure differs from the oracle"
    return ctx, dom, cod, ot, st


CASES = {
    "U1": (SPECS["U1"], [[(-1, 2), (0, 3), (1, 2)], [(-1, 1), (0, 2), (1, 2), (2, 1)]], [[(-2, 1), (-1, 2), (0, 2), (1, 1)], [(0, 3), (1, 2), (-1, 1)]]),
    "SU2": (SPECS["SU2"], [[(1, 2), (2, 3), (3, 1)], [(2, 2), (3, 1)]], [[(1, 1), (2, 2), (3, 2)], [(1, 2), (2, 1)]]),
    "fU1xSU2": (SPECS["fU1xSU2"], [[((0, 1), 2), ((1, 2), 2), ((2, 1), 1)], [((0, 1), 1), ((1, 2), 2)]],
                [[((0, 1), 2), ((1, 2), 1), ((2, 1), 2)], [((0, 1), 1), ((1, 2), 1), ((-1, 2), 1)]]),
}
PERMS = [(0, (1, 0, 2, 3)), (1, (2, 0, 3, 1)), (-1, (0, 2, 3, 1)), (2, (3, 2, 1, 0)), (0, (0, 1, 2, 3)), (-2, (1, 3, 0, 2))]


def simulate(ctx, dom, cod, ot, st, shift, p, world, owner_in, owner_out=None, adjoint=False):
    """all ranks in one process: returns the gathered output image"""
    plans = [PermuteExchangePlan(ctx, dom, cod, shift, p, owner_in, owner_out, r, world, adjoint) for r in range(world)]
    shards = [shard_of(ot.data, st, owner_in, r) for r in range(world)]
    sends = [np_pack(plans[r], shards[r]) for r in range(world)]
    for r in range(world):   # alltoallv bookkeeping: what r sends to q is what q expects from r
        for q in range(world):
            assert plans[r].send_counts[q] == plans[q].recv_counts[r]
        assert plans[r].send_counts[r] == 0
        assert plans[r].owner_out == plans[0].owner_out
    out_dom, out_cod = plans[0].out_legs[:plans[0].out_rank], plans[0].out_legs[plans[0].out_rank:]
    ost = layout_structure(ctx, out_dom, out_cod)
    full = np.full(sum(r * c for r, c in zip(ost[1], ost[2])), np.nan)
    for r in range(world):
        recv = np.zeros(sum(plans[r].recv_counts))
        pos = 0
        for q in range(world):
            o = sum(plans[q].send_counts[:r])
            n = plans[q].send_counts[r]
            recv[pos:pos + n] = sends[q][o:o + n]
            pos += n
        mine = [i for i, o in enumerate(plans[r].owner_out) if o == r]
        assert plans[r].local_sectors(1) == [ost[0][i] for i in mine]
        size = sum(ost[1][i] * ost[2][i] for i in mine)
        out = np_gather(plans[r], shards[r], recv, size)
        pos = 0
        for i in mine:
            n = ost[1][i] * ost[2][i]
            full[ost[3][i]:ost[3][i] + n] = out[pos:pos + n]
            pos += n
    return full, ost, plans


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("world", [1, 2, 3, 4])
def test_exchange_plan_matches_oracle_permute(name, world):
    spec, dom_e, cod_e = CASES[name]
    ctx, dom, cod, ot, st = make_case(name, spec, dom_e, cod_e, 11)
    owner_in = default_owner(ctx, st[1], st[2], world)
    moved = 0
    for shift, p in PERMS:
        ref = ot.permute(shift, p)
        full, ost, plans = simulate(ctx, dom, cod, ot, st, shift, p, world, owner_in)
        assert ost == ref.structure()
        assert not np.isnan(full).any(), "some destination element was never written"
        err = np.linalg.norm(full - ref.data) / max(np.linalg.norm(ref.data), 1e-300)
        assert err < 1e-13, (name, world, shift, p, err)
        moved += sum(sum(pl.send_counts) for pl in plans)
    if world > 1:
        assert moved > 0, "no sub-block ever crossed a rank boundary: the exchange path was not exercised"
    else:
        assert moved == 0


def test_permute_adj_and_explicit_owner_maps():
    spec, dom_e, cod_e = CASES["SU2"]
    ctx, dom, cod, ot, st = make_case("SU2", spec, dom_e, cod_e, 5)
    world = 3
    owner_in = [(2 * i + 1) % world for i in range(len(st[0]))]
    for shift, p in [(1, (2, 0, 3, 1)), (0, (1, 0, 3, 2))]:
        ref = ot.permute_adj(shift, p)
        owner_out = [(i * i) % world for i in range(len(ref.structure()[0]))]
        full, ost, plans = simulate(ctx, dom, cod, ot, st, shift, p, world, owner_in, owner_out, adjoint=True)
        assert ost == ref.structure() and plans[0].owner_out == owner_out
        assert np.linalg.norm(full - ref.data) / np.linalg.norm(ref.data) < 1e-13
    # redistribution = identity permutation with another owner map
    owner_out = [(i + 1) % world for i in range(len(st[0]))]
    full, ost, plans = simulate(ctx, dom, cod, ot, st, 0, (0, 1, 2, 3), world, owner_in, owner_out)
    assert ost == st and np.array_equal(full, ot.data)


def test_owner_lpt_is_deterministic_and_balanced():
    ctx = PlanningContext(SPECS["U1"])
    cost = [9.0, 1.0, 8.0, 2.0, 7.0, 3.0, 7.0, 0.0]
    own = owner_lpt(ctx, cost, 3)
    assert own == owner_lpt(ctx, cost, 3)
    load = [sum(c for c, o in zip(cost, own) if o == r) for r in range(3)]
    assert max(load) - min(load) <= max(cost)
    assert own[0] == 0 and own[2] == 1          # heaviest first, lowest rank on ties
    assert owner_lpt(ctx, cost, 1) == [0] * len(cost)
    with pytest.raises(xb.XbError):
        owner_lpt(ctx, cost, 0)


def test_planning_context_cannot_hold_data():
    ctx = PlanningContext(SPECS["U1"])
    b = ctx.basis([(0, 2), (1, 1)])
    with pytest.raises(xb.XbError) as e:
        ctx.tensor([b], [b])
    assert e.value.status == 2   # XB_ERR_NO_DEVICE: no CPU fallback


# ----------------------------------------------------------------------------------------------- real process group (gloo, world 2)
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, name, out):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        spec, dom_e, cod_e = CASES[name]
        ctx, dom, cod, ot, st = make_case(name, spec, dom_e, cod_e, 3)
        owner_in = default_owner(ctx, st[1], st[2], world)
        worst = 0.0
        for shift, p in PERMS[:4]:
            plan = PermuteExchangePlan(ctx, dom, cod, shift, p, owner_in, None, rank, world)
            mine = shard_of(ot.data, st, owner_in, rank)
            send = torch.from_numpy(np_pack(plan, mine))
            recv = torch.zeros(sum(plan.recv_counts), dtype=torch.float64)
            dist.all_to_all_single(recv, send, plan.recv_counts, plan.send_counts)
            ref = ot.permute(shift, p)
            rst = ref.structure()
            idx = [i for i, o in enumerate(plan.owner_out) if o == rank]
            got = np_gather(plan, mine, recv.numpy(), sum(rst[1][i] * rst[2][i] for i in idx))
            want = shard_of(ref.data, rst, plan.owner_out, rank)
            worst = max(worst, float(np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-300)))
        res = torch.tensor([worst], dtype=torch.float64)
        dist.all_reduce(res, op=dist.ReduceOp.MAX)
        if rank == 0:
            out.put(float(res.item()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["U1", "SU2"])
def test_exchange_over_gloo_world2(name):
    import torch.multiprocessing as mp
    ctxmp = mp.get_context("spawn")
    out = ctxmp.Queue()
    port = _free_port()
    procs = [ctxmp.Process(target=_gloo_worker, args=(r, 2, port, name, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert out.get(timeout=10) < 1e-13


@pytest.mark.parametrize("name", ["SU2", "fU1xSU2"])
def test_recoupling_cache_across_multiplicities(name):
    """Plans for layouts that differ only in the sector multiplicities share the label-keyed recoupling cache of the context
    (Context::recoupling_cache): the second and third tensor are planned from cache hits and must still match the oracle."""
    spec, dom_e, cod_e = CASES[name]
    ctx = PlanningContext(spec)
    osym = xo.Symmetry.from_spec(spec)
    for bump, seed in ((0, 21), (1, 22), (3, 23)):
        de = [[(q, d + bump * (1 + (k + i) % 2)) for k, (q, d) in enumerate(e)] for i, e in enumerate(dom_e)]
        ce = [[(q, d + bump * (1 + (k + i + 1) % 2)) for k, (q, d) in enumerate(e)] for i, e in enumerate(cod_e)]
        dom, cod = [ctx.basis(e) for e in de], [ctx.basis(e) for e in ce]
        ot = xo.Tensor(osym, [obasis(osym, e) for e in de], [obasis(osym, e) for e in ce])
        ot.data = uniform_fill(seed, ot.plainSize())
        st = layout_structure(ctx, dom, cod)
        assert st == ot.structure()
        owner = default_owner(ctx, st[1], st[2], 2)
        for shift, p in PERMS:
            for adjoint in (False, True):
                ref = ot.permute_adj(shift, p) if adjoint else ot.permute(shift, p)
                full, ost, _ = simulate(ctx, dom, cod, ot, st, shift, p, 2, owner, adjoint=adjoint)
                assert ost == ref.structure()
                err = np.linalg.norm(full - ref.data) / max(np.linalg.norm(ref.data), 1e-300)
                assert err < 1e-13, (name, bump, shift, p, adjoint, err)
    ctx.close()


def test_push_plans_exist_exactly_for_fan_in_one():
    """the direct-write form of the exchange (xb_dist_permute_push) needs every destination sub-block to have one source: Abelian plans qualify, SU(2)
    plans with recoupling sums do not (they keep the pack / alltoallv / gather form)"""
    import ctypes as C
    from xped_b200.sharded import PermuteExchangePlan, PlanningContext, default_owner, layout_structure
    for name, entries, shift, p, want in [("U1", [(-1, 2), (0, 3), (1, 2)], 1, (2, 0, 3, 1), 1), ("SU2", [(1, 2), (2, 2), (3, 1)], 1, (2, 0, 3, 1), 0),
                                          ("SU2", [(1, 2), (2, 2), (3, 1)], 0, (1, 0, 2, 3), 1)]:
        ctx = PlanningContext(SPECS[name])
        B = ctx.basis(entries)
        st = layout_structure(ctx, [B, B], [B, B])
        owner = default_owner(ctx, st[1], st[2], 3)
        for rank in range(3):
            plan = PermuteExchangePlan(ctx, [B, B], [B, B], shift, p, owner, None, rank, 3)
            ok = C.c_int(-1)
            assert ctx.lib.xb_dist_plan_pushable(plan.h, C.byref(ok)) == 0
            assert ok.value == want, (name, shift, p)
        ctx.close()


def test_row_panel_assignment_covers_and_balances():
    """xb_dist_multiply_assignment: every row of every sector exactly once, panels start on multiples of the minimal panel, the heaviest rank within a few
    percent of the mean even when one sector alone exceeds a rank's share (W4 over 8 ranks: whole-sector LPT reaches 0.54)"""
    from xped_b200.sharded import PlanningContext, multiply_panel_assignment, owner_lpt
    ctx = PlanningContext(SPECS["U1"])
    side = [(9 - abs(q)) * 576 for q in range(-8, 9)]
    for world in (1, 2, 3, 4, 8):
        own = owner_lpt(ctx, [float(x) ** 3 for x in side], world)
        panels = multiply_panel_assignment(ctx, side, side, side, own, world, 256)
        cover, load = {}, [0.0] * world
        for s, r, r0, rows in panels:
            assert 0 <= r < world and rows > 0 and r0 % 256 == 0
            cover.setdefault(s, []).append((r0, rows))
            load[r] += 2.0 * rows * side[s] * side[s]
        for s, m in enumerate(side):
            pos = 0
            for r0, rows in sorted(cover[s]):
                assert r0 == pos
                pos += rows
            assert pos == m
        assert max(load) <= 1.12 * sum(load) / world, (world, load)
    # a skewed owner map: everything on rank 0 -> the others take panels
    panels = multiply_panel_assignment(ctx, side, side, side, [0] * len(side), 4, 256)
    assert {r for _, r, _, _ in panels} == {0, 1, 2, 3}
    ctx.close()
