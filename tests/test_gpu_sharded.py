"""GPU parity of the sector-sharded path through the C ABI.

(1) One GPU, all ranks of a world of 2..4 simulated in one process: every rank's pack / local / remote launches run on the
    device, the alltoallv is emulated by device-to-device slice copies, the assembled result is compared with the oracle's
    permute of the unsharded tensor (1e-12) and bit-exactly with the unsharded GPU permute.
(2) Real process group: 2 ranks over NCCL (skipped with fewer than 2 GPUs): ShardedTensor permute / multiply / contract / norms.
"""
import ctypes as C
import os
import socket

import numpy as np
import pytest

import xped_oracle as xo

import xped_b200 as xb
from xped_b200.capi import check
from xped_b200.sharded import PermuteExchangePlan, default_owner, layout_structure
from xped_b200.tensor import Tensor, _ints

from helpers import SPECS, obasis, pair
from test_sharded_cpu import CASES, PERMS

pytestmark = pytest.mark.gpu


def _handles(bases):
    return (C.c_void_p * max(len(bases), 1))(*[b.h for b in bases])


def make_shard(ctx, dom, cod, st, owner, rank, seed):
    mine = [i for i, o in enumerate(owner) if o == rank]
    flat = [x for i in mine for x in st[0][i]]
    h = C.c_void_p()
    check(ctx.lib.xb_tensor_create_sectors(ctx.h, _handles(dom), len(dom), _handles(cod), len(cod), _ints(flat), len(mine), C.byref(h)))
    t = Tensor(ctx, h)
    offs = [st[3][i] for i in mine]
    check(ctx.lib.xb_tensor_set_random_at(t.h, seed, (C.c_int64 * max(len(offs), 1))(*offs)))
    return t


def run_simulated(ctx, dom, cod, st, shift, p, world, owner_in, owner_out, seed, adjoint=False):
    import torch
    lib = ctx.lib
    plans = [PermuteExchangePlan(ctx, dom, cod, shift, p, owner_in, owner_out, r, world, adjoint) for r in range(world)]
    shards = [make_shard(ctx, dom, cod, st, owner_in, r, seed) for r in range(world)]
    sends = [torch.zeros(max(sum(pl.send_counts), 1), dtype=torch.float64, device="cuda") for pl in plans]
    for r in range(world):
        check(lib.xb_dist_permute_pack(ctx.h, plans[r].h, shards[r].h, C.c_void_p(sends[r].data_ptr())))
    ctx.sync()
    out_legs = plans[0].out_legs
    ost = layout_structure(ctx, out_legs[:plans[0].out_rank], out_legs[plans[0].out_rank:])
    full = np.full(sum(r * c for r, c in zip(ost[1], ost[2])), np.nan)
    for r in range(world):
        recv = torch.zeros(max(sum(plans[r].recv_counts), 1), dtype=torch.float64, device="cuda")
        pos = 0
        for q in range(world):
            o, n = sum(plans[q].send_counts[:r]), plans[q].send_counts[r]
            recv[pos:pos + n] = sends[q][o:o + n]
            pos += n
        torch.cuda.synchronize()
        h = C.c_void_p()
        check(lib.xb_dist_permute_local(ctx.h, plans[r].h, shards[r].h, C.byref(h)))
        out = Tensor(ctx, h)
        check(lib.xb_dist_permute_remote(ctx.h, plans[r].h, shards[r].h, C.c_void_p(recv.data_ptr()), out.h))
        data = out.download()
        pos = 0
        for i, o in enumerate(plans[r].owner_out):
            if o == r:
                n = ost[1][i] * ost[2][i]
                full[ost[3][i]:ost[3][i] + n] = data[pos:pos + n]
                pos += n
        assert pos == data.size
    return full, ost, plans


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("world", [2, 3, 4])
def test_sharded_permute_simulated_ranks(name, world):
    spec, dom_e, cod_e = CASES[name]
    ctx = xb.Context(spec)
    osym = xo.Symmetry.from_spec(spec)
    gt, ot = pair(ctx, osym, dom_e, cod_e, 21)
    dom, cod = gt.uncoupledDomain(), gt.uncoupledCodomain()
    st = layout_structure(ctx, dom, cod)
    assert st == ot.structure()
    owner_in = default_owner(ctx, st[1], st[2], world)
    moved = 0
    for shift, p in PERMS:
        ref = ot.permute(shift, p)
        full, ost, plans = run_simulated(ctx, dom, cod, st, shift, p, world, owner_in, None, 21)
        assert ost == ref.structure()
        assert not np.isnan(full).any()
        assert np.linalg.norm(full - ref.data) / max(np.linalg.norm(ref.data), 1e-300) < 1e-12
        assert np.array_equal(full, gt.permute(shift, p).download()), "sharded and unsharded permute differ bitwise"
        moved += sum(sum(pl.send_counts) for pl in plans)
    assert moved > 0
    ctx.close()


def test_sharded_permute_adj_and_redistribute():
    spec, dom_e, cod_e = CASES["SU2"]
    ctx = xb.Context(spec)
    osym = xo.Symmetry.from_spec(spec)
    gt, ot = pair(ctx, osym, dom_e, cod_e, 4)
    dom, cod = gt.uncoupledDomain(), gt.uncoupledCodomain()
    st = layout_structure(ctx, dom, cod)
    world = 3
    owner_in = [(2 * i + 1) % world for i in range(len(st[0]))]
    ref = ot.permute_adj(1, (2, 0, 3, 1))
    full, ost, _ = run_simulated(ctx, dom, cod, st, 1, (2, 0, 3, 1), world, owner_in, None, 4, adjoint=True)
    assert ost == ref.structure()
    assert np.linalg.norm(full - ref.data) / np.linalg.norm(ref.data) < 1e-12
    owner_out = [(i + 1) % world for i in range(len(st[0]))]
    full, ost, _ = run_simulated(ctx, dom, cod, st, 0, (0, 1, 2, 3), world, owner_in, owner_out, 4)
    assert np.array_equal(full, ot.data)
    ctx.close()


# ----------------------------------------------------------------------------------------------- real NCCL world
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _products(ctx, comm, osym, name):
    """operator*, lazy adjoint operands and contract on sharded operands (same cases as test_gpu_parity.test_multiply_adjoint_contract)"""
    from xped_b200.sharded import ShardedTensor
    from xped_oracle.rng import uniform_fill
    from helpers import rand_entries
    rng = np.random.default_rng(9)
    B, Cc, D = (rand_entries(SPECS[name], n, rng) for n in (9, 8, 7))

    def both(dom_e, cod_e, seed):
        g = ShardedTensor.zeros(ctx, comm, [ctx.basis(e) for e in dom_e], [ctx.basis(e) for e in cod_e]).setRandom(seed)
        o = xo.Tensor(osym, [obasis(osym, e) for e in dom_e], [obasis(osym, e) for e in cod_e])
        o.data = uniform_fill(seed, o.plainSize())
        return g, o

    def err(g, o, what):
        assert g.structure() == o.structure(), what
        return float(np.linalg.norm(g.gather() - o.data) / max(np.linalg.norm(o.data), 1e-300))

    ga, oa = both([B, Cc], [Cc, D], 1)
    gb, ob = both([Cc, D], [D, B], 2)
    worst = err(ga @ gb, oa @ ob, "operator*")
    worst = max(worst, err(ga.multiply(ga, False, True), oa @ oa.adjoint(), "A * A^dagger"))
    worst = max(worst, err(gb.multiply(ga, True, True), ob.adjoint() @ oa.adjoint(), "B^dagger * A^dagger"))
    worst = max(worst, err(ga.contract([-1, -2, 2, 1], gb, [2, 1, -3, -4], 1), oa.contract([-1, -2, 2, 1], ob, [2, 1, -3, -4], 1), "contract<2 legs>"))
    worst = max(worst, err(ga.contract([1, -1, 2, -2], gb, [2, -3, -4, 1], 1), oa.contract([1, -1, 2, -2], ob, [2, -3, -4, 1], 1), "contract<crossed>"))
    worst = max(worst, err(ga.contract([-3, -1, -4, 1], gb, [-2, 1, -5, -6], 4), oa.contract([-3, -1, -4, 1], ob, [-2, 1, -5, -6], 4), "contract<1 leg>"))
    return worst


def _nccl_worker(rank, world, port, out, peer=False):
    import sys
    import torch
    import torch.distributed as dist
    here = os.path.dirname(os.path.abspath(__file__))
    for pth in (os.path.dirname(here), os.path.join(os.path.dirname(here), "oracle"), here):
        if pth not in sys.path:
            sys.path.insert(0, pth)
    from xped_b200.sharded import Comm, ShardedTensor, enable_peer_memory
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        worst = 0.0
        for name in ("U1", "SU2"):
            spec, dom_e, cod_e = CASES[name]
            ctx = xb.Context(spec, device=rank)
            osym = xo.Symmetry.from_spec(spec)
            comm = Comm()
            if peer:   # CUDA IPC arenas: fused permute + exchange launches for the Abelian case, row-panel products; SU(2) plans (fan-in > 1) keep the NCCL form
                enable_peer_memory(ctx, comm, 1 << 28, name=f"/xb2test_{port}_{name}")
                ctx._panel_rows_min = 2
            dom, cod = [ctx.basis(e) for e in dom_e], [ctx.basis(e) for e in cod_e]
            ot = xo.Tensor(osym, [obasis(osym, e) for e in dom_e], [obasis(osym, e) for e in cod_e])
            from xped_oracle.rng import uniform_fill
            ot.data = uniform_fill(8, ot.plainSize())
            T = ShardedTensor.zeros(ctx, comm, dom, cod).setRandom(8)
            assert T.structure() == ot.structure()
            assert np.array_equal(T.gather(), ot.data)
            for shift, p in PERMS:
                r, orr = T.permute(shift, p), ot.permute(shift, p)
                assert r.structure() == orr.structure()
                worst = max(worst, float(np.linalg.norm(r.gather() - orr.data) / np.linalg.norm(orr.data)))
            worst = max(worst, _products(ctx, comm, osym, name))
            worst = max(worst, abs(T.squaredNorm() - ot.squaredNorm()) / ot.squaredNorm())   # (T * T^dagger).trace(): weighted by the sector degeneracy
            worst = max(worst, abs(T.maxNorm() - float(np.abs(ot.data).max())))
            ctx.close()
        res = torch.tensor([worst], dtype=torch.float64, device="cuda")
        dist.all_reduce(res, op=dist.ReduceOp.MAX)
        if rank == 0:
            out.put(float(res.item()))
    finally:
        dist.destroy_process_group()


def test_sharded_tensor_over_nccl_world2():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctxmp = mp.get_context("spawn")
    out = ctxmp.Queue()
    port = _free_port()
    procs = [ctxmp.Process(target=_nccl_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    assert out.get(timeout=10) < 1e-12


def test_sharded_tensor_peer_memory_world2():
    """the same checks with the peer-memory form of the exchange (csrc/dist.cu): permutes bit-exact / 1e-12 vs the oracle, products incl. row panels"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctxmp = mp.get_context("spawn")
    out = ctxmp.Queue()
    port = _free_port()
    procs = [ctxmp.Process(target=_nccl_worker, args=(r, 2, port, out, True)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    assert out.get(timeout=10) < 1e-12
