"""Sector sharding of the hot path over the GPUs of one box (BASELINE.json configs[4]).

One process per GPU (`torch.distributed`, NCCL over NVLink).  A `ShardedTensor` is, on every rank, a `Tensor` that holds only
the blocks of the sectors this rank owns, plus the owner map over the canonical sector list, which every rank computes
identically (deterministic LPT greedy in the C library) -- no metadata is ever communicated.

  * `A @ B`, element-wise maps, norms: sectors are independent (src/Core/TensorBase.cpp:262-279) -> purely local launches
    (+ a scalar all-reduce for norms / trace).
  * `permute`: a source sub-block of sector q lands in sector(s) q' of the result (src/Core/Tensor.cpp:198-208), possibly
    on another GPU -> ONE exchange step:  pack kernel -> NCCL alltoallv (`all_to_all_single`) -> gather kernel.  The
    destination sub-blocks whose sources are all local are produced by a second launch that overlaps the exchange.

PyTorch is plumbing here (process group, exchange buffers, the collective); every kernel is in libxped_b200.so.
The reference's own distributed path is the MPI/Cyclops storage backend (include/Xped/Util/Mpi.hpp:17-47,
src/Interfaces/TensorInterface_Cyclops_impl.cpp); it shards dense blocks inside CTF, not sectors.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import check
from .tensor import Context, Qbasis, Tensor, _ints


# ----------------------------------------------------------------------------------------------- host metadata
class PlanningContext(Context):
    """Host-metadata-only context (no GPU needed): bases, layouts, owner maps, exchange plans. Cannot hold tensor data."""

    def __init__(self, spec):
        self.lib = capi.load()
        self.device = -1
        self.spec = [(k, int(n), int(f)) for k, n, f in spec]
        from .tensor import KINDS
        facs = (capi.SymFactor * max(len(self.spec), 1))()
        for i, (k, n, f) in enumerate(self.spec):
            facs[i].kind, facs[i].N, facs[i].fermionic = KINDS[k], n, f
        h = C.c_void_p()
        check(self.lib.xb_ctx_create_planning(C.byref(h), facs, len(self.spec)))
        self.h = h
        self.Nq = self.lib.xb_ctx_nq(h)


def _handles(bases):
    return (C.c_void_p * max(len(bases), 1))(*[b.h for b in bases])


def layout_structure(ctx, domain, codomain):
    """([sector q], rows, cols, offsets) of the canonical Tensor(domain, codomain), without allocating it."""
    n = C.c_int()
    d, c = _handles(domain), _handles(codomain)
    check(ctx.lib.xb_layout_structure(ctx.h, d, len(domain), c, len(codomain), C.byref(n), None, None, None, None))
    nq, ns = ctx.Nq, n.value
    q = (C.c_int * max(ns * nq, 1))()
    r, co, o = ((C.c_int64 * max(ns, 1))() for _ in range(3))
    check(ctx.lib.xb_layout_structure(ctx.h, d, len(domain), c, len(codomain), C.byref(n), q, r, co, o))
    return ([tuple(q[i * nq + k] for k in range(nq)) for i in range(ns)], list(r[:ns]), list(co[:ns]), list(o[:ns]))


def owner_lpt(ctx, cost, world):
    """Deterministic LPT greedy: heaviest sector first onto the least-loaded rank (ties: lowest index / lowest rank)."""
    cost = [float(x) for x in cost]
    out = (C.c_int * max(len(cost), 1))()
    check(ctx.lib.xb_dist_owner_lpt((C.c_double * max(len(cost), 1))(*cost), len(cost), int(world), out))
    return list(out[:len(cost)])


def default_owner(ctx, rows, cols, world):
    """LPT on rows*cols*(rows+cols): the cost of a product or decomposition of the block."""
    return owner_lpt(ctx, [float(r) * float(c) * float(r + c) for r, c in zip(rows, cols)], world)


class PermuteExchangePlan:
    """Exchange plan of permute<shift,p> (or permute_adj) of the canonical tensor on (domain | codomain) for one rank."""

    def __init__(self, ctx, domain, codomain, shift, p, owner_in, owner_out, rank, world, adjoint=False):
        self.ctx, self.rank, self.world = ctx, int(rank), int(world)
        h = C.c_void_p()
        oo = _ints(owner_out) if owner_out is not None else None
        check(ctx.lib.xb_dist_permute_plan(ctx.h, _handles(domain), len(domain), _handles(codomain), len(codomain), int(shift), _ints(p), len(p),
                                           int(bool(adjoint)), _ints(owner_in), oo, len(owner_out) if owner_out is not None else 0, self.rank, self.world,
                                           C.byref(h)))
        self.h = h
        sc, rc = (C.c_int64 * self.world)(), (C.c_int64 * self.world)()
        check(ctx.lib.xb_dist_plan_counts(h, sc, rc))
        self.send_counts, self.recv_counts = list(sc), list(rc)
        n, r, c = C.c_int(), C.c_int(), C.c_int()
        check(ctx.lib.xb_dist_plan_out(h, C.byref(n), None, None, C.byref(r), C.byref(c)))
        own = (C.c_int * max(n.value, 1))()
        legs = (C.c_void_p * max(r.value + c.value, 1))()
        check(ctx.lib.xb_dist_plan_out(h, None, own, legs, None, None))
        self.owner_out = list(own[:n.value])
        self.out_rank, self.out_corank = r.value, c.value
        self.out_legs = [Qbasis(ctx, C.c_void_p(legs[i])) for i in range(r.value + c.value)]
        self._buffers = None

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.ctx.lib.xb_dist_plan_free(self.h)
        except Exception:
            pass

    def local_sectors(self, which):
        """sector labels of this rank's input (0) / output (1) shard"""
        n = C.c_int()
        check(self.ctx.lib.xb_dist_plan_local_sectors(self.h, which, C.byref(n), None))
        nq = self.ctx.Nq
        q = (C.c_int * max(n.value * nq, 1))()
        check(self.ctx.lib.xb_dist_plan_local_sectors(self.h, which, C.byref(n), q))
        return [tuple(q[i * nq + k] for k in range(nq)) for i in range(n.value)]

    def records(self, which):
        """flat int64 work lists (see include/xped_b200.h: xb_dist_plan_records) as a (n, width) array"""
        n = C.c_int64()
        check(self.ctx.lib.xb_dist_plan_records(self.h, which, C.byref(n), None))
        width = {0: 36, 1: 36, 2: 20}[which]
        out = np.zeros((n.value, width), dtype=np.int64)
        if n.value:
            check(self.ctx.lib.xb_dist_plan_records(self.h, which, C.byref(n), out.ctypes.data_as(capi.LP)))
        return out


# ----------------------------------------------------------------------------------------------- process group plumbing
class Comm:
    """The process group of the job (one rank per GPU); world == 1 without torch.distributed."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist if dist.is_available() and dist.is_initialized() else None
        self.group = group
        self.rank = self.dist.get_rank(group) if self.dist else 0
        self.world = self.dist.get_world_size(group) if self.dist else 1

    def all_reduce_scalar(self, v, op="sum", device=None):
        if self.world == 1:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64, device=device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM if op == "sum" else self.dist.ReduceOp.MAX, group=self.group)
        return float(t.item())


def enable_peer_memory(ctx, comm, arena_bytes=8 << 30, name=None):
    """Map every rank's device arena into every other rank (CUDA IPC over NVLink; csrc/dist.cu).  Collective.  Afterwards the tensors of `ctx` live in the
    arena, sharded permutes of Abelian tensors run as fused permute + exchange launches (xb_dist_permute_push) and products are balanced by row panels
    (xb_dist_multiply).  Without it (or for plans with recoupling sums) the pack -> NCCL alltoallv -> gather form is used."""
    import os
    if getattr(ctx, "_peer_memory", False):
        return
    if name is None:
        name = f"/xb2dist_{os.environ.get('MASTER_PORT', '0')}_{os.getuid()}"
    if comm.rank == 0:
        check(ctx.lib.xb_dist_shm_create(name.encode(), comm.world))
    if comm.world > 1:
        comm.dist.barrier(group=comm.group)
    check(ctx.lib.xb_dist_init(ctx.h, comm.rank, comm.world, int(arena_bytes), name.encode()))
    ctx._peer_memory = True


def multiply_panel_assignment(ctx, m, n, k, owner, world, panel_rows_min=256):
    """the deterministic (sector, rank, row0, rows) panel list of xb_dist_multiply (host only)"""
    ns = len(m)
    arr = lambda v: (C.c_int64 * max(ns, 1))(*[int(x) for x in v])
    cnt = C.c_int64(0)
    check(ctx.lib.xb_dist_multiply_assignment(arr(m), arr(n), arr(k), _ints(owner), ns, int(world), int(panel_rows_min), C.byref(cnt), None))
    out = (C.c_int64 * max(4 * cnt.value, 1))()
    check(ctx.lib.xb_dist_multiply_assignment(arr(m), arr(n), arr(k), _ints(owner), ns, int(world), int(panel_rows_min), C.byref(cnt), out))
    return [tuple(out[4 * i:4 * i + 4]) for i in range(cnt.value)]


class ShardedTensor:
    """Tensor<double,Rank,CoRank,Symmetry> with its sectors distributed over the ranks of `comm`."""

    def __init__(self, ctx, comm, domain, codomain, owner, local, structure):
        self.ctx, self.comm = ctx, comm
        self.domain, self.codomain = list(domain), list(codomain)
        self.owner = list(owner)
        self.local = local
        self.sectors, self.rows, self.cols, self.offsets = structure

    # -- construction
    @staticmethod
    def zeros(ctx, comm, domain, codomain, owner=None):
        st = layout_structure(ctx, domain, codomain)
        if owner is None:
            owner = default_owner(ctx, st[1], st[2], comm.world)
        assert len(owner) == len(st[0])
        mine = [q for q, o in zip(st[0], owner) if o == comm.rank]
        flat = [x for q in mine for x in q]
        h = C.c_void_p()
        check(ctx.lib.xb_tensor_create_sectors(ctx.h, _handles(domain), len(domain), _handles(codomain), len(codomain), _ints(flat), len(mine), C.byref(h)))
        return ShardedTensor(ctx, comm, domain, codomain, owner, Tensor(ctx, h), st)

    def my_sectors(self):
        return [i for i, o in enumerate(self.owner) if o == self.comm.rank]

    def setRandom(self, seed):
        """the values the unsharded tensor gets from setRandom(seed): counter-based fill at the global storage offsets"""
        offs = [self.offsets[i] for i in self.my_sectors()]
        check(self.ctx.lib.xb_tensor_set_random_at(self.local.h, int(seed), (C.c_int64 * max(len(offs), 1))(*offs)))
        return self

    def plainSize(self):
        return sum(r * c for r, c in zip(self.rows, self.cols))

    def structure(self):
        return (self.sectors, self.rows, self.cols, self.offsets)

    # -- exchange
    def _plan(self, shift, p, owner_out, adjoint):
        """exchange plan (cached in the C library by layout / permutation / owner maps) + the context's exchange buffers: one
        grow-only send / receive pair per context is enough because every use is ordered on the context stream"""
        import torch
        plan = PermuteExchangePlan(self.ctx, self.domain, self.codomain, shift, p, self.owner, owner_out, self.comm.rank, self.comm.world, adjoint)
        bufs = self.ctx.__dict__.setdefault("_exchange_buffers", [None, None])
        dev = torch.device("cuda", self.ctx.device)
        for k, need in enumerate((sum(plan.send_counts), sum(plan.recv_counts))):
            if bufs[k] is None or bufs[k].numel() < need:
                self.ctx.sync()
                bufs[k] = torch.empty(max(need, 1), dtype=torch.float64, device=dev)
        plan._buffers = (bufs[0], bufs[1])
        return plan

    def _permute(self, shift, p, owner_out, adjoint):
        import torch
        ctx, lib = self.ctx, self.ctx.lib
        if getattr(ctx, "_peer_memory", False):
            # host-side plan record cached per (bases, permutation, owner maps): the launches of a pushed permute take ~0.1-0.5 ms, python must not add to that
            key = (tuple(b.h.value for b in self.domain), tuple(b.h.value for b in self.codomain), int(shift), tuple(p), tuple(self.owner),
                   None if owner_out is None else tuple(owner_out), bool(adjoint))
            cache = ctx.__dict__.setdefault("_push_plans", {})
            ent = cache.get(key)
            if ent is None:
                plan = PermuteExchangePlan(ctx, self.domain, self.codomain, shift, p, self.owner, owner_out, self.comm.rank, self.comm.world, adjoint)
                ok = C.c_int(0)
                check(lib.xb_dist_plan_pushable(plan.h, C.byref(ok)))
                dom, cod = plan.out_legs[:plan.out_rank], plan.out_legs[plan.out_rank:]
                ent = (plan, bool(ok.value), dom, cod, layout_structure(ctx, dom, cod), (self.domain, self.codomain))   # (keeps the keyed bases alive)
                if len(cache) >= 256:
                    cache.clear()
                cache[key] = ent
            plan, pushable, dom, cod, st, _ = ent
            if pushable:   # fused permute + exchange over peer memory: no buffers, no collective call
                h = C.c_void_p()
                check(lib.xb_dist_permute_push(ctx.h, plan.h, self.local.h, C.byref(h)))
                return ShardedTensor(ctx, self.comm, dom, cod, plan.owner_out, Tensor(ctx, h), st)
        plan = self._plan(shift, p, owner_out, adjoint)
        send, recv = plan._buffers
        stream = torch.cuda.ExternalStream(lib.xb_ctx_stream(ctx.h), device=torch.device("cuda", ctx.device))
        work = None
        if self.comm.world > 1:   # every rank enters the collective, even with nothing to send or receive
            check(lib.xb_dist_permute_pack(ctx.h, plan.h, self.local.h, C.c_void_p(send.data_ptr())))
            with torch.cuda.stream(stream):
                ns, nr = sum(plan.send_counts), sum(plan.recv_counts)
                work = self.comm.dist.all_to_all_single(recv[:nr], send[:ns], plan.recv_counts, plan.send_counts, group=self.comm.group, async_op=True)
        h = C.c_void_p()
        check(lib.xb_dist_permute_local(ctx.h, plan.h, self.local.h, C.byref(h)))   # overlaps the exchange
        out_local = Tensor(ctx, h)
        if work is not None:
            with torch.cuda.stream(stream):
                work.wait()   # the context stream now waits for the NCCL stream
        check(lib.xb_dist_permute_remote(ctx.h, plan.h, self.local.h, C.c_void_p(recv.data_ptr()), out_local.h))
        dom, cod = plan.out_legs[:plan.out_rank], plan.out_legs[plan.out_rank:]
        return ShardedTensor(ctx, self.comm, dom, cod, plan.owner_out, out_local, layout_structure(ctx, dom, cod))

    def permute(self, shift, p, owner_out=None):
        """Tensor::permute<shift,p...>() (src/Core/Tensor.cpp:335-355) on the shards; owner_out: owner map of the result's sectors"""
        return self._permute(shift, p, owner_out, False)

    def permute_adj(self, shift, p, owner_out=None):
        return self._permute(shift, p, owner_out, True)

    def redistribute(self, owner):
        """same tensor, other owner map (identity permutation, exchange only)"""
        n = len(self.domain) + len(self.codomain)
        return self._permute(0, list(range(n)), list(owner), False)

    # -- products
    def multiply(self, other, adj_self=False, adj_other=False):
        """TensorBase::operator* (src/Core/TensorBase.cpp:243-288), operands optionally read as lazy adjoints: sectors are
        independent, no communication as long as both operands agree on the owner of every common sector (otherwise `other` is
        redistributed first)."""
        ctx = self.ctx
        mine = dict(zip(self.sectors, self.owner))
        theirs = dict(zip(other.sectors, other.owner))
        if any(q in mine and mine[q] != o for q, o in theirs.items()):
            other = other.redistribute([mine.get(q, o) for q, o in theirs.items()])
            theirs = dict(zip(other.sectors, other.owner))
        dom = self.codomain if adj_self else self.domain
        cod = other.domain if adj_other else other.codomain
        st = layout_structure(ctx, dom, cod)
        if getattr(ctx, "_peer_memory", False) and not adj_self and not adj_other:
            # row-panel balanced product: heavy sectors are cut into panels computed by other ranks (operands pulled, result panels pushed over NVLink)
            own = (C.c_int * max(len(st[0]), 1))()
            h = C.c_void_p()
            check(ctx.lib.xb_dist_multiply(ctx.h, self.local.h, _ints(self.owner), other.local.h, _ints(other.owner), int(getattr(ctx, "_panel_rows_min", 256)),
                                           C.byref(h), own, len(st[0])))
            return ShardedTensor(ctx, self.comm, dom, cod, list(own[:len(st[0])]), Tensor(ctx, h), st)
        owner = [mine[q] if q in mine else theirs.get(q, i % self.comm.world) for i, q in enumerate(st[0])]
        my_q = [q for q, o in zip(st[0], owner) if o == self.comm.rank]
        flat = [x for q in my_q for x in q]
        h = C.c_void_p()
        check(ctx.lib.xb_multiply_sectors(ctx.h, self.local.h, int(adj_self), other.local.h, int(adj_other), _ints(flat), len(my_q), C.byref(h)))
        return ShardedTensor(ctx, self.comm, dom, cod, owner, Tensor(ctx, h), st)

    __matmul__ = multiply

    def contract(self, a1, other, a2, res_rank):
        """Tensor::contract<a1,a2,ResRank>() (include/Xped/Core/Tensor.hpp:328-351, Util/Constfct.hpp:55-109): two sharded permutes,
        a local product, one sharded permute.  The owner map of both permuted operands is the LPT of the product's flops."""
        l1, l2 = list(a1), list(a2)
        r1, r2 = len(l1), len(l2)
        nc = max([0] + l1)
        assert nc == max([0] + l2), "contract: contracted labels do not match"
        con1 = [l1.index(k) for k in range(1, nc + 1)]
        p1 = [i for i in range(r1) if i not in con1] + con1
        con2 = [l2.index(k) for k in range(1, nc + 1)]
        p2 = con2 + [i for i in range(r2) if i not in con2]
        opn = [x for x in l1 if x < 0] + [x for x in l2 if x < 0]
        pres = sorted(range(len(opn)), key=lambda i: -opn[i])
        shift1 = len(self.domain) - r1 + nc
        shift2 = len(other.domain) - nc
        shiftres = r1 - nc - res_rank
        ap = self.permute(shift1, p1)
        bp = other.permute(shift2, p2, owner_out=None)
        # align on the product's cost: owner(q) = LPT over m_q * k_q * n_q of the common sectors
        rows_a = dict(zip(ap.sectors, zip(ap.rows, ap.cols)))
        cols_b = dict(zip(bp.sectors, zip(bp.rows, bp.cols)))
        cost = [float(rows_a[q][0]) * rows_a[q][1] * cols_b[q][1] if q in cols_b else 0.0 for q in ap.sectors]
        own_a = owner_lpt(self.ctx, cost, self.comm.world)
        if own_a != ap.owner:
            ap = self.permute(shift1, p1, owner_out=own_a)
        want_b = [dict(zip(ap.sectors, ap.owner)).get(q, o) for q, o in zip(bp.sectors, bp.owner)]
        if want_b != bp.owner:
            bp = other.permute(shift2, p2, owner_out=want_b)
        return ap.multiply(bp).permute(shiftres, pres)

    # -- reductions
    def _dev(self):
        import torch
        return torch.device("cuda", self.ctx.device)

    def squaredNorm(self):
        return self.comm.all_reduce_scalar(self.local.squaredNorm(), "sum", self._dev())

    def norm(self):
        return float(np.sqrt(self.squaredNorm()))

    def trace(self):
        return self.comm.all_reduce_scalar(self.local.trace(), "sum", self._dev())

    def maxNorm(self):
        return self.comm.all_reduce_scalar(self.local.maxNorm(), "max", self._dev())

    # -- tests / IO
    def gather(self):
        """the reference's full `m_data` image on every rank (test / IO helper, not a hot path)"""
        full = np.zeros(self.plainSize(), dtype=np.float64)
        mine = self.local.download()
        pieces = [(self.my_sectors(), mine)]
        if self.comm.world > 1:
            box = [None] * self.comm.world
            self.comm.dist.all_gather_object(box, pieces[0], group=self.comm.group)
            pieces = box
        for secs, data in pieces:
            pos = 0
            for i in secs:
                n = self.rows[i] * self.cols[i]
                full[self.offsets[i]:self.offsets[i] + n] = data[pos:pos + n]
                pos += n
        return full


def permute_plan_stats(ctx, domain, codomain, shift, p):
    """host-side statistics of the cached permute plan (no GPU needed): dict with tasks, tiles, max_fan_in, identity, bulk_tiles, bytes"""
    out = (C.c_int64 * 6)()
    check(ctx.lib.xb_permute_plan_stats(ctx.h, _handles(domain), len(domain), _handles(codomain), len(codomain), int(shift), _ints(p), len(p), out))
    return dict(tasks=out[0], tiles=out[1], max_fan_in=out[2], identity=bool(out[3]), bulk_tiles=out[4], bytes=out[5])


def permute_plan_emulate(ctx, domain, codomain, shift, p, src, n_dst, form, adjoint=False):
    """TEST SUPPORT (planning-only contexts): walks the host tables of the permute plan over the `m_data` image `src` like the kernels do
    (form 0: tile list of the LSU form, 1: table-driven form, 2: cp.async.bulk form) and returns the `m_data` image of the result"""
    src = np.ascontiguousarray(src, dtype=np.float64)
    dst = np.full(int(n_dst), np.nan)
    DP = C.POINTER(C.c_double)
    check(ctx.lib.xb_permute_plan_emulate(ctx.h, _handles(domain), len(domain), _handles(codomain), len(codomain), int(shift), _ints(p), len(p), int(bool(adjoint)),
                                          int(form), src.ctypes.data_as(DP), src.size, dst.ctypes.data_as(DP), dst.size))
    return dst


def multiply_plan_emulate(ctx, a_dom, a_cod, adj_a, b_dom, b_cod, adj_b, a_data, b_data, n_c, sm_count=0):
    """TEST SUPPORT (planning-only contexts): applies the host lists of op(A) * op(B) (tile items, split-K jobs) to the `m_data` images of the operands like
    the grouped GEMM kernels do; returns (m_data image of the product, dict of list sizes)"""
    a_data, b_data = np.ascontiguousarray(a_data, dtype=np.float64), np.ascontiguousarray(b_data, dtype=np.float64)
    c = np.full(int(n_c), np.nan)
    DP = C.POINTER(C.c_double)
    counts = (C.c_int64 * 4)()
    check(ctx.lib.xb_multiply_plan_emulate(ctx.h, _handles(a_dom), len(a_dom), _handles(a_cod), len(a_cod), int(bool(adj_a)), _handles(b_dom), len(b_dom),
                                           _handles(b_cod), len(b_cod), int(bool(adj_b)), int(sm_count), a_data.ctypes.data_as(DP), a_data.size,
                                           b_data.ctypes.data_as(DP), b_data.size, c.ctypes.data_as(DP), c.size, counts))
    return c, dict(tiles128=counts[0], tiles64=counts[1], jobs=counts[2], scratch=counts[3])


def contract_plan_emulate(ctx, a_dom, a_cod, labels_a, b_dom, b_cod, labels_b, res_rank, a_data, b_data, n_c, table_form=True):
    """TEST SUPPORT (planning-only contexts): contract<labels_a, labels_b, res_rank> lowered like the device path (two operand permutes, grouped product,
    permute of the result), every plan / list walked on the host; returns the `m_data` image of the result"""
    a_data, b_data = np.ascontiguousarray(a_data, dtype=np.float64), np.ascontiguousarray(b_data, dtype=np.float64)
    c = np.full(int(n_c), np.nan)
    DP = C.POINTER(C.c_double)
    check(ctx.lib.xb_contract_plan_emulate(ctx.h, _handles(a_dom), len(a_dom), _handles(a_cod), len(a_cod), _ints(labels_a), _handles(b_dom), len(b_dom),
                                           _handles(b_cod), len(b_cod), _ints(labels_b), int(res_rank), int(bool(table_form)), a_data.ctypes.data_as(DP), a_data.size,
                                           b_data.ctypes.data_as(DP), b_data.size, c.ctypes.data_as(DP), c.size))
    return c
