"""Host-side mirror of the reference interface above the C ABI (Python flavour, used by tests and bench).

Names and argument meaning follow Xped: `Qbasis.push_back/conj/combine`, `Tensor.permute(shift, p)`,
`A * B` / `A @ B` (TensorBase::operator*), `contract(a1, other, a2, res_rank)`, `adjoint()`, `tSVD(maxKeep, eps,
PRESERVE_MULTIPLETS)`, `twist(leg)`, `maxNorm()`, `trace()`, `squaredNorm()` (include/Xped/Core/Tensor.hpp:59-444,
TensorBase.hpp:34-148).  Everything numeric happens in libxped_b200.so on the GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import check

KINDS = {"U0": 0, "U1": 1, "ZN": 2, "SU2": 3}


def _ints(seq):
    arr = (C.c_int * max(len(seq), 1))(*[int(x) for x in seq])
    return arr


class Context:
    """One per GPU; carries the symmetry (spec: list of (kind, N, fermionic), e.g. [("SU2", 0, 0)])."""

    def __init__(self, spec, device=0):
        self.lib = capi.load()
        self.device = int(device)
        self.spec = [(k, int(n), int(f)) for k, n, f in spec]
        facs = (capi.SymFactor * max(len(self.spec), 1))()
        for i, (k, n, f) in enumerate(self.spec):
            facs[i].kind, facs[i].N, facs[i].fermionic = KINDS[k], n, f
        h = C.c_void_p()
        check(self.lib.xb_ctx_create(C.byref(h), device, facs, len(self.spec)))
        self.h = h
        self.Nq = self.lib.xb_ctx_nq(h)

    def sync(self):
        check(self.lib.xb_ctx_sync(self.h))

    def stats(self):
        out = (C.c_int64 * 3)()
        check(self.lib.xb_ctx_stats(self.h, out))
        return {"launches": out[0], "plan_hits": out[1], "plan_misses": out[2]}

    # ---- reverse-mode AD tape (include/Xped/AD/ADTensor.hpp): ops issued between tape_begin() and tape_end() record their reverse rule
    def tape_begin(self):
        check(self.lib.xb_tape_begin(self.h))

    def tape_end(self):
        check(self.lib.xb_tape_end(self.h))

    def tape_seed_trace(self, t, weight=1.0):
        check(self.lib.xb_tape_seed_trace(self.h, t.h, float(weight)))

    def tape_seed(self, t, adj):
        check(self.lib.xb_tape_seed(self.h, t.h, adj.h))

    def tape_backward(self):
        check(self.lib.xb_tape_backward(self.h))

    def tape_grad(self, t):
        g = C.c_void_p()
        check(self.lib.xb_tape_grad(self.h, t.h, C.byref(g)))
        return Tensor(self, g) if g.value else None

    def tape_size(self):
        n = C.c_int64()
        check(self.lib.xb_tape_size(self.h, C.byref(n)))
        return n.value

    def set_int(self, key, value):
        check(self.lib.xb_ctx_set_int(self.h, key.encode(), int(value)))

    def get_int(self, key):
        v = C.c_int64()
        check(self.lib.xb_ctx_get_int(self.h, key.encode(), C.byref(v)))
        return v.value

    def flush_l2(self):
        check(self.lib.xb_flush_l2(self.h))

    def timer_start(self):
        check(self.lib.xb_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_double()
        check(self.lib.xb_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def profile(self, on):
        check(self.lib.xb_profile_enable(self.h, int(bool(on))))

    def profile_read(self):
        """{category: (ms, launches, work)} since the last read; work = bytes (permute) / flops (gemm)."""
        ms, cnt, work = (C.c_double * 4)(), (C.c_int64 * 4)(), (C.c_double * 4)()
        check(self.lib.xb_profile_read(self.h, ms, cnt, work))
        return {n: (ms[i], cnt[i], work[i]) for i, n in enumerate(["permute", "gemm", "tsvd", "other"])}

    def close(self):
        if self.h:
            self.lib.xb_ctx_free(self.h)
            self.h = None

    # -- factories
    def basis(self, entries):
        """entries: [(q, dim)], q an int (Nq == 1) or a tuple of Nq ints; push_back order is kept."""
        return Qbasis.from_entries(self, entries)

    def tensor(self, domain, codomain):
        return Tensor.zeros(self, domain, codomain)


class Qbasis:
    def __init__(self, ctx, h):
        self.ctx, self.h = ctx, h

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.ctx.lib.xb_basis_free(self.h)
        except Exception:
            pass

    @staticmethod
    def from_entries(ctx, entries):
        qs, dims = [], []
        for q, d in entries:
            q = (q,) if isinstance(q, (int, np.integer)) else tuple(q)
            assert len(q) == ctx.Nq
            qs += list(q)
            dims.append(int(d))
        h = C.c_void_p()
        check(ctx.lib.xb_basis_create(ctx.h, _ints(qs), (C.c_int64 * max(len(dims), 1))(*dims), len(dims), C.byref(h)))
        return Qbasis(ctx, h)

    def _unary(self, fn):
        h = C.c_void_p()
        check(fn(self.h, C.byref(h)))
        return Qbasis(self.ctx, h)

    def conj(self): return self._unary(self.ctx.lib.xb_basis_conj)
    def sorted(self): return self._unary(self.ctx.lib.xb_basis_sorted)

    def combine(self, other):
        """combine(other).forgetHistory()"""
        h = C.c_void_p()
        check(self.ctx.lib.xb_basis_combine(self.h, other.h, C.byref(h)))
        return Qbasis(self.ctx, h)

    def info(self):
        n, d, fd, cj = C.c_int(), C.c_int64(), C.c_int64(), C.c_int()
        check(self.ctx.lib.xb_basis_size(self.h, C.byref(n), C.byref(d), C.byref(fd), C.byref(cj)))
        return n.value, d.value, fd.value, bool(cj.value)

    def entries(self):
        n = self.info()[0]
        q = (C.c_int * max(n * self.ctx.Nq, 1))()
        d = (C.c_int64 * max(n, 1))()
        check(self.ctx.lib.xb_basis_get(self.h, q, d))
        nq = self.ctx.Nq
        return [(tuple(q[i * nq + k] for k in range(nq)), d[i]) for i in range(n)]

    def dim(self): return self.info()[1]
    def fullDim(self): return self.info()[2]
    def IS_CONJ(self): return self.info()[3]


class Tensor:
    def __init__(self, ctx, h):
        self.ctx, self.h = ctx, h

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.ctx.lib.xb_tensor_free(self.h)
        except Exception:
            pass

    # -- construction / storage
    @staticmethod
    def zeros(ctx, domain, codomain):
        d = (C.c_void_p * max(len(domain), 1))(*[b.h for b in domain])
        c = (C.c_void_p * max(len(codomain), 1))(*[b.h for b in codomain])
        h = C.c_void_p()
        check(ctx.lib.xb_tensor_create(ctx.h, d, len(domain), c, len(codomain), C.byref(h)))
        return Tensor(ctx, h)

    def _out(self, fn, *args):
        h = C.c_void_p()
        check(fn(*args, C.byref(h)))
        return Tensor(self.ctx, h)

    @property
    def rank(self):
        r, c = C.c_int(), C.c_int()
        check(self.ctx.lib.xb_tensor_rank(self.h, C.byref(r), C.byref(c)))
        return r.value

    @property
    def corank(self):
        r, c = C.c_int(), C.c_int()
        check(self.ctx.lib.xb_tensor_rank(self.h, C.byref(r), C.byref(c)))
        return c.value

    def leg_basis(self, leg):
        h = C.c_void_p()
        check(self.ctx.lib.xb_tensor_leg_basis(self.h, leg, C.byref(h)))
        return Qbasis(self.ctx, h)

    def uncoupledDomain(self): return [self.leg_basis(i) for i in range(self.rank)]
    def uncoupledCodomain(self): return [self.leg_basis(self.rank + i) for i in range(self.corank)]
    def plainSize(self): return self.ctx.lib.xb_tensor_plain_size(self.h)

    def structure(self):
        """([sector q], rows, cols, offsets) -- compared bit-exactly against the oracle."""
        n = self.ctx.lib.xb_tensor_num_sectors(self.h)
        nq = self.ctx.Nq
        q = (C.c_int * max(n * nq, 1))()
        r, c, o = ((C.c_int64 * max(n, 1))() for _ in range(3))
        check(self.ctx.lib.xb_tensor_structure(self.h, q, r, c, o))
        return ([tuple(q[i * nq + k] for k in range(nq)) for i in range(n)], list(r[:n]), list(c[:n]), list(o[:n]))

    def tree_pairs(self):
        n = self.ctx.lib.xb_tensor_num_tree_pairs(self.h)
        buf = (C.c_int64 * max(5 * n, 1))()
        check(self.ctx.lib.xb_tensor_tree_pairs(self.h, buf))
        return np.frombuffer(buf, dtype=np.int64, count=5 * n).reshape(n, 5).copy()

    # ---- sub-blocks of one tree pair (Tensor::subBlock / view / subMatrix, Tensor.hpp:191-216); `pair` = index in tree_pairs() order
    def tree_pair_trees(self, pair):
        """(q_uncoupled per leg, q_intermediates, q_coupled, dims per leg) of the fusion-tree pair."""
        nq, R, Cr = max(self.ctx.Nq, 1), self.rank, self.corank
        ni = max(R - 2, 0) + max(Cr - 2, 0)
        qu, qi, qc = (C.c_int * ((R + Cr) * nq))(), (C.c_int * max(ni * nq, 1))(), (C.c_int * nq)()
        dims = (C.c_int64 * (R + Cr))()
        check(self.ctx.lib.xb_tensor_tree_pair_trees(self.h, int(pair), qu, qi, qc, dims))
        n = self.ctx.Nq
        return ([tuple(qu[i * nq:i * nq + n]) for i in range(R + Cr)], [tuple(qi[i * nq:i * nq + n]) for i in range(ni)], tuple(qc[:n]), list(dims))

    def find_tree_pair(self, q_uncoupled, q_intermediates, q_coupled):
        flat = lambda qs: _ints([x for q in qs for x in (q if isinstance(q, (tuple, list)) else (q,))] or [0])
        return int(self.ctx.lib.xb_tensor_find_tree_pair(self.h, flat(q_uncoupled), flat(q_intermediates), flat([q_coupled])))

    def subBlock(self, pair):
        """Dense copy of the sub-block as an array of shape f1.dims + f2.dims (column-major over the legs, Tensor.cpp:962-981)."""
        dims = self.tree_pair_trees(pair)[3]
        n = int(np.prod(dims)) if dims else 1
        out = np.empty(n, dtype=np.float64)
        check(self.ctx.lib.xb_sub_block(self.h, int(pair), out.ctypes.data_as(C.POINTER(C.c_double)), n))
        return out.reshape(dims, order="F")

    def subMatrix(self, pair):
        dims = self.tree_pair_trees(pair)[3]
        r = int(np.prod(dims[:self.rank])) if self.rank else 1
        return self.subBlock(pair).reshape((r, -1), order="F")

    def setSubBlock(self, pair, block):
        b = np.asfortranarray(np.asarray(block, dtype=np.float64)).reshape(-1, order="F")
        b = np.ascontiguousarray(b)
        check(self.ctx.lib.xb_set_sub_block(self.h, int(pair), b.ctypes.data_as(C.POINTER(C.c_double)), b.size))
        return self

    def view(self, pair):
        """(device pointer, leading dimension, rows, cols) of the sub-block in place (Tensor::view)."""
        ptr, ld, r, c = C.c_void_p(), C.c_int64(), C.c_int64(), C.c_int64()
        check(self.ctx.lib.xb_view(self.h, int(pair), C.byref(ptr), C.byref(ld), C.byref(r), C.byref(c)))
        return ptr.value, ld.value, r.value, c.value

    # ---- state interop: native container (xb_tensor_save / xb_tensor_load) and the charge-block (MATLAB) format of IO::loadMatlabTensor
    def save(self, path):
        check(self.ctx.lib.xb_tensor_save(self.ctx.h, self.h, str(path).encode()))

    @staticmethod
    def load(ctx, path):
        h = C.c_void_p()
        check(ctx.lib.xb_tensor_load(ctx.h, str(path).encode(), C.byref(h)))
        return Tensor(ctx, h)

    @staticmethod
    def from_charge_blocks(ctx, charges, sector_dims, rank, voffs, vdat, conj=None, qn_scale=1):
        """IO::loadMatlabTensor (include/Xped/IO/Matlab.hpp:9-95): charges / sector_dims = per leg the two rows of its `charges` dataset,
        voffs = [(combined charge index, offset into vdat)], vdat = flat block data; rank = number of domain legs of the result."""
        n = len(charges)
        ncharge = _ints([len(c) for c in charges])
        ch = [np.ascontiguousarray(c, dtype=np.float64) for c in charges]
        sd = [np.ascontiguousarray(d, dtype=np.float64) for d in sector_dims]
        DPP = C.POINTER(C.c_double) * n
        chp = DPP(*[a.ctypes.data_as(C.POINTER(C.c_double)) for a in ch])
        sdp = DPP(*[a.ctypes.data_as(C.POINTER(C.c_double)) for a in sd])
        idx = np.ascontiguousarray([v[0] for v in voffs], dtype=np.int64)
        off = np.ascontiguousarray([v[1] for v in voffs], dtype=np.int64)
        vd = np.ascontiguousarray(vdat, dtype=np.float64)
        cj = _ints([int(bool(x)) for x in (conj if conj is not None else [0] * n)])
        h = C.c_void_p()
        check(ctx.lib.xb_tensor_from_charge_blocks(ctx.h, n, int(rank), ncharge, chp, sdp, cj, len(voffs), idx.ctypes.data_as(C.POINTER(C.c_int64)),
                                                   off.ctypes.data_as(C.POINTER(C.c_int64)), vd.ctypes.data_as(C.POINTER(C.c_double)), vd.size, int(qn_scale), C.byref(h)))
        return Tensor(ctx, h)

    def upload(self, data):
        a = np.ascontiguousarray(data, dtype=np.float64)
        check(self.ctx.lib.xb_tensor_upload(self.h, a.ctypes.data_as(capi.DP), a.size))
        return self

    def download(self):
        out = np.empty(self.plainSize(), dtype=np.float64)
        check(self.ctx.lib.xb_tensor_download(self.h, out.ctypes.data_as(capi.DP), out.size))
        return out

    def setZero(self):
        check(self.ctx.lib.xb_tensor_set_zero(self.h)); return self

    def setIdentity(self):
        check(self.ctx.lib.xb_tensor_set_identity(self.h)); return self

    def setConstant(self, v):
        check(self.ctx.lib.xb_tensor_set_constant(self.h, float(v))); return self

    def setRandom(self, seed):
        check(self.ctx.lib.xb_tensor_set_random(self.h, int(seed))); return self

    def clone(self): return self._out(self.ctx.lib.xb_tensor_clone, self.h)

    # -- hot path
    def permute(self, shift, p):
        return self._out(self.ctx.lib.xb_permute, self.ctx.h, self.h, int(shift), _ints(p), len(p))

    def permute_adj(self, shift, p):
        return self._out(self.ctx.lib.xb_permute_adj, self.ctx.h, self.h, int(shift), _ints(p), len(p))

    def multiply(self, other, adj_self=False, adj_other=False):
        return self._out(self.ctx.lib.xb_multiply, self.ctx.h, self.h, int(adj_self), other.h, int(adj_other))

    def __matmul__(self, other): return self.multiply(other)

    def __mul__(self, other):
        if isinstance(other, Tensor):
            return self.multiply(other)
        return self.scale(other)

    __rmul__ = __mul__

    def contract(self, a1, other, a2, res_rank):
        return self._out(self.ctx.lib.xb_contract, self.ctx.h, self.h, _ints(a1), len(a1), other.h, _ints(a2), len(a2), int(res_rank))

    def adjoint(self): return self._out(self.ctx.lib.xb_adjoint, self.ctx.h, self.h)
    def twist(self, leg): return self._out(self.ctx.lib.xb_twist, self.ctx.h, self.h, int(leg))
    def scale(self, s): return self._out(self.ctx.lib.xb_axpby, self.ctx.h, float(s), self.h, 0.0, None)
    def axpby(self, alpha, beta, y): return self._out(self.ctx.lib.xb_axpby, self.ctx.h, float(alpha), self.h, float(beta), y.h)
    def __add__(self, other): return self.axpby(1.0, 1.0, other)
    def __sub__(self, other): return self.axpby(1.0, -1.0, other)
    def normalize_max(self): return self._out(self.ctx.lib.xb_normalize_max, self.ctx.h, self.h)
    def normalize_trace(self, w): return self._out(self.ctx.lib.xb_normalize_trace, self.ctx.h, self.h, w.h)
    def diag_sqrt(self): return self._out(self.ctx.lib.xb_diag_map, self.ctx.h, self.h, 0)
    def diag_inv(self): return self._out(self.ctx.lib.xb_diag_map, self.ctx.h, self.h, 1)

    def _scalar(self, fn):
        v = C.c_double()
        check(fn(self.ctx.h, self.h, C.byref(v)))
        return v.value

    def maxNorm(self): return self._scalar(self.ctx.lib.xb_max_norm)
    def trace(self): return self._scalar(self.ctx.lib.xb_trace)
    def squaredNorm(self): return self._scalar(self.ctx.lib.xb_squared_norm)
    def norm(self): return float(np.sqrt(self.squaredNorm()))

    def tSVD(self, maxKeep, eps_svd, PRESERVE_MULTIPLETS=True, return_info=False):
        u, s, v = C.c_void_p(), C.c_void_p(), C.c_void_p()
        tw, ent = C.c_double(0.0), C.c_double(0.0)
        check(self.ctx.lib.xb_tsvd(self.ctx.h, self.h, int(maxKeep), float(eps_svd), int(PRESERVE_MULTIPLETS), C.byref(u), C.byref(s), C.byref(v),
                                   C.byref(tw), C.byref(ent)))
        out = (Tensor(self.ctx, u), Tensor(self.ctx, s), Tensor(self.ctx, v))
        if return_info:
            return out + (dict(truncWeight=tw.value, entropy=ent.value),)
        return out

    def tSVD_backward(self, U, S, Vdag, dU=None, dS=None, dVdag=None):
        return self._out(self.ctx.lib.xb_tsvd_backward, self.ctx.h, self.h, U.h, S.h, Vdag.h, dU.h if dU else None, dS.h if dS else None,
                         dVdag.h if dVdag else None)

    def _pair(self, fn, *args):
        a, b = C.c_void_p(), C.c_void_p()
        check(fn(*args, C.byref(a), C.byref(b)))
        return Tensor(self.ctx, a), Tensor(self.ctx, b)

    def tQR(self, RETURN_LQ=False): return self._pair(self.ctx.lib.xb_tqr, self.ctx.h, self.h, int(bool(RETURN_LQ)))
    def teigh(self): return self._pair(self.ctx.lib.xb_teigh, self.ctx.h, self.h)

    # -- measurement hooks
    def bench_permute(self, shift, p, reps=10):
        ms, nbytes = C.c_double(), C.c_int64()
        check(self.ctx.lib.xb_bench_permute(self.ctx.h, self.h, int(shift), _ints(p), len(p), int(reps), C.byref(ms), C.byref(nbytes)))
        return ms.value, nbytes.value

    def bench_multiply(self, other, reps=5):
        ms, fl = C.c_double(), C.c_double()
        check(self.ctx.lib.xb_bench_multiply(self.ctx.h, self.h, other.h, int(reps), C.byref(ms), C.byref(fl)))
        return ms.value, fl.value


def decompose(T1, T2, chi):
    """PEPSContractions.cpp:18-43"""
    ctx = T1.ctx
    a, b = C.c_void_p(), C.c_void_p()
    check(ctx.lib.xb_decompose(ctx.h, T1.h, T2.h, int(chi), C.byref(a), C.byref(b)))
    return Tensor(ctx, a), Tensor(ctx, b)


def contract_L(Bold, Bra, Ket):
    ctx = Bold.ctx
    h = C.c_void_p()
    check(ctx.lib.xb_contract_L(ctx.h, Bold.h, Bra.h, Ket.h, C.byref(h)))
    return Tensor(ctx, h)


def contract_R(Bold, Bra, Ket):
    ctx = Bold.ctx
    h = C.c_void_p()
    check(ctx.lib.xb_contract_R(ctx.h, Bold.h, Bra.h, Ket.h, C.byref(h)))
    return Tensor(ctx, h)


class CTM:
    """CTM<double,Symmetry,2> on a 1x1 unit cell (include/Xped/PEPS/CTM.hpp:60-312)."""
    NAMES = ["C1", "C2", "C3", "C4", "T1", "T2", "T3", "T4", "A", "Adag"]

    def __init__(self, A: Tensor, chi: int):
        self.ctx = A.ctx
        h = C.c_void_p()
        check(self.ctx.lib.xb_ctm_create(self.ctx.h, A.h, int(chi), C.byref(h)))
        self.h = h

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.ctx.lib.xb_ctm_free(self.h)
        except Exception:
            pass

    def save(self, path):
        check(self.ctx.lib.xb_ctm_save(self.h, str(path).encode()))

    @staticmethod
    def load(ctx, path):
        h = C.c_void_p()
        check(ctx.lib.xb_ctm_load(ctx.h, str(path).encode(), C.byref(h)))
        obj = CTM.__new__(CTM)
        obj.ctx, obj.h = ctx, h
        return obj

    def grow_all(self): check(self.ctx.lib.xb_ctm_step(self.h))
    def move(self, direction): check(self.ctx.lib.xb_ctm_move(self.h, int(direction)))

    def get(self, name):
        h = C.c_void_p()
        check(self.ctx.lib.xb_ctm_get(self.h, self.NAMES.index(name), C.byref(h)))
        return Tensor(self.ctx, h)

    def set(self, name, t):
        check(self.ctx.lib.xb_ctm_set(self.h, self.NAMES.index(name), t.h))

    def contractCorner(self, corner):
        h = C.c_void_p()
        check(self.ctx.lib.xb_ctm_corner(self.h, int(corner), C.byref(h)))
        return Tensor(self.ctx, h)

    def _rdm(self, which):
        h = C.c_void_p()
        check(self.ctx.lib.xb_ctm_rdm(self.h, which, C.byref(h)))
        return Tensor(self.ctx, h)

    def computeRDM_h(self):
        """normalised rho_h of two horizontally neighbouring sites (CTM::computeRDM_h, src/PEPS/CTM.cpp:528-700)"""
        return self._rdm(0)

    def computeRDM_v(self):
        """normalised rho_v of two vertically neighbouring sites (CTM::computeRDM_v, src/PEPS/CTM.cpp:702-874)"""
        return self._rdm(1)


def ctm_energy_grad(A, chi, steps, op_h, op_v, checkpoint=False, want_grad=True):
    """(E, dE/dA) of the 1x1 iPEPS A after `steps` CTM steps (xb_ctm_energy_grad); dE/dA is None when want_grad is False."""
    e, g = C.c_double(), C.c_void_p()
    check(A.ctx.lib.xb_ctm_energy_grad(A.ctx.h, A.h, int(chi), int(steps), op_h.h, op_v.h, int(bool(checkpoint)), C.byref(e), C.byref(g) if want_grad else None))
    return e.value, (Tensor(A.ctx, g) if want_grad else None)


def avg_two_site(rho, op):
    """avg(rho, O, bond) of a two-site observable (src/PEPS/LinearAlgebra.cpp:57-60)"""
    out = C.c_double()
    check(rho.ctx.lib.xb_avg_two_site(rho.ctx.h, rho.h, op.h, C.byref(out)))
    return out.value


def rdm_one_site(rho):
    """rho1 = normalised partial trace over the second site of a two-site density matrix (src/PEPS/CTM.cpp:675-678)"""
    h = C.c_void_p()
    check(rho.ctx.lib.xb_rdm_one_site(rho.ctx.h, rho.h, C.byref(h)))
    return Tensor(rho.ctx, h)


def avg_one_site(rho1, op):
    """avg(rho1, O) of a one-site observable (src/PEPS/LinearAlgebra.cpp:22-24)"""
    out = C.c_double()
    check(rho1.ctx.lib.xb_avg_one_site(rho1.ctx.h, rho1.h, op.h, C.byref(out)))
    return out.value
