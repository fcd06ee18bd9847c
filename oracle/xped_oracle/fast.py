"""TEST / BASELINE INFRASTRUCTURE ONLY -- optimised execution of the oracle for the CPU baseline of bench.py (never imported by the product).

`install()` swaps Tensor.permute / Tensor.__matmul__ for versions that keep the oracle's arithmetic but remove the Python interpreter from the
hot loops, so that the timed CPU arm is what an optimised Eigen + OpenMP + BLAS build of the reference does:

* permute: the (source tree pair -> target tree pair, coefficient) list of Tensor::permute_impl (src/Core/Tensor.cpp:83-214) is computed ONCE per
  (tensor structure, shift, permutation) by the plain oracle code and cached -- the reference's XPED_CACHE_PERMUTE_OUTPUT LRU cache of
  treepair::permute (src/Core/treepair.cpp:170-176,211-215) plays that role -- and applied by the compiled strided-transpose kernel
  oracle/csrc/xped_oracle_kernels.c under OpenMP (one group of items per destination sub-block);
* operator*: the result's structure (coupled bases, fusion trees) is cached per operand structure; the per-sector products are numpy / BLAS GEMMs
  (src/Core/TensorBase.cpp:262-279 calls Eigen/BLAS per block as well).

tests/test_oracle_fast.py checks both against the plain oracle on every symmetry."""
from __future__ import annotations

import ctypes as C
import os
import struct
import subprocess

import numpy as np

from . import tensor as _t
from .tensor import Tensor

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE = os.path.dirname(_HERE)
_LIB = None
MAXL = 8
REC = 4 + 3 * MAXL


def build():
    """gcc -O3 -fopenmp of oracle/csrc -> oracle/_build (also done by __graft_entry__.build())."""
    root = os.path.dirname(_ORACLE)
    subprocess.run(["make", "-C", root, "-f", os.path.join("oracle", "Makefile")], check=True, capture_output=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_ORACLE, "_build", "libxped_oracle_kernels.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.xo_permute_apply.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        _LIB.xo_permute_apply.restype = None
        assert _LIB.xo_rec_size() == REC and _LIB.xo_max_legs() == MAXL
        _LIB.xo_set_threads(int(os.environ.get("XO_PERMUTE_THREADS", "1")))
    return _LIB


def _basis_sig(b):
    return (tuple((tuple(q), n, d) for q, n, d in b.data), bool(b.CONJ), tuple(tuple(t.IS_DUAL) for q, _, _ in b.data for t in b.trees[q]))


def _sig(t: Tensor):
    return (t.sym.name(), tuple(_basis_sig(b) for b in t.uncoupled_domain), tuple(_basis_sig(b) for b in t.uncoupled_codomain), tuple(tuple(q) for q in t.sector))


def _strides(rows, dt, ct):
    """element strides of the legs of one sub-block inside its column-major sector matrix (Tensor.cpp:962-981)"""
    out, s = [], 1
    for d in dt.dims:
        out.append(s)
        s *= d
    s = rows
    for d in ct.dims:
        out.append(s)
        s *= d
    return out


def _fresh(tpl: Tensor) -> Tensor:
    """zero tensor with the template's structure (shared immutable metadata, new storage)"""
    t = Tensor(tpl.sym, tpl.uncoupled_domain, tpl.uncoupled_codomain, tpl.dtype, resize=False, _coupled=(tpl.domain, tpl.codomain))
    t.sector, t.dict, t.offsets = tpl.sector, tpl.dict, tpl.offsets
    t.data = np.zeros(tpl.data.size, dtype=tpl.dtype)
    return t


class _Plan:
    def __init__(self, t: Tensor, shift, pi):
        R = t.rank
        n = R + t.corank
        if n > MAXL:
            raise ValueError("too many legs for the compiled kernel")
        intra = shift == 0 and all(x < R for x in pi[:R]) and all(x >= R for x in pi[R:])
        if intra:
            pd, pc = list(pi[:R]), [x - R for x in pi[R:]]
            out = Tensor(t.sym, [t.uncoupled_domain[k] for k in pd], [t.uncoupled_codomain[k] for k in pc], t.dtype)
        else:
            nd, nc = t._new_bases(shift, pi)
            out = Tensor(t.sym, nd, nc, t.dtype)
        self.template = out
        groups = {}
        for i, q in enumerate(t.sector):
            rows_i = t.shape_of(i)[0]
            for dt in t.domain.tree(q):
                r0 = t.domain.leftOffset(dt)
                pdt = dt.permute(t.sym, pd) if intra else None
                for ct in t.codomain.tree(q):
                    c0 = t.codomain.leftOffset(ct)
                    src_off = t.offsets[i] + r0 + c0 * rows_i
                    sdims = list(dt.dims) + list(ct.dims)
                    sstr = _strides(rows_i, dt, ct)
                    if intra:
                        pct = ct.permute(t.sym, pc)
                        targets = [((ndt, nct), cd * cc) for ndt, cd in pdt.items() for nct, cc in pct.items()]
                    else:
                        targets = list(t._pairs(shift, pi, dt, ct).items())
                    for (ndt, nct), coeff in targets:
                        if abs(coeff) < 1.0e-10:       # Tensor.cpp:127,193
                            continue
                        j = out.dict[ndt.q_coupled]
                        rows_j = out.shape_of(j)[0]
                        dst_off = out.offsets[j] + out.domain.leftOffset(ndt) + out.codomain.leftOffset(nct) * rows_j
                        dstr = _strides(rows_j, ndt, nct)
                        ddims = list(ndt.dims) + list(nct.dims)
                        assert ddims == [sdims[pi[m]] for m in range(n)]
                        rec = [src_off, dst_off, n, struct.unpack("q", struct.pack("d", float(coeff)))[0]]
                        rec += ddims + [1] * (MAXL - n)
                        rec += [sstr[pi[m]] for m in range(n)] + [0] * (MAXL - n)
                        rec += dstr + [0] * (MAXL - n)
                        groups.setdefault(dst_off, []).append(rec)
        recs, gstart = [], [0]
        for key in sorted(groups):
            recs += groups[key]
            gstart.append(len(recs))
        self.records = np.ascontiguousarray(np.array(recs, dtype=np.int64).reshape(-1, REC))
        self.gstart = np.ascontiguousarray(np.array(gstart, dtype=np.int64))

    def apply(self, t: Tensor):
        out = _fresh(self.template)
        src = np.ascontiguousarray(t.data)
        lib().xo_permute_apply(src.ctypes.data, out.data.ctypes.data, self.records.ctypes.data, self.gstart.ctypes.data, len(self.gstart) - 1)
        return out


_PLANS = {}
_MM = {}
_ORIG = {}


def fast_permute(self: Tensor, shift, pi):
    pi = list(pi)
    R, Cr = self.rank, self.corank
    assert R >= shift and Cr >= -shift and sorted(pi) == list(range(R + Cr))
    if shift == 0 and pi == list(range(R + Cr)):
        return self.copy()
    if self.dtype != np.float64 or R + Cr > MAXL:
        return _ORIG["permute"](self, shift, pi)
    key = (_sig(self), shift, tuple(pi))
    plan = _PLANS.get(key)
    if plan is None:
        plan = _PLANS[key] = _Plan(self, shift, pi)
    return plan.apply(self)


def fast_matmul(self: Tensor, other: Tensor):
    assert self.corank == other.rank
    key = (_sig(self), _sig(other))
    tpl = _MM.get(key)
    if tpl is None:
        tpl = _MM[key] = Tensor(self.sym, self.uncoupled_domain, other.uncoupled_codomain, np.result_type(self.dtype, other.dtype),
                                _coupled=(self.domain, other.codomain))
    out = _fresh(tpl)
    for i, q in enumerate(self.sector):
        j = other.dict.get(q)
        if j is None:
            continue
        out.block(out.dict[q])[...] = self.block(i) @ other.block(j)
    return out


def install():
    if "permute" in _ORIG:
        return
    lib()
    _ORIG["permute"], _ORIG["matmul"] = Tensor.permute, Tensor.__matmul__
    Tensor.permute = fast_permute
    Tensor.__matmul__ = fast_matmul


def uninstall():
    if "permute" not in _ORIG:
        return
    Tensor.permute, Tensor.__matmul__ = _ORIG.pop("permute"), _ORIG.pop("matmul")
    _PLANS.clear()
    _MM.clear()
