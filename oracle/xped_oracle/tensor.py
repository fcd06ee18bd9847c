"""TEST INFRASTRUCTURE ONLY -- CPU oracle (never imported by the product).

Block-sparse symmetric tensor with the reference's contiguous storage layout and the numerics of
  src/Core/Tensor.cpp:83-146 (permute_impl, intra-side), :148-214 (permute_impl<shift>),
  :258-333 (permute_adj), :335-355 (permute dispatch), :357-377 (twist), :478-633 (tSVD),
  :757-780 (tQR), :782-803 (teigh), :962-981 (subBlock), :1133-1310 (plainTensor)
  src/Core/TensorBase.cpp:27-80 (trace, maxNorm, squaredNorm), :243-288 (operator*)
  include/Xped/Core/storage/StorageType_contiguous.hpp:79-97 (resize), :161-171 (push_back)
  include/Xped/Core/AdjointOp.hpp:26-69, DiagCoeffUnaryOp.hpp, CoeffUnaryOp.hpp
  include/Xped/Core/Tensor.hpp:328-351 (contract), include/Xped/Util/Constfct.hpp:55-109 (get_permutations)
Dense per-block kernels come from numpy (OpenBLAS gemm, LAPACK gesdd / geqrf / syevd): the reference
delegates the same calls to Eigen+BLAS/LAPACKE (cmake/CompilerOptions.cmake:217-222).
"""
from __future__ import annotations

import math

import numpy as np

from .basis import FusionTree, Qbasis, build_FusionTree, treepair_permute
from .symmetry import Symmetry

_PERMUTE_CACHE = {}   # (symmetry name, shift, pi, t1, t2) -> {(t1', t2'): coeff}; cf. XPED_CACHE_PERMUTE_OUTPUT (treepair.cpp:170-176)


def get_permutations(a1, rank1, a2, rank2, rankres):
    """Constfct.hpp:55-109."""
    r1, r2 = len(a1), len(a2)
    max1 = max(0, max(a1)) if a1 else 0
    max2 = max(0, max(a2)) if a2 else 0
    assert max1 == max2
    p2 = [list(a2).index(k + 1) for k in range(max2)]
    p2 += [i for i in range(r2) if i not in p2]
    con1 = [list(a1).index(k + 1) for k in range(max1)]
    p1 = [i for i in range(r1) if i not in con1] + con1
    tmp = [x for x in a1 if x < 0] + [x for x in a2 if x < 0]
    pres = sorted(range(len(tmp)), key=lambda i: -tmp[i])
    shift1 = rank1 - r1 + max1
    shift2 = rank2 - max2
    shiftres = r1 - max1 - rankres
    return p1, shift1, p2, shift2, pres, shiftres


class Tensor:
    def __init__(self, sym: Symmetry, domain, codomain, dtype=np.float64, resize=True, _coupled=None):
        self.sym = sym
        self.uncoupled_domain = list(domain)
        self.uncoupled_codomain = list(codomain)
        if _coupled is None:
            self.domain = build_FusionTree(sym, self.uncoupled_domain)
            self.codomain = build_FusionTree(sym, self.uncoupled_codomain)
        else:
            self.domain, self.codomain = _coupled
        self.dtype = dtype
        self.sector = []
        self.dict = {}
        self.offsets = []
        self.data = np.zeros(0, dtype=dtype)
        if resize:
            self.resize()

    # ------------------------------------------------------------------ storage
    @property
    def rank(self): return len(self.uncoupled_domain)
    @property
    def corank(self): return len(self.uncoupled_codomain)

    def resize(self):
        """StorageType_contiguous.hpp:79-97."""
        self.sector, self.dict, self.offsets = [], {}, []
        curr = 0
        for q, _, _ in self.domain.data:
            if self.codomain.IS_PRESENT(q):
                self.dict[q] = len(self.sector)
                self.sector.append(q)
                self.offsets.append(curr)
                curr += self.domain.inner_dim(q) * self.codomain.inner_dim(q)
        self.data = np.zeros(curr, dtype=self.dtype)

    def shape_of(self, i):
        q = self.sector[i]
        return self.domain.inner_dim(q), self.codomain.inner_dim(q)

    def block(self, i):
        r, c = self.shape_of(i)
        return self.data[self.offsets[i]:self.offsets[i] + r * c].reshape((r, c), order="F")

    def push_back(self, q, M):
        """StorageType_contiguous.hpp:161-171."""
        if not self.offsets:
            self.offsets.append(0)
        else:
            r, c = self.shape_of(len(self.sector) - 1)
            self.offsets.append(self.offsets[-1] + r * c)
        self.dict[q] = len(self.sector)
        self.sector.append(q)
        r, c = self.domain.inner_dim(q), self.codomain.inner_dim(q)
        assert M.shape == (r, c), (M.shape, (r, c))
        self.data = np.concatenate([self.data, np.asarray(M, dtype=self.dtype).reshape(-1, order="F")])

    def plainSize(self): return self.data.size

    def copy(self):
        t = Tensor(self.sym, self.uncoupled_domain, self.uncoupled_codomain, self.dtype, resize=False, _coupled=(self.domain, self.codomain))
        t.sector, t.dict, t.offsets = list(self.sector), dict(self.dict), list(self.offsets)
        t.data = self.data.copy()
        return t

    def like(self):
        t = self.copy()
        t.data = np.zeros_like(self.data)
        return t

    def structure(self):
        """(sector q's, rows, cols, offsets) -- what the bit-exact structure tests compare."""
        return ([tuple(q) for q in self.sector], [self.shape_of(i)[0] for i in range(len(self.sector))],
                [self.shape_of(i)[1] for i in range(len(self.sector))], list(self.offsets))

    def tree_pairs(self):
        """[(sector idx, row offset, col offset, dims...)] in the reference's loop order."""
        out = []
        for i, q in enumerate(self.sector):
            for dt in self.domain.tree(q):
                for ct in self.codomain.tree(q):
                    out.append((i, self.domain.leftOffset(dt), self.codomain.leftOffset(ct), dt.dims + ct.dims))
        return out

    # ------------------------------------------------------------------ fillers
    def setRandom(self, rng):
        self.data = rng.uniform(-1.0, 1.0, self.data.size).astype(self.dtype)
        if np.issubdtype(self.dtype, np.complexfloating):
            self.data = self.data + 1j * rng.uniform(-1.0, 1.0, self.data.size)
        return self

    def setConstant(self, v):
        self.data[:] = v
        return self

    def setZero(self):
        self.data[:] = 0
        return self

    def setIdentity(self):
        self.data[:] = 0
        for i in range(len(self.sector)):
            b = self.block(i)
            n = min(b.shape)
            b[np.arange(n), np.arange(n)] = 1.0
        return self

    @staticmethod
    def Identity(sym, domain, codomain):
        return Tensor(sym, domain, codomain).setIdentity()

    # ------------------------------------------------------------------ sub-blocks
    def _sub(self, i, dt: FusionTree, ct: FusionTree):
        r0, c0 = self.domain.leftOffset(dt), self.codomain.leftOffset(ct)
        return self.block(i)[r0:r0 + dt.dim, c0:c0 + ct.dim]

    def subBlock(self, dt, ct):
        """Tensor.cpp:962-981: rank-(R+C) column-major tensor of one tree pair."""
        return self._sub(self.dict[dt.q_coupled], dt, ct).reshape(dt.dims + ct.dims, order="F")

    # ------------------------------------------------------------------ permute
    def _new_bases(self, shift, pi):
        """Tensor.cpp:157-174."""
        R = self.rank
        newR = R - shift
        def pick(k, to_domain):
            if pi[k] > R - 1 or R == 0:
                b = self.uncoupled_codomain[pi[k] - R]
                return b.conj() if to_domain else b
            b = self.uncoupled_domain[pi[k]]
            return b if to_domain else b.conj()
        n = R + self.corank
        return [pick(i, True) for i in range(newR)], [pick(i, False) for i in range(newR, n)]

    def _pairs(self, shift, pi, dt, ct):
        key = (self.sym.name(), shift, tuple(pi), dt, ct, dt.IS_DUAL, ct.IS_DUAL)   # (the symmetry by CONTENT: the id of a collected Symmetry object is reused)
        v = _PERMUTE_CACHE.get(key)
        if v is None:
            v = treepair_permute(self.sym, shift, dt, ct, pi)
            _PERMUTE_CACHE[key] = v
        return v

    def permute(self, shift, pi):
        """Tensor.cpp:335-355 (dispatch)."""
        pi = list(pi)
        R, C = self.rank, self.corank
        assert R >= shift and C >= -shift and sorted(pi) == list(range(R + C))
        if shift == 0 and pi == list(range(R + C)):
            return self.copy()
        newR = R - shift
        if shift == 0 and all(x < R for x in pi[:newR]) and all(x >= R for x in pi[newR:]):
            return self._permute_intra(pi[:R], [x - R for x in pi[R:]])
        return self._permute_pair(shift, pi)

    def _permute_intra(self, pd, pc):
        """Tensor.cpp:83-146."""
        R = self.rank
        out = Tensor(self.sym, [self.uncoupled_domain[k] for k in pd], [self.uncoupled_codomain[k] for k in pc], self.dtype)
        total = list(pd) + [x + R for x in pc]
        for i, q in enumerate(self.sector):
            for dt in self.domain.tree(q):
                pdt = dt.permute(self.sym, pd)
                for ct in self.codomain.tree(q):
                    pct = ct.permute(self.sym, pc)
                    T = self.subBlock(dt, ct).transpose(total)
                    for ndt, cd in pdt.items():
                        for nct, cc in pct.items():
                            if abs(cd * cc) < 1.0e-10:
                                continue
                            j = out.dict[q]
                            out._sub(j, ndt, nct)[...] += (cd * cc) * T.reshape((ndt.dim, nct.dim), order="F")
        return out

    def _permute_pair(self, shift, pi):
        """Tensor.cpp:148-214."""
        nd, nc = self._new_bases(shift, pi)
        out = Tensor(self.sym, nd, nc, self.dtype)
        for i, q in enumerate(self.sector):
            for dt in self.domain.tree(q):
                for ct in self.codomain.tree(q):
                    T = self.subBlock(dt, ct).transpose(pi)
                    for (ndt, nct), coeff in self._pairs(shift, pi, dt, ct).items():
                        if abs(coeff) < 1.0e-10:
                            continue
                        j = out.dict[ndt.q_coupled]
                        out._sub(j, ndt, nct)[...] += coeff * T.reshape((ndt.dim, nct.dim), order="F")
        return out

    def permute_adj(self, shift, pi):
        """Tensor.cpp:258-333: transpose of permute<shift,pi> as a linear map (applied to a tensor of
        the *output* type of permute); gather form over the result's tree pairs."""
        pi = list(pi)
        R, C = self.rank, self.corank
        n = R + C
        if shift == 0 and pi == list(range(n)):
            return self.copy()
        pinv = [0] * n
        for i, x in enumerate(pi):
            pinv[x] = i
        newR = R + shift
        def pick(k, to_domain):
            if pinv[k] > R - 1 or R == 0:
                b = self.uncoupled_codomain[pinv[k] - R]
                return b.conj() if to_domain else b
            b = self.uncoupled_domain[pinv[k]]
            return b if to_domain else b.conj()
        out = Tensor(self.sym, [pick(i, True) for i in range(newR)], [pick(i, False) for i in range(newR, n)], self.dtype)
        for i, q in enumerate(out.sector):
            for dt in out.domain.tree(q):
                for ct in out.codomain.tree(q):
                    view = out._sub(i, dt, ct)
                    for (pdt, pct), coeff in out._pairs(shift, pi, dt, ct).items():
                        if abs(coeff) < 1.0e-10:
                            continue
                        t = self.subBlock(pdt, pct).transpose(pinv)
                        view[...] += coeff * t.reshape((dt.dim, ct.dim), order="F")
        return out

    def twist(self, leg):
        """Tensor.cpp:357-377."""
        if not self.sym.fermionic:
            return self.copy()
        out = self.copy()
        R = self.rank
        for i, q in enumerate(out.sector):
            for dt in out.domain.tree(q):
                for ct in out.codomain.tree(q):
                    c = self.sym.coeff_twist(dt.q_uncoupled[leg]) if leg < R else self.sym.coeff_twist(ct.q_uncoupled[leg - R])
                    out._sub(i, dt, ct)[...] *= c
        return out

    # ------------------------------------------------------------------ algebra
    def adjoint(self):
        """AdjointOp.hpp:26-69 + eval(): blocks pushed in the original sector order."""
        out = Tensor(self.sym, self.uncoupled_codomain, self.uncoupled_domain, self.dtype, resize=False,
                     _coupled=(self.codomain, self.domain))
        for i, q in enumerate(self.sector):
            out.push_back(q, self.block(i).conj().T)
        return out

    def __matmul__(self, other):
        """TensorBase.cpp:243-288 (operator*)."""
        assert self.corank == other.rank
        out = Tensor(self.sym, self.uncoupled_domain, other.uncoupled_codomain, np.result_type(self.dtype, other.dtype),
                     _coupled=(self.domain, other.codomain))
        for i, q in enumerate(self.sector):
            j = other.dict.get(q)
            if j is None:
                continue
            k = out.dict[q]
            out.block(k)[...] += self.block(i) @ other.block(j)
        return out

    def contract(self, a1, other, a2, resrank):
        """Tensor.hpp:328-351."""
        p1, s1, p2, s2, pres, sres = get_permutations(list(a1), self.rank, list(a2), other.rank, resrank)
        return (self.permute(s1, p1) @ other.permute(s2, p2)).permute(sres, pres)

    def scale(self, s):
        out = self.copy()
        out.data *= s
        return out

    def __add__(self, other):
        """CoeffBinaryOp.hpp:44-64: pairs the i-th blocks."""
        assert self.sector == other.sector
        out = self.copy()
        out.data = self.data + other.data
        return out

    def __sub__(self, other):
        assert self.sector == other.sector
        out = self.copy()
        out.data = self.data - other.data
        return out

    def trace(self):
        return sum(np.trace(self.block(i)) * self.sym.degeneracy(q) for i, q in enumerate(self.sector))

    def maxNorm(self):
        return float(np.max(np.abs(self.data))) if self.data.size else 0.0

    def squaredNorm(self):
        return float(np.real((self @ self.adjoint()).trace()))

    def norm(self):
        return math.sqrt(self.squaredNorm())

    def _diag_map(self, fn):
        """DiagCoeffUnaryOp.hpp:48-51: diagonal().unaryExpr(f).asDiagonal(), evaluated."""
        out = Tensor(self.sym, self.uncoupled_domain, self.uncoupled_codomain, self.dtype, resize=False, _coupled=(self.domain, self.codomain))
        for i, q in enumerate(self.sector):
            out.push_back(q, np.diag(fn(np.diag(self.block(i)))))
        return out

    def diag_sqrt(self): return self._diag_map(np.sqrt)
    def diag_inv(self): return self._diag_map(lambda d: 1.0 / d)

    # ------------------------------------------------------------------ decompositions
    def tSVD(self, maxKeep, eps_svd, PRESERVE_MULTIPLETS=True, truncWeight=0.0, entropy=0.0, return_info=False):
        """Tensor.cpp:478-633. Returns (U, S, Vdag[, info]); truncWeight/entropy accumulate like the reference's
        in/out arguments. Ties in the global sort are broken by (sector index, k) -- the reference's std::sort
        is unstable there (documented deviation)."""
        sym = self.sym
        Us, Ss, Vs, allSV = [], [], [], []
        for i, q in enumerate(self.sector):
            u, s, vh = np.linalg.svd(self.block(i), full_matrices=False)
            Us.append(u); Ss.append(s); Vs.append(vh)
            allSV += [(q, float(x), i, k) for k, x in enumerate(s)]
        assert allSV
        first_entry = allSV[0]
        allSV.sort(key=lambda e: (-e[1], e[2], e[3]))
        i = 0   # numerically tied values (<= 1e-12 relative) are ordered by (sector index, k), see DESIGN.md
        while i < len(allSV):
            j = i + 1
            while j < len(allSV) and allSV[j - 1][1] - allSV[j][1] <= 1.0e-12 * allSV[i][1]:
                j += 1
            if j - i > 1:
                allSV[i:j] = sorted(allSV[i:j], key=lambda e: (e[2], e[3]))
            i = j
        for e in allSV[maxKeep:]:
            truncWeight += sym.degeneracy(e[0]) * abs(e[1]) ** 2
        allSV = allSV[:min(maxKeep, len(allSV))]
        if eps_svd > 0.0:
            allSV = [e for e in allSV if not (e[1] < eps_svd)]
        if PRESERVE_MULTIPLETS:
            end = -1
            for i in range(len(allSV) - 1, 0, -1):
                rel = 2 * (allSV[i - 1][1] - allSV[i][1]) / (allSV[i - 1][1] + allSV[i][1])
                if rel > 0.1:
                    end = i
                    break
            if end != -1:
                allSV = allSV[:end]
        if not allSV:
            allSV = [first_entry]
        ordered = {}
        for q, s, _, _ in allSV:
            ordered.setdefault(q, []).append(s)
            entropy += -sym.degeneracy(q) * s * s * math.log(s * s)
        trunc = Qbasis(sym)
        for q in sorted(ordered):            # std::map<qType,...> iteration order
            trunc.push_back(q, len(ordered[q]))
        U = Tensor(sym, self.uncoupled_domain, [trunc], self.dtype, resize=False, _coupled=(self.domain, trunc))
        S = Tensor(sym, [trunc], [trunc], np.float64, resize=False, _coupled=(trunc, trunc))
        Vd = Tensor(sym, [trunc], self.uncoupled_codomain, self.dtype, resize=False, _coupled=(trunc, self.codomain))
        for q in sorted(ordered):
            i = self.dict[q]
            n = len(ordered[q])
            U.push_back(q, Us[i][:, :n])
            S.push_back(q, np.diag(Ss[i][:n]))
            Vd.push_back(q, Vs[i][:n, :])
        if return_info:
            return U, S, Vd, dict(truncWeight=truncWeight, entropy=entropy, kept={q: len(v) for q, v in ordered.items()})
        return U, S, Vd

    def tSVD_backward(self, dU=None, dS=None, dVd=None):
        """Reverse rule of tSVD = full_svd_impl + cutoff_matrices (include/Xped/AD/ADTensor.hpp:253-325 and :356-371).
        dU / dS / dVd are the adjoints of the TRUNCATED factors (or None): cutoff_matrices' reverse pass adds them into
        the left columns / top-left corner / top rows of the full factors' adjoints, then the full-SVD rule gives dA."""
        out = self.like()
        out.data[...] = 0.0
        for i, q in enumerate(self.sector):
            u, s, vh = np.linalg.svd(self.block(i), full_matrices=False)
            r = s.size
            dUf, dSf, dVf = np.zeros_like(u), np.zeros((r, r)), np.zeros_like(vh)
            touched = False
            for src, dst, sl in ((dU, dUf, lambda b: (slice(None), slice(0, b.shape[1]))), (dS, dSf, lambda b: (slice(0, b.shape[0]), slice(0, b.shape[1]))),
                                 (dVd, dVf, lambda b: (slice(0, b.shape[0]), slice(None)))):
                if src is not None and q in src.dict:
                    b = src.block(src.dict[q])
                    dst[sl(b)] += b
                    touched = True
            if not touched:
                continue   # :259-260: sectors absent from the (truncated) S contribute nothing
            s2 = s * s
            Finv = s2[None, :] - s2[:, None]                         # vec_diff: res(i,j) = v(j) - v(i)
            small = np.abs(Finv) < 1.0e-12
            with np.errstate(divide="ignore", invalid="ignore"):
                F = np.where(small, Finv / (Finv * Finv + 1.0e-12), 1.0 / np.where(small, 1.0, Finv))
            S = np.diag(s)
            Vdag_dV = vh @ dVf.conj().T
            J = F * (u.conj().T @ dUf)
            K = F * Vdag_dV
            dA = u @ (dSf + (J + J.conj().T) @ S + S @ (K + K.conj().T)) @ vh
            with np.errstate(divide="ignore"):
                Sinv = np.diag(1.0 / s)
            dA += (np.eye(u.shape[0]) - u @ u.conj().T) @ dUf @ Sinv @ vh
            dA += u @ Sinv @ dVf @ (np.eye(vh.shape[1]) - vh.conj().T @ vh)
            out.block(i)[...] += dA
        return out

    def tQR(self, RETURN_LQ=False):
        """Tensor.cpp:757-780."""
        sym = self.sym
        middle = (self.codomain if self.domain.dim() > self.codomain.dim() else self.domain).forgetHistory()
        Q = Tensor(sym, self.uncoupled_domain, [middle], self.dtype, resize=False, _coupled=(self.domain, middle))
        Rt = Tensor(sym, [middle], self.uncoupled_codomain, self.dtype, resize=False, _coupled=(middle, self.codomain))
        for i, q in enumerate(self.sector):
            b = self.block(i)
            qm, rm = np.linalg.qr(b, mode="complete")
            m, n = b.shape
            eye = np.eye(m, n)
            if RETURN_LQ:
                Q.push_back(q, (qm @ eye).conj().T)
                Rt.push_back(q, (eye.T @ rm).conj().T)
            else:
                Q.push_back(q, qm @ eye)
                Rt.push_back(q, eye.T @ rm)
        return Q, Rt

    def teigh(self):
        """Tensor.cpp:782-803."""
        sym = self.sym
        md, mc = self.domain.forgetHistory(), self.codomain.forgetHistory()
        D = Tensor(sym, [md], [mc], np.float64, resize=False, _coupled=(md, mc))
        V = Tensor(sym, self.uncoupled_domain, [mc], self.dtype, resize=False, _coupled=(self.domain, mc))
        for i, q in enumerate(self.sector):
            w, v = np.linalg.eigh(self.block(i))
            D.push_back(q, np.diag(w))
            V.push_back(q, v)
        return D, V

    # ------------------------------------------------------------------ dense expansion (test pin)
    def plainTensor(self):
        """Tensor.cpp:1133-1310: expand with explicit Clebsch-Gordan tensors. Small cases only."""
        sym = self.sym

        def sorted_copy(b):
            c = b.copy()
            c.sort()
            return c
        sd, sc = sorted_copy(self.domain), sorted_copy(self.codomain)
        sud = [sorted_copy(b) for b in self.uncoupled_domain]
        suc = [sorted_copy(b) for b in self.uncoupled_codomain]
        inner = np.zeros((sd.fullDim(), sc.fullDim()), dtype=self.dtype)
        for i, q in enumerate(self.sector):
            deg = sym.degeneracy(q)
            blk = np.kron(self.block(i), np.eye(deg))
            r0, c0 = sd.full_outer_num(q), sc.full_outer_num(q)
            inner[r0:r0 + blk.shape[0], c0:c0 + blk.shape[1]] += blk

        def unitary(coupled, unc):
            n = len(unc)
            U = np.zeros([b.fullDim() for b in unc] + [coupled.fullDim()], dtype=np.float64)
            for q, _, _ in coupled.data:
                degq = sym.degeneracy(q)
                for tree in coupled.tree(q):
                    dims = [unc[i].inner_dim(tree.q_uncoupled[i]) for i in range(n)]
                    ud = int(np.prod(dims)) if n else 1
                    Tid = np.eye(ud).reshape(dims + [ud], order="F")
                    T = tree.asTensor(sym)
                    # tensorProd(Tid, T): per index, Tid is the slow factor (TensorInterface_Eigen_impl.cpp:157-206)
                    full = np.einsum(Tid, list(range(0, 2 * (n + 1), 2)), T, list(range(1, 2 * (n + 1), 2)), list(range(2 * (n + 1))))
                    full = full.reshape([Tid.shape[k] * T.shape[k] for k in range(n + 1)])
                    sl = []
                    for i in range(n):
                        o = unc[i].full_outer_num(tree.q_uncoupled[i])
                        sl.append(slice(o, o + dims[i] * sym.degeneracy(tree.q_uncoupled[i])))
                    o = coupled.full_outer_num(q) + coupled.leftOffset(tree) * degq
                    sl.append(slice(o, o + full.shape[n]))
                    U[tuple(sl)] = full
            return U
        Ud, Uc = unitary(sd, sud), unitary(sc, suc)
        tmp = np.tensordot(Ud, inner, axes=([self.rank], [0]))
        return np.tensordot(tmp, Uc, axes=([self.rank], [self.corank]))
