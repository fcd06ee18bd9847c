"""Host-side mirror of the reference's MPS sweep path (src/MPS/Mps.cpp:181-229, src/MPS/MpsAlgebra.cpp:14-51) on top of the
C ABI.  Only what the hot path needs: the chain of site tensors Ac[l] (Tensor<2,1>: (aux_l, phys_l | aux_{l+1})), the SVD
sweep steps and `dot`.  The reference's `Mps` constructor (gen_maxBasis, Minit/Qinit) is not mirrored: bases are given."""
from __future__ import annotations

import ctypes as C

from .capi import check
from .tensor import Tensor, contract_L, contract_R


class Mps:
    def __init__(self, sites, max_Nsv=10000, eps_svd=1.0e-10):
        self.Ac = list(sites)
        self.ctx = self.Ac[0].ctx
        self.max_Nsv, self.eps_svd = int(max_Nsv), float(eps_svd)     # include/Xped/MPS/Mps.hpp:104-105
        self.truncWeight = [0.0] * len(self.Ac)
        self.S = [0.0] * len(self.Ac)

    def length(self):
        return len(self.Ac)

    def _step(self, fn, loc, nb):
        a, b = C.c_void_p(), C.c_void_p()
        tw, ent = C.c_double(0.0), C.c_double(0.0)
        check(fn(self.ctx.h, self.Ac[loc].h, self.Ac[nb].h if nb is not None else None, self.max_Nsv, self.eps_svd, C.byref(a), C.byref(b),
                 C.byref(tw), C.byref(ent)))
        self.Ac[loc] = Tensor(self.ctx, a)
        if nb is not None:
            self.Ac[nb] = Tensor(self.ctx, b)
        self.truncWeight[loc] = tw.value
        if loc != len(self.Ac) - 1:
            self.S[loc] = ent.value

    def leftSweepStep(self, loc, DISCARD_U=False):
        """Mps::leftSweepStep(loc, DMRG::BROOM::SVD, DISCARD_U) (Mps.cpp:181-202)."""
        self._step(self.ctx.lib.xb_mps_left_sweep_step, loc, loc - 1 if (loc != 0 and not DISCARD_U) else None)

    def rightSweepStep(self, loc, DISCARD_V=False):
        """Mps::rightSweepStep(loc, DMRG::BROOM::SVD, DISCARD_V) (Mps.cpp:204-229)."""
        self._step(self.ctx.lib.xb_mps_right_sweep_step, loc, loc + 1 if (loc != len(self.Ac) - 1 and not DISCARD_V) else None)


def dot(Bra: Mps, Ket: Mps, direction="RIGHT"):
    """dot(Bra, Ket, DIR) (MpsAlgebra.cpp:14-51): chain of contract_L (or contract_R) starting from an identity boundary; returns B.norm()."""
    assert Bra.length() == Ket.length()
    ctx = Ket.ctx
    if direction == "RIGHT":
        B = ctx.tensor([Ket.Ac[0].leg_basis(0)], [Bra.Ac[0].leg_basis(0)]).setIdentity()
        for l in range(Bra.length()):
            B = contract_L(B, Bra.Ac[l], Ket.Ac[l])
        return B.norm()
    B = ctx.tensor([Ket.Ac[-1].leg_basis(2)], [Bra.Ac[-1].leg_basis(2)]).setIdentity()
    for l in range(Bra.length() - 1, -1, -1):
        B = contract_R(B, Bra.Ac[l], Ket.Ac[l])
    return B.norm()


# ----------------------------------------------------------------------------------------------- auxiliary bases of a random MPS
def _vacuum(ctx):
    return tuple(1 if kind == "SU2" else 0 for kind, _, _ in ctx.spec if kind != "U0")


def gen_aux_bases(ctx, L, qloc, Qtarget, Minit, Qinit):
    """Mps::gen_maxBasis + Mps::gen_auxBasis (src/MPS/Mps.cpp:56-127): the bond bases qaux[0..L] of an open chain with local bases
    `qloc` (one [(q, dim)] list, used on every site, or a list of L of them) and total quantum number `Qtarget`: all sectors reachable
    from both ends, reduced to the `Qinit` sectors closest to the linear interpolation 0 -> Qtarget and to at most `Minit` states
    (Qbasis::truncate, src/Core/Qbasis.cpp:278-304).  Host logic on [(q, dim)] lists; the coupled bases come from the C library."""
    if qloc and not isinstance(qloc[0], list):
        qloc = [qloc] * L
    vac = _vacuum(ctx)
    Qtarget = tuple(Qtarget)
    nq = len(vac)

    def combine(a, b):
        return [(tuple(q) if isinstance(q, (tuple, list)) else (q,), d) for q, d in ctx.basis(a).combine(ctx.basis(b)).entries()]

    def intersection(a, b):   # Qbasis::intersection (Qbasis.cpp:398-408): order of `a`, min of the dimensions
        db = dict(b)
        return [(q, min(d, db[q])) for q, d in a if q in db]

    max_basis = [None] * (L + 1)
    max_basis[0] = [(vac, 1)]
    max_basis[L] = [(Qtarget, 1)]
    from_left, from_right = [None] * (L - 1), [None] * (L - 1)
    from_left[0] = combine(max_basis[0], qloc[0])
    for l in range(1, L - 1):
        from_left[l] = combine(from_left[l - 1], qloc[l])
    from_right[L - 2] = combine(max_basis[L], qloc[L - 1])
    for l in range(L - 2, 0, -1):
        from_right[l - 1] = combine(from_right[l], qloc[l])
    for l in range(L - 1):
        max_basis[l + 1] = intersection(from_left[l], from_right[l])

    qaux = []
    for l in range(L + 1):
        entries = list(max_basis[l])
        qs = [q for q, _ in entries]
        if len(qs) > Qinit:   # keep the Qinit sectors closest to the mean (ties keep their order: the reference's std::sort is unstable there)
            mean = [Qtarget[k] * l / L for k in range(nq)]
            qmin, qmax = min(qs), max(qs)

            def dist(q):
                acc = 0.0
                for k in range(nq):
                    delta = qmax[k] - qmin[k]
                    acc += ((q[k] - mean[k]) / delta) ** 2 if delta != 0 else 0.0
                return acc ** 0.5
            kept = set(sorted(qs, key=dist)[:Qinit])
            entries = [(q, d) for q, d in entries if q in kept]
        if Minit < sum(d for _, d in entries):   # Qbasis::truncate: Minit / Nq states per sector, the remainder goes to the vacuum sector
            n = len(entries)
            assert Minit >= n, "cannot truncate the basis and keep at least one state per block"
            D, rem = Minit // n, Minit % n
            entries = [(q, min(d, D + rem if q == vac else D)) for q, d in entries]
        qaux.append(entries)
    return qaux


def random_mps(ctx, L, qloc, Qtarget, Minit, Qinit, seed=0, max_Nsv=10000, eps_svd=1.0e-10):
    """Mps(L, qloc, Qtarget, Mmax, Nqmax, seed) (src/MPS/Mps.cpp:22-41): random site tensors on the generated bond bases (counter-based
    fill, seed + site index, instead of the reference's mt19937)."""
    if qloc and not isinstance(qloc[0], list):
        qloc = [qloc] * L
    qaux = gen_aux_bases(ctx, L, qloc, Qtarget, Minit, Qinit)
    sites = [ctx.tensor([ctx.basis(qaux[l]), ctx.basis(qloc[l])], [ctx.basis(qaux[l + 1])]).setRandom(seed + l) for l in range(L)]
    return Mps(sites, max_Nsv=max_Nsv, eps_svd=eps_svd)
