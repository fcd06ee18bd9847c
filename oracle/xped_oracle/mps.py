"""TEST INFRASTRUCTURE ONLY -- CPU oracle of the MPS transfer-matrix contractions.

Restates src/MPS/MpsContractions.cpp:11-51 (contract_L) and :54-105 (contract_R) loop by loop: sectors of Bra / Ket,
domain trees (Qin, sigma) -> Q, the dims[1] row slices of each tree's sub-matrix, `optimal_prod_add`
(src/Interfaces/MatrixInterface_Eigen_impl.cpp:180-196) as a plain triple product.  Pinned by the reference's own
identity tests/mps_tests.hpp:19-47: contract_L/R must equal the permute + operator* formulation.
Also the SVD sweep steps Mps::leftSweepStep / rightSweepStep (src/MPS/Mps.cpp:181-229) and dot (src/MPS/MpsAlgebra.cpp:14-51);
the reference tests hold no values for them ("parity unpinned"): pinned by invariants (norm conservation, orthonormality).
"""
from __future__ import annotations

import numpy as np

from .tensor import Tensor


def contract_L(Bold: Tensor, Bra: Tensor, Ket: Tensor) -> Tensor:
    sym = Bold.sym
    Bnew = Tensor(sym, [Bra.uncoupled_codomain[0]], [Ket.uncoupled_codomain[0]])
    for i, Q in enumerate(Bra.sector):
        dimQ = Bra.block(i).shape[1]
        if Q not in Ket.dict:
            continue
        k = Ket.dict[Q]
        Mtmp = np.zeros((dimQ, dimQ))
        for tree in Bra.domain.tree(Q):
            r0b, r0k = Bra.domain.leftOffset(tree), Ket.domain.leftOffset(tree)
            Mbra = Bra.block(i)[r0b:r0b + tree.dim, :]
            Mket = Ket.block(k)[r0k:r0k + tree.dim, :]
            Qin = tree.q_uncoupled[0]
            if Qin not in Bold.dict:
                continue
            Bo = Bold.block(Bold.dict[Qin])
            d0 = tree.dims[0]
            for s in range(tree.dims[1]):
                Mtmp += Mbra[s * d0:(s + 1) * d0, :].conj().T @ Bo @ Mket[s * d0:(s + 1) * d0, :]
        Bnew.block(Bnew.dict[Q])[...] += Mtmp
    return Bnew


def contract_R(Bold: Tensor, Bra: Tensor, Ket: Tensor) -> Tensor:
    sym = Bold.sym
    Bnew = Tensor(sym, [Ket.uncoupled_domain[0]], [Bra.uncoupled_domain[0]])
    for i, Q in enumerate(Ket.sector):
        if Q not in Bold.dict or Q not in Bra.dict:
            continue
        Bo = Bold.block(Bold.dict[Q])
        ib = Bra.dict[Q]
        for tree in Ket.domain.tree(Q):
            Qin = tree.q_uncoupled[0]
            d0 = tree.dims[0]
            Mtmp = np.zeros((d0, d0))
            r0b, r0k = Bra.domain.leftOffset(tree), Ket.domain.leftOffset(tree)
            Mbra = Bra.block(ib)[r0b:r0b + tree.dim, :]
            Mket = Ket.block(i)[r0k:r0k + tree.dim, :]
            for s in range(tree.dims[1]):
                Mtmp += Mket[s * d0:(s + 1) * d0, :] @ Bo @ Mbra[s * d0:(s + 1) * d0, :].conj().T
            Mtmp *= sym.coeff_rightOrtho(Q, Qin)
            Bnew.block(Bnew.dict[Qin])[...] += Mtmp
    return Bnew


def left_sweep_step(Ac, loc, max_Nsv=10000, eps_svd=1.0e-10, DISCARD_U=False):
    """Mps::leftSweepStep, DMRG::BROOM::SVD (src/MPS/Mps.cpp:181-202); Ac is the list of site tensors, modified in place."""
    U, S, Vd, info = Ac[loc].permute(+1, (0, 2, 1)).tSVD(max_Nsv, eps_svd, False, return_info=True)
    left = U @ S
    Ac[loc] = Vd.permute(-1, (0, 2, 1))
    if loc != 0 and not DISCARD_U:
        Ac[loc - 1] = Ac[loc - 1] @ left
    return info


def right_sweep_step(Ac, loc, max_Nsv=10000, eps_svd=1.0e-10, DISCARD_V=False):
    """Mps::rightSweepStep, DMRG::BROOM::SVD (src/MPS/Mps.cpp:204-229)."""
    U, S, Vd, info = Ac[loc].tSVD(max_Nsv, eps_svd, False, return_info=True)
    right = S @ Vd
    Ac[loc] = U
    if loc != len(Ac) - 1 and not DISCARD_V:
        Ac[loc + 1] = (right @ Ac[loc + 1].permute(+1, (0, 1, 2))).permute(-1, (0, 1, 2))
    return info


def dot(BraAc, KetAc, direction="RIGHT"):
    """dot (src/MPS/MpsAlgebra.cpp:14-51)."""
    sym = KetAc[0].sym
    if direction == "RIGHT":
        B = Tensor.Identity(sym, [KetAc[0].uncoupled_domain[0]], [BraAc[0].uncoupled_domain[0]])
        for l in range(len(BraAc)):
            B = contract_L(B, BraAc[l], KetAc[l])
        return B.norm()
    B = Tensor.Identity(sym, [KetAc[-1].uncoupled_codomain[0]], [BraAc[-1].uncoupled_codomain[0]])
    for l in range(len(BraAc) - 1, -1, -1):
        B = contract_R(B, BraAc[l], KetAc[l])
    return B.norm()


def gen_aux_bases(sym, L, qloc, Qtarget, Minit, Qinit):
    """Mps::gen_maxBasis + gen_auxBasis (src/MPS/Mps.cpp:56-127), Qbasis::intersection / truncate (src/Core/Qbasis.cpp:398-408,278-304)
    restated on oracle Qbasis objects; returns the bond bases as [(q, dim)] lists.  Ties of the distance sort keep their order (the
    reference's std::sort is unstable there: parity unpinned for ties)."""
    from .basis import Qbasis
    if not isinstance(qloc, list):
        qloc = [qloc] * L
    vac, Qtarget = sym.qvacuum(), tuple(Qtarget)
    nq = len(vac)
    mk = lambda entries: Qbasis.from_list(sym, entries)
    ent = lambda b: [(q, d) for q, _, d in b.data]

    def intersection(a, b):
        out = Qbasis(sym)
        for q, _, d in a.data:
            other = [e for e in b.data if e[0] == q]
            if other:
                out.push_back(q, min(d, other[0][2]))
        return out

    maxb = [None] * (L + 1)
    maxb[0], maxb[L] = mk([(vac, 1)]), mk([(Qtarget, 1)])
    fl, fr = [None] * (L - 1), [None] * (L - 1)
    fl[0] = maxb[0].combine(qloc[0]).forgetHistory()
    for l in range(1, L - 1):
        fl[l] = fl[l - 1].combine(qloc[l]).forgetHistory()
    fr[L - 2] = maxb[L].combine(qloc[L - 1]).forgetHistory()
    for l in range(L - 2, 0, -1):
        fr[l - 1] = fr[l].combine(qloc[l]).forgetHistory()
    for l in range(L - 1):
        maxb[l + 1] = intersection(fl[l], fr[l])
    qaux = []
    for l in range(L + 1):
        entries = ent(maxb[l])
        qs = [q for q, _ in entries]
        if len(qs) > Qinit:
            mean = [Qtarget[k] * l / L for k in range(nq)]
            qmin, qmax = min(qs), max(qs)
            dist = lambda q: sum((((q[k] - mean[k]) / (qmax[k] - qmin[k])) ** 2 if qmax[k] != qmin[k] else 0.0) for k in range(nq)) ** 0.5
            kept = set(sorted(qs, key=dist)[:Qinit])
            entries = [(q, d) for q, d in entries if q in kept]
        if Minit < sum(d for _, d in entries):
            n = len(entries)
            D, rem = Minit // n, Minit % n
            entries = [(q, min(d, D + rem if q == vac else D)) for q, d in entries]
        qaux.append(entries)
    return qaux
