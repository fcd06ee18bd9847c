"""TEST INFRASTRUCTURE ONLY -- CPU oracle of the directional CTMRG step (1x1 unit cell, TRank = 2).

Restates src/PEPS/CTM.cpp:45-244 (init, FROM_A), :385-420 (grow_all), :924-1183 (moves), :1185-1337
(get_projectors, PROJECTION::CORNER), :1339-1641 (renormalize_*), :1643-1760 (contractCorner), :528-874 (computeRDM_h / computeRDM_v,
TRank == 2 branch) with the two-site average of src/PEPS/LinearAlgebra.cpp:36-110,
src/PEPS/PEPSContractions.cpp:18-43 (decompose) and src/PEPS/iPEPS.cpp:82-116,200-215 (init_basis, updateAdags).
The reference has no CTM test of its own: parity of this part is pinned only by invariants ("parity unpinned").
"""
from __future__ import annotations

import numpy as np

from .basis import Qbasis
from .tensor import Tensor


def init_basis(sym, D, phys_basis: Qbasis) -> Qbasis:
    """iPEPS::init lambda (src/PEPS/iPEPS.cpp:82-116)."""
    out = Qbasis(sym)
    nq = phys_basis.Nq()
    dim_per_charge = 1 if D < nq else D // nq
    dim_for_vac = 1 if D < nq else dim_per_charge + D % nq
    if nq == 1:
        dim_per_charge = D // 2
        dim_for_vac = D // 2 + D % 2
    inserted = 0
    phys_qs = phys_basis.qs()
    out.push_back(sym.qvacuum(), dim_for_vac)
    inserted += dim_for_vac
    if inserted == D:
        return out
    for Q in sym.lowest_qs():
        if Q in phys_qs:
            out.push_back(Q, dim_per_charge)
            inserted += dim_per_charge
        if inserted == D:
            return out
    for Q in phys_qs:
        if out.IS_PRESENT(Q):
            continue
        out.push_back(Q, dim_per_charge)
        inserted += dim_per_charge
        if inserted == D:
            break
    assert inserted == D
    out.sort()
    return out


def decompose(T1: Tensor, T2: Tensor, chi: int):
    """PEPSContractions.cpp:18-43."""
    prod = T1 @ T2
    U, S, Vd = prod.tSVD(chi, 1.0e-10, False)
    isqrtS = S.diag_sqrt().diag_inv()
    P1 = (isqrtS @ U.adjoint()) @ T1
    P2 = (T2 @ Vd.adjoint()) @ isqrtS
    return P1, P2


def _tw(t: Tensor, *legs):
    for leg in legs:
        t = t.twist(leg)
    return t


def _normalized(t: Tensor):
    return t.scale(1.0 / t.maxNorm())


class CTM:
    UL, UR, LR, LL = 0, 1, 2, 3

    def __init__(self, A: Tensor, chi: int):
        assert A.rank == 2 and A.corank == 3
        self.sym = A.sym
        self.chi = chi
        self.A = A
        self.Adag = A.adjoint().permute(0, (3, 4, 2, 0, 1))   # iPEPS.cpp:200-215
        self._init_from_A()

    def tensors(self):
        return dict(C1=self.C1, C2=self.C2, C3=self.C3, C4=self.C4, T1=self.T1, T2=self.T2, T3=self.T3, T4=self.T4, A=self.A, Adag=self.Adag)

    def _init_from_A(self):
        """CTM.cpp:158-214."""
        sym, A, Ad = self.sym, self.A, self.Adag
        ketL, ketT, ketR, ketB = A.uncoupled_domain[0], A.uncoupled_domain[1], A.uncoupled_codomain[0], A.uncoupled_codomain[1]
        braL, braT, braR, braB = Ad.uncoupled_domain[0], Ad.uncoupled_domain[1], Ad.uncoupled_codomain[0], Ad.uncoupled_codomain[1]
        fused = lambda k, b: k.combine(b).forgetHistory()
        fuse_ll = Tensor.Identity(sym, [fused(ketL, braL)], [ketL, braL])
        fuse_uu = Tensor.Identity(sym, [fused(ketT, braT)], [ketT, braT])
        fuse_rr = Tensor.Identity(sym, [ketR, braR], [fused(ketR, braR)])
        fuse_dd = Tensor.Identity(sym, [ketB, braB], [fused(ketB, braB)])
        fll, fuu = _tw(fuse_ll, 1, 2), _tw(fuse_uu, 1, 2)
        c = lambda a, la, b, lb, r: a.contract(la, b, lb, r)
        self.C1 = c(c(c(A, [1, 2, -1, -2, 3], _tw(Ad, 3, 4), [1, 2, 3, -3, -4], 2), [1, -1, 2, -2], fuse_rr, [1, 2, -3], 2), [1, 2, -2], fuse_dd, [1, 2, -1], 0)
        self.C2 = c(c(c(A, [-1, 1, 2, -2, 3], _tw(Ad, 4), [-3, 1, 3, 2, -4], 2), [1, -1, 2, -2], fll, [-3, 1, 2], 2), [1, 2, -1], fuse_dd, [1, 2, -2], 1)
        self.C3 = c(c(c(A, [-1, -2, 1, 2, 3], Ad, [-3, -4, 3, 1, 2], 2), [1, -1, 2, -2], fll, [-3, 1, 2], 2), [1, 2, -2], fuu, [-1, 1, 2], 2)
        self.C4 = c(c(c(A, [1, -1, -2, 2, 3], _tw(Ad, 3), [1, -3, 3, -4, 2], 2), [1, -1, 2, -2], fuu, [-3, 1, 2], 2), [1, 2, -1], fuse_rr, [1, 2, -2], 1)
        self.T1 = c(c(c(A, [-1, 1, -2, -3, 2], _tw(Ad, 3, 4), [-4, 1, 2, -5, -6], 3), [1, -1, -2, 2, -3, -4], fll, [-5, 1, 2], 4), [1, -3, 2, -4, -1],
                    fuse_rr, [1, 2, -2], 1)
        self.T2 = c(c(c(A, [-1, -2, 1, -3, 2], _tw(Ad, 4), [-4, -5, 2, 1, -6], 3), [-1, 1, -2, -3, 2, -4], fuu, [-5, 1, 2], 4), [-1, 1, -2, 2, -3],
                    fuse_dd, [1, 2, -4], 3)
        self.T3 = c(c(c(A, [-1, -2, -3, 1, 2], _tw(Ad, 3), [-4, -5, 2, -6, 1], 3), [1, -1, -2, 2, -3, -4], fll, [-5, 1, 2], 4), [-1, 1, -2, 2, -3],
                    fuse_rr, [1, 2, -4], 3)
        self.T4 = c(c(c(A, [1, -1, -2, -3, 2], _tw(Ad, 3, 4), [1, -4, 2, -5, -6], 3), [1, -1, -2, 2, -3, -4], fuu, [-5, 1, 2], 4), [-3, 1, -4, 2, -1],
                    fuse_dd, [1, 2, -2], 1)

    def contractCorner(self, corner):
        """CTM.cpp:1643-1760."""
        c = lambda a, la, b, lb, r: a.contract(la, b, lb, r)
        A, Ad = self.A, self.Adag
        if corner == self.UL:
            C1T1 = c(self.C1, [-1, 1], self.T1, [1, -2, -3, -4], 1)
            T4C1T1 = c(self.T4, [1, -1, -2, -3], _tw(C1T1, 0), [1, -4, -5, -6], 3)
            T4C1T1A = c(T4C1T1, [-1, 1, -2, -3, 2, -4], A, [1, 2, -5, -6, -7], 4)
            return c(T4C1T1A, [-1, 1, -4, 2, -5, -2, 3], _tw(Ad, 3, 4), [1, 2, 3, -6, -3], 3)
        if corner == self.LL:
            C4T3 = c(self.C4, [-1, 1], self.T3, [-2, -3, 1, -4], 1)
            T4C4T3 = c(self.T4, [-1, 1, -2, -3], C4T3, [1, -4, -5, -6], 3)
            T4C4T3A = c(T4C4T3, [-1, 1, -2, 2, -3, -4], _tw(A, 3), [1, -5, -6, 2, -7], 4)
            return c(T4C4T3A, [-4, 1, 2, -1, -5, -2, 3], _tw(Ad, 3), [1, -6, 3, -3, 2], 3)
        if corner == self.UR:
            T1C2 = c(self.T1, [-1, 1, -2, -3], self.C2, [1, -4], 3)
            T1C2T2 = c(T1C2, [-1, -2, -3, 1], self.T2, [-4, -5, 1, -6], 3)
            T1C2T2A = c(T1C2T2, [-1, 1, -2, 2, -3, -4], _tw(A, 2), [-5, 1, 2, -6, -7], 4)
            return c(T1C2T2A, [-1, 1, 2, -4, -2, -5, 3], _tw(Ad, 4), [-3, 1, 3, 2, -6], 3)
        if corner == self.LR:
            C3T3 = c(_tw(self.C3, 1), [-1, 1], self.T3, [-2, -3, -4, 1], 1)
            T2C3T3 = c(self.T2, [-1, -2, -3, 1], C3T3, [1, -4, -5, -6], 3)
            T2C3T3A = c(T2C3T3, [1, -1, -2, 2, -3, -4], _tw(A, 2, 3), [-5, -6, 1, 2, -7], 4)
            return c(T2C3T3A, [1, -1, 2, -4, -5, -2, 3], Ad, [-6, -3, 3, 1, 2], 3)
        raise ValueError(corner)

    def get_projectors(self, direction):
        """CTM.cpp:1196-1318."""
        if direction == 0:
            return decompose(_tw(self.contractCorner(self.LL), 3, 4, 5), self.contractCorner(self.UL), self.chi)
        if direction == 1:
            return decompose(self.contractCorner(self.UR), self.contractCorner(self.LR), self.chi)
        if direction == 2:
            return decompose(self.contractCorner(self.UL), self.contractCorner(self.UR), self.chi)
        if direction == 3:
            return decompose(self.contractCorner(self.LR), _tw(self.contractCorner(self.LL), 0, 1, 2), self.chi)
        raise ValueError(direction)

    def move(self, direction):
        c = lambda a, la, b, lb, r: a.contract(la, b, lb, r)
        A, Ad = self.A, self.Adag
        P1, P2 = self.get_projectors(direction)
        if direction == 0:     # CTM.cpp:1339-1419
            C1T1 = c(self.C1, [-1, 1], self.T1, [1, -4, -2, -3], 3)
            C1n = c(_tw(P1, 1, 2, 3), [-1, 1, 2, 3], C1T1, [1, 2, 3, -2], 0)
            C4T3 = c(self.C4, [-2, 1], self.T3, [-3, -4, 1, -1], 1)
            C4n = c(C4T3, [-2, 1, 2, 3], _tw(P2, 0, 1, 2), [1, 2, 3, -1], 1)
            P1T4 = c(_tw(P1, 1, 2, 3), [-1, 1, -2, -3], self.T4, [-4, 1, -5, -6], 3)
            P1T4A = c(P1T4, [-1, 2, -2, -3, 1, -4], A, [1, -5, -6, 2, -7], 4)
            P1T4AAdag = c(P1T4A, [-1, 2, -4, 1, -5, -2, 3], _tw(Ad, 3, 4), [1, -6, 3, -3, 2], 3)
            T4n = c(P1T4AAdag, [-2, -3, -4, 1, 2, 3], _tw(P2, 0, 1, 2), [1, 2, 3, -1], 1)
            self.C1, self.T4, self.C4 = _normalized(C1n), _normalized(T4n), _normalized(C4n)
        elif direction == 1:   # CTM.cpp:1421-1490
            T1C2 = c(self.T1, [-1, 1, -3, -4], self.C2, [1, -2], 1)
            C2n = c(T1C2, [-1, 1, 2, 3], P2, [1, 2, 3, -2], 1)
            C3T3 = c(_tw(self.C3, 1), [-1, 1], self.T3, [-2, -3, -4, 1], 3)
            C3n = c(P1, [-1, 1, 2, 3], C3T3, [1, 2, 3, -2], 2)
            P1T2 = c(P1, [-1, 1, -2, -3], self.T2, [-4, -5, 1, -6], 3)
            P1T2A = c(P1T2, [-1, 1, -2, 2, -3, -4], _tw(A, 2), [-5, 1, 2, -6, -7], 4)
            P1T2AAdag = c(P1T2A, [-1, 1, 2, -4, -2, -5, 3], _tw(Ad, 4), [-3, 1, 3, 2, -6], 3)
            T2n = c(P1T2AAdag, [-3, -1, -2, 1, 2, 3], P2, [1, 2, 3, -4], 3)
            self.C2, self.T2, self.C3 = _normalized(C2n), _normalized(T2n), _normalized(C3n)
        elif direction == 2:   # CTM.cpp:1492-1558
            T4C1 = c(self.T4, [1, -1, -3, -4], _tw(self.C1, 0), [1, -2], 1)
            C1n = c(T4C1, [-1, 1, 2, 3], P2, [1, 2, 3, -2], 0)
            C2T2 = c(self.C2, [-1, 1], self.T2, [-2, -3, 1, -4], 3)
            C2n = c(P1, [-1, 1, 2, 3], C2T2, [1, 2, 3, -2], 1)
            P1T1 = c(P1, [-1, 1, -2, -3], self.T1, [1, -4, -5, -6], 3)
            P1T1A = c(P1T1, [-1, 1, -2, -3, 2, -4], A, [1, 2, -5, -6, -7], 4)
            P1T1AAdag = c(P1T1A, [-1, 1, -4, 2, -5, -2, 3], _tw(Ad, 3, 4), [1, 2, 3, -6, -3], 3)
            T1n = c(P1T1AAdag, [-1, -3, -4, 1, 2, 3], P2, [1, 2, 3, -2], 1)
            self.C1, self.T1, self.C2 = _normalized(C1n), _normalized(T1n), _normalized(C2n)
        elif direction == 3:   # CTM.cpp:1560-1641
            T4C4 = c(self.T4, [-4, 1, -2, -3], self.C4, [1, -1], 3)
            C4n = c(_tw(P1, 1, 2, 3), [-2, 1, 2, 3], T4C4, [1, 2, 3, -1], 1)
            T2C3 = c(self.T2, [-3, -4, -1, 1], self.C3, [1, -2], 1)
            C3n = c(T2C3, [-1, 1, 2, 3], _tw(P2, 0, 1, 2), [1, 2, 3, -2], 2)
            P1T3 = c(_tw(P1, 1, 2, 3), [-1, 1, -2, -3], _tw(self.T3, 0), [-4, -5, -6, 1], 3)
            P1T3A = c(P1T3, [-1, 1, -2, 2, -3, -4], A, [-5, -6, 1, 2, -7], 4)
            P1T3AAdag = c(P1T3A, [-1, 1, 2, -4, -5, -2, 3], _tw(Ad, 3), [-6, -3, 3, 1, 2], 3)
            T3n = c(P1T3AAdag, [-4, -1, -2, 1, 2, 3], _tw(P2, 0, 1, 2), [1, 2, 3, -3], 3)
            self.C4, self.T3, self.C3 = _normalized(C4n), _normalized(T3n), _normalized(C3n)
        else:
            raise ValueError(direction)

    def grow_all(self):
        """CTM.cpp:385-420."""
        for d in range(4):
            self.move(d)


    # ---- reduced density matrices and bond observables (CTM.cpp:528-874, LinearAlgebra.cpp:36-110) ----------------------------
    def _normalised_rdm(self, rho):
        Id2 = Tensor.Identity(self.sym, rho.uncoupled_codomain, rho.uncoupled_domain)
        norm = rho.contract([1, 2, 3, 4], _tw(Id2, 0, 1), [3, 4, 1, 2], 0).trace()     # CTM.cpp:661-667
        return rho.scale(1.0 / norm)

    def computeRDM_h(self):
        """rho_h (CTM.cpp:585-603 for the network, :660-667 for the normalisation): two horizontally neighbouring sites."""
        c = lambda a, la, b, lb, r: a.contract(la, b, lb, r)
        A, Ad = self.A, self.Adag
        C1T1 = c(self.C1, [-1, 1], self.T1, [1, -2, -3, -4], 1)
        T4C1T1 = c(self.T4, [1, -1, -2, -3], _tw(C1T1, 0), [1, -4, -5, -6], 3)
        T4C1T1A = c(T4C1T1, [-1, 1, -2, -3, 2, -4], A, [1, 2, -5, -6, -7], 4)
        C4T3 = c(self.C4, [-1, 1], self.T3, [-2, -3, 1, -4], 1)
        C4T3Ad = c(C4T3, [-1, -2, 1, -3], _tw(Ad, 3), [-4, -5, -6, -7, 1], 3)
        left_half = c(T4C1T1A, [1, 2, -1, 3, -2, 4, -3], C4T3Ad, [1, 4, -4, 2, 3, -5, -6], 3)
        T1C2 = c(self.T1, [-1, 1, -2, -3], self.C2, [1, -4], 3)
        T1C2T2 = c(T1C2, [-1, -2, -3, 1], self.T2, [-4, -5, 1, -6], 3)
        AT1C2T2 = c(_tw(A, 1), [-1, 1, 2, -2, -3], T1C2T2, [-4, 1, -5, 2, -6, -7], 3)
        T3C3 = c(self.T3, [-1, -2, -3, 1], self.C3, [-4, 1], 3)
        T3C3Ad = c(T3C3, [-1, 1, -2, -3], Ad, [-4, -5, -6, -7, 1], 3)
        right_half = c(AT1C2T2, [-1, 1, -2, -3, 2, 3, 4], T3C3Ad, [1, -4, 4, -5, 2, -6, 3], 3)
        rho = c(left_half, [1, 2, -3, 3, -1, 4], right_half, [2, -4, 1, 3, 4, -2], 2)
        return self._normalised_rdm(rho)

    def computeRDM_v(self):
        """rho_v (CTM.cpp:753-779 for the network): two vertically neighbouring sites."""
        c = lambda a, la, b, lb, r: a.contract(la, b, lb, r)
        A, Ad = self.A, self.Adag
        C1T1 = c(self.C1, [-1, 1], self.T1, [1, -2, -3, -4], 1)
        T4C1T1 = c(self.T4, [1, -1, -2, -3], _tw(C1T1, 0), [1, -4, -5, -6], 3)
        T4C1T1A = c(T4C1T1, [-1, 1, -2, -3, 2, -4], A, [1, 2, -5, -6, -7], 4)
        C2T2 = c(self.C2, [-1, 1], self.T2, [-2, -3, 1, -4], 1)
        AdC2T2 = c(_tw(Ad, 3, 4), [-1, -2, -3, 1, -4], C2T2, [-5, -6, 1, -7], 4)
        upper_half = c(T4C1T1A, [-1, 1, 2, 3, 4, -2, -3], AdC2T2, [1, 3, -4, -5, 2, 4, -6], 3)
        C4T3 = c(self.C4, [-1, 1], self.T3, [-2, -3, 1, -4], 1)
        T4C4T3 = c(self.T4, [-1, 1, -2, -3], C4T3, [1, -4, -5, -6], 3)
        AT4C4T3 = c(_tw(A, 0), [1, -1, -2, 2, -3], T4C4T3, [-4, 1, -5, 2, -6, -7], 3)
        T2C3 = c(self.T2, [-1, -2, -3, 1], self.C3, [1, -4], 3)
        T2C3Ad = c(T2C3, [-1, 1, -2, -3], Ad, [-4, -5, -6, 1, -7], 3)
        lower_half = c(AT4C4T3, [-1, 1, -2, -3, 2, 3, 4], T2C3Ad, [1, -4, 4, 2, -5, -6, 3], 3)
        rho = c(upper_half, [1, 2, -3, -1, 3, 4], lower_half, [2, -4, 1, 4, 3, -2], 2)
        return self._normalised_rdm(rho)


def rdm_one_site(rho: Tensor) -> Tensor:
    """rho1 from a two-site density matrix (CTM.cpp:675-678): trace over the second site, normalised by its twisted trace."""
    sym = rho.sym
    Id = Tensor.Identity(sym, [rho.uncoupled_codomain[1]], [rho.uncoupled_domain[1]])
    rho1 = rho.contract([-1, 1, -2, 2], _tw(Id, 0), [2, 1], 1)
    return rho1.scale(1.0 / _tw(rho1, 0).trace())


def avg_one_site(rho1: Tensor, op: Tensor) -> float:
    """<O> on a site (LinearAlgebra.cpp:22-24): rho1.twist(0).contract<{1,2},{2,1},0>(O).trace()."""
    return float(_tw(rho1, 0).contract([1, 2], op, [2, 1], 0).trace())


def avg_two_site(rho: Tensor, op: Tensor) -> float:
    """<O> on a bond (LinearAlgebra.cpp:57-60): rho.twist(0).twist(1).contract<{1,2,3,4},{3,4,1,2},0>(O).trace()."""
    return float(_tw(rho, 0, 1).contract([1, 2, 3, 4], op, [3, 4, 1, 2], 0).trace())


def heisenberg_bond_SU2(sym, phys: Qbasis, J=1.0) -> Tensor:
    """S_1 . S_2 for two spin-1/2 sites as a Tensor<2,2> on (phys, phys | phys, phys): the reference builds it as
    J * sqrt(3) * tprod(S^dagger, S) (include/Xped/PEPS/Models/Heisenberg.hpp:47); in the coupled basis it is the scalar
    (S_tot(S_tot+1) - 3/2) / 2 per total-spin sector: -3/4 on the singlet (q = 1), +1/4 on the triplet (q = 3)."""
    H = Tensor(sym, [phys, phys], [phys, phys])
    for i, q in enumerate(H.sector):
        stot = (q[0] - 1) / 2.0
        blk = H.block(i)
        assert blk.shape == (1, 1), "heisenberg_bond_SU2 expects one spin-1/2 multiplet per site"
        blk[...] = J * 0.5 * (stot * (stot + 1.0) - 1.5)
    return H
