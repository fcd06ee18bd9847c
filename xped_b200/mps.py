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
