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
