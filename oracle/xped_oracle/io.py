"""TEST INFRASTRUCTURE ONLY -- CPU restatement of IO::loadMatlabTensor (include/Xped/IO/Matlab.hpp:9-95): the charge-block tensor format
(per leg a `charges` dataset with the charge values in row 0 and the sector dimensions in row 1, `vdat` = flat block data, `voffs` = rows of
(combined charge index, offset into vdat)).  The HDF5 container is not read here (no h5py / libhdf5 in this image): the arrays are the input."""
from __future__ import annotations

import numpy as np

from .basis import FusionTree, Qbasis
from .tensor import Tensor


def load_charge_blocks(sym, charges, sector_dims, rank, voffs, vdat, conj=None, qn_scale=1, mod_n=None):
    n = len(charges)
    conj = list(conj) if conj is not None else [False] * n

    def charge_of(c):   # :52  (qn_scale * round(c)) % MOD_N
        q = qn_scale * int(round(float(c)))
        if mod_n:
            q %= mod_n
        return (q,)
    legs = []
    for r in range(n):
        b = Qbasis(sym)
        for c, d in zip(charges[r], sector_dims[r]):
            b.push_back(charge_of(c), int(round(float(d))))
        if conj[r]:      # :54-60: SET_CONJ() + IS_DUAL on the trees, charges untouched
            b.CONJ = True
            b.trees = {q: [t.replace(IS_DUAL=(True,)) for t in ts] for q, ts in b.trees.items()}
        b.sort()         # :61
        legs.append(b)
    tmp = Tensor(sym, legs, [])
    raw = np.asarray(vdat, dtype=np.float64)
    vac = sym.qvacuum()
    for bidx, (combined, begin) in enumerate(voffs):
        end = raw.size if bidx == len(voffs) - 1 else voffs[bidx + 1][1]
        div, qs, dims = int(combined), [], []
        for r in range(n):   # get_indices (:17-33): leg 0 fastest
            idx = div % len(charges[r])
            div //= len(charges[r])
            qs.append(charge_of(charges[r][idx]))
            dims.append(int(round(float(sector_dims[r][idx]))))
        inter, acc = [], qs[0]
        for r in range(1, n - 1):   # computeIntermediates (:73)
            acc = sym.reduceSilent(acc, qs[r])[0]
            inter.append(acc)
        tree = FusionTree(tuple(qs), vac, tuple(dims), tuple(inter), tuple(bool(c) for c in conj))
        i = tmp.dict[vac]
        r0 = tmp.domain.leftOffset(tree)
        blk = tmp.block(i)
        blk[r0:r0 + (end - begin), 0] = raw[begin:end]     # setVal(submatrix, i - begin, 0, raw[i])  (:88)
    return tmp.permute(n - rank, list(range(n)))           # :93
