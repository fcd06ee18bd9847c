#!/usr/bin/env python
"""Generates the committed golden fixtures (run in the build container; the GPU box only reads them):

    python tests/golden/make_golden.py

  wigner.json        6j / 3j symbols from sympy.physics.wigner -- an implementation independent of the oracle and of the
                     product (the reference takes them from GSL 2.7.1, include/Xped/Symmetry/SU2Wrappers.hpp:70-80)
  hotpath.npz        for fixed small bases and counter-based inputs: block structure, permute / permute_adj / operator* /
                     contract outputs, singular values after tSVD, contract_L/R outputs, computed by oracle/xped_oracle
                     (which tests/test_oracle_pins.py pins against the reference's own property / known-answer tests)
"""
import itertools
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import xped_oracle as xo  # noqa: E402
from xped_oracle.mps import contract_L, contract_R  # noqa: E402
from xped_oracle.rng import uniform_fill  # noqa: E402

CASES = {   # name -> (spec, domain entries, codomain entries)
    "U1": ([("U1", 0, 0)], [[((0,), 2), ((1,), 1), ((-1,), 2)], [((0,), 1), ((2,), 2), ((-1,), 1)]], None),
    "SU2": ([("SU2", 0, 0)], [[((1,), 2), ((2,), 3), ((3,), 1)], [((2,), 2), ((3,), 1), ((4,), 1)]], None),
    "fU1xSU2": ([("U1", 0, 1), ("SU2", 0, 0)], [[((0, 1), 2), ((1, 2), 2), ((2, 1), 1)], [((0, 1), 1), ((1, 2), 1), ((-1, 2), 2)]], None),
}
PERMS = [(0, (1, 0, 3, 2)), (1, (2, 0, 3, 1)), (-1, (0, 2, 3, 1)), (2, (3, 2, 1, 0)), (-2, (1, 3, 0, 2)), (0, (0, 1, 3, 2))]


def wigner():
    from sympy import Rational
    from sympy.physics.wigner import wigner_3j, wigner_6j
    out = {"6j": [], "3j": []}
    rng = np.random.default_rng(0)
    seen = set()
    while len(out["6j"]) < 200:
        q = tuple(int(x) for x in rng.integers(1, 7, size=6))   # q = 2S + 1
        if q in seen:
            continue
        seen.add(q)
        j = [Rational(x - 1, 2) for x in q]
        try:
            v = float(wigner_6j(*j))
        except ValueError:
            v = 0.0
        out["6j"].append([list(q), v])
    for q in itertools.product(range(1, 5), repeat=3):
        j = [Rational(x - 1, 2) for x in q]
        for m1 in range(-(q[0] - 1), q[0], 2):
            for m2 in range(-(q[1] - 1), q[1], 2):
                m3 = -(m1 + m2)
                if abs(m3) > q[2] - 1:
                    continue
                try:
                    v = float(wigner_3j(j[0], j[1], j[2], Rational(m1, 2), Rational(m2, 2), Rational(m3, 2)))
                except ValueError:
                    v = 0.0
                out["3j"].append([list(q), [m1, m2, m3], v])
    return out


def main():
    with open(os.path.join(HERE, "wigner.json"), "w") as f:
        json.dump(wigner(), f)
    arrays = {}
    for name, (spec, dom, _) in CASES.items():
        sym = xo.Symmetry.from_spec(spec)
        B, C = (xo.Qbasis.from_list(sym, e) for e in dom)
        t = xo.Tensor(sym, [B, C], [B, C])
        t.data = uniform_fill(3, t.plainSize())
        q, rows, cols, offs = t.structure()
        arrays[f"{name}/structure"] = np.array([list(qq) + [r, c, o] for qq, r, c, o in zip(q, rows, cols, offs)], dtype=np.int64)
        for shift, p in PERMS:
            r = t.permute(shift, p)
            key = f"{name}/permute/{shift}/{''.join(map(str, p))}"
            arrays[key] = r.data
            arrays[key + "/adj"] = r.permute_adj(shift, p).data
        arrays[f"{name}/multiply"] = (t @ t).data
        arrays[f"{name}/contract"] = t.contract([-1, -2, 1, 2], t, [1, 2, -3, -4], 1).data
        _, S, _ = (t @ t).tSVD(5, 1e-10, False)
        arrays[f"{name}/tsvd5/sigma"] = S.data
        arrays[f"{name}/tsvd5/structure"] = np.array([list(qq) + [r] for qq, r in zip(S.structure()[0], S.structure()[1])], dtype=np.int64)
        out = B.combine(C).forgetHistory()
        A1, A2 = xo.Tensor(sym, [B, C], [out]), xo.Tensor(sym, [B, C], [out])
        A1.data, A2.data = uniform_fill(4, A1.plainSize()), uniform_fill(5, A2.plainSize())
        L, R = xo.Tensor(sym, [B], [B]), xo.Tensor(sym, [out], [out])
        L.data, R.data = uniform_fill(6, L.plainSize()), uniform_fill(7, R.plainSize())
        arrays[f"{name}/contract_L"] = contract_L(L, A1, A2).data
        arrays[f"{name}/contract_R"] = contract_R(R, A1, A2).data
    np.savez_compressed(os.path.join(HERE, "hotpath.npz"), **arrays)
    print("wrote", len(arrays), "arrays")


if __name__ == "__main__":
    main()
