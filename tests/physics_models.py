"""Dense restatements of the model Hamiltonians of the reference's tests/physics.cpp (spin-1/2 J1-J2 chain :84-106, Hubbard chain :261-288),
block-diagonalised by their Abelian charges, so that the golden ground-state energies the reference asserts on `H.data.trim<2>().teigh()`
pin `teigh` (oracle and device) including the sorted sector order `Es.block(k)` refers to.  numpy only: test infrastructure."""
import itertools

import numpy as np


def _kron_at(op, site, n, d):
    mats = [np.eye(d)] * n
    mats[site] = op
    out = mats[0]
    for m in mats[1:]:
        out = np.kron(out, m)
    return out


def spin_chain(L=6, J1=1.0, J2=0.8, Bz=0.5):
    """tests/physics.cpp:84-106: H = sum_{i != j} Js(i,j) (Sz_i Sz_j + (S+_i S-_j + S-_i S+_j) / 2) - Bz sum_i Sz_i with Js = J1/2 (|i-j| = 1),
    J2/2 (|i-j| = 2).  Returns (H, charges) with charges[state] = (2 Sz_total,)."""
    sz = np.diag([0.5, -0.5])
    sp = np.array([[0.0, 1.0], [0.0, 0.0]])
    sm = sp.T
    dim = 2 ** L
    H = np.zeros((dim, dim))
    Sz = [_kron_at(sz, i, L, 2) for i in range(L)]
    Sp = [_kron_at(sp, i, L, 2) for i in range(L)]
    Sm = [_kron_at(sm, i, L, 2) for i in range(L)]
    for i in range(L):
        H -= Bz * Sz[i]
        for j in range(L):
            J = 0.5 * J1 if abs(i - j) == 1 else (0.5 * J2 if abs(i - j) == 2 else 0.0)
            if J != 0.0:
                H += J * (Sz[i] @ Sz[j] + 0.5 * (Sp[i] @ Sm[j] + Sm[i] @ Sp[j]))
    charges = [(int(round(2 * sum(0.5 if b == 0 else -0.5 for b in bits))),) for bits in itertools.product((0, 1), repeat=L)]
    return H, charges


def hubbard_chain(L=4, t1=1.0, t2=0.8, Bz=0.5, mu=3.0, U=8.0):
    """tests/physics.cpp:261-288: H = sum_{i != j} ts(i,j) sum_s c+_{i s} c_{j s} + sum_i (-Bz Sz_i - mu n_i + U d_i), ts = -t1 / -t2 on first /
    second neighbours.  Jordan-Wigner over the modes (site 0 up, site 0 down, site 1 up, ...).  charges[state] = (N, 2 Sz_total)."""
    nm = 2 * L
    dim = 2 ** nm
    a = np.array([[0.0, 1.0], [0.0, 0.0]])   # annihilates |1>
    z = np.diag([1.0, -1.0])

    def c(mode):
        mats = [z] * mode + [a] + [np.eye(2)] * (nm - mode - 1)
        out = mats[0]
        for m in mats[1:]:
            out = np.kron(out, m)
        return out
    cs = [c(m) for m in range(nm)]
    n = [x.T @ x for x in cs]
    H = np.zeros((dim, dim))
    for i in range(L):
        up, dn = 2 * i, 2 * i + 1
        H += -Bz * 0.5 * (n[up] - n[dn]) - mu * (n[up] + n[dn]) + U * (n[up] @ n[dn])
        for j in range(L):
            t = -t1 if abs(i - j) == 1 else (-t2 if abs(i - j) == 2 else 0.0)
            if t != 0.0:
                for s in (0, 1):
                    H += t * cs[2 * i + s].T @ cs[2 * j + s]
    charges = []
    for bits in itertools.product((0, 1), repeat=nm):
        N = sum(bits)
        sz2 = sum(bits[2 * i] - bits[2 * i + 1] for i in range(L))
        charges.append((N, sz2))
    return H, charges


def sector_blocks(H, charges):
    """{charge: dense block} of a Hamiltonian that commutes with the charges; asserts that it does."""
    idx = {}
    for k, q in enumerate(charges):
        idx.setdefault(q, []).append(k)
    blocks = {}
    covered = np.zeros_like(H, dtype=bool)
    for q, ii in idx.items():
        ix = np.ix_(ii, ii)
        blocks[q] = H[ix].copy()
        covered[ix] = True
    assert covered.all() or np.abs(H[~covered]).max() == 0.0, "Hamiltonian does not commute with the charges"
    return blocks


GOLDEN = {
    "spin_u1": (-2.43062856688879, 3),        # tests/physics.cpp:106  Es.block(3)(0, 0)
    "hubbard_u1xu1": (-13.378782979745, 13),  # tests/physics.cpp:288  Es.block(13)(0, 0)
    "spin_u0": (-2.43062856688879, 0),        # the same chain without using the symmetry (one block): lowest eigenvalue overall
}
