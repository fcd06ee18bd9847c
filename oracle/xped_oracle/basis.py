"""TEST INFRASTRUCTURE ONLY -- CPU oracle (never imported by the product).

FusionTree, Qbasis, treepair: restates
  include/Xped/Core/FusionTree.hpp:22-115, src/Core/FusionTree.cpp:32-85 (ordering/equality),
  :309-324 (enlarge), :326-351 (permute), :353-412 (swap), :154-307 (asTensor)
  src/Core/Qbasis.cpp:34-91 (push_back), :208-225 (leftOffset), :306-331 (sort), :410-439
  (forgetHistory, conj), :441-495 (combine)
  src/Core/TensorHelper.cpp:15-45 (build_FusionTree)
  src/Core/treepair.cpp:35-217 (turn_right/turn_left/turn/permute)
  include/Xped/Util/Permutations.hpp:195-225 (decompose)
"""
from __future__ import annotations

import numpy as np

from .symmetry import EPS, Symmetry


# --------------------------------------------------------------------------- FusionTree
class FusionTree:
    __slots__ = ("q_uncoupled", "q_coupled", "dims", "q_intermediates", "IS_DUAL", "dim", "_key")

    def __init__(self, q_uncoupled, q_coupled, dims, q_intermediates=(), IS_DUAL=None):
        self.q_uncoupled = tuple(q_uncoupled)
        self.q_coupled = q_coupled
        self.dims = tuple(int(d) for d in dims)
        self.q_intermediates = tuple(q_intermediates)
        self.IS_DUAL = tuple(IS_DUAL) if IS_DUAL is not None else (False,) * len(self.q_uncoupled)
        d = 1
        for x in self.dims:
            d *= x
        self.dim = d
        # operator== (FusionTree.cpp:66-85, non-pedantic): IS_DUAL does not participate.
        self._key = (self.q_uncoupled, self.q_coupled, self.dims, self.q_intermediates)

    @property
    def rank(self):
        return len(self.q_uncoupled)

    def __eq__(self, other):
        return self._key == other._key

    def __hash__(self):
        return hash(self._key)

    def sort_key(self):
        """operator< (FusionTree.cpp:32-64): uncoupled, intermediates, coupled (multiplicities all 0)."""
        return (self.q_uncoupled, self.q_intermediates, self.q_coupled)

    def __repr__(self):
        return f"Tree(unc={self.q_uncoupled}, inter={self.q_intermediates}, c={self.q_coupled}, dims={self.dims})"

    def replace(self, **kw):
        d = dict(q_uncoupled=self.q_uncoupled, q_coupled=self.q_coupled, dims=self.dims,
                 q_intermediates=self.q_intermediates, IS_DUAL=self.IS_DUAL)
        d.update(kw)
        return FusionTree(**d)

    def enlarge(self, other: "FusionTree") -> "FusionTree":
        """FusionTree.cpp:309-324 (q_coupled is set by the caller)."""
        R = self.rank
        inter = list(self.q_intermediates) + [None] * (max(R + 1 - 2, 0) - len(self.q_intermediates))
        if inter:
            inter[-1] = self.q_coupled
        return FusionTree(self.q_uncoupled + (other.q_uncoupled[0],), None, self.dims + (other.dims[0],), inter,
                          self.IS_DUAL + (other.IS_DUAL[0],))

    # -- recoupling
    def swap(self, sym: Symmetry, pos: int):
        """FusionTree.cpp:353-412: exchange legs pos, pos+1; returns {tree: coeff}."""
        R = self.rank
        assert pos < R - 1
        unc = list(self.q_uncoupled); dual = list(self.IS_DUAL); dims = list(self.dims)
        unc[pos], unc[pos + 1] = unc[pos + 1], unc[pos]
        dual[pos], dual[pos + 1] = dual[pos + 1], dual[pos]
        dims[pos], dims[pos + 1] = dims[pos + 1], dims[pos]
        out = {}
        if pos == 0:
            ql, qr = self.q_uncoupled[0], self.q_uncoupled[1]
            qf = self.q_coupled if R == 2 else self.q_intermediates[0]
            c = sym.coeff_swap(ql, qr, qf)
            if abs(c) < EPS:
                return out
            out[FusionTree(unc, self.q_coupled, dims, self.q_intermediates, dual)] = c
            return out
        q1 = self.q_uncoupled[0] if pos == 1 else self.q_intermediates[pos - 2]
        q2 = self.q_uncoupled[pos]
        q3 = self.q_uncoupled[pos + 1]
        Q = self.q_coupled if pos == R - 2 else self.q_intermediates[pos]
        Q12 = self.q_intermediates[pos - 1]
        for Q31 in sym.reduceSilent(q3, q1):
            c = sym.coeff_swap(Q12, q3, Q) * sym.coeff_recouple(q3, q1, q2, Q, Q31, Q12) * sym.coeff_swap(q3, q1, Q31)
            if abs(c) < EPS:
                continue
            inter = list(self.q_intermediates)
            inter[pos - 1] = Q31
            out[FusionTree(unc, self.q_coupled, dims, inter, dual)] = c
        return out

    def permute(self, sym: Symmetry, pi):
        """FusionTree.cpp:326-351: compose adjacent swaps from Permutation::decompose."""
        out = {self: 1.0}
        for pos in decompose(pi):
            tmp = {}
            for tree, coeff in out.items():
                for tree2, coeff2 in tree.swap(sym, pos).items():
                    if tree2 not in tmp:
                        tmp[tree2] = coeff * coeff2
                    elif abs(tmp[tree2] + coeff * coeff2) < EPS:
                        del tmp[tree2]
                    else:
                        tmp[tree2] += coeff * coeff2
            out = tmp
        return out

    def asTensor(self, sym: Symmetry) -> np.ndarray:
        """FusionTree.cpp:154-307: contraction of CGC vertices, column-major semantics irrelevant
        here because numpy arrays are indexed logically. Result has Rank+1 indices."""
        R = self.rank
        if R == 0:
            return np.ones((1,))
        if R == 1:
            out = np.eye(sym.degeneracy(self.q_uncoupled[0]))
        else:
            out = sym.CGC(self.q_uncoupled[0], self.q_uncoupled[1], self.q_coupled if R == 2 else self.q_intermediates[0])
        for i in range(2, R):
            qi = self.q_intermediates[i - 2]
            qo = self.q_coupled if i == R - 1 else self.q_intermediates[i - 1]
            v = sym.CGC(qi, self.q_uncoupled[i], qo)
            out = np.tensordot(out, v, axes=([out.ndim - 1], [0]))
        for i in range(R):
            if self.IS_DUAL[i]:
                one_j = sym.one_j_tensor(self.q_uncoupled[i])
                out = np.moveaxis(np.tensordot(one_j, out, axes=([1], [i])), 0, i)
        return out


def decompose(pi):
    """Permutations.hpp:195-225: adjacent transposition positions, in application order."""
    pi = list(pi)
    N = len(pi)
    a = list(range(N))
    out = []
    if a == pi:
        return out
    for i in range(N):
        index = a.index(pi[i])
        if index == 0:
            continue
        lo = 0 if i == 0 else i
        j = index - 1
        while j >= lo:
            a[j], a[j + 1] = a[j + 1], a[j]
            out.append(j)
            j -= 1
    assert a == pi
    return out


# --------------------------------------------------------------------------- Qbasis
class Qbasis:
    """Graded basis: list of [q, first_index, dim] + trees per q (Qbasis.hpp:37-268)."""

    def __init__(self, sym: Symmetry, depth: int = 1):
        self.sym = sym
        self.depth = depth
        self.data = []          # [q, num, dim]
        self.trees = {}         # q -> [FusionTree]
        self.curr_dim = 0
        self.CONJ = False
        self.IS_SORTED = False

    @staticmethod
    def from_list(sym, entries, sort=False):
        b = Qbasis(sym)
        for q, d in entries:
            b.push_back(tuple(q) if not isinstance(q, int) else (q,), d)
        if sort:
            b.sort()
        return b

    def copy(self):
        b = Qbasis(self.sym, self.depth)
        b.data = [list(e) for e in self.data]
        b.trees = {q: list(t) for q, t in self.trees.items()}
        b.curr_dim = self.curr_dim
        b.CONJ = self.CONJ
        b.IS_SORTED = self.IS_SORTED
        return b

    # -- queries
    def dim(self): return sum(e[2] for e in self.data)
    def fullDim(self): return sum(e[2] * self.sym.degeneracy(e[0]) for e in self.data)
    def Nq(self): return len(self.data)
    def qs(self): return [e[0] for e in self.data]
    def dims(self): return [e[2] for e in self.data]
    def IS_PRESENT(self, q): return any(e[0] == q for e in self.data)

    def inner_dim(self, q):
        for e in self.data:
            if e[0] == q:
                return e[2]
        return 0

    def full_outer_num(self, q):
        out = 0
        for p, _, d in self.data:
            if p == q:
                return out
            out += d * self.sym.degeneracy(p)
        return out

    def tree(self, q): return self.trees[q]

    def leftOffset(self, tree: FusionTree) -> int:
        """Qbasis.cpp:208-225."""
        out = 0
        for t in self.trees[tree.q_coupled]:
            if t == tree:
                return out
            out += t.dim
        raise KeyError(f"tree not in basis: {tree}")

    def signature(self):
        return (tuple((e[0], e[2]) for e in self.data), self.CONJ)

    def __eq__(self, other):
        return self.data == other.data and self.trees == other.trees

    # -- construction
    def push_back(self, q, inner_dim, trees=None):
        """Qbasis.cpp:34-91."""
        if inner_dim == 0 and trees is None:
            return
        idx = next((i for i, e in enumerate(self.data) if e[0] == q), None)
        if trees is not None:
            assert idx is None
            self.data.append([q, self.curr_dim, inner_dim])
            self.trees[q] = list(trees)
            self.IS_SORTED = False
        elif idx is None:
            if self.depth == 1:
                self.trees[q] = [FusionTree((q,), q, (inner_dim,), (), (self.CONJ,))]
            else:  # depth 0
                self.trees[q] = [FusionTree((), q, (), (), ())]
            self.data.append([q, self.curr_dim, inner_dim])
            self.IS_SORTED = False
        else:
            self.data[idx][2] += inner_dim
            for e in self.data[idx + 1:]:
                e[1] += inner_dim
            t = self.trees[q][0]
            self.trees[q][0] = t.replace(dims=(t.dims[0] + inner_dim,))
        self.curr_dim += inner_dim

    def sort(self):
        """Qbasis.cpp:306-331: by Symmetry::compare, on conj(q) for conjugated bases."""
        if len(self.data) < 2:
            return [0]
        keyf = (lambda i: self.sym.conj(self.data[i][0])) if self.CONJ else (lambda i: self.data[i][0])
        index_sort = sorted(range(len(self.data)), key=keyf)
        new = [list(self.data[i]) for i in index_sort]
        acc = 0
        for e in new:
            e[1] = acc
            acc += e[2]
        self.data = new
        self.IS_SORTED = True
        return index_sort

    def forgetHistory(self):
        out = Qbasis(self.sym)
        for q, _, d in self.data:
            out.push_back(q, d)
        return out

    def conj(self):
        """Qbasis.cpp:418-439."""
        assert self.depth == 1
        out = Qbasis(self.sym)
        out.CONJ = not self.CONJ
        for q, _, d in self.data:
            old = self.trees[q][0]
            cq = self.sym.conj(q)
            out.push_back(cq, d, [FusionTree((cq,), cq, old.dims, (), (not old.IS_DUAL[0],))])
        out.sort()
        return out

    def combine(self, other: "Qbasis") -> "Qbasis":
        """Qbasis.cpp:441-495."""
        sym = self.sym
        out = Qbasis(sym, self.depth + 1)
        for q1, _, _ in self.data:
            for q2, _, _ in other.data:
                for q in sym.reduceSilent(q1, q2):
                    for thisTree in self.trees[q1]:
                        for otherTree in other.trees[q2]:
                            total = thisTree.enlarge(otherTree).replace(q_coupled=q)
                            lst = out.trees.setdefault(q, [])
                            if total not in lst:
                                lst.append(total)
        trees = out.trees
        out.trees = {}
        for q, lst in trees.items():
            lst = sorted(lst, key=FusionTree.sort_key)
            out.push_back(q, sum(t.dim for t in lst), lst)
        out.sort()
        return out

    def truncate_copy(self):
        return self.copy()

    def __repr__(self):
        return "Qbasis(" + ", ".join(f"{q}:{d}" for q, _, d in self.data) + (")'" if self.CONJ else ")")


def build_FusionTree(sym: Symmetry, uncoupled) -> Qbasis:
    """TensorHelper.cpp:15-45."""
    if len(uncoupled) == 0:
        b = Qbasis(sym, 0)
        b.push_back(sym.qvacuum(), 1)
        return b
    cur = uncoupled[0]
    for u in uncoupled[1:]:
        cur = cur.combine(u)
    return cur


# --------------------------------------------------------------------------- treepair
def turn_right(sym: Symmetry, t1: FusionTree, t2: FusionTree):
    """treepair.cpp:35-88: bend the last codomain leg to the end of the domain."""
    R, C = t1.rank, t2.rank
    assert C > 0 and t1.q_coupled == t2.q_coupled
    if C == 1:
        a = sym.qvacuum()
    elif C == 2:
        a = t2.q_uncoupled[0]
    else:
        a = t2.q_intermediates[-1]
    inter1 = list(t1.q_intermediates)
    if R + 1 > 2:
        inter1 = inter1 + [None] * (R + 1 - 2 - len(inter1))
        inter1[-1] = t1.q_coupled
    t1p = FusionTree(t1.q_uncoupled + (sym.conj(t2.q_uncoupled[-1]),), a, t1.dims + (t2.dims[-1],), inter1,
                     t1.IS_DUAL + (not t2.IS_DUAL[-1],))
    t2p = FusionTree(t2.q_uncoupled[:-1], a, t2.dims[:-1], t2.q_intermediates[:-1] if C > 2 else (), t2.IS_DUAL[:-1])
    b = t2.q_uncoupled[-1]
    c = t2.q_coupled
    coeff = sym.coeff_turn(a, b, c)
    if t2.IS_DUAL[-1]:
        coeff *= sym.coeff_FS(sym.conj(b))
    if abs(coeff) < EPS:
        return {}
    return {(t1p, t2p): coeff}


def turn_left(sym, t1, t2):
    return {(b, a): c for (a, b), c in turn_right(sym, t2, t1).items()}


def turn(sym, shift: int, t1, t2):
    """treepair.cpp:103-141."""
    if shift == 0:
        return {(t1, t2): 1.0}
    out = {}
    step = turn_left if shift > 0 else turn_right
    nxt = shift - 1 if shift > 0 else shift + 1
    for (a, b), c1 in step(sym, t1, t2).items():
        for key, c2 in turn(sym, nxt, a, b).items():
            out[key] = c1 * c2
    return out


def _abelian_tree(sym, qs, dims, duals):
    """computeIntermediates (FusionTree.cpp:133-150) for Abelian symmetries."""
    n = len(qs)
    if n == 0:
        return FusionTree((), sym.qvacuum(), (), (), ())
    inter = []
    acc = qs[0]
    for i in range(1, n):
        acc = sym.reduceSilent(acc, qs[i])[0]
        if i < n - 1:
            inter.append(acc)
    return FusionTree(qs, acc, dims, inter, duals)


def treepair_permute(sym: Symmetry, shift: int, t1: FusionTree, t2: FusionTree, pi):
    """treepair.cpp:143-217; returns {(domain_tree, codomain_tree): coeff}."""
    R, C = t1.rank, t2.rank
    newR, newC = R - shift, C + shift
    pi = list(pi)
    if not sym.non_abelian and not sym.fermionic:
        def leg(k, to_domain):
            if pi[k] < R:
                q, d, du = t1.q_uncoupled[pi[k]], t1.dims[pi[k]], t1.IS_DUAL[pi[k]]
                return (q, d, du) if to_domain else (sym.conj(q), d, not du)
            q, d, du = t2.q_uncoupled[pi[k] - R], t2.dims[pi[k] - R], t2.IS_DUAL[pi[k] - R]
            return (sym.conj(q), d, not du) if to_domain else (q, d, du)
        dl = [leg(r, True) for r in range(newR)]
        cl = [leg(c + newR, False) for c in range(newC)]
        od = _abelian_tree(sym, tuple(x[0] for x in dl), tuple(x[1] for x in dl), tuple(x[2] for x in dl))
        oc = _abelian_tree(sym, tuple(x[0] for x in cl), tuple(x[1] for x in cl), tuple(x[2] for x in cl))
        if newR == 0:
            od = od.replace(q_coupled=sym.qvacuum())
        if newC == 0:
            oc = oc.replace(q_coupled=sym.qvacuum())
        return {(od, oc): 1.0}
    assert t1.q_coupled == t2.q_coupled
    N = R + C
    pi_id = list(range(N))
    for i in range(C):
        pi_id[i + R] = (C - 1) - i + R
    pi_tmp = [pi_id[pi[i]] for i in range(N)]
    pi_corr = pi_tmp[:newR] + [pi_tmp[(newC - 1) - i + newR] for i in range(newC)]
    out = {}
    for (t1p, trivial), c1 in turn(sym, -C, t1, t2).items():
        for t1pp, c2 in t1p.permute(sym, pi_corr).items():
            for key, c3 in turn(sym, newC, t1pp, trivial).items():
                out[key] = out.get(key, 0.0) + c1 * c2 * c3
    return out
