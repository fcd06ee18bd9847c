"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the Xped hot path (never imported by the product).

Symmetry layer: restates the reference's static `Sym::*` structs as one runtime object.

Reference files followed (all under /root/reference):
  include/Xped/Symmetry/SymBase.hpp:14-102   reduceSilent / compare / default coeff_FS, coeff_turn
  include/Xped/Symmetry/U0.hpp               trivial symmetry (Nq = 0)
  include/Xped/Symmetry/U1.hpp:38-226        U(1): q integer, conj = -q, fusion ql+qr
  include/Xped/Symmetry/ZN.hpp:95-241        Z_N: arithmetic mod N
  include/Xped/Symmetry/SU2.hpp:41-181, src/Symmetry/SU2.cpp:17-119   SU(2), q = 2S+1
  include/Xped/Symmetry/S1xS2.hpp:95-312     products: qarrays concatenated, coefficients multiplied
  include/Xped/Symmetry/SU2Wrappers.hpp:66-81  3j/6j come from GSL 2.7.1 (gsl_sf_coupling_{3j,6j});
      GSL is absent here, so its *published* definition (Racah's single-sum formula) is restated
      with exact integer arithmetic and cross-checked against sympy.physics.wigner in tests/.
"""
from __future__ import annotations

import math
from fractions import Fraction
from functools import lru_cache
from itertools import product as iproduct

EPS = 1.0e-12  # ScalarTraits<double>::epsilon(), include/Xped/Core/ScalarTraits.hpp:17


def phase(q: int) -> float:
    """functions.hpp:76-81 -- (-1)^q (C++ `q % 2` is non-zero for odd negative q too)."""
    return -1.0 if (q % 2) != 0 else 1.0


# --------------------------------------------------------------------------- SU(2) 3nj symbols
def _fact(n: int) -> int:
    return math.factorial(n)


def _delta2(a: int, b: int, c: int) -> Fraction:
    """Triangle coefficient Delta(a b c)^2 with doubled spins a,b,c (all sums below are even)."""
    return Fraction(_fact((a + b - c) // 2) * _fact((a - b + c) // 2) * _fact((-a + b + c) // 2), _fact((a + b + c) // 2 + 1))


def _triad_ok(a: int, b: int, c: int) -> bool:
    return (a + b + c) % 2 == 0 and abs(a - b) <= c <= a + b


@lru_cache(maxsize=None)
def wigner_6j_2j(a: int, b: int, c: int, d: int, e: int, f: int) -> float:
    """{a/2 b/2 c/2; d/2 e/2 f/2}, arguments are doubled spins (GSL convention, SU2Wrappers.hpp:73-76)."""
    if min(a, b, c, d, e, f) < 0:
        return 0.0
    if not (_triad_ok(a, b, c) and _triad_ok(a, e, f) and _triad_ok(d, b, f) and _triad_ok(d, e, c)):
        return 0.0
    pref = _delta2(a, b, c) * _delta2(a, e, f) * _delta2(d, b, f) * _delta2(d, e, c)
    t_min = max(a + b + c, a + e + f, d + b + f, d + e + c) // 2
    t_max = min(a + b + d + e, b + c + e + f, a + c + d + f) // 2
    s = Fraction(0)
    for t in range(t_min, t_max + 1):
        den = (_fact(t - (a + b + c) // 2) * _fact(t - (a + e + f) // 2) * _fact(t - (d + b + f) // 2) * _fact(t - (d + e + c) // 2)
               * _fact((a + b + d + e) // 2 - t) * _fact((b + c + e + f) // 2 - t) * _fact((a + c + d + f) // 2 - t))
        s += Fraction((-1) ** t * _fact(t + 1), den)
    val = s * s * pref  # exact square of the symbol
    root = _sqrt_fraction(val)
    return math.copysign(root, 1.0 if s >= 0 else -1.0) if val != 0 else 0.0


def _sqrt_fraction(x: Fraction) -> float:
    """Correctly rounded-ish sqrt of a non-negative rational (numerators can exceed 2^1024)."""
    if x == 0:
        return 0.0
    n, d = x.numerator, x.denominator
    # scale to keep ~120 significant bits through the integer square root
    shift = max(0, 240 - (n.bit_length() - d.bit_length()))
    shift += shift % 2
    r = math.isqrt((n << shift) // d)
    return math.ldexp(float(r), -shift // 2) if r.bit_length() < 1000 else float(Fraction(r, 1 << (shift // 2)))


@lru_cache(maxsize=None)
def wigner_3j_2j(a: int, b: int, c: int, ma: int, mb: int, mc: int) -> float:
    """(a/2 b/2 c/2; ma/2 mb/2 mc/2) with doubled arguments (GSL convention, SU2Wrappers.hpp:78-81)."""
    if ma + mb + mc != 0 or not _triad_ok(a, b, c):
        return 0.0
    if abs(ma) > a or abs(mb) > b or abs(mc) > c:
        return 0.0
    if (a + ma) % 2 or (b + mb) % 2 or (c + mc) % 2:
        return 0.0
    pref = _delta2(a, b, c) * (_fact((a + ma) // 2) * _fact((a - ma) // 2) * _fact((b + mb) // 2) * _fact((b - mb) // 2)
                               * _fact((c + mc) // 2) * _fact((c - mc) // 2))
    k_min = max(0, (b - c - ma) // 2, (a - c + mb) // 2)
    k_max = min((a + b - c) // 2, (a - ma) // 2, (b + mb) // 2)
    s = Fraction(0)
    for k in range(k_min, k_max + 1):
        den = (_fact(k) * _fact((a + b - c) // 2 - k) * _fact((a - ma) // 2 - k) * _fact((b + mb) // 2 - k)
               * _fact((c - b + ma) // 2 + k) * _fact((c - a - mb) // 2 + k))
        s += Fraction((-1) ** k, den)
    val = s * s * pref
    if val == 0:
        return 0.0
    sign = (1.0 if s >= 0 else -1.0) * (-1.0 if ((a - b - mc) // 2) % 2 else 1.0)
    return sign * _sqrt_fraction(val)


# --------------------------------------------------------------------------- factors
class _Factor:
    nq = 1
    abelian = True

    def __init__(self, fermionic: bool = False):
        self.fermionic = fermionic


class _U0(_Factor):
    """U0.hpp:86-141 -- no quantum number at all."""
    nq = 0
    name = "U0"


class _U1(_Factor):
    name = "U1"

    def vac(self): return 0
    def conj(self, q): return -q
    def deg(self, q): return 1
    def fuse(self, l, r): return [l + r]                       # U1.hpp:189-194
    def triangle(self, a, b, c): return a + b == c                # U1.hpp:219-224
    def odd(self, q): return q % 2 != 0                          # U1.hpp:130-134
    def lowest(self): return [-1, +1]                             # U1.hpp:96-99

    def swap(self, l, r, f):                                      # U1.hpp:174-182
        sign = -1.0 if (self.fermionic and self.odd(l) and self.odd(r)) else 1.0
        return sign if self.triangle(l, r, f) else 0.0

    def six_j(self, q1, q2, q3, q4, q5, q6):                      # U1.hpp:212-216
        t = self.triangle
        return 1.0 if (t(q1, q2, q3) and t(q1, q6, q5) and t(q2, q4, q6) and t(q3, q4, q5)) else 0.0

    def recouple(self, q1, q2, q3, Q, Q12, Q23): return self.six_j(q1, q2, Q12, q3, Q, Q23)   # U1.hpp:162-165
    def turn(self, l, r, f): return 1.0 if self.triangle(l, r, f) else 0.0                      # U1.hpp:158
    def FS(self, q): return 1.0                                                                   # U1.hpp:136
    def twist(self, q): return -1.0 if (self.fermionic and self.odd(q)) else 1.0                 # U1.hpp:130-134
    def right_ortho(self, a, b): return 1.0
    def cgc(self, q1, q2, q3):                                                                    # U1.hpp:196-210
        import numpy as np
        return np.full((1, 1, 1), 1.0 if self.triangle(q1, q2, q3) else 0.0)
    def one_j(self, q):
        import numpy as np
        return np.ones((1, 1))


class _ZN(_U1):
    name = "ZN"

    def __init__(self, N: int, fermionic: bool = False):
        super().__init__(fermionic)
        self.N = N

    def conj(self, q): return (-q) % self.N                      # ZN.hpp:115
    def fuse(self, l, r): return [(l + r) % self.N]              # ZN.hpp:200-205
    def triangle(self, a, b, c): return (a + b) % self.N == c    # ZN.hpp:232-237
    def lowest(self): return [1] if self.N == 2 else [1, self.N - 1]   # ZN.hpp:92-97


class _SU2(_Factor):
    """q = 2S+1 (SU2.hpp:108-112)."""
    name = "SU2"
    abelian = False

    def vac(self): return 1
    def conj(self, q): return q
    def deg(self, q): return q
    def lowest(self): return [2]

    def fuse(self, l, r):                                         # SU2.cpp:17-27
        return list(range(abs(l - r) + 1, l + r - 1 + 1, 2))

    def triangle(self, a, b, c):                                  # SU2.cpp:113-118
        return (c - 1 >= abs(a - b)) and (c - 1 <= a + b - 2)

    def swap(self, l, r, f):                                      # SU2.hpp:166-174
        sign = -1.0 if (self.fermionic and (l - 1) % 2 != 0 and (r - 1) % 2 != 0) else 1.0
        return sign * phase(int((l + r - f - 1) / 2)) if self.triangle(l, r, f) else 0.0

    def six_j(self, q1, q2, q3, q4, q5, q6):                      # SU2.cpp:91-96 -> SU2Wrappers.hpp:73-76
        return wigner_6j_2j(q1 - 1, q2 - 1, q3 - 1, q4 - 1, q5 - 1, q6 - 1)

    def recouple(self, q1, q2, q3, Q, Q12, Q23):                  # SU2.hpp:161-164
        return math.sqrt(Q12 * Q23) * phase(int((q1 + q2 + q3 + Q) / 2.0)) * self.six_j(q1, q2, Q12, q3, Q, Q23)

    def turn(self, l, r, f):                                      # SU2.hpp:139-142
        return self.swap(l, r, f) * math.sqrt(f / l) if self.triangle(l, r, f) else 0.0

    def FS(self, q): return -1.0 if q % 2 == 0 else 1.0          # SU2.cpp:35-39
    def twist(self, q): return (-1.0 if (q - 1) % 2 != 0 else 1.0) if self.fermionic else 1.0   # SU2.hpp:122-126
    def right_ortho(self, a, b): return a / b                     # SU2.hpp:135

    def cgc(self, q1, q2, q3):                                    # SU2.cpp:69-88
        import numpy as np
        out = np.zeros((q1, q2, q3))
        for i1 in range(q1):
            for i2 in range(q2):
                for i3 in range(q3):
                    m1 = -(2 * i1 - (q1 - 1)); m2 = -(2 * i2 - (q2 - 1)); m3 = -(2 * i3 - (q3 - 1))
                    out[i1, i2, i3] = (wigner_3j_2j(q1 - 1, q2 - 1, q3 - 1, m1, m2, -m3)
                                       * phase(int((q1 - q2 + m3) / 2)) * math.sqrt(q3))
        return out

    def one_j(self, q):                                           # SU2.cpp:41-60
        import numpy as np
        out = np.zeros((q, q))
        for i in range(q):
            for j in range(q):
                im = -(2 * i - (q - 1)); jm = -(2 * j - (q - 1))
                if im == -jm:
                    out[i, j] = phase(int((q - im) / 2))
        return out


# --------------------------------------------------------------------------- the symmetry object
class Symmetry:
    """A (product of) symmetry factor(s). Quantum numbers are tuples of ints of length `Nq`
    (qarray<Nq>, qarray.hpp:29-80; ordering = lexicographic, SymBase.hpp:92)."""

    def __init__(self, *factors):
        self.factors = list(factors)
        self.active = [f for f in self.factors if f.nq == 1]
        self.Nq = len(self.active)
        self.non_abelian = any(not f.abelian for f in self.active)
        self.fermionic = any(f.fermionic for f in self.active)
        self.trivial = self.Nq == 0
        self._recouple_cache = {}

    # -- constructors for the symmetries the reference's tests instantiate
    @staticmethod
    def U0(): return Symmetry(_U0())
    @staticmethod
    def U1(fermionic=False): return Symmetry(_U1(fermionic))
    @staticmethod
    def ZN(N, fermionic=False): return Symmetry(_ZN(N, fermionic))
    @staticmethod
    def SU2(fermionic=False): return Symmetry(_SU2(fermionic))

    @staticmethod
    def from_spec(spec):
        """spec: list of (kind, N, fermionic) with kind in {'U0','U1','ZN','SU2'} -- same encoding as the C ABI."""
        fs = []
        for kind, N, ferm in spec:
            fs.append({"U0": lambda: _U0(), "U1": lambda: _U1(bool(ferm)), "ZN": lambda: _ZN(int(N), bool(ferm)),
                       "SU2": lambda: _SU2(bool(ferm))}[kind]())
        return Symmetry(*fs)

    def spec(self):
        return [(f.name, getattr(f, "N", 0), int(f.fermionic)) for f in self.factors]

    def name(self):
        return "x".join(f.name + (str(f.N) if hasattr(f, "N") else "") + ("f" if f.fermionic else "") for f in self.factors)

    # -- SymBase interface
    def qvacuum(self): return tuple(f.vac() for f in self.active)
    def conj(self, q): return tuple(f.conj(x) for f, x in zip(self.active, q))

    def degeneracy(self, q):
        d = 1
        for f, x in zip(self.active, q):
            d *= f.deg(x)
        return d

    def basis_combine(self, ql, qr):
        """S1xS2.hpp:217-229: first factor is the outer loop."""
        parts = [f.fuse(a, b) for f, a, b in zip(self.active, ql, qr)]
        return [tuple(c) for c in iproduct(*parts)]

    reduceSilent = basis_combine

    def triangle(self, q1, q2, q3):
        return all(f.triangle(a, b, c) for f, a, b, c in zip(self.active, q1, q2, q3))

    @staticmethod
    def compare(q1, q2):
        return q1 < q2

    def lowest_qs(self):
        parts = [f.lowest() for f in self.active]
        return [tuple(c) for c in iproduct(*parts)]

    def _prod(self, fn, *qs):
        out = 1.0
        for i, f in enumerate(self.active):
            out *= getattr(f, fn)(*[q[i] for q in qs])
        return out

    def coeff_swap(self, ql, qr, qf): return self._prod("swap", ql, qr, qf)
    def coeff_turn(self, ql, qr, qf): return self._prod("turn", ql, qr, qf)
    def coeff_FS(self, q): return self._prod("FS", q)
    def coeff_twist(self, q): return self._prod("twist", q)
    def coeff_rightOrtho(self, q1, q2): return self._prod("right_ortho", q1, q2)
    def coeff_6j(self, *qs): return self._prod("six_j", *qs)

    def coeff_recouple(self, q1, q2, q3, Q, Q12, Q23):
        key = (q1, q2, q3, Q, Q12, Q23)
        v = self._recouple_cache.get(key)
        if v is None:
            v = self._prod("recouple", q1, q2, q3, Q, Q12, Q23)
            self._recouple_cache[key] = v
        return v

    # -- dense Clebsch-Gordan objects (only for plainTensor, the reference's own absolute pin)
    def CGC(self, q1, q2, q3):
        """Products: tensorProd of the factors' CGCs (S1xS2.hpp:166-176), first factor slow."""
        import numpy as np
        out = np.ones((1, 1, 1))
        for i, f in enumerate(self.active):
            t = f.cgc(q1[i], q2[i], q3[i])
            out = np.einsum("abc,xyz->axbycz", out, t).reshape(out.shape[0] * t.shape[0], out.shape[1] * t.shape[1], out.shape[2] * t.shape[2])
        return out

    def one_j_tensor(self, q):
        import numpy as np
        out = np.ones((1, 1))
        for i, f in enumerate(self.active):
            t = f.one_j(q[i])
            out = np.einsum("ab,xy->axby", out, t).reshape(out.shape[0] * t.shape[0], out.shape[1] * t.shape[1])
        return out
