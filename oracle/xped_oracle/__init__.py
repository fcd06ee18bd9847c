"""TEST INFRASTRUCTURE ONLY -- CPU oracle restating the reference (cpp977/Xped) hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package; the product (xped_b200/) never does.  Parity pin: the oracle reproduces the reference's own
property tests (tests/tree.cpp, tests/tensor.cpp incl. the analytic SU(2) known-answer checks) -- see
tests/test_oracle_pins.py.  tSVD values, tQR and the CTM driver are "parity unpinned" by the reference's
own tests and are pinned only by invariants.
"""
from .symmetry import Symmetry, wigner_3j_2j, wigner_6j_2j  # noqa: F401
from .basis import FusionTree, Qbasis, build_FusionTree, decompose, treepair_permute, turn  # noqa: F401
from .tensor import Tensor, get_permutations  # noqa: F401
