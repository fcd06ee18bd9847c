"""xped_b200 -- B200-native backend for the Xped block-sparse symmetric-tensor hot path.

Product = xped_b200/csrc (CUDA sm_100a kernels + host plan builder + C ABI, include/xped_b200.h) and this thin
ctypes mirror of the reference's Tensor interface. Nothing here imports oracle/.
"""
from .capi import LIB_PATH, SIGNATURES, XbError, load  # noqa: F401
from .tensor import (CTM, Context, Qbasis, Tensor, avg_one_site, avg_two_site, contract_L, contract_R, ctm_energy_grad, decompose,  # noqa: F401
                     rdm_one_site)
from .mps import Mps, dot, gen_aux_bases, random_mps  # noqa: F401
