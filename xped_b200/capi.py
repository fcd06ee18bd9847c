"""ctypes binding of include/xped_b200.h -- the one and only way Python reaches the product.

The shared library is built in-tree by `__graft_entry__.build()` (nvcc, sm_100a). Loading it does not need a
GPU (symbol checks run on CPU); creating a context does, and fails loudly without one: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libxped_b200.so")


class XbError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"xped_b200 status {status}: {msg}")
        self.status = status


class SymFactor(C.Structure):
    _fields_ = [("kind", C.c_int), ("N", C.c_int), ("fermionic", C.c_int)]


P = C.c_void_p
PP = C.POINTER(C.c_void_p)
IP = C.POINTER(C.c_int)
LP = C.POINTER(C.c_int64)
DP = C.POINTER(C.c_double)

# name -> (restype, argtypes); mirrors include/xped_b200.h declaration by declaration
SIGNATURES = {
    "xb_last_error": (C.c_char_p, []),
    "xb_version": (C.c_char_p, []),
    "xb_ctx_create": (C.c_int, [PP, C.c_int, C.POINTER(SymFactor), C.c_int]),
    "xb_ctx_create_planning": (C.c_int, [PP, C.POINTER(SymFactor), C.c_int]),
    "xb_ctx_free": (C.c_int, [P]),
    "xb_ctx_sync": (C.c_int, [P]),
    "xb_tensor_save": (C.c_int, [P, P, C.c_char_p]),
    "xb_tensor_load": (C.c_int, [P, C.c_char_p, PP]),
    "xb_ctm_save": (C.c_int, [P, C.c_char_p]),
    "xb_ctm_load": (C.c_int, [P, C.c_char_p, PP]),
    "xb_tensor_from_charge_blocks": (C.c_int, [P, C.c_int, C.c_int, IP, C.POINTER(DP), C.POINTER(DP), IP, C.c_int64, LP, LP, DP, C.c_int64, C.c_int, PP]),
    "xb_dist_shm_create": (C.c_int, [C.c_char_p, C.c_int]),
    "xb_dist_init": (C.c_int, [P, C.c_int, C.c_int, C.c_int64, C.c_char_p]),
    "xb_dist_shutdown": (C.c_int, [P]),
    "xb_dist_barrier": (C.c_int, [P]),
    "xb_dist_arena_stats": (C.c_int, [P, LP]),
    "xb_dist_plan_pushable": (C.c_int, [P, IP]),
    "xb_dist_permute_push": (C.c_int, [P, P, P, PP]),
    "xb_dist_multiply": (C.c_int, [P, P, IP, P, IP, C.c_int, PP, IP, C.c_int]),
    "xb_dist_multiply_assignment": (C.c_int, [LP, LP, LP, IP, C.c_int, C.c_int, C.c_int, LP, LP]),
    "xb_tape_begin": (C.c_int, [P]),
    "xb_tape_end": (C.c_int, [P]),
    "xb_tape_seed_trace": (C.c_int, [P, P, C.c_double]),
    "xb_tape_seed": (C.c_int, [P, P, P]),
    "xb_tape_backward": (C.c_int, [P]),
    "xb_tape_grad": (C.c_int, [P, P, PP]),
    "xb_tape_size": (C.c_int, [P, LP]),
    "xb_normalize_max": (C.c_int, [P, P, PP]),
    "xb_normalize_trace": (C.c_int, [P, P, P, PP]),
    "xb_ctm_energy_grad": (C.c_int, [P, P, C.c_int64, C.c_int, P, P, C.c_int, C.POINTER(C.c_double), PP]),
    "xb_ctx_set_int": (C.c_int, [P, C.c_char_p, C.c_int64]),
    "xb_ctx_get_int": (C.c_int, [P, C.c_char_p, LP]),
    "xb_ctx_nq": (C.c_int, [P]),
    "xb_ctx_stats": (C.c_int, [P, LP]),
    "xb_basis_create": (C.c_int, [P, IP, LP, C.c_int, PP]),
    "xb_basis_free": (C.c_int, [P]),
    "xb_basis_conj": (C.c_int, [P, PP]),
    "xb_basis_sorted": (C.c_int, [P, PP]),
    "xb_basis_combine": (C.c_int, [P, P, PP]),
    "xb_basis_size": (C.c_int, [P, IP, LP, LP, IP]),
    "xb_basis_get": (C.c_int, [P, IP, LP]),
    "xb_tensor_create": (C.c_int, [P, PP, C.c_int, PP, C.c_int, PP]),
    "xb_tensor_free": (C.c_int, [P]),
    "xb_tensor_clone": (C.c_int, [P, PP]),
    "xb_tensor_rank": (C.c_int, [P, IP, IP]),
    "xb_tensor_leg_basis": (C.c_int, [P, C.c_int, PP]),
    "xb_tensor_plain_size": (C.c_int64, [P]),
    "xb_tensor_num_sectors": (C.c_int, [P]),
    "xb_tensor_structure": (C.c_int, [P, IP, LP, LP, LP]),
    "xb_tensor_num_tree_pairs": (C.c_int64, [P]),
    "xb_tensor_tree_pairs": (C.c_int, [P, LP]),
    "xb_tensor_tree_pair_trees": (C.c_int, [P, C.c_int64, IP, IP, IP, LP]),
    "xb_tensor_find_tree_pair": (C.c_int64, [P, IP, IP, IP]),
    "xb_view": (C.c_int, [P, C.c_int64, C.POINTER(C.c_void_p), LP, LP, LP]),
    "xb_sub_block": (C.c_int, [P, C.c_int64, DP, C.c_int64]),
    "xb_set_sub_block": (C.c_int, [P, C.c_int64, DP, C.c_int64]),
    "xb_tensor_upload": (C.c_int, [P, DP, C.c_int64]),
    "xb_tensor_download": (C.c_int, [P, DP, C.c_int64]),
    "xb_tensor_set_zero": (C.c_int, [P]),
    "xb_tensor_set_identity": (C.c_int, [P]),
    "xb_tensor_set_constant": (C.c_int, [P, C.c_double]),
    "xb_tensor_set_random": (C.c_int, [P, C.c_uint64]),
    "xb_permute": (C.c_int, [P, P, C.c_int, IP, C.c_int, PP]),
    "xb_permute_adj": (C.c_int, [P, P, C.c_int, IP, C.c_int, PP]),
    "xb_multiply": (C.c_int, [P, P, C.c_int, P, C.c_int, PP]),
    "xb_contract": (C.c_int, [P, P, IP, C.c_int, P, IP, C.c_int, C.c_int, PP]),
    "xb_adjoint": (C.c_int, [P, P, PP]),
    "xb_twist": (C.c_int, [P, P, C.c_int, PP]),
    "xb_axpby": (C.c_int, [P, C.c_double, P, C.c_double, P, PP]),
    "xb_diag_map": (C.c_int, [P, P, C.c_int, PP]),
    "xb_max_norm": (C.c_int, [P, P, DP]),
    "xb_trace": (C.c_int, [P, P, DP]),
    "xb_squared_norm": (C.c_int, [P, P, DP]),
    "xb_tsvd": (C.c_int, [P, P, C.c_int64, C.c_double, C.c_int, PP, PP, PP, DP, DP]),
    "xb_tsvd_backward": (C.c_int, [P, P, P, P, P, P, P, P, PP]),
    "xb_tqr": (C.c_int, [P, P, C.c_int, PP, PP]),
    "xb_teigh": (C.c_int, [P, P, PP, PP]),
    "xb_decompose": (C.c_int, [P, P, P, C.c_int64, PP, PP]),
    "xb_contract_L": (C.c_int, [P, P, P, P, PP]),
    "xb_contract_R": (C.c_int, [P, P, P, P, PP]),
    "xb_mps_left_sweep_step": (C.c_int, [P, P, P, C.c_int64, C.c_double, PP, PP, DP, DP]),
    "xb_mps_right_sweep_step": (C.c_int, [P, P, P, C.c_int64, C.c_double, PP, PP, DP, DP]),
    "xb_layout_structure": (C.c_int, [P, PP, C.c_int, PP, C.c_int, IP, IP, LP, LP, LP]),
    "xb_tensor_create_sectors": (C.c_int, [P, PP, C.c_int, PP, C.c_int, IP, C.c_int, PP]),
    "xb_tensor_set_random_at": (C.c_int, [P, C.c_uint64, LP]),
    "xb_dist_owner_lpt": (C.c_int, [DP, C.c_int, C.c_int, IP]),
    "xb_dist_permute_plan": (C.c_int, [P, PP, C.c_int, PP, C.c_int, C.c_int, IP, C.c_int, C.c_int, IP, IP, C.c_int, C.c_int, C.c_int, PP]),
    "xb_dist_plan_free": (C.c_int, [P]),
    "xb_dist_plan_counts": (C.c_int, [P, LP, LP]),
    "xb_dist_plan_out": (C.c_int, [P, IP, IP, PP, IP, IP]),
    "xb_dist_plan_local_sectors": (C.c_int, [P, C.c_int, IP, IP]),
    "xb_dist_plan_records": (C.c_int, [P, C.c_int, LP, LP]),
    "xb_dist_permute_pack": (C.c_int, [P, P, P, C.c_void_p]),
    "xb_dist_permute_local": (C.c_int, [P, P, P, PP]),
    "xb_dist_permute_remote": (C.c_int, [P, P, P, C.c_void_p, P]),
    "xb_multiply_sectors": (C.c_int, [P, P, C.c_int, P, C.c_int, IP, C.c_int, PP]),
    "xb_ctm_create": (C.c_int, [P, P, C.c_int64, PP]),
    "xb_ctm_free": (C.c_int, [P]),
    "xb_ctm_step": (C.c_int, [P]),
    "xb_ctm_get": (C.c_int, [P, C.c_int, PP]),
    "xb_ctm_set": (C.c_int, [P, C.c_int, P]),
    "xb_ctm_corner": (C.c_int, [P, C.c_int, PP]),
    "xb_permute_plan_stats": (C.c_int, [P, PP, C.c_int, PP, C.c_int, C.c_int, IP, C.c_int, LP]),
    "xb_permute_plan_census": (C.c_int, [P, PP, C.c_int, PP, C.c_int, C.c_int, IP, C.c_int, LP]),
    "xb_multiply_plan_emulate": (C.c_int, [P, PP, C.c_int, PP, C.c_int, C.c_int, PP, C.c_int, PP, C.c_int, C.c_int, C.c_int, DP, C.c_int64, DP, C.c_int64, DP, C.c_int64, LP]),
    "xb_contract_plan_emulate": (C.c_int, [P, PP, C.c_int, PP, C.c_int, IP, PP, C.c_int, PP, C.c_int, IP, C.c_int, C.c_int, DP, C.c_int64, DP, C.c_int64, DP, C.c_int64]),
    "xb_permute_plan_emulate": (C.c_int, [P, PP, C.c_int, PP, C.c_int, C.c_int, IP, C.c_int, C.c_int, C.c_int, DP, C.c_int64, DP, C.c_int64]),
    "xb_ctm_rdm": (C.c_int, [P, C.c_int, PP]),
    "xb_avg_two_site": (C.c_int, [P, P, P, DP]),
    "xb_rdm_one_site": (C.c_int, [P, P, PP]),
    "xb_avg_one_site": (C.c_int, [P, P, P, DP]),
    "xb_ctm_move": (C.c_int, [P, C.c_int]),
    "xb_bench_permute": (C.c_int, [P, P, C.c_int, IP, C.c_int, C.c_int, DP, LP]),
    "xb_bench_multiply": (C.c_int, [P, P, P, C.c_int, DP, DP]),
    "xb_flush_l2": (C.c_int, [P]),
    "xb_ctx_stream": (C.c_void_p, [P]),
    "xb_timer_start": (C.c_int, [P]),
    "xb_timer_stop": (C.c_int, [P, DP]),
    "xb_profile_enable": (C.c_int, [P, C.c_int]),
    "xb_profile_read": (C.c_int, [P, DP, LP, DP]),
}

_lib = None


def load():
    """Load libxped_b200.so (no GPU needed). Raises if the extension has not been built: the product never
    substitutes another implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing -- run `python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). "
                              "xped_b200 has no CPU or PyTorch fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status):
    if status != 0:
        raise XbError(status, load().xb_last_error().decode())
