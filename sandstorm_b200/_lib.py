"""ctypes binding of libsandstorm_b200.so (the C ABI in include/sandstorm_b200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is present every
entry point raises.  Nothing in this package imports ``oracle/``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_int, c_size_t, c_uint8, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SS_LIB_PATH") or os.path.join(_HERE, "libsandstorm_b200.so")   # SS_LIB_PATH: A/B kernel variants

SS_OK, SS_ERR_INVALID, SS_ERR_CUDA, SS_ERR_OOM, SS_ERR_UNSUPPORTED = 0, -1, -2, -3, -4
FIELD_FP252, FIELD_GOLDILOCKS = 0, 1
ORDER_NATURAL, ORDER_BITREV, ORDER_BITREV_RC = 0, 1, 3
TREE_KECCAK, TREE_KECCAK_M20, TREE_FRIENDLY, TREE_BLAKE2S_M20, TREE_SHA256 = range(5)

# name -> (restype, argtypes); must list every symbol declared in include/sandstorm_b200.h
SIGNATURES = {
    "ss_version": (c_int, []),
    "ss_create": (c_int, [c_int, POINTER(c_void_p)]),
    "ss_destroy": (None, [c_void_p]),
    "ss_last_error": (c_char_p, [c_void_p]),
    "ss_sync": (c_int, [c_void_p]),
    "ss_kernel_launches": (c_uint64, [c_void_p]),
    "ss_set_option": (c_int, [c_void_p, c_char_p, ctypes.c_int64]),
    "ss_get_option": (ctypes.c_int64, [c_void_p, c_char_p, ctypes.c_int64]),
    "ss_malloc": (c_int, [c_void_p, c_size_t, POINTER(c_void_p)]),
    "ss_free": (c_int, [c_void_p, c_void_p]),
    "ss_host_register": (c_int, [c_void_p, c_void_p, c_size_t]),
    "ss_host_unregister": (c_int, [c_void_p, c_void_p]),
    "ss_memcpy_h2d": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ss_memcpy_d2h": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ss_ntt": (c_int, [c_void_p, c_int, c_void_p, c_uint64, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "ss_lde": (c_int, [c_void_p, c_int, c_void_p, c_uint64, c_int, c_int, c_int, c_void_p, c_uint64, c_void_p, c_uint64, c_int, c_void_p]),
    "ss_merkle_build": (c_int, [c_void_p, c_int, c_int, c_void_p, c_uint64, c_int, c_int, c_int, POINTER(c_void_p), c_void_p]),
    "ss_hash_rows": (c_int, [c_void_p, c_int, c_void_p, c_uint64, c_int, c_int, c_int, c_uint64, c_uint64, c_void_p, c_void_p]),
    "ss_merkle_build_from_leaves": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, POINTER(c_void_p), c_void_p]),
    "ss_bitrev_permute32": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "ss_merkle_root": (c_int, [c_void_p, c_void_p, POINTER(c_uint8)]),
    "ss_merkle_nodes": (c_int, [c_void_p, c_void_p, POINTER(c_uint64), c_size_t, POINTER(c_uint8)]),
    "ss_merkle_leaves": (c_int, [c_void_p, c_void_p, POINTER(c_uint64), c_size_t, POINTER(c_uint8)]),
    "ss_merkle_open": (c_int, [c_void_p, c_void_p, POINTER(c_uint64), c_size_t, POINTER(c_uint8)]),
    "ss_coset_eval": (c_int, [c_void_p, c_int, c_void_p, c_uint64, c_int, c_int, c_void_p, c_void_p, c_uint64, c_void_p]),
    "ss_merkle_combine": (c_int, [c_void_p, c_int, POINTER(c_uint8), c_int, POINTER(c_uint8)]),
    "ss_merkle_combine_open": (c_int, [c_void_p, c_int, POINTER(c_uint8), c_int, c_int, c_uint64, POINTER(c_uint8)]),
    "ss_tree_log_rows": (c_int, [c_void_p]),
    "ss_tree_free": (None, [c_void_p]),
    "ss_pedersen_hash": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ss_rows_gather": (c_int, [c_void_p, c_void_p, c_uint64, c_int, POINTER(c_uint64), c_size_t, c_void_p]),
    "ss_fri_fold": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_uint64, c_uint64, c_void_p, c_void_p]),
    "ss_inv_x_minus_c": (c_int, [c_void_p, c_int, c_int, c_int, c_uint64, c_uint64, c_void_p, c_void_p, c_void_p]),
    "ss_poly_eval": (c_int, [c_void_p, c_int, c_void_p, c_uint64, c_int, c_int, POINTER(ctypes.c_int32), c_void_p, c_size_t, c_void_p]),
    "ss_ood_eval": (c_int, [c_void_p, c_int, c_void_p, c_uint64, c_int, POINTER(ctypes.c_int32), POINTER(c_uint64), c_size_t, c_void_p, c_uint64, c_uint64, c_void_p]),
    "ss_pow_grind": (c_int, [c_void_p, c_int, POINTER(c_uint8), c_int, POINTER(c_uint64)]),
    "ss_ntt_shard": (c_int, [c_void_p, c_int, c_void_p, c_uint64, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_uint64, c_void_p]),
    "ss_shard_dft": (c_int, [c_void_p, c_int, c_void_p, c_uint64, c_void_p, c_uint64, c_uint64, c_int, c_int, c_int, c_uint64, c_void_p]),
    "ss_dist_unique_id": (c_int, [c_void_p, POINTER(c_uint8)]),
    "ss_dist_init": (c_int, [c_void_p, POINTER(c_uint8), c_int, c_int]),
    "ss_dist_finalize": (c_int, [c_void_p]),
    "ss_dist_lde": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ss_dist_halo": (c_int, [c_void_p, c_void_p, c_uint64, c_int, c_int, c_uint64, c_void_p]),
    "ss_dist_allgather": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "ss_dist_commit": (c_int, [c_void_p, c_int, c_int, c_void_p, c_uint64, c_int, c_int, POINTER(c_uint8), POINTER(c_void_p), POINTER(c_uint8), c_void_p]),
    "ss_dist_open": (c_int, [c_void_p, c_int, c_int, c_void_p, POINTER(c_uint8), POINTER(c_uint64), c_size_t, POINTER(c_uint8), c_void_p]),
    "ss_dist_gather_rows": (c_int, [c_void_p, c_void_p, c_uint64, c_int, c_int, POINTER(c_uint64), c_size_t, c_void_p, c_void_p]),
    "ss_perm_product": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_uint64, c_uint64, c_void_p, c_void_p, c_void_p, c_uint64, c_void_p]),
    "ss_diluted_aggregate": (c_int, [c_void_p, c_int, c_void_p, c_uint64, c_uint64, c_void_p, c_void_p, c_void_p, c_uint64, c_void_p]),
    "ss_constraint_eval": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_uint64, c_int, c_int, c_int, c_uint64, c_uint64, c_int, c_void_p, c_void_p]),
}

_lib = None


class SandstormError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"sandstorm_b200 status {status}: {message}")
        self.status = status


def load() -> ctypes.CDLL:
    """Loads the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  sandstorm_b200 has no CPU fallback."
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
