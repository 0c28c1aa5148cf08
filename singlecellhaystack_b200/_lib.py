"""ctypes binding of libhaystack_b200.so (C ABI in include/haystack_b200.h).

There is deliberately no fallback: if the shared library is missing, or no B200 is usable,
loading / context creation raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("HAYSTACK_B200_LIB", _PKG / "lib" / "libhaystack_b200.so"))

HS_OK, HS_ERR_INVALID, HS_ERR_STATE, HS_ERR_CUDA, HS_ERR_NOMEM = 0, 1, 2, 3, 4
ABI_VERSION = 1

_i64, _i32, _int = C.c_int64, C.c_int32, C.c_int
_vp = C.c_void_p
_dp = C.POINTER(C.c_double)

# name -> (restype, argtypes); mirrors include/haystack_b200.h line by line
SIGNATURES = {
    "hs_abi_version": (_int, []),
    "hs_create": (_int, [_int, C.POINTER(_vp)]),
    "hs_destroy": (None, [_vp]),
    "hs_last_error": (C.c_char_p, [_vp]),
    "hs_set_stream": (_int, [_vp, _vp]),
    "hs_synchronize": (_int, [_vp]),
    "hs_density": (_int, [_vp, _vp, _i64, _int, _vp, _int, _vp, _vp, _vp, _vp, _vp]),
    "hs_density_buffers": (_int, [_vp, _i64, _int, C.POINTER(_vp), C.POINTER(_i64), C.POINTER(_vp), C.POINTER(_i64)]),
    "hs_density_commit": (_int, [_vp, _i64, _int]),
    "hs_density_shape": (_int, [_vp, C.POINTER(_i64), C.POINTER(_int)]),
    "hs_dkl_dense": (_int, [_vp, _vp, _i64, _i64, _vp]),
    "hs_dkl_csr": (_int, [_vp, _vp, _int, _vp, _vp, _i64, _vp]),
    "hs_dkl_csr_f32": (_int, [_vp, _vp, _int, _vp, _vp, _i64, _vp]),
    "hs_dkl_perm_csr_begin": (_int, [_vp, _vp, _i64, _vp, _vp, _int, _int]),
    "hs_dkl_perm_csr_wait": (_int, [_vp, _vp]),
    "hs_dkl_perm_csr_f32": (_int, [_vp, _vp, _i64, _vp, _vp, _int, _int, _vp]),
    "hs_dkl_perm_dense": (_int, [_vp, _vp, _vp, _int, _int, _vp]),
    "hs_dkl_perm_csr": (_int, [_vp, _vp, _i64, _vp, _vp, _int, _int, _vp]),
    "hs_dkl_perm_csr_seeded": (_int, [_vp, _vp, _i64, _vp, _int, C.c_uint64, C.c_uint64, _vp]),
    "hs_dkl_perm_dense_seeded": (_int, [_vp, _vp, _int, C.c_uint64, C.c_uint64, _vp]),
    "hs_dkl_perm_dense_seeded_batch": (_int, [_vp, _vp, _i64, _vp, _i64, _int, C.c_uint64, C.c_uint64, _vp]),
    "hs_dkl_perm_dense_seeded_dev": (_int, [_vp, _vp, _int, C.c_uint64, C.c_uint64, _vp]),
    "hs_dkl_csr_randomized": (_int, [_vp, _vp, _int, _vp, _vp, _i64, _vp, _i64, _int, C.c_uint64, C.c_uint64, _vp, _vp]),
    "hs_perm_sample": (_int, [C.c_uint64, C.c_uint64, _i64, _i64, _int, _vp]),
    "hs_dkl_csr_dev": (_int, [_vp, _vp, _vp, _vp, _i64, _vp]),
    "hs_dkl_perm_csr_seeded_dev": (_int, [_vp, _vp, _i64, _vp, _int, C.c_uint64, C.c_uint64, _vp]),
    "hs_dkl_perm_csr_dev": (_int, [_vp, _vp, _i64, _vp, _vp, _int, _int, _vp]),
    "hs_dkl_perm_csr_seeded_units_dev": (_int, [_vp, _vp, _i64, _vp, _int, C.c_uint64, _vp, _vp]),
    "hs_density_shard_dist_dev": (_int, [_vp, _vp, _i64, _int, _vp, _int, _int, _int]),
    "hs_density_shard_buffers": (_int, [_vp, _i64, _int, _int, C.POINTER(_vp), C.POINTER(_i64), C.POINTER(_vp),
                                        C.POINTER(_i64), C.POINTER(_vp), C.POINTER(_i64)]),
    "hs_density_shard_dens_dev": (_int, [_vp, _i64, _int, _vp, _int, _int]),
    "hs_density_shard_finish": (_int, [_vp, _i64, _int, _vp, _vp]),
    "hs_dkl_dense_dev": (_int, [_vp, _vp, _i64, _i64, _vp]),
    "hs_dkl_perm_dense_dev": (_int, [_vp, _vp, _vp, _int, _int, _vp]),
    "hs_density_dev": (_int, [_vp, _vp, _i64, _int, _vp, _int, _vp, _vp, _vp]),
    "hs_profile_csr": (_int, [_vp, _vp, _int, _vp, _vp, _i64, _vp]),
    "hs_profile_dense": (_int, [_vp, _vp, _i64, _i64, _vp]),
    "hs_row_stats_csr": (_int, [_vp, _vp, _int, _vp, _i64, _i64, _vp, _vp]),
    "hs_row_stats_csr_dev": (_int, [_vp, _vp, _vp, _i64, _i64, _vp, _vp]),
    "hs_grid_seed_begin": (_int, [_vp, _vp, _i64, _int]),
    "hs_grid_seed_step": (_int, [_vp, _i64, _vp, _vp]),
    "hs_select_by_cv_dev": (_int, [_vp, _vp, _vp, _i64, _vp, _i64, _vp, _vp]),
    "hs_csc_to_csr": (_int, [_vp, _vp, _int, _vp, _vp, _i64, _i64, _vp, _vp, _vp]),
    "hs_csc_to_csr_dev": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp]),
    "hs_r_rng_create": (_int, [C.c_uint32, _int, C.POINTER(_vp)]),
    "hs_r_rng_destroy": (None, [_vp]),
    "hs_r_unif_rand": (C.c_double, [_vp]),
    "hs_r_sample_int": (_int, [_vp, _i32, _i32, _i32, _vp]),
    "hs_r_sample_prob_one": (_int, [_vp, _i32, _vp, _vp]),
    "hs_r_kmeans": (_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "hs_launch_count": (_i64, [_vp]),
    "hs_last_timers": (_int, [_vp, C.POINTER(C.c_float), _int]),
    "hs_enable_timers": (_int, [_vp, _int]),
}

_lib = None


class HaystackLibraryError(RuntimeError):
    """The CUDA library is missing, could not be loaded, or a call into it failed."""


def load() -> C.CDLL:
    """Load libhaystack_b200.so (once) and declare every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise HaystackLibraryError(
            f"{LIB_PATH} not found. Build it with `python -m singlecellhaystack_b200.build` "
            "(needs nvcc). There is no CPU fallback.")
    try:
        lib = C.CDLL(str(LIB_PATH))
    except OSError as e:  # pragma: no cover - depends on the machine
        raise HaystackLibraryError(f"could not load {LIB_PATH}: {e}") from e
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means header and library disagree
        fn.restype = res
        fn.argtypes = args
    if lib.hs_abi_version() != ABI_VERSION:
        raise HaystackLibraryError(
            f"ABI mismatch: library reports {lib.hs_abi_version()}, binding expects {ABI_VERSION}")
    _lib = lib
    return lib


def last_error(handle) -> str:
    msg = load().hs_last_error(handle)
    return msg.decode("utf-8", "replace") if msg else ""
