"""Loader of libcntmc.so (the CUDA engine behind include/cntmc.h).

There is no fallback: if the shared library is missing or cannot be loaded, importing anything that computes raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CNTMC_LIB") or os.path.join(HERE, "libcntmc.so")  # override: kernel experiments only

V, I32, I64, U64, D, CP = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double, C.c_char_p

# name -> (restype, argtypes); mirrors include/cntmc.h one to one (tests check that every symbol is exported)
SIGNATURES = {
    "cntmc_create": (C.c_int, [CP, C.POINTER(V)]),
    "cntmc_destroy": (None, [V]),
    "cntmc_last_error": (CP, [V]),
    "cntmc_version": (CP, []),
    "cntmc_set_device": (C.c_int, [V, C.c_int]),
    "cntmc_set_stream": (C.c_int, [V, V]),
    "cntmc_load_mesh": (C.c_int, [V, CP]),
    "cntmc_set_mesh": (C.c_int, [V, I64, I64, V, V]),
    "cntmc_set_rate_table": (C.c_int, [V, V, V, V, V, V, V]),
    "cntmc_get_rate_table_dims": (C.c_int, [V, V]),
    "cntmc_get_rate_table": (C.c_int, [V, V, V, V, V, V]),
    "cntmc_save_rate_table": (C.c_int, [V, CP]),
    "cntmc_load_rate_table": (C.c_int, [V, CP]),
    "cntmc_kubo_init": (C.c_int, [V]),
    "cntmc_kubo_create_particles": (C.c_int, [V, I64, U64, U64]),
    "cntmc_kubo_create_particles_replay": (C.c_int, [V, I64, V, V, V]),
    "cntmc_kubo_step": (C.c_int, [V, D, I64, V]),
    "cntmc_kubo_step_dev": (C.c_int, [V, D, I64, V]),
    "cntmc_kubo_step_host_state": (C.c_int, [V, D, I64, I64, V, V, V, V, V, V, V]),
    "cntmc_time": (D, [V]),
    "cntmc_kubo_max_time": (D, [V]),
    "cntmc_time_step": (D, [V]),
    "cntmc_number_of_particles": (I64, [V]),
    "cntmc_hops": (I64, [V]),
    "cntmc_reinjections": (I64, [V]),
    "cntmc_crossings": (I64, [V]),
    "cntmc_probes": (I64, [V]),
    "cntmc_init": (C.c_int, [V, I64, I64, U64, I64]),
    "cntmc_init_replay": (C.c_int, [V, I64, I64, I64, V, V, V]),
    "cntmc_get_gids": (C.c_int, [V, V]),
    "cntmc_step": (C.c_int, [V, D, I64, V, V]),
    "cntmc_step_dev": (C.c_int, [V, D, I64, V]),
    "cntmc_get_area": (C.c_int, [V, V]),
    "cntmc_num_contact_sites": (C.c_int, [V, C.c_int, V]),
    "cntmc_get_contact_sites": (C.c_int, [V, C.c_int, V]),
    "cntmc_number_of_segments": (C.c_int, [V]),
    "cntmc_get_scatterer_statistics": (C.c_int, [V, V]),
    "cntmc_track_particle": (C.c_int, [V, D, U64, U64, I64, V, V, I64, V, V, V]),
    "cntmc_num_sites": (C.c_int, [V, V]),
    "cntmc_get_sites": (C.c_int, [V, V, V, V, V, V, V]),
    "cntmc_get_domain": (C.c_int, [V, V]),
    "cntmc_get_removal_domain": (C.c_int, [V, V]),
    "cntmc_num_inject": (C.c_int, [V, V]),
    "cntmc_get_inject": (C.c_int, [V, V]),
    "cntmc_csr_nnz": (C.c_int, [V, V]),
    "cntmc_get_csr": (C.c_int, [V, V, V, V]),
    "cntmc_get_csr_row": (C.c_int, [V, I64, I64, V, V, V]),
    "cntmc_csr_midpoint_guards": (I64, [V]),
    "cntmc_csr_build_seconds": (D, [V]),
    "cntmc_get_particles": (C.c_int, [V, V, V, V, V, V, V]),
    "cntmc_trace_enable": (C.c_int, [V, I32]),
    "cntmc_trace_get": (C.c_int, [V, V, V]),
    "cntmc_set_option": (C.c_int, [V, CP, I64]),
    "cntmc_get_option": (I64, [V, CP]),
    "cntmc_last_step_ms": (D, [V]),
    "cntmc_last_step_launches": (I64, [V]),
    "cntmc_sync": (C.c_int, [V]),
    "cntmc_last_kernel_ms": (D, [V]),
    "cntmc_last_kernel_launches": (I64, [V]),
    "cntmc_dbg_walk_arith": (C.c_int, [C.c_int, I64, C.c_uint64, V]),
    "cntmc_multi_create": (C.c_int, [CP, C.c_int, V, C.POINTER(V)]),
    "cntmc_multi_destroy": (None, [V]),
    "cntmc_multi_last_error": (CP, [V]),
    "cntmc_multi_num_devices": (C.c_int, [V]),
    "cntmc_multi_handle": (V, [V, C.c_int]),
    "cntmc_multi_nccl_version": (C.c_int, []),
    "cntmc_multi_load_mesh": (C.c_int, [V, CP]),
    "cntmc_multi_set_mesh": (C.c_int, [V, I64, I64, V, V]),
    "cntmc_multi_set_option": (C.c_int, [V, CP, I64]),
    "cntmc_multi_kubo_init": (C.c_int, [V]),
    "cntmc_multi_kubo_create_particles": (C.c_int, [V, I64, U64]),
    "cntmc_multi_kubo_step": (C.c_int, [V, D, I64, V]),
    "cntmc_multi_get_particles": (C.c_int, [V, V, V, V, V, V, V]),
    "cntmc_multi_number_of_particles": (I64, [V]),
    "cntmc_multi_hops": (I64, [V]),
    "cntmc_multi_time": (D, [V]),
    "cntmc_multi_init": (C.c_int, [V, I64, I64, U64]),
    "cntmc_multi_step": (C.c_int, [V, D, I64, V, V]),
    "cntmc_davoody_last_error": (CP, []),
    "cntmc_tube_create": (V, [C.c_int, C.c_int, C.c_int]),
    "cntmc_tube_destroy": (None, [V]),
    "cntmc_tube_info": (C.c_int, [V, V, V]),
    "cntmc_tube_exciton_dims": (C.c_int, [V, C.c_int, V]),
    "cntmc_tube_exciton_energy": (C.c_int, [V, C.c_int, V]),
    "cntmc_transfer_create": (V, [V, V, D, D, C.c_int]),
    "cntmc_transfer_destroy": (None, [V]),
    "cntmc_transfer_info": (C.c_int, [V, V, V]),
    "cntmc_transfer_pair_factors": (C.c_int, [V, V, V, V]),
    "cntmc_transfer_first_order": (C.c_int, [V, I64, V, V, V, V, V]),
    "cntmc_transfer_table": (C.c_int, [V, V, V, V, V, V, V]),
    "cntmc_create_davoody_table": (C.c_int, [V, V, V, V, V, V, V]),
    "cntmc_fp64_peak": (C.c_int, [C.c_int, V]),
    "cntmc_hermitian_eig": (C.c_int, [C.c_int, V, V, V]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen libcntmc.so and declare every entry point.  Raises if the engine has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build the CUDA engine with `python -m cnt_film_monte_carlo_b200.build` "
                "(there is no CPU fallback)"
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            if os.environ.get("CNTMC_LIB") and not hasattr(lib, name):
                continue  # kernel experiments: an older build of the library
            fn = getattr(lib, name)  # AttributeError here = header and library disagree
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib
