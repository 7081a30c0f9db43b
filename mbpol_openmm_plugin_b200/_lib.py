"""ctypes binding of the C ABI in include/mbpol_b200.h (libmbpol_b200.so, built in-tree by build.py).

This module is the *only* way Python reaches the compute path; there is no CPU fallback: if the shared library is missing
or no CUDA device is present, calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libmbpol_b200.so")

ABI_VERSION = 2
OK, EINVAL, ECUDA, EBOX, ECHANGED, ENOMEM, ECOMM = 0, -1, -2, -3, -4, -5, -6
NOCUTOFF, CUTOFF_PERIODIC, CUTOFF_NONPERIODIC = 0, 1, 2
DO_2B, DO_3B, KEEP_LISTS, OVERWRITE_FORCES = 1, 2, 4, 8
DIST_ALLREDUCE, DIST_NO_BCAST = 16, 32
COMM_ID_BYTES = 128
FLOP_PER_DIMER, FLOP_PER_TRIMER = 30601, 111818

# every symbol include/mbpol_b200.h declares (checked by tests/test_abi.py)
EXPORTS = (
    "mbpol_create", "mbpol_destroy", "mbpol_last_error", "mbpol_update_particles", "mbpol_set_box",
    "mbpol_set_partition", "mbpol_eval", "mbpol_eval_host", "mbpol_add_forces_fixed_point", "mbpol_get_pairs",
    "mbpol_get_triplets", "mbpol_get_stats", "mbpol_set_timing", "mbpol_pip_eval", "mbpol_eval_tuples",
    "mbpol_measure_fp64_peak", "mbpol_pip_instruction_count", "mbpol_eval_onebody", "mbpol_eval_onebody_host",
    "mbpol_energy_device", "mbpol_get_energy", "mbpol_eval_openmm", "mbpol_comm_unique_id", "mbpol_comm_init_rank",
    "mbpol_comm_adopt", "mbpol_comm_destroy", "mbpol_eval_distributed", "mbpol_eval_distributed_host", "mbpol_set_skin", "mbpol_nve_run", "mbpol_eval_dispersion", "mbpol_eval_dispersion_host",
)


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("n_molecules", C.c_int32), ("n_particles", C.c_int32),
                ("particle_indices", C.POINTER(C.c_int32)), ("cutoff_2b_nm", C.c_double), ("cutoff_3b_nm", C.c_double),
                ("method_2b", C.c_int32), ("method_3b", C.c_int32), ("device", C.c_int32), ("reserved", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("n_pairs", C.c_int64), ("n_triplets", C.c_int64), ("n_pairs_owned", C.c_int64),
                ("n_triplets_owned", C.c_int64), ("ms_lists", C.c_float), ("ms_2b", C.c_float), ("ms_3b", C.c_float),
                ("ms_reduce", C.c_float), ("ms_pip_2b", C.c_float), ("ms_pip_3b", C.c_float), ("n_launches", C.c_int64),
                ("ms_bcast", C.c_float), ("ms_comm_reduce", C.c_float), ("lists_sliced", C.c_int32),
                ("lists_reused", C.c_int32), ("n_repartitions", C.c_int64), ("n_row_builds", C.c_int64)]


class NveConfig(C.Structure):
    _fields_ = [("dt_ps", C.c_double), ("flags", C.c_uint32), ("onebody_periodic", C.c_int32),
                ("msite_index", C.POINTER(C.c_int32)), ("msite_weights", C.c_double * 3)]


NVE_ONEBODY = 64


class MBPolError(RuntimeError):
    """Raised for any non-zero return of the C ABI (the adapter's OpenMMException analogue)."""

    def __init__(self, code, msg):
        super().__init__(f"[mbpol_b200 error {code}] {msg}")
        self.code = code
        self.message = msg


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, dp, ip = C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)
    L.mbpol_create.argtypes = [C.POINTER(vp), C.POINTER(Config)]
    L.mbpol_destroy.argtypes = [vp]
    L.mbpol_destroy.restype = None
    L.mbpol_last_error.argtypes = [vp]
    L.mbpol_last_error.restype = C.c_char_p
    L.mbpol_update_particles.argtypes = [vp, C.c_int32, ip]
    L.mbpol_set_box.argtypes = [vp, C.POINTER(C.c_double)]
    L.mbpol_set_partition.argtypes = [vp, C.c_int32, C.c_int32]
    L.mbpol_set_skin.argtypes = [vp, C.c_double]
    L.mbpol_eval_dispersion.argtypes = [vp, dp, dp, C.POINTER(C.c_double), C.c_double, C.c_int32, vp]
    L.mbpol_eval_dispersion_host.argtypes = [vp, dp, dp, C.POINTER(C.c_double), C.c_double, C.c_int32]
    L.mbpol_nve_run.argtypes = [vp, dp, dp, dp, C.POINTER(NveConfig), C.c_int32, dp, vp]
    L.mbpol_eval.argtypes = [vp, dp, dp, C.POINTER(C.c_double), C.c_uint32, vp]
    L.mbpol_eval_host.argtypes = [vp, dp, dp, C.POINTER(C.c_double), C.c_uint32]
    L.mbpol_eval_onebody.argtypes = [vp, dp, dp, C.POINTER(C.c_double), C.c_int32, vp]
    L.mbpol_eval_onebody_host.argtypes = [vp, dp, dp, C.POINTER(C.c_double), C.c_int32]
    L.mbpol_add_forces_fixed_point.argtypes = [vp, vp, C.c_int32, vp]
    L.mbpol_get_pairs.argtypes = [vp, vp, C.POINTER(C.c_size_t)]
    L.mbpol_get_triplets.argtypes = [vp, vp, C.POINTER(C.c_size_t)]
    L.mbpol_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.mbpol_set_timing.argtypes = [vp, C.c_int32]
    L.mbpol_pip_eval.argtypes = [C.c_int, dp, dp, C.c_int64, C.c_int64, vp]
    L.mbpol_eval_tuples.argtypes = [C.c_int, dp, C.c_int32, vp, C.c_int64, C.POINTER(C.c_double), C.c_int, dp, dp, vp]
    L.mbpol_measure_fp64_peak.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.mbpol_pip_instruction_count.argtypes = [C.c_int]
    L.mbpol_energy_device.argtypes = [vp]
    L.mbpol_energy_device.restype = C.c_void_p
    L.mbpol_get_energy.argtypes = [vp, C.POINTER(C.c_double), vp]
    L.mbpol_eval_openmm.argtypes = [vp, dp, C.c_int32, dp, C.c_int32, dp, C.c_int32, dp, C.c_int32, C.c_uint32, vp]
    L.mbpol_comm_unique_id.argtypes = [C.c_char_p]
    L.mbpol_comm_init_rank.argtypes = [vp, C.c_char_p, C.c_int32, C.c_int32]
    L.mbpol_comm_adopt.argtypes = [vp, vp]
    L.mbpol_comm_destroy.argtypes = [vp]
    L.mbpol_eval_distributed.argtypes = [vp, dp, dp, C.POINTER(C.c_double), C.c_uint32, vp]
    L.mbpol_eval_distributed_host.argtypes = [vp, dp, dp, C.POINTER(C.c_double), C.c_uint32]
    for name in EXPORTS:
        fn = getattr(L, name)
        if name not in ("mbpol_destroy", "mbpol_last_error", "mbpol_energy_device"):
            fn.restype = C.c_int
    _lib = L
    return L


def check(rc, ctx=None):
    if rc != OK:
        msg = load().mbpol_last_error(ctx)
        raise MBPolError(rc, msg.decode() if msg else "unknown error")
