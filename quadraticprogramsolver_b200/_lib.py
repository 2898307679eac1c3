"""ctypes binding of libqpb200.so (the C ABI of include/qpb200.h).

The library is built in-tree (``quadraticprogramsolver_b200/libqpb200.so``) by
``__graft_entry__.build()`` / ``csrc/Makefile``.  There is NO fallback: if the shared library is
missing, importing a solver entry point raises; if no sm_100 GPU is present, ``*_create`` returns
``QPB200_ERR_DEVICE`` which surfaces as :class:`QPB200Error`.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QPB200_LIB", os.path.join(_HERE, "libqpb200.so"))   # QPB200_LIB: A/B builds

QPB200_OK = 0
ERR_ARG, ERR_NONFINITE, ERR_CUDA, ERR_NCCL, ERR_FACTOR, ERR_DEVICE = -1, -2, -3, -4, -5, -6
LINSOLVE_PCG, LINSOLVE_CHOLESKY = 0, 1
PRECOND_NONE, PRECOND_JACOBI = 0, 1


class QPB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libqpb200 error {code}: {msg}")
        self.code = code


class Settings(C.Structure):
    """``qpb200_settings`` (include/qpb200.h) -- the kwargs of SolveQuadraticProgram.jl:15-17."""
    _fields_ = [("max_iter", C.c_int64), ("eps_abs", C.c_double), ("eps_rel", C.c_double), ("rho", C.c_double),
                ("sigma", C.c_double), ("alpha", C.c_double), ("delta", C.c_double), ("adaptive_rho", C.c_int32),
                ("lin_solver", C.c_int32), ("rho_factor", C.c_double), ("check_every", C.c_int64),
                ("polish_iter", C.c_int64), ("minres_eps", C.c_double), ("minres_iter", C.c_int64),
                ("pcg_eps", C.c_double), ("pcg_max_iter", C.c_int64), ("pcg_rel_eps", C.c_double),
                ("precond", C.c_int32), ("device", C.c_int32), ("spmv_loader", C.c_int32),
                ("reserved_i", C.c_int32 * 7), ("reserved_d", C.c_double * 4)]


class Info(C.Structure):
    """``qpb200_info`` (include/qpb200.h)."""
    _fields_ = [("conv_flag", C.c_int32), ("polish_status", C.c_int32), ("iterations", C.c_int64),
                ("rho_final", C.c_double), ("res_prim", C.c_double), ("res_dual", C.c_double),
                ("rho_updates", C.c_int64), ("pcg_iters_total", C.c_int64), ("pcg_maxed", C.c_int64),
                ("solve_ms", C.c_double), ("setup_ms", C.c_double), ("kernel_launches", C.c_int64),
                ("polish_minres_iters", C.c_int64), ("polish_active", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class ProxReport(C.Structure):
    """``qpb200_proxqp_report`` (include/qpb200.h)."""
    _fields_ = [("converged", C.c_int32), ("reserved", C.c_int32), ("iterations", C.c_int64), ("rho", C.c_double),
                ("sigma", C.c_double), ("res_prim", C.c_double), ("res_dual", C.c_double), ("rho_updates", C.c_int64),
                ("solve_ms", C.c_double), ("kernel_launches", C.c_int64)]


# every symbol include/qpb200.h declares
EXPORTS = [
    "qpb200_version", "qpb200_device_count", "qpb200_default_settings", "qpb200_last_error",
    "qpb200_create", "qpb200_solve", "qpb200_update_vectors", "qpb200_update_settings", "qpb200_set_rho_scale",
    "qpb200_destroy", "qpb200_proxqp_default_settings", "qpb200_proxqp_solve",
    "qpb200_apply", "qpb200_time_apply", "qpb200_apply_bytes",
    "qpb200_batch_create", "qpb200_batch_solve", "qpb200_batch_update_vectors", "qpb200_batch_destroy",
    "qpb200_batch_solve_once", "qpb200_batch_create_shared",
    "qpb200_dist_unique_id", "qpb200_dist_create", "qpb200_dist_create_full", "qpb200_dist_rows", "qpb200_dist_solve",
    "qpb200_debug_tile_plan", "qpb200_debug_tile_nnz", "qpb200_debug_equilibrate", "qpb200_debug_assemble_h",
    "qpb200_debug_partition", "qpb200_debug_slice",
]

_lib = None


def load():
    """Load libqpb200.so (once).  Raises if it has not been built -- there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a).  quadraticprogramsolver_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    p64, pd, pv = C.POINTER(C.c_int64), C.POINTER(C.c_double), C.c_void_p
    lib.qpb200_version.restype = C.c_int
    lib.qpb200_device_count.restype = C.c_int
    lib.qpb200_last_error.restype = C.c_char_p
    lib.qpb200_default_settings.argtypes = [C.POINTER(Settings)]
    lib.qpb200_default_settings.restype = None
    lib.qpb200_create.argtypes = [C.POINTER(pv), C.c_int64, C.c_int64, p64, p64, pd, p64, p64, pd, pd, pd, pd,
                                  C.POINTER(Settings), C.c_int32]
    lib.qpb200_solve.argtypes = [pv, pd, pd, pd, C.POINTER(Info)]
    lib.qpb200_update_vectors.argtypes = [pv, pd, pd, pd]
    lib.qpb200_update_settings.argtypes = [pv, C.POINTER(Settings)]
    lib.qpb200_set_rho_scale.argtypes = [pv, pd]
    lib.qpb200_proxqp_default_settings.argtypes = [C.POINTER(Settings)]
    lib.qpb200_proxqp_default_settings.restype = None
    lib.qpb200_proxqp_solve.argtypes = [pv, C.c_int64, C.POINTER(Settings), pd, pd, pd, pd, C.c_int32, C.POINTER(ProxReport)]
    lib.qpb200_destroy.argtypes = [pv]
    lib.qpb200_destroy.restype = None
    lib.qpb200_apply.argtypes = [pv, C.c_int32, pd, pd]
    lib.qpb200_time_apply.argtypes = [pv, C.c_int32, C.c_int32, C.c_int32, pd]
    lib.qpb200_apply_bytes.argtypes = [pv, C.c_int32]
    lib.qpb200_apply_bytes.restype = C.c_int64
    lib.qpb200_batch_create.argtypes = [C.POINTER(pv), C.c_int64, C.c_int64, C.c_int64, pd, pd, pd, pd, pd,
                                        C.POINTER(Settings)]
    lib.qpb200_batch_create_shared.argtypes = [C.POINTER(pv), C.c_int64, C.c_int64, C.c_int64, pd, pd, pd, pd, pd,
                                               C.POINTER(Settings)]
    lib.qpb200_batch_solve.argtypes = [pv, pd, C.POINTER(C.c_int32), p64, C.POINTER(Info)]
    lib.qpb200_batch_solve_once.argtypes = [C.c_int64, C.c_int64, C.c_int64, pd, pd, pd, pd, pd, C.POINTER(Settings), pd,
                                            C.POINTER(C.c_int32), p64, C.POINTER(Info)]
    lib.qpb200_batch_update_vectors.argtypes = [pv, pd, pd, pd]
    lib.qpb200_batch_destroy.argtypes = [pv]
    lib.qpb200_batch_destroy.restype = None
    lib.qpb200_dist_unique_id.argtypes = [pv]
    lib.qpb200_dist_create.argtypes = [C.POINTER(pv), C.c_int32, C.c_int32, pv, C.c_int64, C.c_int64,
                                       p64, p64, pd, p64, p64, pd, pd, pd, pd, C.POINTER(Settings), C.c_int32]
    lib.qpb200_dist_create_full.argtypes = [C.POINTER(pv), C.c_int32, C.c_int32, pv, C.c_int64, C.c_int64,
                                            p64, p64, pd, p64, p64, pd, pd, pd, pd, C.POINTER(Settings), C.c_int32]
    lib.qpb200_dist_rows.argtypes = [pv, p64, p64]
    lib.qpb200_dist_solve.argtypes = [pv, pd, pd, pd, C.POINTER(Info)]
    p32 = C.POINTER(C.c_int32)
    lib.qpb200_debug_tile_plan.argtypes = [C.c_int32, p32, C.c_int32, p32, C.c_int64, p32, p32]
    lib.qpb200_debug_tile_plan.restype = C.c_int64
    lib.qpb200_debug_tile_nnz.restype = C.c_int32
    lib.qpb200_debug_equilibrate.argtypes = [C.c_int64, C.c_int64, p64, p64, pd, p64, p64, pd, pd, C.c_int32, C.c_int32,
                                             pd, pd, pd, pd, pd, pd]
    lib.qpb200_debug_equilibrate.restype = C.c_int
    lib.qpb200_debug_assemble_h.argtypes = [C.c_int64, C.c_int64, p64, p64, pd, p64, p64, pd, C.c_int32, p32, p32, p32, pd, pd, pd]
    lib.qpb200_debug_assemble_h.restype = C.c_int
    lib.qpb200_debug_partition.argtypes = [C.c_int64, C.c_int64, p64, p64, p64, C.c_int32, C.c_int32, p64, p64]
    lib.qpb200_debug_partition.restype = C.c_int64
    lib.qpb200_debug_slice.argtypes = [C.c_int64, C.c_int64, p64, p64, p64, pd, C.c_int32, C.c_int32, C.c_int32, p64, p64, p64,
                                       p64, p64, pd, C.c_int64]
    lib.qpb200_debug_slice.restype = C.c_int
    _lib = lib
    return lib


def check(rc: int):
    if rc != QPB200_OK:
        raise QPB200Error(rc, load().qpb200_last_error().decode("utf-8", "replace"))


def default_settings() -> Settings:
    s = Settings()
    load().qpb200_default_settings(C.byref(s))
    return s
