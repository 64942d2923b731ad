"""ctypes binding of librve_b200.so (include/rve_b200.h).  Fails loudly: there is no CPU fallback."""
import ctypes
import os

import numpy as np

from .build import LIB

c_int32_p = ctypes.POINTER(ctypes.c_int32)
c_int64_p = ctypes.POINTER(ctypes.c_int64)
c_double_p = ctypes.POINTER(ctypes.c_double)

MODEL = {"lin": 0, "dnn": 1, "per": 2, "MR": 3}
PRECOND = {"jacobi": 0, "twolevel": 1}
OPT_SYM_B, OPT_OPERATOR, OPT_RESIDUAL_REPLACE = 1, 2, 3
OPERATOR = {"assembled": 0, "matfree": 1}

# every symbol include/rve_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = ["rve_create", "rve_destroy", "rve_last_error", "rve_set_batch", "rve_set_batch_shared", "rve_set_batch_morphed", "rve_assemble", "rve_get_sizes",
           "rve_get_csr", "rve_get_rowmap", "rve_get_nodemap", "rve_get_rhs", "rve_get_solution",
           "rve_get_coarse_dim", "rve_get_coarse_dense", "rve_solve", "rve_epilogue_tangent",
           "rve_epilogue_snapshot", "rve_epilogue_homogenisation", "rve_set_option", "rve_project", "rve_timings", "rve_bench_spmv", "rve_apply_operator", "rve_host_symbolic"]

_lib = None


class RVEError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("librve_b200 error %d: %s" % (code, msg))
        self.code = code


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB):
        raise ImportError("librve_b200.so is not built: run `python -m deepbnd_b200.build` "
                          "(or __graft_entry__.build()); there is no CPU fallback")
    lib = ctypes.CDLL(LIB)
    for s in SYMBOLS:
        getattr(lib, s)
    lib.rve_last_error.restype = ctypes.c_char_p
    lib.rve_last_error.argtypes = [ctypes.c_void_p]
    lib.rve_create.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
    lib.rve_destroy.argtypes = [ctypes.c_void_p]
    lib.rve_set_batch.argtypes = [ctypes.c_void_p, ctypes.c_int32, c_int64_p, c_int64_p, c_double_p, c_int32_p,
                                  c_int32_p, c_double_p, ctypes.c_int32, ctypes.c_int32, c_double_p]
    lib.rve_set_batch_shared.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, c_double_p, c_int32_p,
                                         c_int32_p, c_double_p, ctypes.c_int32, ctypes.c_int32, c_double_p]
    lib.rve_set_batch_morphed.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, c_double_p, c_int32_p,
                                          c_int32_p, c_double_p, ctypes.c_int32, ctypes.c_int32, c_double_p, ctypes.c_int32, c_int32_p,
                                          c_int32_p, c_double_p, c_double_p, ctypes.c_int32, c_double_p, ctypes.c_double,
                                          ctypes.c_double, c_double_p]
    lib.rve_assemble.argtypes = [ctypes.c_void_p]
    lib.rve_get_sizes.argtypes = [ctypes.c_void_p, ctypes.c_int32, c_int64_p, c_int64_p, c_int64_p, c_int64_p]
    lib.rve_get_csr.argtypes = [ctypes.c_void_p, ctypes.c_int32, c_int32_p, c_int32_p, c_double_p]
    lib.rve_get_rowmap.argtypes = [ctypes.c_void_p, ctypes.c_int32, c_int32_p, c_int32_p]
    lib.rve_get_nodemap.argtypes = [ctypes.c_void_p, ctypes.c_int32, c_int32_p, c_int32_p]
    lib.rve_get_rhs.argtypes = [ctypes.c_void_p, ctypes.c_int32, c_double_p]
    lib.rve_get_solution.argtypes = [ctypes.c_void_p, ctypes.c_int32, c_double_p]
    lib.rve_get_coarse_dim.argtypes = [ctypes.c_void_p, ctypes.c_int32, c_int32_p, c_int32_p]
    lib.rve_get_coarse_dense.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, c_double_p]
    lib.rve_solve.argtypes = [ctypes.c_void_p, ctypes.c_int32, c_double_p, c_double_p, ctypes.c_double,
                              ctypes.c_int32, ctypes.c_int32, c_int32_p, c_int32_p]
    lib.rve_epilogue_tangent.argtypes = [ctypes.c_void_p, c_double_p]
    lib.rve_epilogue_snapshot.argtypes = [ctypes.c_void_p] + [c_double_p] * 5
    lib.rve_epilogue_homogenisation.argtypes = [ctypes.c_void_p] + [c_double_p] * 5
    lib.rve_set_option.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32]
    lib.rve_project.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, c_double_p, c_double_p, c_double_p,
                                c_double_p]
    lib.rve_timings.argtypes = [ctypes.c_void_p, c_double_p, c_double_p]
    lib.rve_apply_operator.argtypes = [ctypes.c_void_p, ctypes.c_int32, c_double_p, c_double_p]
    lib.rve_bench_spmv.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, c_double_p, c_double_p]
    lib.rve_host_symbolic.argtypes = [c_double_p, ctypes.c_int32, c_int32_p, c_int32_p, ctypes.c_int32, ctypes.c_int32,
                                      ctypes.c_int32, ctypes.c_int32, ctypes.c_int32] + [c_int32_p] * 10 + [c_double_p]
    _lib = lib
    return lib


def dptr(a):
    return None if a is None else a.ctypes.data_as(c_double_p)


def iptr(a):
    return None if a is None else a.ctypes.data_as(c_int32_p)


def lptr(a):
    return None if a is None else a.ctypes.data_as(c_int64_p)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)
