"""ctypes binding of librodeo_cuda.so (the C ABI in include/rodeo_cuda.h).

There is no fallback: if the library is missing or cannot be loaded this module raises,
and so does everything built on it.
"""
import ctypes
import os

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_PKG, "lib", "librodeo_cuda.so")

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int_p = ctypes.POINTER(ctypes.c_int32)


class ProblemDesc(ctypes.Structure):
    _fields_ = [("n_state", ctypes.c_int32), ("n_meas", ctypes.c_int32), ("n_eval", ctypes.c_int32),
                ("ode", ctypes.c_int32), ("interrogate", ctypes.c_int32), ("flags", ctypes.c_int32),
                ("tmin", ctypes.c_double), ("tmax", ctypes.c_double),
                ("wgt_meas", ctypes.c_void_p), ("wgt_state", ctypes.c_void_p),
                ("mu_state", ctypes.c_void_p), ("var_state", ctypes.c_void_p)]


# every symbol include/rodeo_cuda.h declares: (name, restype, argtypes)
_V, _I, _L, _D, _Z = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_size_t
SYMBOLS = [
    ("rodeo_cuda_abi_version", _I, []),
    ("rodeo_cuda_last_error", ctypes.c_char_p, []),
    ("rodeo_cuda_ode_count", _I, []),
    ("rodeo_cuda_ode_n_meas", _I, [_I]),
    ("rodeo_cuda_ode_n_theta", _I, [_I]),
    ("rodeo_cuda_supported", _I, [_I, _I, _I, _I]),
    ("rodeo_cuda_supported_blocks", _I, [_I, _I, _V, _I]),
    ("rodeo_cuda_kernel_set_count", _I, []),
    ("rodeo_cuda_kernel_set_info", _I, [_I, c_int_p, c_int_p, c_int_p, c_int_p, c_int_p,
                                        ctypes.POINTER(ctypes.c_char_p)]),
    ("rodeo_cuda_problem_create", _I, [ctypes.POINTER(ProblemDesc), ctypes.POINTER(_V)]),
    ("rodeo_cuda_problem_destroy", _I, [_V]),
    ("rodeo_cuda_problem_set", _I, [_V, _I, _V]),
    ("rodeo_cuda_problem_info", _I, [_V, c_int_p, c_int_p, c_int_p]),
    ("rodeo_cuda_problem_structured", _I, [_V]),
    ("rodeo_cuda_history_bytes_per_system", _Z, [_V]),
    ("rodeo_cuda_set_workspace_limit", _I, [_V, _Z]),
    ("rodeo_cuda_set_checkpoint", _I, [_V, _I]),
    ("rodeo_cuda_get_checkpoint", _I, [_V]),
    ("rodeo_cuda_set_var_layout", _I, [_V, _I]),
    ("rodeo_cuda_get_var_layout", _I, [_V]),
    ("rodeo_cuda_var_doubles_per_step", _I, [_V]),
    ("rodeo_cuda_set_shared_cov", _I, [_V, _I]),
    ("rodeo_cuda_get_shared_cov", _I, [_V]),
    ("rodeo_cuda_reserve", _I, [_V, _L]),
    ("rodeo_cuda_solve", _I, [_V, _I, _L, _V, _V, _V, _L, _V, _V, _V, _V, _V]),
    ("rodeo_cuda_solve_batched_prior", _I, [_V, _I, _L, _V, _V, _V, _L, _V, _V, _V, _V, _V, _V, _V, _V, _V]),
    ("rodeo_cuda_last_batched_prior_blockdiag", _I, [_V]),
    ("rodeo_cuda_loglik_batched_prior", _I, [_V, _I, _L, _V, _V, _V, _L, _V, _V, _V, _V, _V, ctypes.c_int32, _V,
                                             ctypes.c_int32, _V, _D, _V, _V, _V]),
    ("rodeo_cuda_solve_host", _I, [_V, _I, _L, _V, _V, _V, _L, _V, _V, _V, _V]),
    ("rodeo_cuda_loglik", _I, [_V, _I, _L, _V, _V, _V, _L, _V, ctypes.c_int32, _V, ctypes.c_int32,
                               _V, _D, _V, _V, _V]),
    ("rodeo_cuda_debug_copy_history", _I, [_V, _V, _Z, _V]),
    ("rodeo_cuda_debug_set_row_limit", _I, [_V, _D]),
    ("rodeo_cuda_launch_count", _L, [_V]),
    ("rodeo_cuda_reset_launch_count", _I, [_V]),
    ("rodeo_cuda_last_kernel_ms", _I, [_V, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]),
    ("rodeo_cuda_kernel_ms_total", _I, [_V, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                        c_int_p]),
    ("rodeo_cuda_set_timing", _I, [_V, _I]),
]

_lib = None


class RodeoCudaError(RuntimeError):
    pass


def load():
    """Load librodeo_cuda.so; raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m rodeo_legacy_b200.build` "
            "(nvcc, sm_100a).  rodeo_legacy_b200.cuda has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)          # AttributeError if the ABI symbol is absent
        fn.restype = res
        fn.argtypes = args
    if lib.rodeo_cuda_abi_version() != 1:
        raise RuntimeError("librodeo_cuda.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().rodeo_cuda_last_error().decode()
        if rc == -2:
            raise NotImplementedError(msg)
        if rc == -1:
            raise ValueError(msg)
        if rc == -4:
            raise MemoryError(msg)
        raise RodeoCudaError(msg)
