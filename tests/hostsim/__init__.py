"""ctypes driver for tests/hostsim/hostsim.cpp (the CUDA kernel bodies compiled for the CPU).
TEST INFRASTRUCTURE ONLY -- see the header of hostsim.cpp."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libhostsim.so")
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_lib = None
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)
KINDS = {"mv": 0, "sim": 1, "both": 2, "ll_mean": 3, "ll_draw": 4}
ODE_IDS = {"chkrebtii": 0, "fitz": 1, "lorenz63": 2, "mseir": 3}


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "hostsim.cpp")
        deps = [src] + [os.path.join(_ROOT, "rodeo_legacy_b200", "csrc", f)
                        for f in ("kalman_tps.cuh", "kalman_tps.inl", "kalman_math.inl", "kalman_common.cuh",
                                  "ode_functors.cuh")]
        if not os.path.exists(_LIB) or any(os.path.getmtime(d) > os.path.getmtime(_LIB) for d in deps):
            os.makedirs(os.path.dirname(_LIB), exist_ok=True)
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w",
                                   "-ffp-contract=off", "-o", _LIB, src])
        _lib = ctypes.CDLL(_LIB)
    return _lib


def _p(a, t=_dp):
    return None if a is None else a.ctypes.data_as(t)


def solve(ode, W, tmin, tmax, N, Q, c, R, x0, theta=None, z=None, kind="mv", ig=0, unroll=True, blockdiag=False, ck=1, shared_cov=False,
          thin_index=None, state_index=None, Y=None, gamma=1.0):
    W, Q, R = (np.asfortranarray(np.asarray(a, dtype=np.float64)) for a in (W, Q, R))
    c = np.ascontiguousarray(c, dtype=np.float64)
    m, n = W.shape
    x0 = np.ascontiguousarray(x0, dtype=np.float64).reshape(-1, n)
    B = x0.shape[0]
    theta = None if theta is None else np.ascontiguousarray(theta, dtype=np.float64).reshape(B, -1)
    z_stride = 0
    if z is not None:
        z = np.asarray(z, dtype=np.float64)
        if z.ndim == 3:
            z = np.ascontiguousarray(np.transpose(z, (0, 2, 1)))
            z_stride = n * N
        else:
            z = np.asfortranarray(z)
    k = KINDS[kind]
    x = np.full((B, N + 1, n), np.nan) if k in (1, 2) else None
    mu = np.full((B, N + 1, n), np.nan) if k in (0, 2) else None
    var = np.full((B, N + 1, n, n), np.nan) if k in (0, 2) else None
    if shared_cov and var is not None:      # ONE variance array for the whole batch in this mode
        var = np.full((1, N + 1, n, n), np.nan)
    ll = np.full(B, np.nan) if k >= 3 else None
    status = np.zeros(B, dtype=np.int32)
    ti = None if thin_index is None else np.ascontiguousarray(thin_index, dtype=np.int32)
    si = None if state_index is None else np.ascontiguousarray(state_index, dtype=np.int32)
    Yc = None if Y is None else np.ascontiguousarray(Y, dtype=np.float64)
    rc = lib().hostsim_solve(
        ctypes.c_int(ODE_IDS[ode]), ctypes.c_int(n), ctypes.c_int(m), ctypes.c_int(N),
        ctypes.c_double(tmin), ctypes.c_double(tmax), _p(W), _p(Q), _p(c), _p(R), ctypes.c_int(ig),
        ctypes.c_int(k), ctypes.c_int(int(unroll)), ctypes.c_int(int(blockdiag)), ctypes.c_int(int(ck)), ctypes.c_int(int(shared_cov)), ctypes.c_long(B), _p(x0), _p(theta), _p(z),
        ctypes.c_long(z_stride), _p(x), _p(mu), _p(var), _p(status, _ip), _p(ti, _ip),
        ctypes.c_int(0 if ti is None else len(ti)), _p(si, _ip),
        ctypes.c_int(0 if si is None else len(si)), _p(Yc), ctypes.c_double(gamma), _p(ll))
    assert rc == 0, rc
    return {"x": x, "mu": mu, "var": var, "status": status, "loglik": ll}
