"""CPU oracle for the KalmanODE hot path -- TEST INFRASTRUCTURE ONLY.

ctypes binding of ``oracle/kalman_oracle.c`` (a plain-C restatement of the reference's
``rodeo/eigen/KalmanTV.h`` + ``rodeo/eigen/KalmanTVODE.h`` arithmetic driven in the call
order of ``rodeo/cython/KalmanODE.pyx``).  Only ``tests/``, ``__graft_entry__.smoke()`` and
the CPU-baseline legs of ``bench.py`` may import this package; the product package
``rodeo_legacy_b200`` never does.

Parity pin: checked against golden vectors generated from the reference's own pure-Python
path (``tests/golden/make_golden.py``), see ``tests/test_oracle_golden.py``.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

ODE_IDS = {"chkrebtii": 0, "fitz": 1, "lorenz63": 2, "mseir": 3}
MODE_MV, MODE_SIM, MODE_BOTH = 1, 2, 3
INTERROGATE = {"probde": 0, "schobert": 1, "chkrebtii": 2}

_lib = None
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


def build(force=False):
    """Compile liboracle.so with gcc (idempotent)."""
    src = os.path.join(_HERE, "kalman_oracle.c")
    fast = os.path.join(_HERE, "_build", "liboracle_fast_v4.so")
    if (not force and os.path.exists(_LIB_PATH) and os.path.exists(fast)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)
            and os.path.getmtime(fast) >= os.path.getmtime(src)):
        return _LIB_PATH
    subprocess.check_call(["make", "-s", "-C", _HERE, "-B"])
    return _LIB_PATH


def select_timing_build():
    """Switch this process to the fast-math timing build of the oracle (the reference's own compiler
    flags, setup.py:37-42) at the widest ISA level the host supports.  For bench.py's CPU-baseline legs
    ONLY: results may differ from the parity build in the last bits.  Returns a description."""
    global _lib, _LIB_PATH
    build()
    flags = ""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    flags = line
                    break
    except OSError:
        pass
    v4 = all((" " + k) in flags for k in ("avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl"))
    name = "liboracle_fast_v4.so" if v4 else "liboracle_fast_v3.so"
    path = os.path.join(_HERE, "_build", name)
    if not os.path.exists(path):
        return "gcc -O3 -fno-fast-math (parity build; timing build missing)"
    _LIB_PATH, _lib = path, None
    return "gcc -O3 -ffast-math -march=x86-64-v%d -fopenmp (the reference's setup.py flags, ISA level pinned)" % (4 if v4 else 3)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.oracle_solve.restype = ctypes.c_int
        _lib.oracle_solve_batch.restype = ctypes.c_int
        _lib.oracle_max_threads.restype = ctypes.c_int
        _lib.oracle_loglik.restype = ctypes.c_double
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _f(a):
    """column-major float64 copy (what the C side expects for matrices)."""
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


class Problem:
    """The shared part of a batch of solves: prior (Q, c, R), W, grid and RHS name."""

    def __init__(self, ode, W, tmin, tmax, n_eval, wgt_state, mu_state, var_state,
                 interrogate="probde", block_sizes=None):
        """block_sizes: the n_deriv_prior list when the blocks are UNEQUAL (the right-hand side then reads
        variable k where its block starts); None = the reference functors' equal split."""
        self.offsets = None if block_sizes is None else np.concatenate(
            [[0], np.cumsum(block_sizes)[:-1]]).astype(np.int32)
        self.ode = ODE_IDS[ode]
        self.interrogate = INTERROGATE[interrogate]
        self.W = _f(W)
        self.m, self.n = self.W.shape
        self.tmin, self.tmax, self.N = float(tmin), float(tmax), int(n_eval)
        self.Q, self.R = _f(wgt_state), _f(var_state)
        self.c = np.ascontiguousarray(mu_state, dtype=np.float64)
        assert self.Q.shape == (self.n, self.n) and self.R.shape == (self.n, self.n)


def _set_offsets(prob):
    off = getattr(prob, "offsets", None)
    if off is None:
        lib().oracle_set_var_offsets(None, ctypes.c_int(0))
    else:
        lib().oracle_set_var_offsets(off.ctypes.data_as(_ip), ctypes.c_int(len(off)))


def solve(prob, x0, theta=None, z=None, mode=MODE_MV, return_hist=False):
    """One solve.  z is (n, N) as the reference's z_state.  Returns dict."""
    _set_offsets(prob)
    n, N = prob.n, prob.N
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    theta = None if theta is None else np.ascontiguousarray(theta, dtype=np.float64)
    zf = None if z is None else _f(z)
    if (mode & MODE_SIM or prob.interrogate == 2) and zf is None:
        raise ValueError("z required")
    x = np.zeros((N + 1, n)) if mode & MODE_SIM else None
    mu = np.zeros((N + 1, n)) if mode & MODE_MV else None
    var = np.zeros((N + 1, n, n)) if mode & MODE_MV else None
    hist = np.zeros((N + 1) * (2 * n + 2 * n * n)) if return_hist else None
    st = lib().oracle_solve(
        ctypes.c_int(prob.ode), ctypes.c_int(prob.interrogate), ctypes.c_int(n),
        ctypes.c_int(prob.m), ctypes.c_int(N), ctypes.c_double(prob.tmin),
        ctypes.c_double(prob.tmax), _p(prob.W), _p(prob.Q), _p(prob.c), _p(prob.R),
        _p(x0), _p(theta), _p(zf), ctypes.c_int(mode), _p(x), _p(mu), _p(var), _p(hist))
    out = {"status": st, "x": x, "mu": mu, "var": var}
    if return_hist:
        T = N + 1
        o = 0
        out["mu_pred"] = hist[o:o + T * n].reshape(T, n); o += T * n
        # stored column-major per step: [t][j][i] = Sigma(i, j)
        out["var_pred"] = hist[o:o + T * n * n].reshape(T, n, n); o += T * n * n
        out["mu_filt"] = hist[o:o + T * n].reshape(T, n); o += T * n
        out["var_filt"] = hist[o:o + T * n * n].reshape(T, n, n)
    return out


def solve_batch(prob, x0, theta=None, z=None, mode=MODE_MV, nthreads=0):
    """Batch of solves, OpenMP over systems.  x0 (B,n); theta (B,k); z (n,N) or (B,n,N)."""
    _set_offsets(prob)
    n, N = prob.n, prob.N
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    B = x0.shape[0]
    n_theta = 0
    if theta is not None:
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        n_theta = theta.shape[1]
    z_stride = 0
    if z is not None:
        z = np.asarray(z, dtype=np.float64)
        if z.ndim == 3:
            z = np.ascontiguousarray(np.transpose(z, (0, 2, 1)))  # (B, N, n): column-major per system
            z_stride = n * N
        else:
            z = _f(z)
    x = np.zeros((B, N + 1, n)) if mode & MODE_SIM else None
    mu = np.zeros((B, N + 1, n)) if mode & MODE_MV else None
    var = np.zeros((B, N + 1, n, n)) if mode & MODE_MV else None
    status = np.zeros(B, dtype=np.int32)
    lib().oracle_solve_batch(
        ctypes.c_int(prob.ode), ctypes.c_int(prob.interrogate), ctypes.c_int(n),
        ctypes.c_int(prob.m), ctypes.c_int(N), ctypes.c_double(prob.tmin),
        ctypes.c_double(prob.tmax), _p(prob.W), _p(prob.Q), _p(prob.c), _p(prob.R),
        ctypes.c_long(B), _p(x0), _p(theta), ctypes.c_int(n_theta), _p(z),
        ctypes.c_long(z_stride), ctypes.c_int(mode), _p(x), _p(mu), _p(var),
        status.ctypes.data_as(_ip), ctypes.c_int(nthreads))
    return {"status": status, "x": x, "mu": mu, "var": var}


def filter_step(prob, t, mu_filt, var_filt, theta=None):
    """One filter step t -> t+1 (probde) from given filtered moments; var (n, n) symmetric.
    Returns (mu_filt, var_filt) at time index t+1."""
    _set_offsets(prob)
    n = prob.n
    mu_in = np.ascontiguousarray(mu_filt, dtype=np.float64)
    var_in = _f(var_filt)
    th = None if theta is None else np.ascontiguousarray(theta, dtype=np.float64)
    mu_out, var_out = np.zeros(n), np.zeros((n, n), order="F")
    lib().oracle_filter_step.restype = ctypes.c_int
    lib().oracle_filter_step(
        ctypes.c_int(prob.ode), ctypes.c_int(n), ctypes.c_int(prob.m), ctypes.c_int(prob.N),
        ctypes.c_double(prob.tmin), ctypes.c_double(prob.tmax), _p(prob.W), _p(prob.Q), _p(prob.c),
        _p(prob.R), _p(th), ctypes.c_int(int(t)), _p(mu_in), _p(var_in), _p(mu_out), _p(var_out), None, None)
    return mu_out, np.ascontiguousarray(var_out)


def max_threads():
    return lib().oracle_max_threads()


def ode_eval(ode, X, t, theta):
    X = np.ascontiguousarray(X, dtype=np.float64)
    th = None if theta is None else np.ascontiguousarray(theta, dtype=np.float64)
    lib().oracle_set_var_offsets(None, ctypes.c_int(0))
    m = lib().oracle_ode_n_meas(ctypes.c_int(ODE_IDS[ode]))
    out = np.zeros(m)
    lib().oracle_ode_eval(ctypes.c_int(ODE_IDS[ode]), _p(X), ctypes.c_int(len(X)),
                          ctypes.c_double(t), _p(th), _p(out))
    return out


def thinning_index(data_t, ode_t):
    data_t = np.ascontiguousarray(data_t, dtype=np.float64)
    ode_t = np.ascontiguousarray(ode_t, dtype=np.float64)
    ind = np.zeros(len(data_t), dtype=np.int32)
    lib().oracle_thinning_index(_p(data_t), ctypes.c_int(len(data_t)), _p(ode_t),
                                ctypes.c_int(len(ode_t)), ind.ctypes.data_as(_ip))
    return ind


def loglik(X, ind, state_ind, Y, gamma):
    X = np.ascontiguousarray(X, dtype=np.float64)
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    ind = np.ascontiguousarray(ind, dtype=np.int32)
    state_ind = np.ascontiguousarray(state_ind, dtype=np.int32)
    return lib().oracle_loglik(_p(X), ctypes.c_int(X.shape[1]), ind.ctypes.data_as(_ip),
                               ctypes.c_int(len(ind)), state_ind.ctypes.data_as(_ip),
                               ctypes.c_int(len(state_ind)), _p(Y), ctypes.c_double(gamma))
