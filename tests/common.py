"""Shared test utilities: the parity metric, golden loaders, synthetic workloads."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as f:
        return {k: f[k] for k in f.files}


def load_mixed(name):
    """One case of tests/golden/mixed_blocks.npz (unequal block sizes; keys are `case__field`)."""
    with np.load(os.path.join(GOLDEN, "mixed_blocks.npz")) as f:
        return {k.split("__")[1]: f[k] for k in f.files if k.startswith(name + "__")}


def traj_err(a, b, axis=-2):
    """Trajectory-scaled max error per state component:
    max_t |a - b| / max_t |b|, time on `axis` (default: second-to-last, i.e. arrays
    shaped (..., N+1, n)).  Components whose reference trajectory is identically zero are
    compared absolutely."""
    a, b = np.asarray(a), np.asarray(b)
    num = np.max(np.abs(a - b), axis=axis)
    den = np.max(np.abs(b), axis=axis)
    return np.where(den > 0, num / np.where(den > 0, den, 1.0), num)


def var_err(a, b):
    """Same metric for covariance histories (..., T, n, n): time is axis -3."""
    return traj_err(a, b, axis=-3)


def ode_name(golden_name):
    for k in ("chkrebtii", "fitz", "lorenz", "mseir"):
        if golden_name.startswith(k):
            return "lorenz63" if k == "lorenz" else k
    raise KeyError(golden_name)
