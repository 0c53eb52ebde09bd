"""CPU-side checks: the C-ABI library loads and exports every symbol include/rodeo_cuda.h
declares; host-side logic that does not need a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from tests.common import ROOT


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "rodeo_cuda.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(rodeo_cuda_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from rodeo_legacy_b200.cuda import _lib
    lib = _lib.load()
    names = _header_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), name
    assert sorted(s[0] for s in _lib.SYMBOLS) == names
    assert lib.rodeo_cuda_abi_version() == 1


def test_registry_and_error_strings_without_gpu():
    from rodeo_legacy_b200.cuda import _lib
    lib = _lib.load()
    assert [lib.rodeo_cuda_ode_n_meas(i) for i in range(4)] == [1, 2, 3, 5]
    assert [lib.rodeo_cuda_ode_n_theta(i) for i in range(4)] == [0, 3, 3, 6]
    assert lib.rodeo_cuda_ode_n_meas(9) == -1
    for ode, n, m in [(0, 4, 1), (1, 6, 2), (2, 9, 3), (3, 15, 5)]:
        assert lib.rodeo_cuda_supported(ode, n, m, 0) == 3      # dense and block-diagonal sets
    assert lib.rodeo_cuda_supported(2, 12, 3, 0) == 2           # block-diagonal only (p = 4)
    assert lib.rodeo_cuda_supported(1, 7, 2, 0) == 0
    assert lib.rodeo_cuda_supported(1, 6, 3, 0) == 0
    # invalid descriptor is rejected before any CUDA call
    h = ctypes.c_void_p()
    assert lib.rodeo_cuda_problem_create(None, ctypes.byref(h)) == -1
    assert b"null" in lib.rodeo_cuda_last_error()
    import numpy as np
    W = np.zeros((2, 7), order="F")
    Q = np.asfortranarray(np.eye(7))
    c = np.zeros(7)
    d = _lib.ProblemDesc(7, 2, 10, 1, 0, 0, 0.0, 1.0, W.ctypes.data, Q.ctypes.data, c.ctypes.data,
                         Q.ctypes.data)
    assert lib.rodeo_cuda_problem_create(ctypes.byref(d), ctypes.byref(h)) == -2
    assert b"no compiled kernel" in lib.rodeo_cuda_last_error()


def test_front_end_fails_loudly_without_cuda():
    import torch
    from rodeo_legacy_b200.cuda import KalmanODE
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        KalmanODE(np.zeros((1, 4), order="F"), 0, 1, 10, "chkrebtii", np.zeros(4),
                  np.eye(4, order="F"), np.eye(4, order="F"))


def test_ode_name_resolution():
    from rodeo_legacy_b200.cuda.KalmanODE import _ode_id
    assert _ode_id("fitz") == 1 and _ode_id("Lorenz") == 2 and _ode_id("lorenz63") == 2

    def f(x, t, theta, out):
        pass
    with pytest.raises(TypeError, match="registered device functor"):
        _ode_id(f)
    f.rodeo_cuda_name = "mseir"
    assert _ode_id(f) == 3
    with pytest.raises(ValueError):
        _ode_id("vanderpol")


def test_shard_ranges_cover_batch():
    from rodeo_legacy_b200.dist import shard_range, shard_sizes
    for B in (0, 1, 7, 8, 65536, 1048576, 1048577):
        for w in (1, 2, 4, 8):
            sizes = shard_sizes(B, w)
            assert sum(sizes) == B and max(sizes) - min(sizes) <= 1
            edges = [shard_range(B, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == B
            assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))
