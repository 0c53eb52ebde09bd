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


def test_unequal_block_sizes_are_enumerated_without_gpu():
    """n_deriv_prior lists with unequal entries (the reference accepts any list, rodeo/utils/utils.py:97-105):
    served by dense sets compiled per list (csrc/kernels_mixed.cu), found from where wgt_meas puts the variables."""
    import numpy as np
    from rodeo_legacy_b200 import ibm_init, indep_init, zero_pad
    from rodeo_legacy_b200.cuda import _lib, compiled_shapes, supports
    lib = _lib.load()
    shapes = compiled_shapes()
    assert len(shapes) == lib.rodeo_cuda_kernel_set_count()
    mixed = {(s["ode"], tuple(s["block_sizes"])) for s in shapes if len(set(s["block_sizes"])) > 1}
    assert mixed == {("fitz", (2, 3)), ("fitz", (3, 4)), ("fitz", (4, 3)), ("lorenz63", (2, 3, 3))}
    assert all(s["family"] == "dense" for s in shapes if len(set(s["block_sizes"])) > 1)
    assert {"ode": "fitz", "n_state": 6, "block_sizes": [3, 3], "family": "structured", "name": "fitz/bds3"} in shapes \
        or any(s["ode"] == "fitz" and s["block_sizes"] == [3, 3] and s["family"] == "structured" for s in shapes)
    assert supports("fitz", [3, 4]) and supports("fitz", [4, 3]) and supports("lorenz63", [2, 3, 3])
    assert supports("fitz", [3, 3]) and supports("mseir", [3] * 5)
    assert not supports("fitz", [2, 5]) and not supports("lorenz63", [3, 4]) and not supports("fitz", [3, 0])

    def create(n_prior, ode=1):
        n_deriv = [1] * len(n_prior)
        n = sum(n_prior)
        Wm = np.zeros((len(n_prior), sum(n_deriv) + len(n_deriv)))
        for i in range(len(n_deriv)):
            Wm[i, 2 * i + 1] = 1
        W = np.asfortranarray(zero_pad(Wm, n_deriv, n_prior))
        k = indep_init(ibm_init(0.1, n_prior, [0.1] * len(n_prior)), n_prior)
        Q, R, c = np.asfortranarray(k["wgt_state"]), np.asfortranarray(k["var_state"]), np.zeros(n)
        d = _lib.ProblemDesc(n, len(n_prior), 10, ode, 0, 0, 0.0, 1.0, W.ctypes.data, Q.ctypes.data, c.ctypes.data,
                             R.ctypes.data)
        h = ctypes.c_void_p()
        return lib.rodeo_cuda_problem_create(ctypes.byref(d), ctypes.byref(h)), lib.rodeo_cuda_last_error()

    rc, err = create([2, 4])      # 6 = 2 * 3, but the variables are NOT at the equal split: no silent wrong answer
    assert rc == -2 and b"block sizes [2, 4]" in err
    rc, err = create([5, 2])
    assert rc == -2 and b"block sizes [5, 2]" in err


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
