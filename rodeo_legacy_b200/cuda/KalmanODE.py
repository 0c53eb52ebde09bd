"""Batched KalmanODE front end over the CUDA library.

Mirrors the reference's backend class (``rodeo/cython/KalmanODE.pyx``): constructor
``KalmanODE(W, tmin, tmax, n_eval, ode_fun, mu_state, wgt_state, var_state, z_state=None)``
(:170-172), properties ``wgt_meas / mu_state / var_state / wgt_state / z_state`` with
shape- and contiguity-checked setters and the ``z_state`` deleter that re-zeros (:220-278),
methods ``solve / solve_sim / solve_mv (x0, W=None, theta=None)`` (:335, :396, :429).

Differences, all forced by batching on a GPU:
  * ``ode_fun`` is the NAME of a registered device functor ("chkrebtii", "fitz", "lorenz63",
    "mseir") -- there is no Python callback per step.  A callable carrying a
    ``rodeo_cuda_name`` attribute is accepted too.
  * ``x0`` may be (n,) or (B, n); ``theta`` (k,) or (B, k); ``z_state`` (n, N) shared by the
    batch or (B, n, N).  A 1-D ``x0`` returns the reference's shapes (N+1, n) / (N+1, n, n);
    a 2-D one returns batch-major arrays (B, N+1, n) / (B, N+1, n, n).
  * NumPy in -> NumPy out through the host-buffer entry point; torch CUDA tensors in ->
    torch CUDA tensors out through the device-pointer entry point (nothing leaves HBM).
  * ``var[..., 0, :, :]`` is zeros (the compiled reference backends leave it uninitialised).
  * ``force_dense=True`` keeps the general dense kernels even when the prior is block diagonal
    (the block-diagonal family is an exact specialisation: zeros stay zeros).
  * ``solve_mv`` never touches the RNG; ``solve_sim`` / ``solve`` fill an all-zero
    ``z_state`` from ``rand_mat`` (global NumPy RNG) exactly like the reference (:285-286).
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ..utils import rand_mat

ODE_FUNS = {"chkrebtii": 0, "fitz": 1, "lorenz63": 2, "mseir": 3}
_ALIASES = {"lorenz": "lorenz63", "fitzhugh_nagumo": "fitz", "fitzhugh-nagumo": "fitz", "fn": "fitz"}
_INTERROGATE = {"probde": 0, "schobert": 1, "chkrebtii": 2}
MODE_MV, MODE_SIM, MODE_BOTH = 1, 2, 3
_FIELDS = {"wgt_meas": 0, "mu_state": 1, "wgt_state": 2, "var_state": 3}
_VAR_LAYOUTS = {"full": 0, "blocks": 1, "diag": 2}
_VAR_LAYOUTS_INV = {v: k for k, v in _VAR_LAYOUTS.items()}


def register_names():
    """Names of the compiled right-hand sides."""
    return sorted(ODE_FUNS)


def compiled_shapes():
    """Every kernel set compiled into the library: a list of dicts ``ode`` (name), ``n_state``, ``block_sizes``
    (the ``n_deriv_prior`` it serves: one entry per ODE variable), ``family`` ("dense" takes any prior of that
    shape, "blockdiag" / "structured" are chosen automatically for block-diagonal priors) and ``name``.  Works
    without a GPU (rodeo_cuda_kernel_set_info)."""
    lib = _lib.load()
    names = {v: k for k, v in ODE_FUNS.items()}
    out = []
    for i in range(lib.rodeo_cuda_kernel_set_count()):
        ode, n, bd, st = (ctypes.c_int32() for _ in range(4))
        sizes = (ctypes.c_int32 * 8)()
        name = ctypes.c_char_p()
        _lib.check(lib.rodeo_cuda_kernel_set_info(i, ctypes.byref(ode), ctypes.byref(n), ctypes.byref(bd),
                                                  ctypes.byref(st), sizes, ctypes.byref(name)))
        m = lib.rodeo_cuda_ode_n_meas(ode.value)
        out.append(dict(ode=names.get(ode.value, ode.value), n_state=n.value, block_sizes=list(sizes[:m]),
                        family="structured" if st.value else "blockdiag" if bd.value else "dense",
                        name=name.value.decode()))
    return out


def supports(ode_fun, n_deriv_prior, interrogate="probde"):
    """True when a kernel set is compiled for this right-hand side and list of block sizes (equal or not)."""
    sizes = (ctypes.c_int32 * len(n_deriv_prior))(*[int(v) for v in n_deriv_prior])
    return _lib.load().rodeo_cuda_supported_blocks(_ode_id(ode_fun), len(n_deriv_prior), sizes,
                                                   _INTERROGATE[interrogate]) != 0


def _ode_id(ode_fun):
    name = getattr(ode_fun, "rodeo_cuda_name", ode_fun)
    if isinstance(name, str):
        name = _ALIASES.get(name.lower(), name.lower())
        if name in ODE_FUNS:
            return ODE_FUNS[name]
        raise ValueError(f"unknown ode_fun {name!r}; registered: {register_names()}")
    if isinstance(name, (int, np.integer)) and 0 <= int(name) < _lib.load().rodeo_cuda_ode_count():
        return int(name)
    raise TypeError("rodeo.cuda runs ode_fun on the GPU: pass the name of a registered device "
                    f"functor ({register_names()}), not a Python callable")


def _copynm(dest, source, name):
    """The reference's checked copy (KalmanODE.pyx:104-110)."""
    source = np.asarray(source)
    if not source.flags.f_contiguous:
        raise TypeError('{} is not f contiguous.'.format(name))
    if not source.shape == dest.shape:
        raise TypeError('{} has incorrect shape.'.format(name))
    dest[...] = source
    return source


def _readonly(a):
    """Property getters hand out READ-ONLY views.  The reference's getters expose its internal buffers and
    ``kode.z_state[:] = z`` works there; here the C handle (and the cached device copy of z) hold their own
    copies, made by the setters, so an in-place write would be silently ignored -- it raises instead:
    assign through the property (``kode.z_state = z``)."""
    v = a.view()
    v.flags.writeable = False
    return v


def _check_out(a, shape, name):
    """Caller-provided host output buffer: float64, C-contiguous, exactly the expected shape (the C side
    copies raw bytes into it)."""
    if a is None:
        raise TypeError(f'{name} is required.')
    if not isinstance(a, np.ndarray) or a.dtype != np.float64:
        raise TypeError(f'{name} must be a float64 NumPy array.')
    if tuple(a.shape) != tuple(shape):
        raise TypeError('{} has incorrect shape.'.format(name))
    if not a.flags.c_contiguous or not a.flags.writeable:
        raise TypeError(f'{name} must be C-contiguous and writeable.')


def _ptr(t):
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        return ctypes.c_void_p(t.data_ptr())
    return ctypes.c_void_p(t.ctypes.data)


class KalmanODE:
    def __init__(self, W, tmin, tmax, n_eval, ode_fun, mu_state, wgt_state, var_state,
                 z_state=None, interrogate="probde", device=None, force_dense=False, checkpoint=0,
                 shared_cov=False, use_structure=True, var_layout="full"):
        self._lib = _lib.load()                      # raises if the CUDA library is missing
        if not torch.cuda.is_available():
            raise RuntimeError("rodeo.cuda needs a CUDA device; there is no CPU fallback")
        W = np.asarray(W, dtype=np.float64)
        self.n_meas, self.n_state = W.shape
        self.tmin, self.tmax = float(tmin), float(tmax)
        self.n_eval = int(n_eval)
        self.n_steps = self.n_eval + 1
        self.ode_fun = ode_fun
        self._ode = _ode_id(ode_fun)
        self._ig = _INTERROGATE[interrogate]
        self.n_theta = self._lib.rodeo_cuda_ode_n_theta(self._ode)
        if self._lib.rodeo_cuda_ode_n_meas(self._ode) != self.n_meas:
            raise TypeError('{} has incorrect shape.'.format('W'))
        n, m = self.n_state, self.n_meas
        self._wgt_meas = np.empty((m, n), order='F')
        self._mu_state = np.empty(n, order='F')
        self._wgt_state = np.empty((n, n), order='F')
        self._var_state = np.empty((n, n), order='F')
        self._z_state = np.zeros((n, self.n_eval), order='F')
        self._wgt_meas[:] = W
        self._mu_state[:] = mu_state
        self._wgt_state[:] = wgt_state
        self._var_state[:] = var_state
        if z_state is not None:
            self.z_state = z_state
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None \
            else torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("rodeo.cuda needs a CUDA device; there is no CPU fallback")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        desc = _lib.ProblemDesc(n, m, self.n_eval, self._ode, self._ig,
                                int(bool(force_dense)) | (0 if use_structure else 2),
                                self.tmin, self.tmax,
                                self._wgt_meas.ctypes.data, self._wgt_state.ctypes.data,
                                self._mu_state.ctypes.data, self._var_state.ctypes.data)
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.rodeo_cuda_problem_create(ctypes.byref(desc), ctypes.byref(h)))
        self._h = h
        if checkpoint != 0:
            self.checkpoint = checkpoint
        if shared_cov:
            self.shared_cov = True
        if var_layout != "full":
            self.var_layout = var_layout
        self._z_dev = None          # cached device copy of the shared z
        self._ll_cache = {}

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self._lib.rodeo_cuda_problem_destroy(h)
            except Exception:
                pass

    # ---- properties (KalmanODE.pyx:220-278) ---------------------------------------------
    def _push(self, name):
        arr = getattr(self, "_" + name)
        _lib.check(self._lib.rodeo_cuda_problem_set(self._h, _FIELDS[name], arr.ctypes.data))

    @property
    def wgt_meas(self):
        return _readonly(self._wgt_meas)

    @wgt_meas.setter
    def wgt_meas(self, v):
        _copynm(self._wgt_meas, v, 'wgt_meas')
        self._push("wgt_meas")

    @property
    def mu_state(self):
        return _readonly(self._mu_state)

    @mu_state.setter
    def mu_state(self, v):
        _copynm(self._mu_state, v, 'mu_state')
        self._push("mu_state")

    @property
    def var_state(self):
        return _readonly(self._var_state)

    @var_state.setter
    def var_state(self, v):
        _copynm(self._var_state, v, 'var_state')
        self._push("var_state")

    @property
    def wgt_state(self):
        return _readonly(self._wgt_state)

    @wgt_state.setter
    def wgt_state(self, v):
        _copynm(self._wgt_state, v, 'wgt_state')
        self._push("wgt_state")

    @property
    def z_state(self):
        return _readonly(self._z_state)

    @z_state.setter
    def z_state(self, v):
        v = np.asarray(v, dtype=np.float64)
        if v.ndim == 3:                      # per-system draws (B, n, N): batched extension
            if v.shape[1:] != (self.n_state, self.n_eval):
                raise TypeError('{} has incorrect shape.'.format('z_state'))
            self._z_state = v
        else:
            if self._z_state.ndim != 2:
                self._z_state = np.zeros((self.n_state, self.n_eval), order='F')
            _copynm(self._z_state, v, 'z_state')
        self._z_dev = None

    @z_state.deleter
    def z_state(self):
        self._z_state = np.zeros((self.n_state, self.n_eval), order='F')
        self._z_dev = None

    # ---- workspace / introspection ----------------------------------------------------------
    def reserve(self, max_batch):
        with torch.cuda.device(self.device):
            _lib.check(self._lib.rodeo_cuda_reserve(self._h, int(max_batch)))

    def set_workspace_limit(self, nbytes):
        _lib.check(self._lib.rodeo_cuda_set_workspace_limit(self._h, int(nbytes)))

    @property
    def kernel_family(self):
        """('blockdiag' | 'dense', 'registers' | 'rolled'): the kernel family in use.  The
        block-diagonal family is selected automatically when wgt_state / var_state are block
        diagonal with n_meas equal blocks and row a of wgt_meas lies inside block a (what
        ibm_init / car_init + indep_init + zero_pad produce); ``force_dense=True`` disables it."""
        bd, rg, _ = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        _lib.check(self._lib.rodeo_cuda_problem_info(self._h, ctypes.byref(bd), ctypes.byref(rg), ctypes.byref(_)))
        return ("blockdiag" if bd.value else "dense", ("rolled", "registers", "wide")[rg.value])

    @property
    def var_layout(self):
        """Layout of the returned variance: ``"full"`` (default, the reference's (n, n) matrices), ``"blocks"``
        ((n_meas, p, p): the diagonal blocks only) or ``"diag"`` ((n,): marginal variances).  The compact layouts
        are opt-in for block-diagonal priors with at least two blocks; they hold the same numbers as the
        corresponding entries of the full output and cut the bytes written and copied to the host (FitzHugh-Nagumo solve_mv:
        336 -> 192 / 96 bytes per system and step).  Not applied to ``priors=`` calls or shared-covariance mode."""
        return _VAR_LAYOUTS_INV[int(self._lib.rodeo_cuda_get_var_layout(self._h))]

    @var_layout.setter
    def var_layout(self, name):
        if name not in _VAR_LAYOUTS:
            raise ValueError(f"var_layout must be one of {sorted(_VAR_LAYOUTS)}")
        _lib.check(self._lib.rodeo_cuda_set_var_layout(self._h, _VAR_LAYOUTS[name]))

    def _var_shape(self, B):
        n, m, T = self.n_state, self.n_meas, self.n_steps
        lay = self.var_layout
        p = n // m
        return (B, T, m, p, p) if lay == "blocks" else (B, T, n) if lay == "diag" else (B, T, n, n)

    @property
    def last_batched_prior_blockdiag(self):
        """True when the last ``solve*(..., priors=...)`` call ran on the block-diagonal thread-per-block kernels
        (every system's own prior block diagonal), False when it used the dense family."""
        return bool(self._lib.rodeo_cuda_last_batched_prior_blockdiag(self._h))

    @property
    def structured(self):
        """True when the block-diagonal kernels specialised for ibm_init / zero_pad priors are in use (every Q
        block upper triangular, every W row a unit vector): they skip the exactly-zero terms, same bits.
        ``use_structure=False`` keeps the general block-diagonal kernels."""
        return bool(self._lib.rodeo_cuda_problem_structured(self._h))

    @property
    def shared_cov(self):
        """Shared-covariance mode (off by default).  All systems solved by one KalmanODE object
        share its prior, so their covariances are identical (they do not depend on theta, x0 or
        the ODE).  With ``shared_cov = True`` they are computed once per prior and the batch
        kernels run the mean recursions only; ``var`` is then returned as a read-only broadcast
        view of ONE (N+1, n, n) array (same shape as in the general mode, no per-system copy)."""
        return bool(self._lib.rodeo_cuda_get_shared_cov(self._h))

    @shared_cov.setter
    def shared_cov(self, enabled):
        _lib.check(self._lib.rodeo_cuda_set_shared_cov(self._h, int(bool(enabled))))

    @property
    def checkpoint(self):
        """Checkpoint interval of the filter history in effect (set 0 = auto, 1, 2 or 4; reads
        back the interval actually used): with k > 1 only every
        k-th filtered record goes to HBM and the smoother re-runs the filter in between
        (bit-identical results, k-fold less history traffic)."""
        return int(self._lib.rodeo_cuda_get_checkpoint(self._h))

    @checkpoint.setter
    def checkpoint(self, interval):
        _lib.check(self._lib.rodeo_cuda_set_checkpoint(self._h, int(interval)))

    @property
    def history_bytes_per_system(self):
        return int(self._lib.rodeo_cuda_history_bytes_per_system(self._h))

    @property
    def launch_count(self):
        return int(self._lib.rodeo_cuda_launch_count(self._h))

    def set_timing(self, enabled=True):
        _lib.check(self._lib.rodeo_cuda_set_timing(self._h, int(bool(enabled))))

    def last_kernel_ms(self):
        f, b = ctypes.c_float(), ctypes.c_float()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.rodeo_cuda_last_kernel_ms(self._h, ctypes.byref(f), ctypes.byref(b)))
        return f.value, b.value

    def kernel_ms_total(self):
        """(forward ms, backward ms, solves) summed over every solve since ``set_timing(True)``."""
        f, b, k = ctypes.c_double(), ctypes.c_double(), ctypes.c_int32()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.rodeo_cuda_kernel_ms_total(self._h, ctypes.byref(f), ctypes.byref(b), ctypes.byref(k)))
        return f.value, b.value, k.value

    # ---- argument handling --------------------------------------------------------------------
    def _prep(self, x0, W, theta, need_z):
        if W is not None:
            self._wgt_meas[:] = np.asarray(W, dtype=np.float64)     # KalmanODE.pyx:441-442
            self._push("wgt_meas")
        on_dev = isinstance(x0, torch.Tensor)
        if on_dev:
            if x0.dtype != torch.float64 or not x0.is_cuda:
                raise TypeError("x0 must be a float64 CUDA tensor (or a NumPy array)")
            if x0.device != self.device:
                raise ValueError(f"x0 is on {x0.device} but this KalmanODE was created on {self.device}")
            single = x0.dim() == 1
            x0b = x0.reshape(-1, self.n_state).contiguous() if x0.numel() % self.n_state == 0 else None
        else:
            x0a = np.asarray(x0, dtype=np.float64)
            if x0a.ndim == 1 and not x0a.flags.f_contiguous:
                raise TypeError('{} is not f contiguous.'.format('x0'))
            single = x0a.ndim == 1
            x0b = np.ascontiguousarray(x0a.reshape(-1, self.n_state)) if x0a.size % self.n_state == 0 else None
        if x0b is None or x0b.shape[-1] != self.n_state or (single and x0b.shape[0] != 1):
            raise TypeError('{} has incorrect shape.'.format('x0'))
        B = x0b.shape[0]
        th = None
        if self.n_theta:
            if theta is None:
                raise TypeError("theta is required by this ode_fun")
            if on_dev:
                th = torch.as_tensor(theta, dtype=torch.float64, device=x0b.device)
                th = th.reshape(-1, self.n_theta)
                th = (th.expand(B, self.n_theta) if th.shape[0] == 1 else th).contiguous()
            else:
                th = np.asarray(theta, dtype=np.float64).reshape(-1, self.n_theta)
                th = np.array(np.broadcast_to(th, (B, self.n_theta))) if th.shape[0] == 1 else np.ascontiguousarray(th)
            if th.shape[0] != B:
                raise TypeError('{} has incorrect shape.'.format('theta'))
        z, z_stride = None, 0
        if need_z:
            if not np.any(self._z_state):                               # KalmanODE.pyx:285-286
                self._z_state = rand_mat(self.n_eval, self.n_state)
                self._z_dev = None
            zs = self._z_state
            if zs.ndim == 3:
                if zs.shape[0] != B:
                    raise TypeError('{} has incorrect shape.'.format('z_state'))
                # per system column-major (n x N)  ==  C-order (B, N, n)
                zc = np.ascontiguousarray(np.transpose(zs, (0, 2, 1)))
                z_stride = self.n_state * self.n_eval
            else:
                zc = np.asfortranarray(zs)
            if on_dev:
                if self._z_dev is None or self._z_dev.device != x0b.device:
                    # column-major bytes, whatever the torch shape says
                    flat = zc.ravel(order="K") if zs.ndim == 3 else zc.ravel(order="F")
                    self._z_dev = torch.from_numpy(np.ascontiguousarray(flat)).to(x0b.device)
                z = self._z_dev
            else:
                z = zc
        return on_dev, single, B, x0b, th, z, z_stride

    def _batched_prior(self, priors, B, dev):
        """Device arrays of per-system priors: each (B, ...) array becomes batch-major with every
        system column-major (what the C ABI expects); missing keys stay None (= shared value)."""
        shapes = {"wgt_meas": (self.n_meas, self.n_state), "mu_state": (self.n_state,),
                  "wgt_state": (self.n_state, self.n_state), "var_state": (self.n_state, self.n_state)}
        out = {}
        for key, shp in shapes.items():
            v = priors.get(key)
            if v is None:
                out[key] = None
                continue
            t = torch.as_tensor(v, dtype=torch.float64, device=dev)
            if tuple(t.shape) != (B,) + shp:
                raise TypeError('{} has incorrect shape.'.format(key))
            out[key] = (t.transpose(1, 2) if len(shp) == 2 else t).contiguous()
        unknown = set(priors) - set(shapes)
        if unknown:
            raise TypeError(f"unknown prior fields {sorted(unknown)}")
        return out

    def _run(self, mode, x0, W, theta, out=None, priors=None):
        need_z = bool(mode & MODE_SIM) or self._ig == 2
        on_dev, single, B, x0b, th, z, z_stride = self._prep(x0, W, theta, need_z)
        if priors is not None:
            return self._run_batched_prior(mode, priors, on_dev, single, B, x0b, th, z, z_stride)
        n, T = self.n_state, self.n_steps
        mv, sim = bool(mode & MODE_MV), bool(mode & MODE_SIM)
        shared = self.shared_cov
        vshape = (T, n, n) if shared else self._var_shape(B)     # shared: ONE array for the batch
        if on_dev:
            dev = x0b.device
            mk = lambda *s: torch.empty(s, dtype=torch.float64, device=dev)
            x = mk(B, T, n) if sim else None
            mu = mk(B, T, n) if mv else None
            var = mk(*vshape) if mv else None
            status = torch.zeros(B, dtype=torch.int32, device=dev)
            with torch.cuda.device(dev):
                stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
                _lib.check(self._lib.rodeo_cuda_solve(
                    self._h, mode, B, _ptr(x0b), _ptr(th), _ptr(z), z_stride, _ptr(x), _ptr(mu),
                    _ptr(var), _ptr(status), stream))
        else:
            if out is not None:
                x, mu, var = out
                if sim:
                    _check_out(x, (B, T, n), 'x_out')
                if mv:
                    _check_out(mu, (B, T, n), 'mu_out')
                    _check_out(var, vshape, 'var_out')     # shared-covariance mode: ONE (N+1, n, n) array
            else:
                x = np.empty((B, T, n)) if sim else None
                mu = np.empty((B, T, n)) if mv else None
                var = np.empty(vshape) if mv else None
            status = np.zeros(B, dtype=np.int32)
            with torch.cuda.device(self.device):
                _lib.check(self._lib.rodeo_cuda_solve_host(
                    self._h, mode, B, _ptr(x0b), _ptr(th), _ptr(z), z_stride, _ptr(x), _ptr(mu),
                    _ptr(var), _ptr(status)))
        self.status = status
        if shared and var is not None and out is None:     # broadcast view: same shape, no copies
            var = var.unsqueeze(0).expand(B, T, n, n) if on_dev else np.broadcast_to(var, (B, T, n, n))
        if single:
            x, mu, var = (None if a is None else a[0] for a in (x, mu, var))
        return x, mu, var

    def _run_batched_prior(self, mode, priors, on_dev, single, B, x0b, th, z, z_stride):
        dev = x0b.device if on_dev else self.device
        to_dev = lambda a: None if a is None else (a if isinstance(a, torch.Tensor) else torch.from_numpy(
            np.ascontiguousarray(a.ravel(order="K") if (a.ndim == 3 and z_stride) else
                                 (a.ravel(order="F") if a.ndim == 2 and a is z else a))).to(dev))
        x0d, thd, zd = to_dev(x0b), to_dev(th), to_dev(z)
        bp = self._batched_prior(priors, B, dev)
        n, T = self.n_state, self.n_steps
        mv, sim = bool(mode & MODE_MV), bool(mode & MODE_SIM)
        mk = lambda *s: torch.empty(s, dtype=torch.float64, device=dev)
        x = mk(B, T, n) if sim else None
        mu = mk(B, T, n) if mv else None
        var = mk(B, T, n, n) if mv else None
        status = torch.zeros(B, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(self._lib.rodeo_cuda_solve_batched_prior(
                self._h, mode, B, _ptr(x0d), _ptr(thd), _ptr(zd), z_stride, _ptr(bp["wgt_meas"]),
                _ptr(bp["mu_state"]), _ptr(bp["wgt_state"]), _ptr(bp["var_state"]), _ptr(x), _ptr(mu),
                _ptr(var), _ptr(status), stream))
        self.status = status if on_dev else status.cpu().numpy()
        if not on_dev:
            x, mu, var = (None if a is None else a.cpu().numpy() for a in (x, mu, var))
        if single:
            x, mu, var = (None if a is None else a[0] for a in (x, mu, var))
        return x, mu, var

    # ---- the reference's three methods -----------------------------------------------------
    # `priors` (batched extension): dict with any of wgt_meas (B, m, n), mu_state (B, n),
    # wgt_state (B, n, n), var_state (B, n, n) -- one prior PER SYSTEM (e.g. a sweep over sigma);
    # missing fields use the object's shared value.
    def solve(self, x0, W=None, theta=None, priors=None):
        """(kalman_sim, kalman_mu, kalman_var) -- KalmanODE.pyx:335-394."""
        return self._run(MODE_BOTH, x0, W, theta, priors=priors)

    def solve_sim(self, x0, W=None, theta=None, priors=None):
        """A draw from the solution posterior -- KalmanODE.pyx:396-427."""
        return self._run(MODE_SIM, x0, W, theta, priors=priors)[0]

    def solve_mv(self, x0, W=None, theta=None, priors=None):
        """(posterior mean, posterior variance) -- KalmanODE.pyx:429-463."""
        return self._run(MODE_MV, x0, W, theta, priors=priors)[1:]

    # Host path writing into caller-provided NumPy buffers (float64, C-contiguous, exact shape; checked).  A fresh
    # pageable np.empty result costs its first-touch page faults and a staged copy on every call (FN solve_sim,
    # 65,536 systems: 280 ms for 1.26 GB); buffers the caller allocates once -- ideally pinned, e.g.
    # torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy() -- take the copy at PCIe speed (25 ms).
    def solve_mv_into(self, x0, theta, mu_out, var_out):
        self._run(MODE_MV, x0, None, theta, out=(None, mu_out, var_out))

    def solve_sim_into(self, x0, theta, x_out):
        self._run(MODE_SIM, x0, None, theta, out=(x_out, None, None))

    def solve_into(self, x0, theta, x_out, mu_out, var_out):
        self._run(MODE_BOTH, x0, None, theta, out=(x_out, mu_out, var_out))

    # ---- fused thinning + Gaussian log-likelihood (examples/inference.py:54-56, 66-94) ------
    def loglik(self, x0, theta, Y, thin_index, state_ind, gamma, use_draw=False, priors=None):
        """One log-likelihood per system, no per-step output materialised.

        sum_d sum_k log N(Y[d, k]; X[thin_index[d], state_ind[k]], sd=gamma) with X the
        posterior mean (use_draw=False) or a posterior draw (use_draw=True, as
        ``kalman_nlpost`` does with solve_sim).  Inputs NumPy or torch CUDA; returns the same kind.
        """
        need_z = bool(use_draw) or self._ig == 2
        on_dev, single, B, x0b, th, z, z_stride = self._prep(x0, None, theta, need_z)
        dev = x0b.device if on_dev else self.device
        thin = np.ascontiguousarray(thin_index, dtype=np.int32)
        si = np.ascontiguousarray(state_ind, dtype=np.int32)
        Yc = np.ascontiguousarray(Y.detach().cpu().numpy() if isinstance(Y, torch.Tensor) else Y,
                                  dtype=np.float64)
        if Yc.shape != (len(thin), len(si)):
            raise TypeError('{} has incorrect shape.'.format('Y'))
        if thin.size and (thin.min() < 0 or thin.max() > self.n_eval or np.any(np.diff(thin) < 0)):
            raise ValueError("thin_index must be ascending grid indices in [0, n_eval]")
        if si.size and (si.min() < 0 or si.max() >= self.n_state):
            raise ValueError("state_ind out of range")
        key = (thin.tobytes(), si.tobytes(), Yc.tobytes(), str(dev))
        if key not in self._ll_cache:
            self._ll_cache.clear()
            self._ll_cache[key] = tuple(torch.from_numpy(a).to(dev) for a in (thin, si, Yc))
        d_thin, d_si, d_Y = self._ll_cache[key]
        if not on_dev:
            x0b = torch.from_numpy(x0b).to(dev)
            th = None if th is None else torch.from_numpy(th).to(dev)
            z = None if z is None else torch.from_numpy(
                np.ascontiguousarray(z.ravel(order="K" if z_stride else "F"))).to(dev)
        out = torch.empty(B, dtype=torch.float64, device=dev)
        status = torch.zeros(B, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            if priors is None:
                _lib.check(self._lib.rodeo_cuda_loglik(
                    self._h, int(bool(use_draw)), B, _ptr(x0b), _ptr(th), _ptr(z), z_stride,
                    _ptr(d_thin), len(thin), _ptr(d_si), len(si), _ptr(d_Y), float(gamma),
                    _ptr(out), _ptr(status), stream))
            else:       # one prior per system (e.g. a likelihood sweep over the IBM scale sigma)
                bp = self._batched_prior(priors, B, dev)
                _lib.check(self._lib.rodeo_cuda_loglik_batched_prior(
                    self._h, int(bool(use_draw)), B, _ptr(x0b), _ptr(th), _ptr(z), z_stride,
                    _ptr(bp["wgt_meas"]), _ptr(bp["mu_state"]), _ptr(bp["wgt_state"]), _ptr(bp["var_state"]),
                    _ptr(d_thin), len(thin), _ptr(d_si), len(si), _ptr(d_Y), float(gamma),
                    _ptr(out), _ptr(status), stream))
        self.status = status if on_dev else status.cpu().numpy()
        res = out if on_dev else out.cpu().numpy()
        return res[0] if single else res
