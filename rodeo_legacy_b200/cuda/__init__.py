"""``rodeo.cuda`` -- the B200 backend module, beside the reference's rodeo.cython / rodeo.eigen /
rodeo.numba: ``from rodeo_legacy_b200.cuda import KalmanODE``."""
from .KalmanODE import KalmanODE, ODE_FUNS, register_names, compiled_shapes, supports
from ._lib import RodeoCudaError

__all__ = ["KalmanODE", "ODE_FUNS", "register_names", "compiled_shapes", "supports", "RodeoCudaError"]
