"""Test-side view of the BASELINE workloads: everything in rodeo_legacy_b200.workloads plus the oracle
problem builder (the product package never imports oracle/)."""
from rodeo_legacy_b200.workloads import *          # noqa: F401,F403
from rodeo_legacy_b200.workloads import _mseir_rhs, _wmat, _pack   # noqa: F401


def make_oracle_problem(w, **kw):
    import oracle
    return oracle.Problem(w["ode"], w["W"], w["tmin"], w["tmax"], w["N"], w["Q"], w["c"], w["R"], **kw)
