"""Generate the golden fixtures under tests/golden/ from the REFERENCE's own code.

Run in the build container only (needs /root/reference, which does not exist on the GPU
box):   python tests/golden/make_golden.py

What is imported from the reference (read-only tree, nothing copied):
  * its pure-Python implementation of the KalmanODE path
    tests/KalmanODE_py.py -> tests/depreciated/kalman/kalman_ode_higher.py
    -> tests/depreciated/kalman/KalmanTV.py  (+ rodeo.utils.solveV / norm_sim)
  * its prior helpers rodeo.ibm.ibm_init, rodeo.car.car_init, rodeo.utils.{zero_pad,
    indep_init, rand_mat}
  * examples/inference.py `thinning` + `loglike` (with stubs for the plotting imports)
Two shims are applied from outside the tree: the stale package name
`probDE` is aliased to `rodeo`, and `np.math` is pointed at `math` (NumPy >= 2).

The compiled reference backends (rodeo.cython / rodeo.eigen / rodeo.numba) cannot be
built here: `kalmantv` and Eigen are absent.  The pure-Python path implements the same
algorithm with SciPy Cholesky (cho_factor / cho_solve / cholesky).
"""
import math
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _import_reference():
    np.math = math                                     # shim 2
    sys.path.insert(0, os.path.join(REF, "tests"))
    sys.path.insert(0, REF)
    import rodeo                                       # the reference package
    import rodeo.utils
    import rodeo.utils.utils
    assert rodeo.__file__.startswith(REF), rodeo.__file__
    sys.modules["probDE"] = rodeo                      # shim 1
    sys.modules["probDE.utils"] = rodeo.utils
    sys.modules["probDE.utils.utils"] = rodeo.utils.utils
    from KalmanODE_py import KalmanODE_py
    from rodeo.ibm import ibm_init
    from rodeo.car import car_init
    from rodeo.utils import zero_pad, indep_init, rand_mat
    # examples/inference.py imports plotting / numdifftools at module scope: stub them
    for name in ["numdifftools", "seaborn", "matplotlib", "matplotlib.pyplot",
                 "matplotlib.patches", "matplotlib.lines"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, os.path.join(REF, "examples"))
    import scipy.stats  # noqa: F401  (inference.loglike uses sp.stats)
    from inference import inference
    return dict(KalmanODE_py=KalmanODE_py, ibm_init=ibm_init, car_init=car_init,
                zero_pad=zero_pad, indep_init=indep_init, rand_mat=rand_mat,
                inference=inference)


# ---- the four right-hand sides, in the reference's Python form --------------------
def chkrebtii(x, t, theta=None, out=None):      # tests/utils.py:9-12
    out[0] = math.sin(2 * t) - x[0]


def fitz(X, t, theta, out=None):                # examples/fitz.py:7-16
    a, b, c = theta
    p = len(X) // 2
    V, R = X[0], X[p]
    out[0] = c * (V - V * V * V / 3 + R)
    out[1] = -1 / c * (V - a + b * R)
    return out


def lorenz(X, t, theta, out=None):              # examples/lorenz.py:10-19
    rho, sigma, beta = theta
    p = len(X) // 3
    x, y, z = X[0], X[p], X[2 * p]
    out[0] = -sigma * x + sigma * y
    out[1] = rho * x - y - x * z
    out[2] = -beta * z + x * y
    return out


def mseir(X, t, theta, out=None):               # examples/mseir.py:7-20
    if out is None:
        out = np.empty(5)
    p = len(X) // 5
    M, S, E, I, R = X[::p]
    N = M + S + E + I + R
    Lambda, delta, beta, mu, epsilon, gamma = theta
    out[0] = Lambda - delta * M - mu * M
    out[1] = delta * M - beta * S * I / N - mu * S
    out[2] = beta * S * I / N - (epsilon + mu) * E
    out[3] = epsilon * E - (gamma + mu) * I
    out[4] = gamma * I - mu * R
    return out


def _wmat(n_deriv):
    W = np.zeros((len(n_deriv), sum(n_deriv) + len(n_deriv)))
    for i in range(len(n_deriv)):
        W[i, sum(n_deriv[:i]) + i + 1] = 1
    return W


def _solve_all(ref, fun, W, tmin, tmax, N, kinit, z, x0, theta):
    """mv and sim from the reference's pure-Python solver (one object per call: it
    mutates nothing, but keep the calls independent)."""
    k = ref["KalmanODE_py"](W, tmin, tmax, N, fun, **kinit)
    k.z_state = z
    mu, var = k.solve_mv(x0, W, theta)
    x = k.solve_sim(x0, W, theta)
    return np.array(x), np.array(mu), np.array(var)


def main():
    ref = _import_reference()
    zero_pad, indep_init, ibm_init, car_init = (ref["zero_pad"], ref["indep_init"],
                                                ref["ibm_init"], ref["car_init"])
    out = {}

    # ---- C1 chkrebtii (tests/test_kalmanode.py:15-38; README / docs n_eval) --------
    for N in (100, 200):
        n_deriv, n_prior = [2], [4]
        tmin, tmax = 0.0, 10.0
        W = zero_pad(np.array([[0.0, 0.0, 1.0]]), n_deriv, n_prior)
        x0 = zero_pad(np.array([-1.0, 0.0, 1.0]), n_deriv, n_prior)
        kinit = ibm_init((tmax - tmin) / N, n_prior, [0.5])
        np.random.seed(100 + N)
        z = ref["rand_mat"](N, 4)
        x, mu, var = _solve_all(ref, chkrebtii, W, tmin, tmax, N, kinit, z, x0, None)
        out[f"chkrebtii_n{N}"] = dict(
            W=W, x0=x0, z=z, tmin=tmin, tmax=tmax, N=N, Q=kinit["wgt_state"],
            c=kinit["mu_state"], R=kinit["var_state"], x=x, mu=mu, var=var)

    # ---- C2 FitzHugh-Nagumo (tests/timings/fitz_time.py:57-85) -----------------------
    n_deriv, n_prior = [1, 1], [3, 3]
    tmin, tmax, N = 0.0, 40.0, 400
    W = zero_pad(np.array([[0.0, 1.0, 0.0, 0.0], [0.0, 0.0, 0.0, 1.0]]), n_deriv, n_prior)
    x0 = zero_pad(np.array([-1, 1, 1, 1 / 3]), n_deriv, n_prior)
    kinit = indep_init(ibm_init((tmax - tmin) / N, n_prior, [0.1, 0.1]), n_prior)
    rng = np.random.default_rng(0)
    thetas = np.exp(np.log([0.2, 0.2, 3.0]) + 0.1 * rng.standard_normal((4, 3)))
    thetas[0] = [0.2, 0.2, 3.0]
    np.random.seed(2)
    z = ref["rand_mat"](N, 6)
    xs, mus, vars_ = zip(*[_solve_all(ref, fitz, W, tmin, tmax, N, kinit, z, x0, th)
                           for th in thetas])
    out["fitz_n400"] = dict(W=W, x0=x0, z=z, tmin=tmin, tmax=tmax, N=N, theta=thetas,
                            Q=kinit["wgt_state"], c=kinit["mu_state"], R=kinit["var_state"],
                            x=np.array(xs), mu=np.array(mus), var=np.array(vars_))
    # thinning + log-likelihood (examples/inference.py:54-56, 66-94), gamma = 0.2
    inf = ref["inference"]([0, 3], int(tmin), int(tmax), fitz)
    data_t = np.linspace(1, tmax, int(tmax - tmin))
    ode_t = np.linspace(tmin, tmax, N + 1)
    Y = np.array(xs[0])[::10][1:, [0, 3]] + 0.2 * np.random.default_rng(5).standard_normal((40, 2))
    ll = [inf.loglike(Y, inf.thinning(data_t, ode_t, np.array(x))[:, [0, 3]], 0.2) for x in xs]
    llmu = [inf.loglike(Y, inf.thinning(data_t, ode_t, np.array(m))[:, [0, 3]], 0.2) for m in mus]
    # recover the indices thinning picked (X rows are distinct, so match rows)
    probe = np.arange(N + 1, dtype=float)[:, None] * np.ones((1, 2))
    ind = inf.thinning(data_t, ode_t, probe)[:, 0].astype(int)
    out["fitz_n400"].update(Y=Y, loglik_x=np.array(ll), loglik_mu=np.array(llmu),
                            thin_ind=ind, data_t=data_t, gamma=0.2)
    # a second thinning case with a grid that does not hit the data times exactly
    ode_t2 = np.linspace(0, 40, 131)
    out["thinning_n130"] = dict(
        data_t=data_t, ode_t=ode_t2,
        ind=inf.thinning(data_t, ode_t2, np.arange(131.0)[:, None] * np.ones((1, 2)))[:, 0].astype(int))

    # ---- C3 Lorenz63 (examples/lorenz.py:21-53; IBM prior tests/timings/lorenz_time.py:84)
    n_deriv, n_prior = [1, 1, 1], [3, 3, 3]
    tmin, tmax, N = 0.0, 20.0, 5000
    W = zero_pad(_wmat(n_deriv), n_deriv, n_prior)
    x0 = zero_pad(np.array([-12, 70, -5, 125, 38, -124 / 3]), n_deriv, n_prior)
    theta = np.array([28, 10, 8 / 3])
    kinit = indep_init(ibm_init((tmax - tmin) / N, n_prior, [0.5] * 3), n_prior)
    np.random.seed(2)
    z = ref["rand_mat"](N, 9)
    x, mu, var = _solve_all(ref, lorenz, W, tmin, tmax, N, kinit, z, x0, theta)
    out["lorenz_n5000"] = dict(W=W, x0=x0, z=z, tmin=tmin, tmax=tmax, N=N, theta=theta,
                               Q=kinit["wgt_state"], c=kinit["mu_state"],
                               R=kinit["var_state"], x=x, mu=mu, var_every=50,
                               var=var[::50])
    # Lorenz with the CAR(p) prior of examples/lorenz.py:44-57 (dense Q blocks), short
    N = 600
    tmax = 2.4
    np.random.seed(7)
    X0 = np.column_stack([[-12, -5, 38], [70, 125, -124 / 3]])
    ode_init, x0c = car_init((tmax - tmin) / N, n_prior, np.array([1.3] * 3),
                             np.array([0.5] * 3), X0)
    kinit = indep_init(ode_init, n_prior)
    z = ref["rand_mat"](N, 9)
    x, mu, var = _solve_all(ref, lorenz, W, tmin, tmax, N, kinit, z, x0c, theta)
    out["lorenz_car_n600"] = dict(W=W, x0=x0c, z=z, tmin=tmin, tmax=tmax, N=N, theta=theta,
                                  Q=kinit["wgt_state"], c=kinit["mu_state"],
                                  R=kinit["var_state"], x=x, mu=mu, var_every=10,
                                  var=var[::10], tau=1.3, sigma=0.5,
                                  car_wgt=ode_init["wgt_state"][0],
                                  car_var=ode_init["var_state"][0])

    # ---- C4 MSEIR (examples/mseir.py:22-47) ------------------------------------------
    n_deriv, n_prior = [1] * 5, [3] * 5
    tmin, tmax, N = 0.0, 40.0, 1000
    theta_true = np.array([1.1, 0.7, 0.4, 0.005, 0.02, 0.03])
    x0v = np.array([1000.0, 100, 50, 3, 3])
    v0 = mseir(x0v, 0, theta_true)
    x0 = zero_pad(np.ravel([x0v, v0], "F"), n_deriv, n_prior)
    W = zero_pad(_wmat(n_deriv), n_deriv, n_prior)
    kinit = indep_init(ibm_init((tmax - tmin) / N, n_prior, [0.1] * 5), n_prior)
    rng = np.random.default_rng(3)
    thetas = theta_true * np.exp(0.1 * rng.standard_normal((2, 6)))
    thetas[0] = theta_true
    np.random.seed(4)
    z = ref["rand_mat"](N, 15)
    xs, mus, vars_ = zip(*[_solve_all(ref, mseir, W, tmin, tmax, N, kinit, z, x0, th)
                           for th in thetas])
    out["mseir_n1000"] = dict(W=W, x0=x0, z=z, tmin=tmin, tmax=tmax, N=N, theta=thetas,
                              Q=kinit["wgt_state"], c=kinit["mu_state"], R=kinit["var_state"],
                              x=np.array(xs), mu=np.array(mus), var_every=25,
                              var=np.array(vars_)[:, ::25])

    # ---- prior helpers ---------------------------------------------------------------
    ib = ibm_init(0.1, [3, 4], [0.1, 0.7])
    out["helpers"] = dict(
        ibm_wgt0=ib["wgt_state"][0], ibm_var0=ib["var_state"][0],
        ibm_wgt1=ib["wgt_state"][1], ibm_var1=ib["var_state"][1],
        ibm_single_wgt=ibm_init(0.05, [4], [0.5])["wgt_state"],
        ibm_single_var=ibm_init(0.05, [4], [0.5])["var_state"],
        indep_wgt=indep_init(ib, [3, 4])["wgt_state"],
        indep_var=indep_init(ib, [3, 4])["var_state"],
        zero_pad_vec=zero_pad(np.array([1.0, 2, 3, 4, 5]), [1, 2], [3, 5]),
        zero_pad_mat=zero_pad(np.array([[0.0, 1, 0, 0, 0], [0, 0, 0, 0, 1]]), [1, 2], [3, 5]),
        car_wgt=car_init(0.01, [4], [1.3], [0.5])["wgt_state"],
        car_var=car_init(0.01, [4], [1.3], [0.5])["var_state"],
    )

    for name, d in out.items():
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **{k: np.asarray(v) for k, v in d.items()})
        print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
