"""Golden fixtures for priors with UNEQUAL block sizes (n_deriv_prior = [3, 4], [4, 3], [2, 3], [2, 3, 3]), from the
REFERENCE's own pure-Python path.  Same imports and shims as make_golden.py (which it reuses); run in the build
container only:   python tests/golden/make_golden_mixed.py

The reference accepts any n_deriv_prior list (rodeo/utils/utils.py:97-105 zero_pad, rodeo/ibm/ibm_init.py:54-58,
rodeo/utils/utils.py indep_init); its Python `ode_fun` indexes the state itself, so the right-hand sides below read
variable k at the start of block k (examples/fitz.py:7-16 does the same with its own p).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402


def fitz_at(off):
    def f(X, t, theta, out=None):            # examples/fitz.py:7-16 with explicit offsets
        a, b, c = theta
        V, R = X[off[0]], X[off[1]]
        out[0] = c * (V - V * V * V / 3 + R)
        out[1] = -1 / c * (V - a + b * R)
        return out
    return f


def lorenz_at(off):
    def f(X, t, theta, out=None):            # examples/lorenz.py:10-19 with explicit offsets
        rho, sigma, beta = theta
        x, y, z = X[off[0]], X[off[1]], X[off[2]]
        out[0] = -sigma * x + sigma * y
        out[1] = rho * x - y - x * z
        out[2] = -beta * z + x * y
        return out
    return f


def main():
    ref = mg._import_reference()
    zero_pad, indep_init, ibm_init = ref["zero_pad"], ref["indep_init"], ref["ibm_init"]
    out = {}
    # p = 4 blocks: dt = 1/6 keeps the draw covariance V~ = Sigma_f - T~^T T (KalmanTV.h:522-523) factorisable in every
    # FP64 implementation; at dt = 0.1 its last Cholesky pivot is rounding noise (the oracle itself fails from
    # dt = 1/15 on), see DESIGN section 3
    cases = [("fitz", [3, 4], 120, 20.0), ("fitz", [4, 3], 120, 20.0), ("fitz", [2, 3], 200, 20.0),
             ("lorenz", [2, 3, 3], 300, 1.2)]
    for ode, n_prior, N, tmax in cases:
        m, n = len(n_prior), sum(n_prior)
        n_deriv = [1] * m
        off = np.concatenate([[0], np.cumsum(n_prior)[:-1]]).astype(int)
        tmin = 0.0
        W = zero_pad(mg._wmat(n_deriv), n_deriv, n_prior)
        if ode == "fitz":
            x0 = zero_pad(np.array([-1, 1, 1, 1 / 3]), n_deriv, n_prior)
            theta = np.array([[0.2, 0.2, 3.0], [0.25, 0.15, 2.7]])
            fun, sigma = fitz_at(off), [0.1, 0.2]
        else:
            x0 = zero_pad(np.array([-12, 70, -5, 125, 38, -124 / 3]), n_deriv, n_prior)
            theta = np.array([[28, 10, 8 / 3], [27.0, 9.5, 2.5]])
            fun, sigma = lorenz_at(off), [0.5, 0.4, 0.6]
        kinit = indep_init(ibm_init((tmax - tmin) / N, n_prior, sigma), n_prior)
        np.random.seed(11 + n)
        z = ref["rand_mat"](N, n)
        xs, mus, vars_ = zip(*[mg._solve_all(ref, fun, W, tmin, tmax, N, kinit, z, x0, th) for th in theta])
        name = f"{ode}_blocks_" + "_".join(str(v) for v in n_prior)
        out[name] = dict(W=W, x0=x0, z=z, tmin=tmin, tmax=tmax, N=N, theta=theta, block_sizes=np.array(n_prior),
                         Q=kinit["wgt_state"], c=kinit["mu_state"], R=kinit["var_state"],
                         x=np.array(xs), mu=np.array(mus), var=np.array(vars_))
    path = os.path.join(HERE, "mixed_blocks.npz")
    np.savez_compressed(path, **{f"{k}__{f}": np.asarray(v) for k, d in out.items() for f, v in d.items()})
    print(f"mixed_blocks.npz: {os.path.getsize(path) / 1024:.0f} KiB; cases {list(out)}")


if __name__ == "__main__":
    main()
