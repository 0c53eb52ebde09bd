"""Host-side helpers of the KalmanODE call sequence (NumPy, one-off per problem).

Same names, argument meaning and return layout as the reference's ``rodeo.utils``
(``rodeo/utils/utils.py``): ``zero_pad`` :49-75, ``indep_init`` :77-109, ``rand_mat``
:128-145, ``mvncond`` :6-31, ``solveV`` :34-47, ``norm_sim`` :111-126.  They run once per
problem on the host; the per-solve arithmetic lives in the CUDA kernels.
"""
import numpy as np
import scipy.linalg as _sl

__all__ = ["mvncond", "solveV", "zero_pad", "indep_init", "norm_sim", "rand_mat"]


def mvncond(mu, Sigma, icond):
    """A, b, V with  y[~icond] | y[icond] ~ N(A y[icond] + b, V)  (utils.py:6-31)."""
    icond = np.asarray(icond, dtype=bool)
    keep = ~icond
    S22 = Sigma[np.ix_(icond, icond)]
    S12 = Sigma[np.ix_(keep, icond)]
    A = S12 @ _sl.cho_solve(_sl.cho_factor(S22), np.eye(int(icond.sum())))
    b = mu[keep] - A @ mu[icond]
    V = Sigma[np.ix_(keep, keep)] - A @ Sigma[np.ix_(icond, keep)]
    return A, b, V


def solveV(V, B):
    """X = V^{-1} B for a variance matrix V via Cholesky (utils.py:34-47)."""
    return _sl.cho_solve(_sl.cho_factor(V), B)


def zero_pad(x, n_deriv, n_deriv_prior):
    """Spread each variable's (q_i+1) known derivatives into the head of its p_i-block.

    ``x`` is a vector (stacked per-variable derivatives) or a matrix with one row per
    variable (e.g. W).  Returned array is F-ordered like the reference's (utils.py:49-75).
    """
    x = np.asarray(x, dtype=float)
    have = np.asarray(n_deriv, dtype=int) + 1          # entries supplied per variable
    src = np.concatenate(([0], np.cumsum(have)))
    dst = np.concatenate(([0], np.cumsum(n_deriv_prior)))
    p = int(dst[-1])
    if x.ndim == 1:
        out = np.zeros(p, order="F")
        for i, k in enumerate(have):
            out[dst[i]:dst[i] + k] = x[src[i]:src[i] + k]
    else:
        out = np.zeros((len(have), p), order="F")
        for i, k in enumerate(have):
            out[i, dst[i]:dst[i] + k] = x[i, src[i]:src[i] + k]
    return out


def indep_init(init, n_deriv_prior):
    """Assemble per-variable prior blocks into block-diagonal Q and R (utils.py:77-109)."""
    blocks_Q, blocks_R = init["wgt_state"], init["var_state"]
    if isinstance(blocks_Q, np.ndarray) and blocks_Q.ndim == 2:   # single variable
        blocks_Q, blocks_R = [blocks_Q], [blocks_R]
    p = int(np.sum(n_deriv_prior))
    Q = np.zeros((p, p), order="F")
    R = np.zeros((p, p), order="F")
    o = 0
    for k, bQ, bR in zip(n_deriv_prior, blocks_Q, blocks_R):
        Q[o:o + k, o:o + k] = bQ
        R[o:o + k, o:o + k] = bR
        o += k
    return {"wgt_state": Q, "mu_state": init["mu_state"], "var_state": R}


def norm_sim(z, mu, V):
    """x = mu + chol_lower(V) z  (utils.py:111-126)."""
    return _sl.cholesky(V, lower=True) @ z + mu


def rand_mat(n, p, pd=True):
    """(p, n) F-ordered N(0,1) matrix from the global NumPy RNG; V V^T when p == n and pd
    (utils.py:128-145 -- the quirk only matters when n_eval == n_state)."""
    V = np.asfortranarray(np.random.randn(p, n))
    if p == n and pd:
        V = np.asfortranarray(V @ V.T)
    return V
