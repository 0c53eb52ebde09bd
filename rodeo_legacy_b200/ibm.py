"""q-times integrated Brownian motion prior (host, NumPy).

API of the reference's ``rodeo.ibm`` (``rodeo/ibm/ibm_init.py``: ``ibm_state`` :3-33,
``ibm_init`` :35-65): per variable with block size p, step dt and scale sigma

    A[i, j] = dt^(j-i) / (j-i)!                                   (i <= j, else 0)
    Q[i, j] = sigma^2 dt^(2p+1-i-j) / ((2p+1-i-j) (p-i)! (p-j)!)

(0-based i, j; note q = block size p, not p-1, as in the reference).  Uses
``math.factorial`` -- the reference's ``np.math.factorial`` no longer exists in NumPy 2.
"""
from math import factorial

import numpy as np

__all__ = ["ibm_state", "ibm_init"]


def ibm_state(dt, q, sigma):
    idx = np.arange(q)
    d = idx[None, :] - idx[:, None]                       # j - i
    fact = np.array([factorial(k) for k in range(q + 1)], dtype=float)
    A = np.where(d >= 0, dt ** np.maximum(d, 0) / fact[np.maximum(d, 0)], 0.0)
    e = 2 * q + 1 - idx[:, None] - idx[None, :]           # 2q+1-i-j
    Q = sigma ** 2 * dt ** e / (e * fact[q - idx][:, None] * fact[q - idx][None, :])
    return np.asfortranarray(A), np.asfortranarray(Q)


def ibm_init(dt, n_deriv_prior, sigma):
    """dict(wgt_state, mu_state, var_state); lists of per-variable blocks when there is
    more than one variable (feed to ``indep_init``), bare matrices when there is one."""
    blocks = [ibm_state(dt, p, s) for p, s in zip(n_deriv_prior, sigma)]
    wgt = [b[0] for b in blocks]
    var = [b[1] for b in blocks]
    if len(blocks) == 1:
        wgt, var = wgt[0], var[0]
    return {"wgt_state": wgt, "mu_state": np.zeros(int(np.sum(n_deriv_prior))),
            "var_state": var}
