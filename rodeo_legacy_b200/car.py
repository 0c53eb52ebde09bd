"""CAR(p) (continuous autoregressive) prior (host, NumPy).

API of the reference's ``rodeo.car`` package: ``root_gen`` / ``car_initial_draw`` /
``car_state`` / ``car_init`` (``rodeo/car/car_init.py`` :8-20, :22-49, :51-73, :75-120),
``car_mou`` (``car_mou.py`` :3-42), ``car_cov`` (``car_cov.py`` :5-44), ``car_var``
(``car_var.py`` :5-31).  With roots r, Vandermonde-like Q[i, k] = (-r_k)^i and
Sigma~ = Q^-1 diag(0..0, sigma^2) Q^-T:

    V~(t)[i, j] = Sigma~[i, j] / (r_i + r_j) (1 - exp(-(r_i + r_j) t)),  V(t) = Q V~ Q^T
    wgt_state   = Q diag(exp(-r dt)) Q^-1,   var_state = V(dt)
"""
import numpy as np

from .utils import mvncond

__all__ = ["root_gen", "car_mou", "car_cov", "car_var", "car_state",
           "car_initial_draw", "car_initial_draw_batch", "car_init"]


def root_gen(tau, p):
    return 10.0 * np.concatenate(([1.0 / tau], np.linspace(1 + 1 / (10 * (p - 1)), 1.1, p - 1)))


def car_mou(roots, sigma=1.0, test=False):
    roots = np.asarray(roots, dtype=float)
    p = len(roots)
    Q = np.empty((p, p))          # Q[i, k] = (-r_k)^i by repeated multiplication
    row = np.ones(p)
    for i in range(p):
        Q[i] = row
        row = -row * roots
    Q_inv = np.linalg.pinv(Q)
    if test:
        Sigma = np.zeros((p, p))
        Sigma[-1, -1] = sigma * sigma
        Gamma = np.zeros((p, p))
        Gamma[np.arange(p - 1), np.arange(1, p)] = -1.0
        Gamma[-1] = np.linalg.multi_dot([Q[-1], np.diag(roots), Q_inv])
        return Gamma, Sigma
    diag = np.zeros(p)
    diag[-1] = sigma * sigma
    return (Q_inv * diag) @ Q_inv.T, Q


def _v_tilde(Sigma_tilde, roots, t=None):
    rs = roots[:, None] + roots[None, :]
    V = Sigma_tilde / rs
    if t is not None:
        V = V * (1.0 - np.exp(-rs * t))
    iu = np.triu_indices(len(roots))
    out = np.zeros_like(V)
    out[iu] = V[iu]
    return out + np.triu(out, 1).T           # symmetrise from the upper triangle


def car_var(tseq, roots, sigma=1.0):
    roots = np.asarray(roots, dtype=float)
    Sigma_tilde, Q = car_mou(roots, sigma)
    var = np.zeros((len(roots), len(roots), len(tseq)), order="F")
    for k, t in enumerate(tseq):
        var[:, :, k] = np.linalg.multi_dot([Q, _v_tilde(Sigma_tilde, roots, t), Q.T])
    return var


def car_cov(tseq, roots, sigma=1.0, corr=False, v_infinity=False):
    roots = np.asarray(roots, dtype=float)
    Sigma_tilde, Q = car_mou(roots, sigma)
    Q_inv = np.linalg.pinv(Q)
    V_inf = np.linalg.multi_dot([Q, _v_tilde(Sigma_tilde, roots), Q.T])
    if v_infinity:
        return V_inf
    sd = np.sqrt(np.diag(V_inf))
    cov = np.zeros((len(tseq), len(roots), len(roots)))
    for k, t in enumerate(tseq):
        cov[k] = V_inf @ ((Q_inv.T * np.exp(-roots * t)) @ Q.T)
        if corr:
            cov[k] = cov[k] / sd[:, None] / sd[None, :]
    return cov


def car_state(delta_t, roots, sigma):
    roots = np.asarray(roots, dtype=float)
    _, Q = car_mou(roots, sigma)
    var_state = car_var(delta_t, roots, sigma)[:, :, 0]
    wgt_state = np.asfortranarray((Q * np.exp(-roots * delta_t[0])) @ np.linalg.pinv(Q))
    return wgt_state, var_state


def car_initial_draw(roots, sigma, x0, p):
    """X0 with the q+1 known derivatives fixed and the rest drawn from the stationary
    CAR law conditioned on them (global NumPy RNG, as the reference)."""
    x0 = np.asarray(x0, dtype=float)
    q = len(x0) - 1
    if p == q + 1:
        return x0
    V_inf = car_cov([], roots, sigma, v_infinity=True)
    icond = np.arange(p) <= q
    A, b, V = mvncond(np.zeros(p), V_inf, icond)
    out = np.zeros(p)
    out[:q + 1] = x0
    out[q + 1:] = np.random.multivariate_normal(A @ x0 + b, V)
    return out


def car_initial_draw_batch(roots, sigma, X0, p, z=None):
    """car_initial_draw for a batch: X0 is (B, q+1) -> (B, p), each row's unknown derivatives an
    independent draw from the stationary CAR law conditioned on the known ones
    (rodeo/car/car_init.py:22-49, vectorised).  NumPy in -> NumPy out (global NumPy
    RNG unless standard normals `z` (B, p-q-1) are given); torch in -> torch out on the input's device
    (torch RNG), so x0 for a batch of solves can be produced next to the solver without a host
    round trip.  The conditional law does not depend on the system: one (A, b, chol V) serves the batch."""
    is_torch = hasattr(X0, "device") and not isinstance(X0, np.ndarray)
    q = X0.shape[1] - 1
    if p == q + 1:
        return X0
    V_inf = car_cov([], roots, sigma, v_infinity=True)
    A, b, V = mvncond(np.zeros(p), V_inf, np.arange(p) <= q)
    L = np.linalg.cholesky(V)
    if is_torch:
        import torch
        kw = dict(dtype=torch.float64, device=X0.device)
        At, bt, Lt = torch.as_tensor(A, **kw), torch.as_tensor(b, **kw), torch.as_tensor(L, **kw)
        zz = torch.randn(X0.shape[0], p - q - 1, **kw) if z is None else torch.as_tensor(z, **kw)
        return torch.cat([X0.to(torch.float64), X0.to(torch.float64) @ At.T + bt + zz @ Lt.T], dim=1)
    X0 = np.asarray(X0, dtype=float)
    zz = np.random.standard_normal((X0.shape[0], p - q - 1)) if z is None else np.asarray(z, dtype=float)
    return np.concatenate([X0, X0 @ A.T + b + zz @ L.T], axis=1)


def car_init(dt, n_deriv_prior, tau, sigma, x0=None):
    delta_t = np.array([dt])
    wgt, var = [], []
    x0_state = np.zeros(int(np.sum(n_deriv_prior))) if x0 is not None else None
    o = 0
    for i, p in enumerate(n_deriv_prior):
        roots = root_gen(tau[i], p)
        if x0 is not None:
            x0_state[o:o + p] = car_initial_draw(roots, sigma[i], x0[i], p)
        A, V = car_state(delta_t, roots, sigma[i])
        wgt.append(A)
        var.append(V)
        o += p
    if len(wgt) == 1:
        wgt, var = wgt[0], var[0]
    init = {"wgt_state": wgt, "mu_state": np.zeros(int(np.sum(n_deriv_prior))),
            "var_state": var}
    return (init, x0_state) if x0 is not None else init
