"""Batched parameter inference on top of ``rodeo.cuda`` -- the caller side of the hot path.

Mirrors the reference's ``examples/inference.py`` class (same constructor, ``loglike``,
``simulate``, ``thinning``, ``kalman_nlpost``, ``phi_fit``, ``theta_sample``; plotting left out):
the reference evaluates the negative log-posterior of phi = log(theta) one solve at a time
(``kalman_nlpost``: ``solve_sim`` -> ``thinning`` -> Gaussian log-likelihood + log-normal prior,
examples/inference.py:85-94) inside Nelder-Mead and a numerical Hessian (``phi_fit`` :117-134).
Here ``kalman_nlpost`` takes a whole batch of phi and makes ONE fused solve + thinning + likelihood
call (``KalmanODE.loglik``, no per-step output ever materialised), and the Hessian of ``phi_fit`` is
a central-difference stencil evaluated in a single batched call.
"""
import numpy as np
import scipy.linalg
import scipy.optimize
import scipy.stats
from scipy.integrate import odeint

__all__ = ["inference", "thin_index"]


def thin_index(data_tseq, ode_tseq):
    """Grid index the reference's ``thinning`` picks for every data time (examples/inference.py:66-83):
    the solver time nearest to it, scanning upward, ties and the ``diff`` bookkeeping as there."""
    ind = np.zeros(len(data_tseq), dtype=int)
    data_i = ode_i = 0
    diff = 1000
    while data_i < len(data_tseq) and ode_i < len(ode_tseq):
        gap = abs(data_tseq[data_i] - ode_tseq[ode_i])
        if data_tseq[data_i] > ode_tseq[ode_i]:
            diff = min(diff, gap)
            ode_i += 1
        else:
            ind[data_i] = ode_i if diff > gap else ode_i - 1
            data_i += 1
            diff = 1000
    return ind


class inference:
    """Mode / quadrature inference for theta with the batched KalmanODE solver.

    Args as the reference class: ``state_ind`` (index of each variable's 0-th derivative), ``tmin``,
    ``tmax``, ``fun`` (host RHS, only used by ``simulate``), ``W``, ``kode`` (a
    ``rodeo_legacy_b200.cuda.KalmanODE``).
    """

    def __init__(self, state_ind, tmin, tmax, fun, W=None, kode=None):
        self.state_ind = state_ind
        self.tmin = tmin
        self.tmax = tmax
        self.fun = fun
        self.W = W
        self.kode = kode

    def loglike(self, x, mean, var):
        """Sum of normal log-densities; ``var`` is passed as the SCALE, as in the reference (:54-56)."""
        return np.sum(scipy.stats.norm.logpdf(x=x, loc=mean, scale=var))

    def simulate(self, fun, x0, theta_true, gamma, rng=None):
        """Noisy observations at t = tmin+1, ..., tmax from ``odeint`` (:58-64)."""
        tseq = np.linspace(self.tmin, self.tmax, int(self.tmax - self.tmin) + 1)
        X_t = odeint(fun, x0, tseq, args=(theta_true,))[1:, ]
        rng = np.random.default_rng() if rng is None else rng
        return X_t + gamma * rng.normal(size=X_t.shape), X_t

    def thinning(self, data_tseq, ode_tseq, X):
        """Rows of the solver output nearest to the data times (:66-83)."""
        return X[thin_index(data_tseq, ode_tseq), :]

    def _grids(self, step_size):
        data_tseq = np.linspace(1, self.tmax, int(self.tmax - self.tmin))
        ode_tseq = np.linspace(self.tmin, self.tmax, int((self.tmax - self.tmin) / step_size) + 1)
        return data_tseq, ode_tseq

    def kalman_nlpost(self, phi, Y_t, x0, step_size, phi_mean, phi_sd, gamma, use_draw=True):
        """Negative log-posterior of phi = log(theta); ``phi`` (k,) -> float or (B, k) -> (B,).

        Per system: theta = exp(phi); X = solve_sim (``use_draw=True``, the reference's choice, with
        the solver object's z_state) or the posterior mean; thinned to the data times; Gaussian
        log-likelihood with scale gamma plus the normal log-prior on phi (:85-94).  One fused GPU
        call for the whole batch.
        """
        phi = np.asarray(phi, dtype=float)
        single = phi.ndim == 1
        phib = np.atleast_2d(phi)
        data_tseq, ode_tseq = self._grids(step_size)
        ind = thin_index(data_tseq, ode_tseq)
        x0b = np.tile(np.asarray(x0, dtype=float), (len(phib), 1))
        if self.W is not None:
            self.kode.wgt_meas = np.asfortranarray(self.W)
        ll = self.kode.loglik(x0b, np.exp(phib), np.asarray(Y_t, dtype=float), ind, self.state_ind, gamma,
                              use_draw=use_draw)
        lp = np.asarray(ll) + np.sum(scipy.stats.norm.logpdf(phib, loc=phi_mean, scale=phi_sd), axis=1)
        return -lp[0] if single else -lp

    def phi_fit(self, Y_t, x0, step_size, theta_true, phi_sd, gamma, use_draw=True, fd_step=1e-3):
        """Posterior mode of phi (Nelder-Mead, as :117-127) and its covariance from the inverse of a
        central-difference Hessian whose 2k^2+1 stencil points are evaluated in ONE batched call
        (the reference uses numdifftools, one solve per point, :128-133)."""
        phi_mean = np.log(theta_true)
        k = len(phi_mean)
        args = (Y_t, x0, step_size, phi_mean, phi_sd, gamma, use_draw)
        res = scipy.optimize.minimize(self.kalman_nlpost, phi_mean + 0.1, args=args, method="Nelder-Mead")
        phi_hat = res.x
        h = fd_step * np.maximum(1.0, np.abs(phi_hat))
        pts = [phi_hat]
        for i in range(k):
            for j in range(i, k):
                for si, sj in ((1, 1), (1, -1), (-1, 1), (-1, -1)):
                    p = phi_hat.copy()
                    p[i] += si * h[i]
                    p[j] += sj * h[j] if j != i else 0.0
                    pts.append(p)
        f = self.kalman_nlpost(np.array(pts), *args)
        H = np.zeros((k, k))
        q = 1
        for i in range(k):
            for j in range(i, k):
                fpp, fpm, fmp, fmm = f[q:q + 4]
                q += 4
                if i == j:                      # points are phi +- h e_i (listed twice each)
                    H[i, i] = (fpp - 2 * f[0] + fmm) / h[i] ** 2
                else:
                    H[i, j] = H[j, i] = (fpp - fpm - fmp + fmm) / (4 * h[i] * h[j])
        cho = scipy.linalg.cho_factor(H)
        return phi_hat, scipy.linalg.cho_solve(cho, np.eye(k))

    def theta_sample(self, phi_hat, phi_var, n_samples, rng=None):
        """theta = exp(phi), phi ~ N(phi_hat, phi_var) (:136-140)."""
        rng = np.random.default_rng() if rng is None else rng
        return np.exp(rng.multivariate_normal(phi_hat, phi_var, n_samples))
