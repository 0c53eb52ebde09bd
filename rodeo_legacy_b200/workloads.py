"""The BASELINE.json workloads, built with this package's host helpers: synthetic problem
definitions and seeded batches for bench.py, tools/, examples/ and the tests."""
import numpy as np

from . import ibm_init, indep_init, zero_pad


def _wmat(n_deriv):
    W = np.zeros((len(n_deriv), sum(n_deriv) + len(n_deriv)))
    for i in range(len(n_deriv)):
        W[i, sum(n_deriv[:i]) + i + 1] = 1
    return W


def _pack(ode, W, tmin, tmax, N, kinit, x0, **kw):
    d = dict(ode=ode, W=np.asfortranarray(W), tmin=float(tmin), tmax=float(tmax), N=int(N),
             Q=kinit["wgt_state"], c=kinit["mu_state"], R=kinit["var_state"], x0=x0)
    d.update(kw)
    return d


def chkrebtii(N=200):
    """C1: tests/test_kalmanode.py:15-32 with n_eval = 200."""
    n_deriv, n_prior = [2], [4]
    W = zero_pad(np.array([[0.0, 0.0, 1.0]]), n_deriv, n_prior)
    x0 = zero_pad(np.array([-1.0, 0.0, 1.0]), n_deriv, n_prior)
    kinit = indep_init(ibm_init(10.0 / N, n_prior, [0.5]), n_prior)
    return _pack("chkrebtii", W, 0, 10, N, kinit, x0, n_theta=0)


def fitz(N=400):
    """C2 / C5: tests/timings/fitz_time.py:57-85."""
    n_deriv, n_prior = [1, 1], [3, 3]
    W = zero_pad(np.array([[0.0, 1.0, 0.0, 0.0], [0.0, 0.0, 0.0, 1.0]]), n_deriv, n_prior)
    x0 = zero_pad(np.array([-1, 1, 1, 1 / 3]), n_deriv, n_prior)
    kinit = indep_init(ibm_init(40.0 / N, n_prior, [0.1, 0.1]), n_prior)
    return _pack("fitz", W, 0, 40, N, kinit, x0, n_theta=3, theta0=np.array([0.2, 0.2, 3.0]),
                 state_ind=[0, 3])


def fitz_batch(B, seed=0):
    """theta_b = exp(log(0.2, 0.2, 3) + 0.1 eps_b), default_rng(seed); x0 shared."""
    w = fitz()
    rng = np.random.default_rng(seed)
    theta = np.exp(np.log(w["theta0"]) + 0.1 * rng.standard_normal((B, 3)))
    return w, np.tile(w["x0"], (B, 1)), theta


def lorenz(N=5000, tmax=20.0):
    """C3: examples/lorenz.py:21-53 with the IBM prior of tests/timings/lorenz_time.py:84."""
    n_deriv, n_prior = [1, 1, 1], [3, 3, 3]
    W = zero_pad(_wmat(n_deriv), n_deriv, n_prior)
    x0 = zero_pad(np.array([-12, 70, -5, 125, 38, -124 / 3]), n_deriv, n_prior)
    kinit = indep_init(ibm_init(tmax / N, n_prior, [0.5] * 3), n_prior)
    return _pack("lorenz63", W, 0, tmax, N, kinit, x0, n_theta=3,
                 theta0=np.array([28.0, 10.0, 8 / 3]))


def lorenz_batch(B, N=5000, tmax=20.0, seed=1):
    """x0_b: positions perturbed by 1e-2 eps_b, velocities recomputed from the RHS."""
    w = lorenz(N, tmax)
    rng = np.random.default_rng(seed)
    pos = np.array([-12.0, -5.0, 38.0]) + 1e-2 * rng.standard_normal((B, 3))
    rho, sigma, beta = w["theta0"]
    vel = np.stack([-sigma * pos[:, 0] + sigma * pos[:, 1],
                    rho * pos[:, 0] - pos[:, 1] - pos[:, 0] * pos[:, 2],
                    -beta * pos[:, 2] + pos[:, 0] * pos[:, 1]], axis=1)
    x0 = np.zeros((B, 9))
    x0[:, 0::3] = pos
    x0[:, 1::3] = vel
    return w, x0, np.tile(w["theta0"], (B, 1))


def _mseir_rhs(x, th):
    M, S, E, I, R = x
    N = M + S + E + I + R
    Lambda, delta, beta, mu, epsilon, gamma = th
    return np.array([Lambda - delta * M - mu * M, delta * M - beta * S * I / N - mu * S,
                     beta * S * I / N - (epsilon + mu) * E, epsilon * E - (gamma + mu) * I,
                     gamma * I - mu * R])


def mseir(N=1000):
    """C4: examples/mseir.py:22-47."""
    n_deriv, n_prior = [1] * 5, [3] * 5
    theta0 = np.array([1.1, 0.7, 0.4, 0.005, 0.02, 0.03])
    x0v = np.array([1000.0, 100, 50, 3, 3])
    x0 = zero_pad(np.ravel([x0v, _mseir_rhs(x0v, theta0)], "F"), n_deriv, n_prior)
    W = zero_pad(_wmat(n_deriv), n_deriv, n_prior)
    kinit = indep_init(ibm_init(40.0 / N, n_prior, [0.1] * 5), n_prior)
    return _pack("mseir", W, 0, 40, N, kinit, x0, n_theta=6, theta0=theta0)


def mseir_batch(B, N=1000, seed=3):
    w = mseir(N)
    rng = np.random.default_rng(seed)
    theta = w["theta0"] * np.exp(0.1 * rng.standard_normal((B, 6)))
    return w, np.tile(w["x0"], (B, 1)), theta


def make_kode(w, z_state=None, **kw):
    from .cuda import KalmanODE
    return KalmanODE(w["W"], w["tmin"], w["tmax"], w["N"], w["ode"], w["c"], w["Q"], w["R"],
                     z_state=z_state, **kw)


def fitz_observations(kode, w, seed=5, gamma=0.2):
    """Config 5's data: one noisy FitzHugh-Nagumo trajectory observed at the integer times 1..40
    (examples/inference.py:58-64): the posterior mean at theta0 plus N(0, gamma^2) noise, components V and R.
    Returns (Y (40, 2), thin_index (40,), state_ind)."""
    import torch
    ind = np.arange(w["N"] // 40, w["N"] + 1, w["N"] // 40)
    rng = np.random.default_rng(seed)
    mu, _ = kode.solve_mv(torch.from_numpy(w["x0"][None]).to(kode.device),
                          theta=torch.from_numpy(w["theta0"][None]).to(kode.device))
    Y = mu[0].cpu().numpy()[ind][:, w["state_ind"]] + gamma * rng.standard_normal((len(ind), 2))
    return Y, ind, list(w["state_ind"])


def fitz_sweep_theta(total, lo, hi, w=None, seed=0):
    """Rows [lo, hi) of config 5's global theta stream (one default_rng(seed) stream of `total` draws, so the
    slices of all ranks tile the same sweep whatever the world size)."""
    w = w or fitz()
    rng = np.random.default_rng(seed)
    eps = rng.standard_normal((total, 3))[lo:hi]
    return np.exp(np.log(w["theta0"]) + 0.1 * eps)
