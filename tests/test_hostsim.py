"""The CUDA kernel bodies (thread-per-system), executed on the CPU by tests/hostsim,
against the C oracle: catches algebra / indexing / layout errors without a GPU."""
import numpy as np
import pytest

import oracle
from tests import hostsim
from tests.common import load_golden, traj_err, var_err, ode_name

# (golden case, mean tol, var tol, draw tol): floors measured in tests/test_oracle_golden.py --
# the CAR-prior Lorenz case is ill-conditioned (cond R = 2.5e12) and chaotic.
CASES = [("chkrebtii_n200", 1e-10, 1e-8, 1e-5), ("fitz_n400", 1e-10, 1e-9, 1e-9),
         ("lorenz_car_n600", 1e-7, 1e-6, 1e-7), ("mseir_n1000", 1e-10, 1e-8, 1e-7)]


def _setup(name, B=3):
    g = load_golden(name)
    ode = ode_name(name)
    p = oracle.Problem(ode, g["W"], float(g["tmin"]), float(g["tmax"]), int(g["N"]), g["Q"], g["c"], g["R"])
    rng = np.random.default_rng(11)
    x0 = np.tile(g["x0"], (B, 1))
    theta = None
    if "theta" in g:
        th = np.atleast_2d(g["theta"])
        theta = th[np.arange(B) % len(th)] * np.exp(0.02 * rng.standard_normal((B, th.shape[1])))
    return g, ode, p, x0, theta


@pytest.mark.parametrize("blockdiag", [False, True], ids=["dense", "blockdiag"])
@pytest.mark.parametrize("name,mtol,vtol,xtol", CASES)
def test_bodies_match_oracle(name, mtol, vtol, xtol, blockdiag):
    g, ode, p, x0, theta = _setup(name)
    ref = oracle.solve_batch(p, x0, theta, g["z"], oracle.MODE_BOTH)
    args = (ode, g["W"], p.tmin, p.tmax, p.N, g["Q"], g["c"], g["R"], x0, theta, g["z"])
    both = hostsim.solve(*args, kind="both", blockdiag=blockdiag)
    assert not both["status"].any() and not ref["status"].any()
    assert traj_err(both["mu"], ref["mu"]).max() < mtol
    assert var_err(both["var"][:, 1:], ref["var"][:, 1:]).max() < vtol
    assert not both["var"][:, 0].any()
    assert traj_err(both["x"], ref["x"]).max() < xtol
    mv = hostsim.solve(*args, kind="mv", blockdiag=blockdiag)
    sim = hostsim.solve(*args, kind="sim", blockdiag=blockdiag)
    np.testing.assert_array_equal(mv["mu"], both["mu"])
    np.testing.assert_array_equal(mv["var"], both["var"])
    np.testing.assert_array_equal(sim["x"], both["x"])


@pytest.mark.parametrize("blockdiag", [False, True], ids=["dense", "blockdiag"])
def test_per_system_z_and_loglik(blockdiag):
    g, ode, p, x0, theta = _setup("fitz_n400", B=4)
    rng = np.random.default_rng(3)
    z = rng.standard_normal((4, p.n, p.N))
    ref = oracle.solve_batch(p, x0, theta, z, oracle.MODE_BOTH)
    args = (ode, g["W"], p.tmin, p.tmax, p.N, g["Q"], g["c"], g["R"], x0, theta, z)
    sim = hostsim.solve(*args, kind="sim", blockdiag=blockdiag)
    assert traj_err(sim["x"], ref["x"]).max() < 1e-9
    ind, Y, gam = g["thin_ind"], g["Y"], float(g["gamma"])
    for kind, key in (("ll_mean", "mu"), ("ll_draw", "x")):
        r = hostsim.solve(*args, kind=kind, thin_index=ind, state_index=[0, 3], Y=Y, gamma=gam,
                          blockdiag=blockdiag)
        want = [oracle.loglik(ref[key][b], ind, [0, 3], Y, gam) for b in range(4)]
        np.testing.assert_allclose(r["loglik"], want, rtol=1e-9)


@pytest.mark.parametrize("blockdiag", [False, True], ids=["dense", "blockdiag"])
@pytest.mark.parametrize("ig,name", [(1, "schobert"), (2, "chkrebtii")])
def test_interrogation_variants(ig, name, blockdiag):
    g, ode, p, x0, theta = _setup("fitz_n400", B=2)
    po = oracle.Problem(ode, g["W"], p.tmin, p.tmax, p.N, g["Q"], g["c"], g["R"], interrogate=name)
    # schobert (H = 0) leaves Sigma_f singular, so its sampling pass is degenerate (the
    # reference ships it disabled): compare mean/variance only for it.
    mode, kind = (oracle.MODE_MV, "mv") if ig == 1 else (oracle.MODE_BOTH, "both")
    theta = np.tile(g["theta"][0], (2, 1))       # schobert diverges for other draws
    ref = oracle.solve_batch(po, x0, theta, g["z"], mode)
    assert np.isfinite(ref["mu"]).all()
    r = hostsim.solve(ode, g["W"], p.tmin, p.tmax, p.N, g["Q"], g["c"], g["R"], x0, theta, g["z"],
                      kind=kind, ig=ig, blockdiag=blockdiag)
    assert traj_err(r["mu"], ref["mu"]).max() < 1e-9
    if ig == 1:   # measured components have exactly zero variance: compare on the global scale
        assert np.abs(r["var"] - ref["var"]).max() < 1e-9 * np.abs(ref["var"]).max()
    else:
        assert var_err(r["var"][:, 1:], ref["var"][:, 1:]).max() < 1e-8
    if ig == 2:
        assert traj_err(r["x"], ref["x"]).max() < 1e-8


@pytest.mark.parametrize("ck", [2, 3, 4])
@pytest.mark.parametrize("blockdiag", [False, True], ids=["dense", "blockdiag"])
@pytest.mark.parametrize("name,N", [("fitz_n400", None), ("chkrebtii_n200", None), ("fitz_n400", 7), ("fitz_n400", 1),
                                    ("fitz_n400", 2)])
def test_checkpointed_history_gives_identical_results(name, N, blockdiag, ck):
    """Storing every ck-th filtered record and re-running the filter in the backward pass must
    give bit-identical outputs to storing everything (same code, same inputs), for n_eval that
    is and is not a multiple of ck, including the degenerate n_eval = 1, 2."""
    g, ode, p, x0, theta = _setup(name, B=2)
    N = int(g["N"]) if N is None else N
    tmax = p.tmax * N / int(g["N"])
    z = g["z"][:, :N]
    args = (ode, g["W"], p.tmin, tmax, N, g["Q"], g["c"], g["R"], x0, theta, z)
    full = hostsim.solve(*args, kind="both", blockdiag=blockdiag)
    chk = hostsim.solve(*args, kind="both", blockdiag=blockdiag, ck=ck)
    for key in ("mu", "var", "x"):
        np.testing.assert_array_equal(chk[key], full[key])
    if name.startswith("fitz"):
        ind = np.unique(np.clip(np.arange(0, N + 1, max(1, N // 5)), 0, N))
        Y = np.zeros((len(ind), 2))
        a = hostsim.solve(*args, kind="ll_draw", blockdiag=blockdiag, thin_index=ind, state_index=[0, 3], Y=Y, gamma=0.3)
        b = hostsim.solve(*args, kind="ll_draw", blockdiag=blockdiag, ck=ck, thin_index=ind, state_index=[0, 3], Y=Y,
                          gamma=0.3)
        np.testing.assert_array_equal(a["loglik"], b["loglik"])


@pytest.mark.parametrize("blockdiag", [False, True], ids=["dense", "blockdiag"])
@pytest.mark.parametrize("name,mtol,vtol,xtol", CASES)
def test_shared_covariance_bodies_match_oracle(name, mtol, vtol, xtol, blockdiag):
    """Shared-covariance mode (tables once + mean-only recursions) against the oracle, which
    computes every system's covariance itself."""
    if name.startswith("lorenz") or name.startswith("mseir"):
        if not blockdiag:
            pytest.skip("dense register family exists for n_state <= 6 only")
    g, ode, p, x0, theta = _setup(name)
    ref = oracle.solve_batch(p, x0, theta, g["z"], oracle.MODE_BOTH)
    args = (ode, g["W"], p.tmin, p.tmax, p.N, g["Q"], g["c"], g["R"], x0, theta, g["z"])
    both = hostsim.solve(*args, kind="both", blockdiag=blockdiag, shared_cov=True)
    assert not both["status"].any()
    assert traj_err(both["mu"], ref["mu"]).max() < mtol
    assert traj_err(both["x"], ref["x"]).max() < xtol
    assert both["var"].shape[0] == 1
    assert var_err(both["var"][:, 1:], ref["var"][:1, 1:]).max() < vtol
    assert not both["var"][:, 0].any()
    mv = hostsim.solve(*args, kind="mv", blockdiag=blockdiag, shared_cov=True)
    sim = hostsim.solve(*args, kind="sim", blockdiag=blockdiag, shared_cov=True)
    np.testing.assert_array_equal(mv["mu"], both["mu"])
    np.testing.assert_array_equal(sim["x"], both["x"])
    if name.startswith("fitz"):
        ind, Y, gam = g["thin_ind"], g["Y"], float(g["gamma"])
        for kind, key in (("ll_mean", "mu"), ("ll_draw", "x")):
            r = hostsim.solve(*args, kind=kind, blockdiag=blockdiag, shared_cov=True, thin_index=ind,
                              state_index=[0, 3], Y=Y, gamma=gam)
            want = [oracle.loglik(ref[key][b], ind, [0, 3], Y, gam) for b in range(len(x0))]
            np.testing.assert_allclose(r["loglik"], want, rtol=1e-8)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_dense_problems(seed):
    """Random dense priors (nothing block diagonal, nothing triangular, W dense, c nonzero): the
    dense kernel bodies against the oracle, including checkpointed history."""
    rng = np.random.default_rng(seed)
    n, m, N = 6, 2, 60
    A = rng.standard_normal((n, n))
    R = 1e-3 * (A @ A.T + n * np.eye(n))
    Q = np.eye(n) + 0.05 * rng.standard_normal((n, n))
    W = rng.standard_normal((m, n))
    c = 0.01 * rng.standard_normal(n)
    B = 3
    x0 = rng.standard_normal((B, n))
    theta = np.exp(np.log([0.2, 0.2, 3.0]) + 0.1 * rng.standard_normal((B, 3)))
    z = rng.standard_normal((n, N))
    p = oracle.Problem("fitz", W, 0.0, 3.0, N, Q, c, R)
    ref = oracle.solve_batch(p, x0, theta, z, oracle.MODE_BOTH)
    assert not ref["status"].any()
    for ck in (1, 2, 4):
        r = hostsim.solve("fitz", W, 0.0, 3.0, N, Q, c, R, x0, theta, z, kind="both", ck=ck)
        assert not r["status"].any()
        assert traj_err(r["mu"], ref["mu"]).max() < 1e-10
        assert traj_err(r["x"], ref["x"]).max() < 1e-9
        assert np.abs(r["var"] - ref["var"]).max() < 1e-11 * np.abs(ref["var"]).max()
    sc = hostsim.solve("fitz", W, 0.0, 3.0, N, Q, c, R, x0, theta, z, kind="both", shared_cov=True)
    assert traj_err(sc["mu"], ref["mu"]).max() < 1e-10
    assert traj_err(sc["x"], ref["x"]).max() < 1e-9
    assert np.abs(sc["var"][0] - ref["var"][0]).max() < 1e-11 * np.abs(ref["var"]).max()
    # a prior that is not positive definite is reported, not hidden
    bad = hostsim.solve("fitz", W, 0.0, 3.0, N, Q, c, -R, x0, theta, z, kind="mv")
    assert bad["status"].all()


@pytest.mark.parametrize("name", ["chkrebtii_n200", "fitz_n400", "mseir_n1000"])
def test_structured_bodies_are_bit_identical(name):
    """The structured block-diagonal arithmetic (upper-triangular Q blocks, w = e_order: exact zeros skipped,
    kalman_math.inl) against the general block-diagonal bodies on the IBM-prior golden cases: same bits for every
    output, every entry point, checkpoint 1 and 2 -- on the CPU, no GPU needed."""
    g, ode, p, x0, theta = _setup(name)
    args = (ode, g["W"], p.tmin, p.tmax, p.N, g["Q"], g["c"], g["R"], x0, theta, g["z"])
    for ck in (1, 2):
        gen = hostsim.solve(*args, kind="both", blockdiag=1, ck=ck)
        st = hostsim.solve(*args, kind="both", blockdiag=2, ck=ck)
        for key in ("x", "mu", "var", "status"):
            np.testing.assert_array_equal(st[key], gen[key])
    if name == "fitz_n400":
        ind, Y, gam = g["thin_ind"], g["Y"], float(g["gamma"])
        for kind in ("ll_mean", "ll_draw"):
            a = hostsim.solve(*args, kind=kind, blockdiag=1, thin_index=ind, state_index=[0, 3], Y=Y, gamma=gam)
            b = hostsim.solve(*args, kind=kind, blockdiag=2, thin_index=ind, state_index=[0, 3], Y=Y, gamma=gam)
            np.testing.assert_array_equal(a["loglik"], b["loglik"])
