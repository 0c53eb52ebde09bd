"""GPU parity tests: the CUDA path, called through the C ABI (ctypes) by the Python front
end, against (a) the golden vectors of the reference's own Python path, (b) the C oracle on
the same seeded inputs, (c) size-independent properties at BASELINE.json's full sizes.

Tolerances are the trajectory-scaled FP64 noise floors established in
tests/test_oracle_golden.py (north-star: mean/variance rtol 1e-10, draws 1e-9 where the
problem's conditioning allows two FP64 implementations to agree that closely)."""
import numpy as np
import pytest
import torch

import oracle
from tests import workloads as wl
from tests.common import load_golden, traj_err, var_err, ode_name

pytestmark = pytest.mark.gpu


def _kode_from_golden(g, name, **kw):
    from rodeo_legacy_b200.cuda import KalmanODE
    return KalmanODE(np.asfortranarray(g["W"]), float(g["tmin"]), float(g["tmax"]), int(g["N"]),
                     ode_name(name), g["c"], np.asfortranarray(g["Q"]), np.asfortranarray(g["R"]), **kw)


# (case, mean tol, var tol, draw tol, horizon for the tight mean/draw check)
GOLDEN = [("chkrebtii_n100", 1e-10, 1e-8, 1e-5, None), ("chkrebtii_n200", 1e-10, 1e-8, 1e-5, None),
          ("fitz_n400", 1e-10, 1e-10, 1e-9, None), ("lorenz_n5000", 1e-9, 1e-7, 1e-9, 2500),
          ("lorenz_car_n600", 1e-7, 1e-6, 1e-7, None), ("mseir_n1000", 1e-10, 1e-8, 1e-7, None)]


@pytest.mark.parametrize("force_dense", [False, True], ids=["auto", "dense"])
@pytest.mark.parametrize("name,mtol,vtol,xtol,hor", GOLDEN)
def test_golden_vectors_host_and_device_paths(name, mtol, vtol, xtol, hor, force_dense):
    g = load_golden(name)
    k = _kode_from_golden(g, name, z_state=np.asfortranarray(g["z"]), force_dense=force_dense)
    # every golden prior is block diagonal: "auto" must pick that family, "dense" must not
    assert k.kernel_family[0] == ("dense" if force_dense else "blockdiag")
    batched = g["mu"].ndim == 3
    B = g["mu"].shape[0] if batched else 1
    x0 = np.tile(g["x0"], (B, 1)) if batched else np.asfortranarray(g["x0"])
    theta = g["theta"] if "theta" in g else None
    every = int(g["var_every"]) if "var_every" in g else 1
    sl = slice(None) if hor is None else slice(0, hor + 1)
    # host path (NumPy in / NumPy out), all three methods
    x, mu, var = k.solve(x0, theta=theta)
    mu2, var2 = k.solve_mv(x0, theta=theta)
    x2 = k.solve_sim(x0, theta=theta)
    assert x.shape == g["x"].shape and mu.shape == g["mu"].shape
    assert not np.asarray(k.status).any()
    np.testing.assert_array_equal(mu, mu2)
    np.testing.assert_array_equal(var, var2)
    np.testing.assert_array_equal(x, x2)
    vthin = var[..., ::every, :, :]
    assert traj_err(mu[..., sl, :], g["mu"][..., sl, :]).max() < mtol
    assert traj_err(x[..., sl, :], g["x"][..., sl, :]).max() < xtol
    assert var_err(vthin[..., 1:, :, :], g["var"][..., 1:, :, :]).max() < vtol
    assert not var[..., 0, :, :].any()
    if hor is not None:                     # chaotic tail: loose, as between any two FP64 codes
        assert traj_err(mu, g["mu"]).max() < 1e-3 and traj_err(x, g["x"]).max() < 1e-3
    # device path (torch CUDA in / out) gives the same bits
    x0d = torch.from_numpy(np.ascontiguousarray(x0)).cuda()
    thd = None if theta is None else torch.from_numpy(np.ascontiguousarray(theta)).cuda()
    xd, mud, vard = k.solve(x0d, theta=thd)
    assert xd.is_cuda and mud.shape == mu.shape
    np.testing.assert_array_equal(mud.cpu().numpy(), mu)
    np.testing.assert_array_equal(vard.cpu().numpy(), var)
    np.testing.assert_array_equal(xd.cpu().numpy(), x)


@pytest.mark.parametrize("force_dense", [False, True], ids=["auto", "dense"])
@pytest.mark.parametrize("maker,B,mtol,vtol,xtol", [
    (wl.fitz_batch, 1500, 1e-10, 1e-10, 1e-9),
    (lambda B: wl.mseir_batch(B, N=200), 96, 1e-10, 1e-8, 1e-7),
    (lambda B: wl.lorenz_batch(B, N=500, tmax=2.0), 130, 1e-9, 1e-7, 1e-8)])
def test_seeded_batch_against_oracle(maker, B, mtol, vtol, xtol, force_dense):
    w, x0, theta = maker(B)
    rng = np.random.default_rng(42)
    z = np.asfortranarray(rng.standard_normal((len(w["x0"]), w["N"])))
    k = wl.make_kode(w, z_state=z, force_dense=force_dense)
    ref = oracle.solve_batch(wl.make_oracle_problem(w), x0, theta, z, oracle.MODE_BOTH)
    x, mu, var = k.solve(torch.from_numpy(x0).cuda(), theta=torch.from_numpy(theta).cuda())
    x, mu, var = x.cpu().numpy(), mu.cpu().numpy(), var.cpu().numpy()
    assert not k.status.any().item() and not ref["status"].any()
    assert traj_err(mu, ref["mu"]).max() < mtol
    assert var_err(var[:, 1:], ref["var"][:, 1:]).max() < vtol
    assert traj_err(x, ref["x"]).max() < xtol
    np.testing.assert_array_equal(var, np.swapaxes(var, -1, -2))     # exactly symmetric


@pytest.mark.parametrize("force_dense", [False, True], ids=["auto", "dense"])
def test_per_system_draws_and_ragged_chunks(force_dense):
    """z per system; batch not a multiple of anything; workspace limit forcing 3 chunks."""
    w, x0, theta = wl.fitz_batch(2500, seed=9)
    rng = np.random.default_rng(1)
    z = rng.standard_normal((2500, 6, 400))
    k = wl.make_kode(w, z_state=z, force_dense=force_dense)
    k.set_workspace_limit(1024 * k.history_bytes_per_system)          # 1024 systems per chunk
    x = k.solve_sim(torch.from_numpy(x0).cuda(), theta=torch.from_numpy(theta).cuda()).cpu().numpy()
    assert k.launch_count == 6
    pick = [0, 1, 1023, 1024, 2047, 2048, 2499]
    ref = oracle.solve_batch(wl.make_oracle_problem(w), x0[pick], theta[pick], z[pick], oracle.MODE_SIM)
    assert traj_err(x[pick], ref["x"]).max() < 1e-9
    # host path, chunked the same way, same bits
    xh = k.solve_sim(x0, theta=theta)
    np.testing.assert_array_equal(xh, x)


def test_per_system_draws_wide_state():
    """z per system through the warp-per-block smoother (Lorenz63, 3 blocks), partial last tile."""
    w, x0, theta = wl.lorenz_batch(75, N=300, tmax=1.2)
    z = np.random.default_rng(3).standard_normal((75, 9, 300))
    k = wl.make_kode(w, z_state=z)
    assert k.kernel_family == ("blockdiag", "registers")
    x, mu, var = k.solve(torch.from_numpy(x0).cuda(), theta=torch.from_numpy(theta).cuda())
    ref = oracle.solve_batch(wl.make_oracle_problem(w), x0, theta, z, oracle.MODE_BOTH)
    assert traj_err(x.cpu().numpy(), ref["x"]).max() < 1e-8
    assert traj_err(mu.cpu().numpy(), ref["mu"]).max() < 1e-9
    assert var_err(var.cpu().numpy()[:, 1:], ref["var"][:, 1:]).max() < 1e-7
    xh = k.solve_sim(x0, theta=theta)
    np.testing.assert_array_equal(xh, x.cpu().numpy())


@pytest.mark.parametrize("force_dense", [False, True], ids=["auto", "dense"])
def test_fitz_full_size_properties(force_dense):
    """BASELINE config 2 (65,536 theta draws, n_eval 400, solve_mv) through properties that
    do not need the oracle at full size, plus an oracle spot-check of 48 systems."""
    B = 65536
    w, x0, theta = wl.fitz_batch(B)
    k = wl.make_kode(w, force_dense=force_dense)
    x0d, thd = torch.from_numpy(x0).cuda(), torch.from_numpy(theta).cuda()
    mu, var = k.solve_mv(x0d, theta=thd)
    torch.cuda.synchronize()
    assert mu.shape == (B, 401, 6) and var.shape == (B, 401, 6, 6)
    assert int(k.status.abs().sum()) == 0
    assert bool(torch.isfinite(mu).all()) and bool(torch.isfinite(var).all())
    # F5: the covariance path does not depend on theta -> identical bits across the batch
    assert bool((var == var[0:1]).all())
    assert bool((mu[:, 0] == x0d).all())
    # idempotence / determinism: further solves (same and fresh handles) give the same bits.
    # Repeated on purpose: this is the test that caught a 1-in-10^4-tiles WAR race between late
    # shared-memory reads of a ring slot and its TMA refill (fixed with fence.proxy.async).
    for rep in range(6):
        kk = k if rep % 2 == 0 else wl.make_kode(w, force_dense=force_dense)
        mu2, var2 = kk.solve_mv(x0d, theta=thd)
        assert bool((mu2 == mu).all()) and bool((var2 == var).all()), rep
        del mu2, var2
    # permutation equivariance: solving a shuffled batch shuffles the answers
    perm = torch.randperm(B, device="cuda", generator=torch.Generator("cuda").manual_seed(0))
    mu3, _ = k.solve_mv(x0d[perm], theta=thd[perm])
    assert bool((mu3 == mu[perm]).all())
    del mu3
    # the oracle on 4,096 systems: the first tiles, a block in the middle, the last tiles (north-star: 1e-10)
    pick = np.r_[0:1024, B // 2 - 17:B // 2 - 17 + 2048, B - 1024:B]
    ref = oracle.solve_batch(wl.make_oracle_problem(w), x0[pick], theta[pick], None, oracle.MODE_MV)
    pk = torch.from_numpy(pick).cuda()
    assert traj_err(mu[pk].cpu().numpy(), ref["mu"]).max() < 1e-10
    assert var_err(var[pk].cpu().numpy()[:, 1:], ref["var"][:, 1:]).max() < 1e-10
    if not force_dense:      # the two families agree with each other to rounding as well
        kd = wl.make_kode(w, force_dense=True)
        mud, vard = kd.solve_mv(x0d[:4096], theta=thd[:4096])
        assert float((mud - mu[:4096]).abs().max()) < 1e-10
        assert float((vard - var[:4096]).abs().max()) < 1e-12 * float(var.abs().max()) + 1e-22
    # smoothed variance never exceeds the filtered one at the end point, and is PSD-diagonal
    assert bool((torch.diagonal(var[0], dim1=-2, dim2=-1) >= 0).all())


@pytest.mark.parametrize("maker,B,mode", [
    (lambda B: wl.lorenz_batch(B), 16384, "sim"),
    (lambda B: wl.mseir_batch(B), 32768, "mv")])
def test_wide_state_configs_at_full_size(maker, B, mode):
    """BASELINE configs 3 (Lorenz63, 16,384 x0, n_eval 5000, solve_sim) and 4 (MSEIR, 32,768
    theta, n_eval 1000, solve_mv) at FULL size, with an oracle spot-check and the
    size-independent properties (finite, status clean, covariance theta-invariant,
    permutation equivariance)."""
    w, x0, theta = maker(B)
    n = len(w["x0"])
    np.random.seed(2)
    z = np.asfortranarray(np.random.randn(n, w["N"]))
    k = wl.make_kode(w, z_state=z)
    x0d, thd = torch.from_numpy(x0).cuda(), torch.from_numpy(theta).cuda()
    # oracle coverage: 256 (Lorenz63) / 512 (MSEIR) systems from the first, a middle and the last tiles
    q = 64 if mode == "sim" else 128
    pick = np.r_[0:q, B // 2 - 5:B // 2 - 5 + 2 * q, B - q:B]
    po = wl.make_oracle_problem(w)
    if mode == "sim":
        x = k.solve_sim(x0d, theta=thd)
        assert bool(torch.isfinite(x).all()) and int(k.status.abs().sum()) == 0
        ref = oracle.solve_batch(po, x0[pick], theta[pick], z, oracle.MODE_SIM)
        got = x[torch.from_numpy(pick).cuda()].cpu().numpy()
        # Lorenz63 amplifies rounding differences at its Lyapunov rate (e^{0.9 t}: x 8,000 by t = 10): tight over
        # the first 1,000 steps, 2e-8 at 2,500 (worst of these 256 systems: 6e-9), loose at the end of the horizon;
        # per-step parity itself is pinned at 1e-12 by test_lorenz_one_step_consistency
        assert traj_err(got[:, :1001], ref["x"][:, :1001]).max() < 1e-11
        assert traj_err(got[:, :2501], ref["x"][:, :2501]).max() < 2e-8
        assert traj_err(got, ref["x"]).max() < 1e-3
        perm = torch.arange(B - 1, -1, -1, device="cuda")
        x = x[perm]
        assert bool((k.solve_sim(x0d[perm], theta=thd[perm]) == x).all())
    else:
        mu, var = k.solve_mv(x0d, theta=thd)          # 63 GB of variance output at full size
        assert bool(torch.isfinite(mu).all()) and int(k.status.abs().sum()) == 0
        for lo in range(0, B, 4096):                  # chunked compare: no 63 GB temporaries
            assert bool((var[lo:lo + 4096] == var[0:1]).all())
        ref = oracle.solve_batch(po, x0[pick], theta[pick], None, oracle.MODE_MV)
        pk = torch.from_numpy(pick).cuda()
        assert traj_err(mu[pk].cpu().numpy(), ref["mu"]).max() < 1e-10
        # cond(R) ~ 1e10 at dt = 0.04: two FP64 implementations agree to 1.4e-9 here
        assert var_err(var[pk].cpu().numpy()[:, 1:], ref["var"][:, 1:]).max() < 1e-8


@pytest.mark.parametrize("maker,B,eps,reps", [
    (lambda B: wl.lorenz_batch(B, N=1000, tmax=4.0), 255, 1e-6, 1e-7),   # chaotic: a larger R perturbation diverges
    (lambda B: wl.mseir_batch(B, N=250), 256, 1e-3, 1e-3)])
def test_wide_state_dense_family(maker, B, eps, reps):
    """Dense kernels for n_state 9 and 15 (half a warp per system, kalman_dw.cuh): any Q / R / W.
    All three entry points, an odd batch (half-filled last warp), and a genuinely dense prior."""
    w, x0, theta = maker(B)
    n = len(w["x0"])
    rng = np.random.default_rng(8)
    z = np.asfortranarray(rng.standard_normal((n, w["N"])))
    k = wl.make_kode(w, z_state=z, force_dense=True)
    assert k.kernel_family == ("dense", "wide")
    x0d, thd = torch.from_numpy(x0).cuda(), torch.from_numpy(theta).cuda()
    pick = [0, 100, B - 1]
    po = wl.make_oracle_problem(w)
    x = k.solve_sim(x0d, theta=thd)
    mu, var = k.solve_mv(x0d, theta=thd)
    x2, mu2, var2 = k.solve(x0d, theta=thd)
    ref = oracle.solve_batch(po, x0[pick], theta[pick], z, oracle.MODE_BOTH)
    assert traj_err(x[pick].cpu().numpy(), ref["x"]).max() < 1e-8
    assert traj_err(mu[pick].cpu().numpy(), ref["mu"]).max() < 1e-10
    assert var_err(var[pick].cpu().numpy()[:, 1:], ref["var"][:, 1:]).max() < 1e-8
    assert torch.equal(x, x2) and torch.equal(mu, mu2) and torch.equal(var, var2)
    assert bool((var[:, 0] == 0).all()) and torch.equal(mu[:, 0], x0d)
    # a prior with no structure at all (dense Q, SPD R, W touching every state)
    A = rng.standard_normal((n, n))
    wd = dict(w)
    wd["R"] = np.asfortranarray(w["R"] + reps * np.abs(w["R"]).max() * (A @ A.T) / n)
    wd["Q"] = np.asfortranarray(w["Q"] + eps * rng.standard_normal((n, n)))
    wd["W"] = np.asfortranarray(w["W"] + 10 * eps * rng.standard_normal(w["W"].shape))
    wd["c"] = eps * rng.standard_normal(n)
    kd = wl.make_kode(wd, z_state=z)
    assert kd.kernel_family == ("dense", "wide")
    xd, mud, vard = kd.solve(x0d[:64], theta=thd[:64])
    refd = oracle.solve_batch(wl.make_oracle_problem(wd), x0[:64], theta[:64], z, oracle.MODE_BOTH)
    assert np.isfinite(refd["mu"]).all() and np.isfinite(refd["x"]).all()
    # the perturbed problem is worse conditioned; tolerances = 5x what two independent FP64 codes
    # (oracle vs the thread-per-system kernels) differ by on it
    assert traj_err(mud.cpu().numpy(), refd["mu"]).max() < 2e-9
    assert np.abs(vard.cpu().numpy() - refd["var"]).max() < 1e-8 * np.abs(refd["var"]).max()
    assert traj_err(xd.cpu().numpy(), refd["x"]).max() < 1e-7
    np.testing.assert_array_equal(vard.cpu().numpy(), np.swapaxes(vard.cpu().numpy(), -1, -2))
    # per-system priors on the same family: each of 20 systems its own (dense) Q, R, W, c
    nb = 20
    m_ = w["W"].shape[0]
    Qb = np.stack([np.asarray(w["Q"]) + 0.1 * eps * rng.standard_normal((n, n)) for _ in range(nb)])
    Rb = np.stack([np.asarray(w["R"]) * (1 + 0.02 * i) + 0.1 * reps * np.abs(w["R"]).max() * (A @ A.T) / n for i in range(nb)])
    Wb = np.stack([np.asarray(w["W"]) + eps * rng.standard_normal((m_, n)) for _ in range(nb)])
    cb = 0.1 * eps * rng.standard_normal((nb, n))
    xp, mup_, varp = k.solve(x0d[:nb], theta=thd[:nb], priors=dict(wgt_state=Qb, var_state=Rb, wgt_meas=Wb, mu_state=cb))
    for b in (0, 7, nb - 1):
        po_b = oracle.Problem(w["ode"], Wb[b], w["tmin"], w["tmax"], w["N"], Qb[b], cb[b], Rb[b])
        rb = oracle.solve(po_b, x0[b], theta[b], z, oracle.MODE_BOTH)
        assert traj_err(mup_[b].cpu().numpy(), rb["mu"]).max() < 2e-9
        assert np.abs(varp[b].cpu().numpy() - rb["var"]).max() < 1e-8 * np.abs(rb["var"]).max()
        assert traj_err(xp[b].cpu().numpy(), rb["x"]).max() < 1e-7
    # fused log-likelihood epilogue of the same family: equals the density of the returned mean / draw
    ind = np.arange(10, w["N"] + 1, 10)
    comp = [0, n - 1, 0]                     # a repeated component is legal
    Y = mu[0, ind][:, comp].cpu().numpy() + 0.1
    cst = -Y.size * (np.log(0.3) + 0.5 * np.log(2 * np.pi))
    for draw, traj in ((False, mu), (True, x)):
        got = k.loglik(x0d, thd, Y, ind, comp, 0.3, use_draw=draw).cpu().numpy()
        tc = traj[:, ind][:, :, comp].cpu().numpy()
        want = cst - 0.5 * (((Y[None] - tc) / 0.3) ** 2).sum(axis=(1, 2))
        assert got.shape == (B,) and np.abs(got - want).max() < 1e-9 * np.abs(want).max()


def test_dense_prior_uses_dense_family():
    """A prior that is NOT block diagonal (random SPD R, dense Q) must run on the dense
    kernels automatically and still match the oracle."""
    w, x0, theta = wl.fitz_batch(300, seed=2)
    rng = np.random.default_rng(12)
    A = rng.standard_normal((6, 6))
    w = dict(w)
    w["R"] = np.asfortranarray(w["R"] + 1e-9 * (A @ A.T))
    w["Q"] = np.asfortranarray(w["Q"] + 1e-3 * rng.standard_normal((6, 6)))
    w["W"] = np.asfortranarray(w["W"] + 1e-2 * rng.standard_normal((2, 6)))
    w["c"] = 1e-3 * rng.standard_normal(6)
    z = np.asfortranarray(rng.standard_normal((6, 400)))
    k = wl.make_kode(w, z_state=z)
    assert k.kernel_family == ("dense", "registers")
    ref = oracle.solve_batch(wl.make_oracle_problem(w), x0, theta, z, oracle.MODE_BOTH)
    x, mu, var = k.solve(torch.from_numpy(x0).cuda(), theta=torch.from_numpy(theta).cuda())
    assert traj_err(mu.cpu().numpy(), ref["mu"]).max() < 1e-9
    assert var_err(var.cpu().numpy()[:, 1:], ref["var"][:, 1:]).max() < 1e-8
    assert traj_err(x.cpu().numpy(), ref["x"]).max() < 1e-8
    # setting a structured prior back re-selects the block-diagonal family
    w0 = wl.fitz()
    k.wgt_state, k.var_state, k.wgt_meas, k.mu_state = w0["Q"], w0["R"], w0["W"], w0["c"]
    assert k.kernel_family[0] == "blockdiag"


def test_loglik_matches_oracle_and_sharded_sweep_layout():
    g = load_golden("fitz_n400")
    w, x0, theta = wl.fitz_batch(3000, seed=4)
    z = np.asfortranarray(g["z"])
    k = wl.make_kode(w, z_state=z)
    ind, Y, gam = g["thin_ind"], g["Y"], float(g["gamma"])
    x0d, thd = torch.from_numpy(x0).cuda(), torch.from_numpy(theta).cuda()
    ll_mean = k.loglik(x0d, thd, Y, ind, [0, 3], gam)
    ll_draw = k.loglik(x0d, thd, Y, ind, [0, 3], gam, use_draw=True)
    pick = [0, 1, 999, 2999]
    ref = oracle.solve_batch(wl.make_oracle_problem(w), x0[pick], theta[pick], z, oracle.MODE_BOTH)
    for got, key in ((ll_mean, "mu"), (ll_draw, "x")):
        want = [oracle.loglik(ref[key][i], ind, [0, 3], Y, gam) for i in range(len(pick))]
        np.testing.assert_allclose(got[pick].cpu().numpy(), want, rtol=1e-9)
    # golden: the reference's own thinning + loglike on its own solver output
    ll_g = k.loglik(np.tile(g["x0"], (4, 1)), g["theta"], Y, ind, [0, 3], gam)
    np.testing.assert_allclose(ll_g, g["loglik_mu"], rtol=1e-8)
    # equals the likelihood of the materialised mean
    mu, _ = k.solve_mv(x0d[:64], theta=thd[:64])
    manual = [oracle.loglik(m, ind, [0, 3], Y, gam) for m in mu.cpu().numpy()]
    np.testing.assert_allclose(ll_mean[:64].cpu().numpy(), manual, rtol=1e-12)


def test_reference_unit_test_pattern():
    """tests/test_kalmanode.py:13-66 and tests/test_deterministic.py:11-46 of the reference,
    with rodeo.cuda in place of the compiled backends."""
    from scipy import integrate
    rel_err = lambda a, b: np.sum(np.abs((a.ravel() - b.ravel()) / a.ravel()))
    w = wl.chkrebtii(100)
    np.random.seed(0)
    from rodeo_legacy_b200 import rand_mat
    z = rand_mat(100, 4)
    k = wl.make_kode(w)
    k.z_state = z
    ksim = k.solve_sim(w["x0"])
    ref = oracle.solve(wl.make_oracle_problem(w), w["x0"], None, z, oracle.MODE_SIM)["x"]
    assert ksim.shape == (101, 4)
    assert rel_err(ksim[:, 0], ref[:, 0]) <= 0.001
    assert rel_err(ksim[1:, 1], ref[1:, 1]) <= 0.001
    w3 = wl.chkrebtii(300)
    k3 = wl.make_kode(w3)
    sim = k3.solve_sim(w3["x0"])            # z auto-filled from the global RNG, like the reference
    assert np.any(k3.z_state)
    det = integrate.odeint(lambda x, t: [x[1], np.sin(2 * t) - x[0]], [-1, 0], np.linspace(0, 10, 301))
    assert rel_err(sim[:, 0], det[:, 0]) <= 10.0
    assert rel_err(sim[1:, 1], det[1:, 1]) <= 10.0
    del k3.z_state
    assert not np.any(k3.z_state)
    mu, var = k3.solve_mv(w3["x0"])          # solve_mv must not touch the RNG / z
    assert not np.any(k3.z_state) and mu.shape == (301, 4) and var.shape == (301, 4, 4)


def test_error_behaviour():
    w = wl.fitz()
    k = wl.make_kode(w)
    with pytest.raises(TypeError, match="incorrect shape"):
        k.solve_mv(np.zeros(5), theta=w["theta0"])
    with pytest.raises(TypeError, match="not f contiguous"):
        k.wgt_state = np.ascontiguousarray(np.arange(36.0).reshape(6, 6))
    with pytest.raises(TypeError, match="incorrect shape"):
        k.var_state = np.zeros((5, 5), order="F")
    with pytest.raises(TypeError, match="theta"):
        k.solve_mv(w["x0"])
    from rodeo_legacy_b200.cuda import KalmanODE
    with pytest.raises(NotImplementedError, match="no compiled kernel"):
        KalmanODE(np.zeros((2, 14), order="F"), 0, 1, 10, "fitz", np.zeros(14), np.eye(14), np.eye(14))
    with pytest.raises(TypeError, match="registered device functor"):
        KalmanODE(w["W"], 0, 40, 400, lambda *a: None, w["c"], w["Q"], w["R"])
    # a non positive-definite prior is reported per system, not raised
    kb = wl.make_kode(w)
    kb.var_state = np.asfortranarray(-np.eye(6))
    kb.solve_mv(w["x0"], theta=w["theta0"])
    assert np.asarray(kb.status).any()
    # property setter takes effect: W override per call (KalmanODE.pyx:441-442)
    mu_a, _ = k.solve_mv(w["x0"], theta=w["theta0"])
    mu_b, _ = k.solve_mv(w["x0"], W=w["W"], theta=w["theta0"])
    np.testing.assert_array_equal(mu_a, mu_b)
    # C ABI: device outputs must be 16-byte aligned (vector / TMA stores) -- refused, not faulted
    import ctypes
    from rodeo_legacy_b200.cuda import _lib
    lib = _lib.load()
    buf = torch.zeros(2 * 401 * 42 + 8, dtype=torch.float64, device="cuda")
    x0d = torch.from_numpy(np.tile(w["x0"], (2, 1))).cuda()
    thd = torch.from_numpy(np.tile(w["theta0"], (2, 1))).cuda()
    st = torch.zeros(2, dtype=torch.int32, device="cuda")
    mu_p, var_p = buf.data_ptr() + 8, buf.data_ptr() + 8 + 2 * 401 * 6 * 8
    rc = lib.rodeo_cuda_solve(k._h, 1, 2, ctypes.c_void_p(x0d.data_ptr()), ctypes.c_void_p(thd.data_ptr()), None, 0,
                              None, ctypes.c_void_p(mu_p), ctypes.c_void_p(var_p), ctypes.c_void_p(st.data_ptr()), None)
    assert rc == -1 and b"16-byte aligned" in lib.rodeo_cuda_last_error()
    # empty batch
    mu_e, var_e = k.solve_mv(np.zeros((0, 6)), theta=np.zeros((0, 3)))
    assert mu_e.shape == (0, 401, 6) and var_e.shape == (0, 401, 6, 6)


@pytest.mark.parametrize("force_dense", [False, True], ids=["auto", "dense"])
@pytest.mark.parametrize("name", ["schobert", "chkrebtii"])
def test_interrogation_variants_on_gpu(name, force_dense):
    """The other two interrogation methods of the reference (forecast_sch / forecast,
    rodeo/cython/KalmanODE.pyx:13-67; selected there by commenting code) against the oracle."""
    g = load_golden("fitz_n400")
    w = wl.fitz()
    B = 64
    theta = np.tile(g["theta"][0], (B, 1))          # schobert diverges for other draws
    x0 = np.tile(w["x0"], (B, 1))
    if name == "chkrebtii":     # (schobert amplifies rounding 1e4-fold once x0 moves: keep it fixed there)
        x0 = x0 * (1 + 1e-3 * np.random.default_rng(3).standard_normal((B, 1)))
    z = np.asfortranarray(g["z"])
    k = wl.make_kode(w, z_state=z, interrogate=name, force_dense=force_dense)
    po = wl.make_oracle_problem(w, interrogate=name)
    x0d, thd = torch.from_numpy(x0).cuda(), torch.from_numpy(theta).cuda()
    if name == "schobert":      # H = 0 leaves Sigma_f singular: the sampling pass is degenerate
        ref = oracle.solve_batch(po, x0, theta, z, oracle.MODE_MV)
        mu, var = (a.cpu().numpy() for a in k.solve_mv(x0d, theta=thd))
        assert np.isfinite(ref["mu"]).all()
        assert traj_err(mu, ref["mu"]).max() < 1e-8
        assert np.abs(var - ref["var"]).max() < 1e-8 * np.abs(ref["var"]).max()
    else:
        ref = oracle.solve_batch(po, x0, theta, z, oracle.MODE_BOTH)
        x, mu, var = (a.cpu().numpy() for a in k.solve(x0d, theta=thd))
        assert traj_err(mu, ref["mu"]).max() < 1e-9
        assert traj_err(x, ref["x"]).max() < 1e-8
        assert var_err(var[:, 1:], ref["var"][:, 1:]).max() < 1e-8


@pytest.mark.parametrize("ck", [2, 4])
@pytest.mark.parametrize("maker,B,N", [
    (wl.fitz_batch, 1000, 400), (wl.fitz_batch, 70, 7), (wl.fitz_batch, 40, 2), (wl.fitz_batch, 33, 1),
    (lambda B: wl.lorenz_batch(B, N=301, tmax=1.2), 100, 301),
    (lambda B: wl.mseir_batch(B, N=150), 64, 150)])
def test_checkpointed_history_is_bit_identical(maker, B, N, ck):
    """checkpoint = k stores every k-th filtered record and re-runs the filter inside the
    smoother: outputs must equal the checkpoint = 1 outputs bit for bit (solve, loglik), for
    n_eval that is / is not a multiple of k, on both register families."""
    w, x0, theta = maker(B)
    if w["N"] != N:                      # shorten the FitzHugh-Nagumo grid
        w = dict(w, N=N, tmax=w["tmax"] * N / w["N"])
    n = len(w["x0"])
    z = np.asfortranarray(np.random.default_rng(4).standard_normal((n, N)))
    x0d, thd = torch.from_numpy(x0).cuda(), torch.from_numpy(theta).cuda()
    for force_dense in ([False, True] if n <= 6 else [False]):
        k1 = wl.make_kode(w, z_state=z, force_dense=force_dense, checkpoint=1)
        kk = wl.make_kode(w, z_state=z, force_dense=force_dense, checkpoint=ck)
        assert k1.checkpoint == 1
        if n > 6:       # warp-per-block kernels keep the full history (kalman_bdw.cuh)
            assert kk.checkpoint == 1
            continue
        assert kk.checkpoint == ck
        assert wl.make_kode(w, force_dense=force_dense).checkpoint == (1 if force_dense else 2)   # auto
        assert kk.history_bytes_per_system < k1.history_bytes_per_system or N <= 2
        for a, b in zip(k1.solve(x0d, theta=thd), kk.solve(x0d, theta=thd)):
            assert bool((a == b).all())
        mu1, var1 = k1.solve_mv(x0d, theta=thd)
        muk, vark = kk.solve_mv(x0d, theta=thd)
        assert bool((mu1 == muk).all()) and bool((var1 == vark).all())
        assert bool((k1.solve_sim(x0d, theta=thd) == kk.solve_sim(x0d, theta=thd)).all())
        if n == 6:
            ind = np.unique(np.clip(np.arange(0, N + 1, max(1, N // 7)), 0, N))
            Y = np.random.default_rng(1).standard_normal((len(ind), 2))
            for draw in (False, True):
                l1 = k1.loglik(x0d, thd, Y, ind, [0, 3], 0.3, use_draw=draw)
                lk = kk.loglik(x0d, thd, Y, ind, [0, 3], 0.3, use_draw=draw)
                assert bool((l1 == lk).all())
    # other interrogation methods silently keep the full history
    ks = wl.make_kode(w, z_state=z, interrogate="schobert", checkpoint=ck)
    assert ks.checkpoint == 1


@pytest.mark.parametrize("maker,B,mtol,vtol,xtol,dense_too", [
    (wl.fitz_batch, 3000, 1e-10, 1e-10, 1e-9, True),
    (lambda B: wl.lorenz_batch(B, N=500, tmax=2.0), 130, 1e-9, 1e-7, 1e-8, False),
    (lambda B: wl.mseir_batch(B, N=200), 96, 1e-10, 1e-8, 1e-7, False),
    # n_eval + 1 even / a multiple of 4: the blocked output runs take their general path; ragged last tile
    (lambda B: wl.lorenz_batch(B, N=501, tmax=2.0), 45, 1e-9, 1e-7, 1e-8, False),
    (lambda B: wl.mseir_batch(B, N=203), 33, 1e-10, 1e-8, 1e-7, False)])
def test_shared_covariance_mode(maker, B, mtol, vtol, xtol, dense_too):
    """shared_cov=True: covariance path computed once per prior, batch kernels run the mean
    recursions only.  Against the oracle (which computes every system's covariance) and against
    this backend's own general mode; var comes back as a broadcast view of one array."""
    w, x0, theta = maker(B)
    n = len(w["x0"])
    z = np.asfortranarray(np.random.default_rng(6).standard_normal((n, w["N"])))
    ref = oracle.solve_batch(wl.make_oracle_problem(w), x0, theta, z, oracle.MODE_BOTH)
    x0d, thd = torch.from_numpy(x0).cuda(), torch.from_numpy(theta).cuda()
    for force_dense in ([False, True] if dense_too else [False]):
        k = wl.make_kode(w, z_state=z, force_dense=force_dense, shared_cov=True)
        assert k.shared_cov
        n0 = k.launch_count
        x, mu, var = k.solve(x0d, theta=thd)
        assert k.launch_count == n0 + 3          # tables (once) + mean filter + mean smoother
        assert var.shape == (B, w["N"] + 1, n, n) and var.stride(0) == 0
        assert int(k.status.abs().sum()) == 0
        assert traj_err(mu.cpu().numpy(), ref["mu"]).max() < mtol
        assert traj_err(x.cpu().numpy(), ref["x"]).max() < xtol
        assert var_err(var[0].cpu().numpy()[1:], ref["var"][0, 1:]).max() < vtol
        assert not bool(var[0, 0].any())
        mu2, var2 = k.solve_mv(x0d, theta=thd)
        assert k.launch_count == n0 + 5          # tables are cached
        assert bool((mu2 == mu).all())
        assert bool((k.solve_sim(x0d, theta=thd) == x).all())
        # general mode of the same backend
        kg = wl.make_kode(w, z_state=z, force_dense=force_dense)
        xg, mug, varg = kg.solve(x0d, theta=thd)
        scale = float(mug.abs().amax())
        assert float((mug - mu).abs().max()) < 1e-9 * scale and float((xg - x).abs().max()) < 1e-7 * scale
        # host path: one (N+1, n, n) variance crosses PCIe, returned as a broadcast view
        xh, muh, varh = k.solve(x0, theta=theta)
        np.testing.assert_array_equal(muh, mu.cpu().numpy())
        np.testing.assert_array_equal(xh, x.cpu().numpy())
        assert varh.shape == (B, w["N"] + 1, n, n) and varh.strides[0] == 0
        np.testing.assert_array_equal(varh[0], var[0].cpu().numpy())
        # changing the prior invalidates the tables
        k.var_state = np.asfortranarray(4.0 * np.asarray(w["R"]))
        _, var4 = k.solve_mv(x0d, theta=thd)
        assert float(var4[0, -1, 0, 0]) > 1.5 * float(var[0, -1, 0, 0])
    if n == 6:
        g = load_golden("fitz_n400")
        ks = wl.make_kode(w, z_state=z, shared_cov=True)
        ind, Y, gam = g["thin_ind"], g["Y"], float(g["gamma"])
        for draw, key in ((False, "mu"), (True, "x")):
            ll = ks.loglik(x0d, thd, Y, ind, [0, 3], gam, use_draw=draw).cpu().numpy()
            want = [oracle.loglik(ref[key][i], ind, [0, 3], Y, gam) for i in range(0, B, 97)]
            np.testing.assert_allclose(ll[::97], want, rtol=1e-8)
    with pytest.raises(NotImplementedError):
        wl.make_kode(w, z_state=z, interrogate="chkrebtii", shared_cov=True)


def test_per_system_priors():
    """every system brings its own prior (here: a sweep over the IBM scale sigma
    and a perturbed W / c per system).  Against the oracle solving each system with its own prior."""
    from rodeo_legacy_b200 import ibm_init, indep_init
    w, x0, theta = wl.fitz_batch(40, seed=5)
    B, n, N = 40, 6, w["N"]
    rng = np.random.default_rng(2)
    sig = 0.05 + 0.2 * rng.random((B, 2))
    Q = np.empty((B, n, n)); R = np.empty((B, n, n)); W = np.empty((B, 2, n)); c = np.empty((B, n))
    for b in range(B):
        kin = indep_init(ibm_init(40.0 / N, [3, 3], list(sig[b])), [3, 3])
        Q[b], R[b] = kin["wgt_state"], kin["var_state"]
        W[b] = w["W"] + (1e-3 * rng.standard_normal((2, n)) if b % 2 else 0.0)
        c[b] = 1e-4 * rng.standard_normal(n) if b % 3 == 0 else 0.0
    z = np.asfortranarray(rng.standard_normal((n, N)))
    k = wl.make_kode(w, z_state=z)
    x0d, thd = torch.from_numpy(x0).cuda(), torch.from_numpy(theta).cuda()
    pri = dict(wgt_state=Q, var_state=R, wgt_meas=W, mu_state=c)
    x, mu, var = k.solve(x0d, theta=thd, priors=pri)
    assert int(k.status.abs().sum()) == 0
    for b in range(B):
        po = oracle.Problem("fitz", W[b], w["tmin"], w["tmax"], N, Q[b], c[b], R[b])
        ref = oracle.solve(po, x0[b], theta[b], z, oracle.MODE_BOTH)
        assert traj_err(mu[b].cpu().numpy(), ref["mu"]).max() < 1e-9
        assert traj_err(x[b].cpu().numpy(), ref["x"]).max() < 1e-7
        # the perturbed W couples the two blocks by ~1e-14-sized entries: compare on the matrix scale
        vb = var[b].cpu().numpy()
        assert np.abs(vb - ref["var"]).max() < 1e-9 * np.abs(ref["var"]).max()
        assert var_err(np.diagonal(vb, axis1=-2, axis2=-1)[1:, None, :], np.diagonal(ref["var"], axis1=-2, axis2=-1)[1:, None, :]).max() < 1e-8
    # only some fields batched; NumPy in -> NumPy out; equals the shared-prior solve when the
    # batched prior repeats the shared one
    mu_s, var_s = k.solve_mv(x0, theta=theta)
    mu_p, var_p = k.solve_mv(x0, theta=theta, priors=dict(var_state=np.tile(np.asarray(w["R"]), (B, 1, 1))))
    assert isinstance(mu_p, np.ndarray)
    assert np.abs(mu_p - mu_s).max() < 1e-10 * np.abs(mu_s).max()
    assert np.abs(var_p - var_s).max() < 1e-10 * np.abs(var_s).max()
    with pytest.raises(TypeError, match="incorrect shape"):
        k.solve_mv(x0, theta=theta, priors=dict(var_state=np.zeros((B, 5, 5))))
    # wide state: dense rolled family
    w9, x09, th9 = wl.lorenz_batch(24, N=200, tmax=0.8)
    k9 = wl.make_kode(w9)
    R9 = np.stack([np.asarray(w9["R"]) * (1 + 0.5 * i / 24) for i in range(24)])
    mu9, var9 = k9.solve_mv(torch.from_numpy(x09).cuda(), theta=torch.from_numpy(th9).cuda(), priors=dict(var_state=R9))
    for b in (0, 11, 23):
        po = oracle.Problem("lorenz63", w9["W"], w9["tmin"], w9["tmax"], 200, w9["Q"], w9["c"], R9[b])
        ref = oracle.solve(po, x09[b], th9[b], None, oracle.MODE_MV)
        assert traj_err(mu9[b].cpu().numpy(), ref["mu"]).max() < 1e-9
        assert var_err(var9[b].cpu().numpy()[1:], ref["var"][1:]).max() < 1e-7


def test_config5_sweep():
    """BASELINE config 5 on one GPU at reduced size: a theta sweep through the fused
    log-likelihood (chunked workspace, as the 1,048,576-theta run of tools/sweep_c5.py), the
    sharding helpers on a single rank, and an oracle check of a strided sample."""
    from rodeo_legacy_b200.dist import gather_batch, shard_range
    total = 40000
    w = wl.fitz()
    rng = np.random.default_rng(0)
    theta = np.exp(np.log(w["theta0"]) + 0.1 * rng.standard_normal((total, 3)))
    x0 = np.tile(w["x0"], (total, 1))
    ind = np.arange(10, 401, 10)
    k = wl.make_kode(w)
    k.set_workspace_limit(8192 * k.history_bytes_per_system)        # 5 chunks
    mu_true, _ = k.solve_mv(w["x0"], theta=w["theta0"])
    Y = mu_true[ind][:, [0, 3]] + 0.2 * np.random.default_rng(5).standard_normal((40, 2))
    lo, hi = shard_range(total, 0, 1)
    assert (lo, hi) == (0, total)
    n0 = k.launch_count
    ll = gather_batch(k.loglik(torch.from_numpy(x0).cuda(), torch.from_numpy(theta).cuda(), Y, ind, [0, 3], 0.2), total)
    assert k.launch_count - n0 == 10 and ll.shape == (total,)
    pick = np.linspace(0, total - 1, 24).astype(int)
    ref = oracle.solve_batch(wl.make_oracle_problem(w), x0[pick], theta[pick], None, oracle.MODE_MV)
    want = [oracle.loglik(m, ind, [0, 3], Y, 0.2) for m in ref["mu"]]
    np.testing.assert_allclose(ll[pick].cpu().numpy(), want, rtol=1e-8)
    # the sweep finds the neighbourhood of the true parameters
    best = theta[int(ll.argmax())]
    assert np.all(np.abs(np.log(best / w["theta0"])) < 0.25)


def test_batched_inference_front_end():
    """The caller side: inference.kalman_nlpost batched over phi (examples/inference.py:85-94) and
    phi_fit (Nelder-Mead + a Hessian stencil evaluated in one batched call)."""
    import scipy.stats
    from rodeo_legacy_b200.inference import inference, thin_index
    g = load_golden("fitz_n400")
    w = wl.fitz()
    z = np.asfortranarray(g["z"])
    k = wl.make_kode(w, z_state=z)
    inf = inference([0, 3], 0, 40, None, W=w["W"], kode=k)
    Y, gam = g["Y"], float(g["gamma"])
    phi_mean, phi_sd = np.log(w["theta0"]), np.ones(3)
    rng = np.random.default_rng(8)
    phi = phi_mean + 0.1 * rng.standard_normal((37, 3))
    nl = inf.kalman_nlpost(phi, Y, w["x0"], 0.1, phi_mean, phi_sd, gam)          # draws, as the reference
    nl_mean = inf.kalman_nlpost(phi, Y, w["x0"], 0.1, phi_mean, phi_sd, gam, use_draw=False)
    assert nl.shape == (37,)
    po = wl.make_oracle_problem(w)
    ref = oracle.solve_batch(po, np.tile(w["x0"], (37, 1)), np.exp(phi), z, oracle.MODE_BOTH)
    ind = thin_index(np.linspace(1, 40, 40), np.linspace(0, 40, 401))
    for key, got in (("x", nl), ("mu", nl_mean)):
        want = [-(inf.loglike(Y, ref[key][b][ind][:, [0, 3]], gam) + inf.loglike(phi[b], phi_mean, phi_sd))
                for b in range(37)]
        np.testing.assert_allclose(got, want, rtol=1e-9)
    # a single phi returns a float, equal to the batched value
    assert np.isclose(inf.kalman_nlpost(phi[3], Y, w["x0"], 0.1, phi_mean, phi_sd, gam), nl[3], rtol=1e-12)
    # the golden observations were generated around theta_true: the fit must land near it
    phi_hat, phi_var = inf.phi_fit(Y, w["x0"], 0.1, w["theta0"], phi_sd, gam, use_draw=False)
    assert np.all(np.abs(phi_hat - phi_mean) < 0.5)
    assert np.all(np.linalg.eigvalsh(phi_var) > 0) and np.all(np.sqrt(np.diag(phi_var)) < 1.0)
    th = inf.theta_sample(phi_hat, phi_var, 1000, rng)
    assert th.shape == (1000, 3) and np.all(th > 0)


def test_ragged_batch_every_system_against_oracle():
    """A batch that is a multiple of nothing (12,345 systems: partial last warp, partial last 16-system
    row of the granule stores, partial last 256-system block of the permuted history order): EVERY system
    against the oracle, all three outputs, on both FitzHugh-Nagumo kernel generations."""
    B = 12345
    w, x0, theta = wl.fitz_batch(B, seed=11)
    z = np.asfortranarray(np.random.default_rng(7).standard_normal((6, 400)))
    ref = oracle.solve_batch(wl.make_oracle_problem(w), x0, theta, z, oracle.MODE_BOTH)
    k = wl.make_kode(w, z_state=z)
    x0d, thd = torch.from_numpy(x0).cuda(), torch.from_numpy(theta).cuda()
    x, mu, var = k.solve(x0d, theta=thd)
    assert int(k.status.abs().sum()) == 0 and not ref["status"].any()
    assert traj_err(mu.cpu().numpy(), ref["mu"]).max() < 1e-10
    assert traj_err(x.cpu().numpy(), ref["x"]).max() < 1e-9
    assert var_err(var.cpu().numpy()[:, 1:], ref["var"][:, 1:]).max() < 1e-10
    assert bool((var == var[0:1]).all()) and bool((var[:, 0] == 0).all())
    mu2, var2 = k.solve_mv(x0d, theta=thd)
    assert torch.equal(mu2, mu) and torch.equal(var2, var)
    assert torch.equal(k.solve_sim(x0d, theta=thd), x)
    # tiny batches: fewer systems than one 16-system row / one warp
    for b in (1, 5, 17, 255, 257):
        xs, ms, vs = k.solve(x0d[:b], theta=thd[:b])
        assert torch.equal(ms, mu[:b]) and torch.equal(vs, var[:b]) and torch.equal(xs, x[:b]), b
    # odd n_eval (n_eval + 1 even: every trajectory starts on a different granule phase), short grids
    for N in (1, 2, 3, 7, 64, 399):
        wn = dict(w, N=N, tmax=w["tmax"] * N / w["N"])
        zn = np.asfortranarray(z[:, :N])
        kn = wl.make_kode(wn, z_state=zn)
        refn = oracle.solve_batch(wl.make_oracle_problem(wn), x0[:300], theta[:300], zn, oracle.MODE_BOTH)
        xn, mn, vn = kn.solve(x0d[:300], theta=thd[:300])
        assert traj_err(mn.cpu().numpy(), refn["mu"]).max() < 1e-10, N
        assert traj_err(xn.cpu().numpy(), refn["x"]).max() < 1e-9, N
        assert np.abs(vn.cpu().numpy() - refn["var"]).max() < 1e-10 * np.abs(refn["var"]).max() + 1e-300, N


def _unpack_bd_record(rec, n, nb):
    """(mu (n,), Sigma (n, n)) from one history record of the block-diagonal family: mu, then per block the
    packed upper triangle, row-major (csrc/kalman_common.cuh)."""
    p = n // nb
    mu = rec[:n].copy()
    S = np.zeros((n, n))
    o = n
    for b in range(nb):
        for i in range(p):
            for j in range(i, p):
                S[b * p + i, b * p + j] = S[b * p + j, b * p + i] = rec[o]
                o += 1
    return mu, S


def test_lorenz_one_step_consistency():
    """SURVEY section 7 'hard parts': Lorenz63 is chaotic, so trajectories of two FP64 codes part at the Lyapunov
    rate -- ONE step does not.  The filter history the GPU wrote (rodeo_cuda_debug_copy_history) is fed to the
    oracle's one-step filter at time index t; the GPU's record for t+1 must equal the result to 1e-12."""
    import ctypes
    B, N = 64, 5000
    w, x0, theta = wl.lorenz_batch(B, N=N)
    k = wl.make_kode(w, checkpoint=1)
    x0d, thd = torch.from_numpy(x0).cuda(), torch.from_numpy(theta).cuda()
    mu, var = k.solve_mv(x0d, theta=thd)
    torch.cuda.synchronize()
    n, nb = 9, 3
    NE = n + nb * 6
    nbytes = (B // 32) * N * NE * 32 * 8
    hist = torch.empty(nbytes // 8, dtype=torch.float64, device="cuda")
    from rodeo_legacy_b200.cuda import _lib
    _lib.check(_lib.load().rodeo_cuda_debug_copy_history(k._h, ctypes.c_void_p(hist.data_ptr()), nbytes, None))
    torch.cuda.synchronize()
    h = hist.cpu().numpy().reshape(B // 32, N, NE, 32)      # [tile][record r = time N - r][element][lane]
    po = wl.make_oracle_problem(w)
    worst = 0.0
    for b in (0, 31, 32, 63):
        for t in (1, 10, 1234, 2500, 4000, 4998):
            m0, S0 = _unpack_bd_record(h[b // 32, N - t, :, b % 32], n, nb)
            m1, S1 = _unpack_bd_record(h[b // 32, N - (t + 1), :, b % 32], n, nb)
            mo, So = oracle.filter_step(po, t, m0, S0, theta[b])
            worst = max(worst, np.abs(mo - m1).max() / np.abs(m1).max(), np.abs(So - S1).max() / np.abs(S1).max())
    assert worst < 1e-12, worst
    # the terminal smoothed moments ARE the last filtered record
    m_end, S_end = _unpack_bd_record(h[0, 0, :, 5], n, nb)
    np.testing.assert_array_equal(mu[5, N].cpu().numpy(), m_end)
    np.testing.assert_array_equal(var[5, N].cpu().numpy(), S_end)


def test_config5_full_size_sweep():
    """BASELINE config 5 at its full 1,048,576 theta draws on one GPU (the sharded form runs under torchrun in
    bench.py): fused log-likelihoods, an oracle check of a strided 256-sample, determinism."""
    total = 1 << 20
    w = wl.fitz()
    theta = wl.fitz_sweep_theta(total, 0, total, w)
    x0 = np.tile(w["x0"], (total, 1))
    k = wl.make_kode(w)
    Y, ind, sind = wl.fitz_observations(k, w)
    x0d, thd = torch.from_numpy(x0).cuda(), torch.from_numpy(theta).cuda()
    ll = k.loglik(x0d, thd, Y, ind, sind, 0.2)
    assert ll.shape == (total,) and bool(torch.isfinite(ll).all()) and int(k.status.abs().sum()) == 0
    pick = np.linspace(0, total - 1, 256).astype(int)
    ref = oracle.solve_batch(wl.make_oracle_problem(w), x0[pick], theta[pick], None, oracle.MODE_MV)
    want = [oracle.loglik(m, ind, sind, Y, 0.2) for m in ref["mu"]]
    np.testing.assert_allclose(ll[torch.from_numpy(pick).cuda()].cpu().numpy(), want, rtol=1e-8)
    assert torch.equal(k.loglik(x0d, thd, Y, ind, sind, 0.2), ll)
    best = theta[int(ll.argmax())]
    assert np.all(np.abs(np.log(best / w["theta0"])) < 0.25)


def test_advice_round1_regressions():
    """Round-1 advisor findings: (1) shared-covariance host path with odd B * (N+1) * n (Lorenz63, n = 9, odd
    batch, even n_eval, solve()); (2) property getters are read-only views; (3) caller-provided host output
    buffers are validated; (4) the batched-prior entry point refuses misaligned device outputs."""
    import ctypes
    w, x0, theta = wl.lorenz_batch(33, N=300, tmax=1.2)
    z = np.asfortranarray(np.random.default_rng(3).standard_normal((9, 300)))
    ks = wl.make_kode(w, z_state=z, shared_cov=True)
    ref = oracle.solve_batch(wl.make_oracle_problem(w), x0, theta, z, oracle.MODE_BOTH)
    for b in (33, 1):                                   # odd batch, and the reference's own call shape (1-D x0)
        xa, ma, va = ks.solve(x0[:b] if b > 1 else np.asfortranarray(x0[0]), theta=theta[:b] if b > 1 else theta[0])
        assert traj_err(ma, ref["mu"][:b] if b > 1 else ref["mu"][0]).max() < 1e-9
        assert traj_err(xa, ref["x"][:b] if b > 1 else ref["x"][0]).max() < 1e-8
    w15, x15, th15 = wl.mseir_batch(7, N=100)
    z15 = np.asfortranarray(np.random.default_rng(4).standard_normal((15, 100)))
    k15 = wl.make_kode(w15, z_state=z15, shared_cov=True)
    r15 = oracle.solve_batch(wl.make_oracle_problem(w15), x15, th15, z15, oracle.MODE_BOTH)
    x_, m_, v_ = k15.solve(x15, theta=th15)
    assert traj_err(m_, r15["mu"]).max() < 1e-10 and traj_err(x_, r15["x"]).max() < 1e-7
    # (2)
    wf = wl.fitz()
    k = wl.make_kode(wf)
    for name in ("wgt_meas", "mu_state", "wgt_state", "var_state", "z_state"):
        with pytest.raises(ValueError, match="read-only"):
            getattr(k, name)[...] = 0.0
    k.mu_state = np.full(6, 1e-3)                       # the setter is the way in; it reaches the device
    mu_c, _ = k.solve_mv(wf["x0"], theta=wf["theta0"])
    k.mu_state = np.zeros(6)
    mu_0, _ = k.solve_mv(wf["x0"], theta=wf["theta0"])
    assert np.abs(mu_c - mu_0).max() > 1e-6
    # (3)
    x0b, thb = np.tile(wf["x0"], (4, 1)), np.tile(wf["theta0"], (4, 1))
    good_mu, good_var = np.empty((4, 401, 6)), np.empty((4, 401, 6, 6))
    k.solve_mv_into(x0b, thb, good_mu, good_var)
    np.testing.assert_array_equal(good_mu[0], mu_0)
    for bad_mu, bad_var in ((np.empty((4, 401, 6), dtype=np.float32), good_var), (good_mu, np.empty((4, 401, 6, 6))[:, ::2]),
                            (good_mu, np.empty((3, 401, 6, 6))), (np.empty((6, 401, 4)).transpose(2, 1, 0), good_var)):
        with pytest.raises(TypeError):
            k.solve_mv_into(x0b, thb, bad_mu, bad_var)
    ksc = wl.make_kode(wf, shared_cov=True)
    with pytest.raises(TypeError, match="incorrect shape"):
        ksc.solve_mv_into(x0b, thb, good_mu, good_var)          # shared mode wants ONE (N+1, n, n) array
    ksc.solve_mv_into(x0b, thb, good_mu, np.empty((401, 6, 6)))
    # (4)
    lib = _lib_load()
    buf = torch.zeros(2 * 401 * 42 + 8, dtype=torch.float64, device="cuda")
    x0d = torch.from_numpy(np.tile(wf["x0"], (2, 1))).cuda()
    thd = torch.from_numpy(np.tile(wf["theta0"], (2, 1))).cuda()
    st = torch.zeros(2, dtype=torch.int32, device="cuda")
    mu_p, var_p = buf.data_ptr() + 8, buf.data_ptr() + 8 + 2 * 401 * 6 * 8
    V = ctypes.c_void_p
    rc = lib.rodeo_cuda_solve_batched_prior(k._h, 1, 2, V(x0d.data_ptr()), V(thd.data_ptr()), None, 0, None, None, None,
                                            None, None, V(mu_p), V(var_p), V(st.data_ptr()), None)
    assert rc == -1 and b"16-byte aligned" in lib.rodeo_cuda_last_error()


def _lib_load():
    from rodeo_legacy_b200.cuda import _lib
    return _lib.load()


@pytest.mark.parametrize("maker,B", [
    (wl.fitz_batch, 2000), (lambda B: wl.lorenz_batch(B, N=400, tmax=1.6), 130), (lambda B: wl.mseir_batch(B, N=150), 70),
    (lambda B: (wl.chkrebtii(150), np.tile(wl.chkrebtii(150)["x0"], (B, 1)) * (1 + 1e-3 * np.arange(B)[:, None]), None), 40)])
def test_structured_prior_kernels_are_bit_identical(maker, B):
    """ibm_init + zero_pad priors (upper-triangular Q blocks, unit-vector W rows) select the structured
    block-diagonal kernels, which skip the exactly-zero terms: the outputs must equal the general block-diagonal
    kernels' bit for bit (solve, loglik, checkpoint 1 and auto), and a CAR prior (dense Q blocks) must not select
    them."""
    w, x0, theta = maker(B)
    n = len(w["x0"])
    z = np.asfortranarray(np.random.default_rng(9).standard_normal((n, w["N"])))
    x0d = torch.from_numpy(x0).cuda()
    thd = None if theta is None else torch.from_numpy(theta).cuda()
    for ck in (0, 1):
        ks = wl.make_kode(w, z_state=z, checkpoint=ck)
        kg = wl.make_kode(w, z_state=z, checkpoint=ck, use_structure=False)
        assert ks.structured and not kg.structured and ks.kernel_family == kg.kernel_family == ("blockdiag", "registers")
        for a, b in zip(ks.solve(x0d, theta=thd), kg.solve(x0d, theta=thd)):
            assert torch.equal(a, b)
        for a, b in zip(ks.solve_mv(x0d, theta=thd), kg.solve_mv(x0d, theta=thd)):
            assert torch.equal(a, b)
        assert torch.equal(ks.status, kg.status)
        if n == 6:
            ind = np.arange(10, w["N"] + 1, 10)
            Y = np.random.default_rng(1).standard_normal((len(ind), 2))
            for draw in (False, True):
                assert torch.equal(ks.loglik(x0d, thd, Y, ind, [0, 3], 0.3, use_draw=draw),
                                   kg.loglik(x0d, thd, Y, ind, [0, 3], 0.3, use_draw=draw))
    # a W row that is not the unit vector, or a Q block with a non-zero below the diagonal: general kernels
    w2 = dict(w)
    w2["W"] = np.asfortranarray(np.asarray(w["W"]) * 2.0)
    assert not wl.make_kode(w2, z_state=z).structured
    w3 = dict(w)
    Q3 = np.array(w["Q"], order="F")
    Q3[1, 0] = 1e-6
    w3["Q"] = Q3
    k3 = wl.make_kode(w3, z_state=z)
    assert not k3.structured and k3.kernel_family[0] == "blockdiag"


def test_structured_flag_on_car_prior():
    g = load_golden("lorenz_car_n600")
    k = _kode_from_golden(g, "lorenz_car_n600", z_state=np.asfortranarray(g["z"]))
    assert k.kernel_family[0] == "blockdiag" and not k.structured


@pytest.mark.parametrize("maker,B,N", [
    (lambda B: wl.lorenz_batch(B, N=301, tmax=1.2), 107, 301), (lambda B: wl.lorenz_batch(B, N=300, tmax=1.2), 20, 300),
    (lambda B: wl.lorenz_batch(B, N=2, tmax=0.01), 41, 2), (lambda B: wl.lorenz_batch(B, N=1, tmax=0.01), 9, 1),
    (lambda B: wl.mseir_batch(B, N=150), 64, 150), (lambda B: wl.mseir_batch(B, N=77), 13, 77)])
def test_draw_only_smoother_for_three_or_more_blocks(maker, B, N):
    """solve_sim on priors with >= 3 blocks runs on the thread-per-block, two-systems-per-lane kernel
    (kalman_tpx.cuh) with checkpoint interval 2 by default: against the oracle, bit-identical to the full
    history (checkpoint=1), to the draws of solve() (warp-per-block kernel) and to the host path; shared and
    per-system z; batches that fill no warp."""
    w, x0, theta = maker(B)
    n = len(w["x0"])
    rng = np.random.default_rng(10)
    z = np.asfortranarray(rng.standard_normal((n, N)))
    x0d, thd = torch.from_numpy(x0).cuda(), torch.from_numpy(theta).cuda()
    k2, k1 = wl.make_kode(w, z_state=z), wl.make_kode(w, z_state=z, checkpoint=1)
    x2 = k2.solve_sim(x0d, theta=thd)
    x1 = k1.solve_sim(x0d, theta=thd)
    assert torch.equal(x1, x2) and int(k2.status.abs().sum()) == 0
    xb, _, _ = k2.solve(x0d, theta=thd)
    assert torch.equal(xb, x2)
    np.testing.assert_array_equal(k2.solve_sim(x0, theta=theta), x2.cpu().numpy())
    ref = oracle.solve_batch(wl.make_oracle_problem(w), x0, theta, z, oracle.MODE_SIM)
    assert traj_err(x2.cpu().numpy(), ref["x"]).max() < (1e-8 if n == 9 else 1e-7)
    assert torch.equal(x2[:, 0], x0d)
    # general (unstructured) kernels give the same bits
    assert torch.equal(wl.make_kode(w, z_state=z, use_structure=False).solve_sim(x0d, theta=thd), x2)
    # per-system draws
    zb = rng.standard_normal((B, n, N))
    kz = wl.make_kode(w, z_state=zb)
    xz = kz.solve_sim(x0d, theta=thd)
    refz = oracle.solve_batch(wl.make_oracle_problem(w), x0, theta, zb, oracle.MODE_SIM)
    assert traj_err(xz.cpu().numpy(), refz["x"]).max() < (1e-8 if n == 9 else 1e-7)


def test_per_system_blockdiag_priors_and_sigma_sweep_likelihood():
    """SURVEY section 8(f) #4's motivating use: a sweep over the IBM scale sigma.  Every system has its own (block
    diagonal) R and c: the thread-per-block kernels run with a per-lane prior (checkpoint 2, granule stores);
    rodeo_cuda_loglik_batched_prior gives the likelihood of each sigma without materialising trajectories."""
    from rodeo_legacy_b200 import ibm_init, indep_init
    B = 333                                           # ragged: partial 16-system row, partial 256 block
    w, x0, theta = wl.fitz_batch(B, seed=6)
    n, N = 6, w["N"]
    rng = np.random.default_rng(3)
    sig = 0.02 + 0.3 * rng.random((B, 2))
    R = np.empty((B, n, n)); c = np.zeros((B, n))
    for b in range(B):
        R[b] = indep_init(ibm_init(40.0 / N, [3, 3], list(sig[b])), [3, 3])["var_state"]
        if b % 4 == 0:
            c[b] = 1e-4 * rng.standard_normal(n)
    z = np.asfortranarray(rng.standard_normal((n, N)))
    k = wl.make_kode(w, z_state=z)
    x0d, thd = torch.from_numpy(x0).cuda(), torch.from_numpy(theta).cuda()
    pri = dict(var_state=R, mu_state=c)
    x, mu, var = k.solve(x0d, theta=thd, priors=pri)
    assert k.last_batched_prior_blockdiag and int(k.status.abs().sum()) == 0
    pick = list(range(0, B, 7)) + [B - 1]
    for b in pick:
        po = oracle.Problem("fitz", w["W"], w["tmin"], w["tmax"], N, w["Q"], c[b], R[b])
        ref = oracle.solve(po, x0[b], theta[b], z, oracle.MODE_BOTH)
        assert traj_err(mu[b].cpu().numpy(), ref["mu"]).max() < 1e-10
        assert var_err(var[b].cpu().numpy()[1:], ref["var"][1:]).max() < 1e-9
        assert traj_err(x[b].cpu().numpy(), ref["x"]).max() < 1e-8
    # the three entry points agree bit for bit; NumPy in -> NumPy out
    mu2, var2 = k.solve_mv(x0d, theta=thd, priors=pri)
    assert torch.equal(mu2, mu) and torch.equal(var2, var)
    assert torch.equal(k.solve_sim(x0d, theta=thd, priors=pri), x)
    mu_h, _ = k.solve_mv(x0, theta=theta, priors=pri)
    np.testing.assert_array_equal(mu_h, mu.cpu().numpy())
    # repeating the shared prior per system reproduces the shared-prior solve exactly (same kernels, same bits)
    mu_s, var_s = k.solve_mv(x0d, theta=thd)
    mu_r, var_r = k.solve_mv(x0d, theta=thd, priors=dict(var_state=np.tile(np.asarray(w["R"]), (B, 1, 1))))
    assert k.last_batched_prior_blockdiag
    assert torch.equal(mu_r, mu_s) and torch.equal(var_r, var_s)
    # a prior that couples the blocks for ONE system sends the call to the dense family
    R2 = R.copy()
    R2[200, 0, 3] = R2[200, 3, 0] = 1e-12
    mu_d, _ = k.solve_mv(x0d, theta=thd, priors=dict(var_state=R2, mu_state=c))
    assert not k.last_batched_prior_blockdiag
    assert float((mu_d[:200] - mu[:200]).abs().max()) < 1e-9 * float(mu.abs().max())
    # likelihood of each sigma: rodeo_cuda_loglik_batched_prior against the oracle's solve + loglik
    g = load_golden("fitz_n400")
    ind, Y, gam = g["thin_ind"], g["Y"], float(g["gamma"])
    for draw, key in ((False, "mu"), (True, "x")):
        ll = k.loglik(x0d, thd, Y, ind, [0, 3], gam, use_draw=draw, priors=pri).cpu().numpy()
        for b in pick[::5]:
            po = oracle.Problem("fitz", w["W"], w["tmin"], w["tmax"], N, w["Q"], c[b], R[b])
            ref = oracle.solve(po, x0[b], theta[b], z, oracle.MODE_BOTH)
            assert np.isclose(ll[b], oracle.loglik(ref[key], ind, [0, 3], Y, gam), rtol=1e-8)
    ll_np = k.loglik(x0, theta, Y, ind, [0, 3], gam, priors=pri)
    assert isinstance(ll_np, np.ndarray) and ll_np.shape == (B,)


@pytest.mark.parametrize("B,N", [(3000, 400), (77, 33), (1, 2)])
def test_compact_variance_layouts(B, N):
    """var_layout="blocks" / "diag" (opt-in): the diagonal (p, p) blocks / the marginal variances of the reference's
    (n, n) output, bit for bit, through the device and the host entry points, solve_mv and solve; means and draws
    unchanged; against the oracle's full matrices."""
    w, x0, theta = wl.fitz_batch(B, seed=12)
    if N != w["N"]:
        w = dict(w, N=N, tmax=w["tmax"] * N / w["N"])
    z = np.asfortranarray(np.random.default_rng(2).standard_normal((6, N)))
    x0d, thd = torch.from_numpy(x0).cuda(), torch.from_numpy(theta).cuda()
    kf = wl.make_kode(w, z_state=z)
    xf, muf, varf = kf.solve(x0d, theta=thd)
    ref = oracle.solve_batch(wl.make_oracle_problem(w), x0, theta, None, oracle.MODE_MV)
    blocks_of = lambda v: torch.stack([v[..., 0:3, 0:3], v[..., 3:6, 3:6]], dim=-3)
    for lay in ("blocks", "diag"):
        k = wl.make_kode(w, z_state=z, var_layout=lay)
        assert k.var_layout == lay
        want = blocks_of(varf) if lay == "blocks" else torch.diagonal(varf, dim1=-2, dim2=-1)
        mu, var = k.solve_mv(x0d, theta=thd)
        assert var.shape == ((B, N + 1, 2, 3, 3) if lay == "blocks" else (B, N + 1, 6))
        assert torch.equal(mu, muf) and torch.equal(var, want.contiguous())
        x, mu2, var2 = k.solve(x0d, theta=thd)
        assert torch.equal(x, xf) and torch.equal(mu2, muf) and torch.equal(var2, var)
        assert torch.equal(k.solve_sim(x0d, theta=thd), xf)
        muh, varh = k.solve_mv(x0, theta=theta)                      # host buffers: fewer bytes cross PCIe
        np.testing.assert_array_equal(varh, var.cpu().numpy())
        np.testing.assert_array_equal(muh, muf.cpu().numpy())
        ro = ref["var"]
        wo = (np.stack([ro[..., 0:3, 0:3], ro[..., 3:6, 3:6]], axis=-3) if lay == "blocks"
              else np.diagonal(ro, axis1=-2, axis2=-1))
        got = var.cpu().numpy()
        assert np.abs(got - wo).max() <= 1e-10 * np.abs(wo).max()
        # 1-D x0: the reference's single-solve shapes
        m1, v1 = k.solve_mv(np.asfortranarray(x0[0]), theta=theta[0])
        assert v1.shape == ((N + 1, 2, 3, 3) if lay == "blocks" else (N + 1, 6))
    kf.var_layout = "diag"                                           # settable after construction, and back
    assert kf.solve_mv(x0d, theta=thd)[1].shape == (B, N + 1, 6)
    kf.var_layout = "full"
    assert torch.equal(kf.solve_mv(x0d, theta=thd)[1], varf)
    with pytest.raises(ValueError):
        kf.var_layout = "upper"
    # dense and single-block priors refuse instead of silently returning the full layout
    with pytest.raises(NotImplementedError):
        wl.make_kode(w, force_dense=True, var_layout="diag")
    with pytest.raises(NotImplementedError):
        wl.make_kode(wl.chkrebtii(50), var_layout="blocks")
    ks = wl.make_kode(w, shared_cov=True, var_layout="blocks")
    with pytest.raises(NotImplementedError):
        ks.solve_mv(x0d, theta=thd)


def test_blocks_with_different_priors_and_offsets():
    """Blocks that are NOT copies of each other (different sigma per variable, a non-zero mu_state): the
    thread-per-block kernels take their per-lane prior path (BlockPrior<false>) instead of immediate constant-bank
    operands.  FitzHugh-Nagumo (two blocks, granule stores) and Lorenz63 (three blocks, draw-only kernel), structured
    and general arithmetic, against the oracle."""
    from rodeo_legacy_b200 import ibm_init, indep_init
    rng = np.random.default_rng(21)
    for maker, B, sig, xtol in ((wl.fitz_batch, 700, [0.07, 0.31], 1e-9),
                                (lambda B: wl.lorenz_batch(B, N=400, tmax=1.6), 90, [0.3, 0.5, 0.9], 1e-8)):
        w, x0, theta = maker(B)
        n, N = len(w["x0"]), w["N"]
        nb = len(sig)
        kin = indep_init(ibm_init((w["tmax"] - w["tmin"]) / N, [3] * nb, sig), [3] * nb)
        w = dict(w, Q=kin["wgt_state"], R=kin["var_state"], c=1e-4 * rng.standard_normal(n))
        z = np.asfortranarray(rng.standard_normal((n, N)))
        ref = oracle.solve_batch(wl.make_oracle_problem(w), x0, theta, z, oracle.MODE_BOTH)
        x0d, thd = torch.from_numpy(x0).cuda(), torch.from_numpy(theta).cuda()
        outs = []
        for st in (True, False):
            k = wl.make_kode(w, z_state=z, use_structure=st)
            assert k.structured == st and k.kernel_family == ("blockdiag", "registers")
            x, mu, var = k.solve(x0d, theta=thd)
            xs = k.solve_sim(x0d, theta=thd)
            assert int(k.status.abs().sum()) == 0 and torch.equal(xs, x)
            assert traj_err(mu.cpu().numpy(), ref["mu"]).max() < 1e-9
            assert traj_err(x.cpu().numpy(), ref["x"]).max() < xtol
            assert np.abs(var.cpu().numpy() - ref["var"]).max() < 1e-9 * np.abs(ref["var"]).max()
            outs.append((x, mu, var))
        for a, b in zip(*outs):
            assert torch.equal(a, b)


def test_very_long_trajectories_fall_back_to_the_staged_smoother():
    """The thread-per-block kernels keep 32-bit byte offsets inside a 16-system output row; beyond ~466,000 steps
    (n_state 6) the previous-generation kernels take over.  (1) the hand-over itself, forced on an ordinary size
    through the test hook: bit-identical outputs; (2) a 470,000-step solve: runs, clean status, and agrees with the
    oracle as far as two FP64 codes can over such a horizon (the FitzHugh-Nagumo limit cycle is neutrally stable in
    phase: 1e-5 after 470,000 steps, 1e-9 over the first 5,000)."""
    w, x0, theta = wl.fitz_batch(300, seed=2)
    z = np.asfortranarray(np.random.default_rng(4).standard_normal((6, 400)))
    x0d, thd = torch.from_numpy(x0).cuda(), torch.from_numpy(theta).cuda()
    k, kf = wl.make_kode(w, z_state=z), wl.make_kode(w, z_state=z)
    _lib_check(_lib_load().rodeo_cuda_debug_set_row_limit(kf._h, 1024.0))
    for a, b in zip(k.solve(x0d, theta=thd), kf.solve(x0d, theta=thd)):
        assert torch.equal(a, b)
    N = 470000
    w = wl.fitz(N=400)
    w = dict(w, N=N, tmax=w["tmax"] * N / 400)
    from rodeo_legacy_b200 import ibm_init, indep_init
    kin = indep_init(ibm_init(w["tmax"] / N, [3, 3], [0.1, 0.1]), [3, 3])
    w.update(Q=kin["wgt_state"], R=kin["var_state"], c=kin["mu_state"])
    B = 3
    x0 = np.tile(w["x0"], (B, 1))
    theta = w["theta0"] * np.exp(0.05 * np.random.default_rng(1).standard_normal((B, 3)))
    kl = wl.make_kode(w)
    mu, var = kl.solve_mv(torch.from_numpy(x0).cuda(), theta=torch.from_numpy(theta).cuda())
    ref = oracle.solve_batch(wl.make_oracle_problem(w), x0, theta, None, oracle.MODE_MV)
    assert int(kl.status.abs().sum()) == 0 and bool(torch.isfinite(mu).all())
    mu = mu.cpu().numpy()
    assert traj_err(mu[:, :5001], ref["mu"][:, :5001]).max() < 1e-8      # measured 1.1e-9
    assert traj_err(mu, ref["mu"]).max() < 1e-3                        # measured 1.0e-5
    # the covariance recursion of the probde interrogation (measurement variance = W Sigma_p W^T, a function of the
    # state covariance itself) drifts apart slowly between two FP64 codes: 1.4e-5 of the largest entry after
    # 470,000 steps, FP64 rounding over the first 5,000
    got, want = var.cpu().numpy(), ref["var"]
    assert np.abs(got[:, 1:5001] - want[:, 1:5001]).max() < 1e-9 * np.abs(want).max()
    assert np.abs(got[:, 1:] - want[:, 1:]).max() < 1e-4 * np.abs(want).max()


def _lib_check(rc):
    from rodeo_legacy_b200.cuda import _lib
    _lib.check(rc)


@pytest.mark.parametrize("maker,B", [(lambda B: wl.lorenz_batch(B, N=300, tmax=1.2), 75), (lambda B: wl.mseir_batch(B, N=120), 41)])
def test_compact_variance_layouts_three_or_more_blocks(maker, B):
    """var_layout on the warp-per-block kernels (Lorenz63: 3 blocks, MSEIR: 5): the diagonal blocks / marginal
    variances of the full output bit for bit, solve_mv and solve, device and host entry points."""
    w, x0, theta = maker(B)
    n, N = len(w["x0"]), w["N"]
    nb, p = n // 3, 3
    z = np.asfortranarray(np.random.default_rng(5).standard_normal((n, N)))
    x0d, thd = torch.from_numpy(x0).cuda(), torch.from_numpy(theta).cuda()
    kf = wl.make_kode(w, z_state=z)
    xf, muf, varf = kf.solve(x0d, theta=thd)
    for lay in ("blocks", "diag"):
        k = wl.make_kode(w, z_state=z, var_layout=lay)
        want = (torch.stack([varf[..., a * p:(a + 1) * p, a * p:(a + 1) * p] for a in range(nb)], dim=-3)
                if lay == "blocks" else torch.diagonal(varf, dim1=-2, dim2=-1)).contiguous()
        mu, var = k.solve_mv(x0d, theta=thd)
        assert var.shape == ((B, N + 1, nb, p, p) if lay == "blocks" else (B, N + 1, n))
        assert torch.equal(mu, muf) and torch.equal(var, want)
        x, mu2, var2 = k.solve(x0d, theta=thd)
        assert torch.equal(x, xf) and torch.equal(mu2, muf) and torch.equal(var2, want)
        muh, varh = k.solve_mv(x0, theta=theta)
        np.testing.assert_array_equal(varh, want.cpu().numpy())


@pytest.mark.parametrize("name,family", [("fitz_blocks_2_3", "registers"), ("fitz_blocks_3_4", "wide"),
                                         ("fitz_blocks_4_3", "wide"), ("lorenz_blocks_2_3_3", "wide")])
def test_unequal_block_sizes(name, family):
    """n_deriv_prior with unequal entries (VERDICT r1 missing #2): the handle finds the variables from wgt_meas and
    runs the dense set compiled for that list (csrc/kernels_mixed.cu).  Against the reference's Python path
    (tests/golden/mixed_blocks.npz) and, on a seeded batch, the oracle with the same offsets."""
    from tests.common import load_mixed
    from rodeo_legacy_b200.cuda import supports
    g = load_mixed(name)
    sizes = [int(v) for v in g["block_sizes"]]
    assert supports(ode_name(name), sizes)
    k = _kode_from_golden(g, name, z_state=np.asfortranarray(g["z"]))
    assert k.kernel_family == ("dense", family)
    p4 = 4 in sizes
    # a p = 4 IBM block: the draws of two FP64 codes agree to ~1e-5 only (cancellation in V~)
    mtol, vtol, xtol = (1e-9, 1e-8, 5e-5) if p4 else (1e-11, 1e-10, 1e-9)
    B0 = len(g["theta"])
    x, mu, var = k.solve(np.tile(g["x0"], (B0, 1)), theta=g["theta"])            # host path
    assert traj_err(mu, g["mu"]).max() < mtol
    assert var_err(var[:, 1:], g["var"][:, 1:]).max() < vtol
    assert traj_err(x, g["x"]).max() < xtol
    # a seeded batch (ragged: not a multiple of 32) through the device path, every system against the oracle
    B = 203
    rng = np.random.default_rng(21)
    theta = g["theta"][0] * np.exp(0.05 * rng.standard_normal((B, g["theta"].shape[1])))
    x0 = np.tile(g["x0"], (B, 1)) * (1 + 0.01 * rng.standard_normal((B, len(g["x0"]))))
    po = oracle.Problem(ode_name(name), g["W"], float(g["tmin"]), float(g["tmax"]), int(g["N"]), g["Q"], g["c"],
                        g["R"], block_sizes=sizes)
    ref = oracle.solve_batch(po, x0, theta, g["z"], oracle.MODE_BOTH)
    x0d, thd = torch.from_numpy(x0).cuda(), torch.from_numpy(theta).cuda()
    xd, mud, vard = k.solve(x0d, theta=thd)
    mu2, var2 = k.solve_mv(x0d, theta=thd)
    x2 = k.solve_sim(x0d, theta=thd)
    assert torch.equal(mud, mu2) and torch.equal(vard, var2) and torch.equal(xd, x2)
    # GPU vs oracle: two FP64 codes on the same inputs (tighter than either against SciPy's Cholesky)
    assert traj_err(mud.cpu().numpy(), ref["mu"]).max() < (1e-9 if p4 else 1e-11)
    assert var_err(vard.cpu().numpy()[:, 1:], ref["var"][:, 1:]).max() < (1e-8 if p4 else 1e-10)
    assert not k.status.any().item() and not ref["status"].any()
    assert traj_err(xd.cpu().numpy(), ref["x"]).max() < (5e-5 if p4 else 1e-9)
    # fused log-likelihood of the same family equals the density of the returned mean
    n, N = len(g["x0"]), int(g["N"])
    ind = np.arange(20, N + 1, 20)
    comp = [0, sizes[0]]
    Y = mud[0, ind][:, comp].cpu().numpy() + 0.05
    got = k.loglik(x0d, thd, Y, ind, comp, 0.2).cpu().numpy()
    tc = mud[:, ind][:, :, comp].cpu().numpy()
    want = -Y.size * (np.log(0.2) + 0.5 * np.log(2 * np.pi)) - 0.5 * (((Y[None] - tc) / 0.2) ** 2).sum(axis=(1, 2))
    assert np.abs(got - want).max() < 1e-9 * np.abs(want).max()
    # per-system priors keep the offsets of the handle (dense family with a thread-private prior)
    nb = 12
    Rb = np.stack([np.asarray(g["R"]) * (1 + 0.1 * i) for i in range(nb)])
    mup, varp = k.solve_mv(x0d[:nb], theta=thd[:nb], priors=dict(var_state=Rb))
    for b in (0, 5, nb - 1):
        pb = oracle.Problem(ode_name(name), g["W"], float(g["tmin"]), float(g["tmax"]), N, g["Q"], g["c"], Rb[b],
                            block_sizes=sizes)
        rb = oracle.solve(pb, x0[b], theta[b], None, oracle.MODE_MV)
        assert traj_err(mup[b].cpu().numpy(), rb["mu"]).max() < (1e-9 if p4 else 1e-11)
        assert var_err(varp[b].cpu().numpy()[1:], rb["var"][1:]).max() < (1e-8 if p4 else 1e-10)


def test_unequal_block_sizes_without_a_compiled_set_fail_loudly():
    from rodeo_legacy_b200 import ibm_init, indep_init, zero_pad
    from rodeo_legacy_b200.cuda import KalmanODE, RodeoCudaError
    n_prior = [2, 4]      # n_state 6 = 2 * 3, but the second variable sits at index 2, not 3
    W = zero_pad(np.array([[0.0, 1.0, 0.0, 0.0], [0.0, 0.0, 0.0, 1.0]]), [1, 1], n_prior)
    kinit = indep_init(ibm_init(0.1, n_prior, [0.1, 0.1]), n_prior)
    with pytest.raises((RodeoCudaError, NotImplementedError), match="block sizes"):
        KalmanODE(W, 0.0, 40.0, 400, "fitz", **kinit)


_FORWARD_AB = r"""
import sys, numpy as np, torch
sys.path.insert(0, {root!r})
from tests import workloads as wl
out = {{}}
for name, maker, B, kw in [("fitz", wl.fitz_batch, 700, {{}}), ("fitz_sch", wl.fitz_batch, 300, dict(interrogate="schobert")),
                           ("fitz_chk", wl.fitz_batch, 300, dict(interrogate="chkrebtii")),
                           ("mseir", lambda B: wl.mseir_batch(B, N=120), 70, {{}}),
                           ("lorenz", lambda B: wl.lorenz_batch(B, N=200, tmax=0.8), 45, {{}})]:
    w, x0, theta = maker(B)
    n = len(w["x0"])
    z = np.asfortranarray(np.random.default_rng(11).standard_normal((n, w["N"])))
    k = wl.make_kode(w, z_state=z, **kw)
    x, mu, var = k.solve(torch.from_numpy(x0).cuda(), theta=torch.from_numpy(theta).cuda())
    out[name + "_x"], out[name + "_mu"], out[name + "_var"] = x.cpu().numpy(), mu.cpu().numpy(), var.cpu().numpy()
np.savez({path!r}, **out)
"""


def test_forward_kernels_are_bit_identical(tmp_path):
    """The warp-per-block forward filter (kalman_wpb.cuh, the default for two-block priors and for full-history
    passes with three or more blocks) against the kernels it replaced (RODEO_CUDA_FORWARD=legacy: thread per
    block / thread per system) and forced everywhere (=wpb: also the checkpointed passes with three or more
    blocks): every output of `solve` bit for bit, for the three interrogation variants.  The preference is read
    once per process, so each setting runs in its own interpreter."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for mode in ("default", "legacy", "wpb"):
        path = str(tmp_path / f"{mode}.npz")
        env = dict(os.environ)
        env.pop("RODEO_CUDA_FORWARD", None)
        if mode != "default":
            env["RODEO_CUDA_FORWARD"] = mode
        r = subprocess.run([sys.executable, "-c", _FORWARD_AB.format(root=root, path=path)], env=env,
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        res[mode] = dict(np.load(path))
    for key, ref in res["legacy"].items():
        if key.endswith("_mu") and "_sch" not in key:
            assert np.isfinite(ref).all(), key
        # (schobert: H = 0 leaves Sigma_f singular and diverges for most theta draws -- the NaNs must match too)
        for mode in ("default", "wpb"):
            assert np.array_equal(ref, res[mode][key], equal_nan=True), (mode, key)


def test_into_variants_write_caller_buffers():
    """solve_sim_into / solve_into / solve_mv_into (host path into buffers the caller allocated once, e.g. pinned):
    the same numbers as the returning methods, bad buffers refused before anything is copied."""
    w, x0, theta = wl.fitz_batch(1500, seed=8)
    z = np.asfortranarray(np.random.default_rng(8).standard_normal((6, w["N"])))
    k = wl.make_kode(w, z_state=z)
    x, mu, var = k.solve(x0, theta=theta)
    T = w["N"] + 1
    pin = lambda *s: torch.empty(s, dtype=torch.float64, pin_memory=True).numpy()
    xo, mo, vo = pin(1500, T, 6), pin(1500, T, 6), pin(1500, T, 6, 6)
    k.solve_into(x0, theta, xo, mo, vo)
    assert np.array_equal(xo, x) and np.array_equal(mo, mu) and np.array_equal(vo, var)
    xo[:] = 0.0
    k.solve_sim_into(x0, theta, xo)
    assert np.array_equal(xo, k.solve_sim(x0, theta=theta)) and np.array_equal(xo, x)
    mo[:] = 0.0
    k.solve_mv_into(x0, theta, mo, vo)
    assert np.array_equal(mo, mu)
    for bad in (np.empty((1500, T, 6), dtype=np.float32), np.empty((1499, T, 6)), np.empty((1500, T, 12))[:, :, ::2]):
        with pytest.raises(TypeError):
            k.solve_sim_into(x0, theta, bad)


@pytest.mark.parametrize("maker,B,mtol,vtol", [
    (lambda B: wl.lorenz_batch(B, N=301, tmax=1.2), 37, 1e-9, 1e-7),
    (lambda B: wl.mseir_batch(B, N=150), 21, 1e-9, 1e-7)])
def test_chkrebtii_interrogation_on_the_wide_dense_kernels(maker, B, mtol, vtol):
    """forecast (the chkrebtii interrogation, rodeo/cython/KalmanODE.pyx:13-49: the right-hand side is evaluated at a
    draw from N(mu_pred, Sigma_pred) with z[:, t]) on the half-warp-per-system kernels for dense wide states
    (kalman_dw.cuh; they used to hand this variant to the rolled kernels), against the oracle."""
    w, x0, theta = maker(B)
    n = len(w["x0"])
    z = np.asfortranarray(np.random.default_rng(6).standard_normal((n, w["N"])))
    k = wl.make_kode(w, z_state=z, interrogate="chkrebtii", force_dense=True)
    assert k.kernel_family == ("dense", "wide")           # the wide kernels, not the rolled ones
    ref = oracle.solve_batch(wl.make_oracle_problem(w, interrogate="chkrebtii"), x0, theta, z, oracle.MODE_MV)
    mu, var = (a.cpu().numpy() for a in k.solve_mv(torch.from_numpy(x0).cuda(), theta=torch.from_numpy(theta).cuda()))
    assert np.isfinite(mu).all() and int(np.count_nonzero(k.status.cpu().numpy())) == 0
    assert traj_err(mu, ref["mu"]).max() < mtol
    assert var_err(var[:, 1:], ref["var"][:, 1:]).max() < vtol
