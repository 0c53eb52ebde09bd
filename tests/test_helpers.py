"""Host-side prior helpers against values produced by the reference's own helpers
(tests/golden/helpers.npz, made by tests/golden/make_golden.py)."""
import numpy as np

from rodeo_legacy_b200 import ibm_init, car_init, indep_init, zero_pad, rand_mat
from rodeo_legacy_b200.utils import norm_sim, solveV, mvncond
from tests.common import load_golden

TOL = dict(rtol=1e-13, atol=0)


def test_ibm_init_matches_reference():
    g = load_golden("helpers")
    ib = ibm_init(0.1, [3, 4], [0.1, 0.7])
    np.testing.assert_allclose(ib["wgt_state"][0], g["ibm_wgt0"], **TOL)
    np.testing.assert_allclose(ib["var_state"][0], g["ibm_var0"], **TOL)
    np.testing.assert_allclose(ib["wgt_state"][1], g["ibm_wgt1"], **TOL)
    np.testing.assert_allclose(ib["var_state"][1], g["ibm_var1"], **TOL)
    assert ib["mu_state"].shape == (7,) and not ib["mu_state"].any()
    one = ibm_init(0.05, [4], [0.5])
    np.testing.assert_allclose(one["wgt_state"], g["ibm_single_wgt"], **TOL)
    np.testing.assert_allclose(one["var_state"], g["ibm_single_var"], **TOL)
    # SURVEY appendix A check value: p=3, dt=0.1, sigma=0.1 -> Q[0,0]
    np.testing.assert_allclose(ib["var_state"][0][0, 0], 3.96825397e-12, rtol=1e-8)


def test_indep_init_and_zero_pad():
    g = load_golden("helpers")
    k = indep_init(ibm_init(0.1, [3, 4], [0.1, 0.7]), [3, 4])
    np.testing.assert_allclose(k["wgt_state"], g["indep_wgt"], **TOL)
    np.testing.assert_allclose(k["var_state"], g["indep_var"], **TOL)
    assert k["wgt_state"].flags.f_contiguous
    v = zero_pad(np.array([1.0, 2, 3, 4, 5]), [1, 2], [3, 5])
    np.testing.assert_array_equal(v, g["zero_pad_vec"])
    m = zero_pad(np.array([[0.0, 1, 0, 0, 0], [0, 0, 0, 0, 1]]), [1, 2], [3, 5])
    np.testing.assert_array_equal(m, g["zero_pad_mat"])
    # single-variable dict passes through indep_init too
    one = indep_init(ibm_init(0.05, [4], [0.5]), [4])
    np.testing.assert_allclose(one["wgt_state"], g["ibm_single_wgt"], **TOL)


def test_car_init_matches_reference():
    g = load_golden("helpers")
    c = car_init(0.01, [4], [1.3], [0.5])
    np.testing.assert_allclose(c["wgt_state"], g["car_wgt"], rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(c["var_state"], g["car_var"], rtol=1e-9, atol=1e-20)
    lg = load_golden("lorenz_car_n600")
    np.random.seed(7)
    X0 = np.column_stack([[-12, -5, 38], [70, 125, -124 / 3]])
    init, x0 = car_init(2.4 / 600, [3, 3, 3], np.array([1.3] * 3), np.array([0.5] * 3), X0)
    np.testing.assert_allclose(init["wgt_state"][0], lg["car_wgt"], rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(init["var_state"][0], lg["car_var"], rtol=1e-9, atol=1e-20)
    # same seed, same draw order -> same conditional initial draw
    np.testing.assert_allclose(x0, lg["x0"], rtol=1e-9, atol=1e-12)


def test_car_initial_draw_batch():
    """Batched initial draw: row b equals the single-system draw made with the same normals; the
    conditional law (mean A x0 + b, variance V) is the reference's (rodeo/car/car_init.py:22-49)."""
    import torch
    from rodeo_legacy_b200.car import car_initial_draw_batch, car_cov, root_gen
    from rodeo_legacy_b200.utils import mvncond
    roots, sigma, p = root_gen(1.5, 4), 0.3, 4
    rng = np.random.default_rng(3)
    X0 = rng.standard_normal((50, 2))
    z = rng.standard_normal((50, 2))
    out = car_initial_draw_batch(roots, sigma, X0, p, z=z)
    assert out.shape == (50, 4) and np.array_equal(out[:, :2], X0)
    A, b, V = mvncond(np.zeros(p), car_cov([], roots, sigma, v_infinity=True), np.arange(p) <= 1)
    L = np.linalg.cholesky(V)
    np.testing.assert_allclose(out[:, 2:], X0 @ A.T + b + z @ L.T, rtol=1e-13)
    # all derivatives known: nothing to draw
    assert car_initial_draw_batch(roots, sigma, rng.standard_normal((3, 4)), p).shape == (3, 4)
    # torch in -> torch out, same numbers
    out_t = car_initial_draw_batch(roots, sigma, torch.from_numpy(X0), p, z=z)
    assert isinstance(out_t, torch.Tensor)
    np.testing.assert_allclose(out_t.numpy(), out, rtol=1e-13)
    # sample moments of fresh draws match the conditional law
    np.random.seed(0)
    big = car_initial_draw_batch(roots, sigma, np.tile(X0[:1], (20000, 1)), p)[:, 2:]
    assert np.abs(big.mean(0) - (A @ X0[0] + b)).max() < 4 * np.sqrt(np.diag(V).max() / 20000)
    assert np.abs(np.cov(big.T) - V).max() < 0.05 * np.abs(V).max()


def test_rand_mat_shape_and_quirk():
    np.random.seed(0)
    z = rand_mat(10, 4)
    assert z.shape == (4, 10) and z.flags.f_contiguous
    np.random.seed(0)
    np.testing.assert_array_equal(z, np.random.randn(4, 10))
    sq = rand_mat(4, 4)                    # p == n: V V^T (reference quirk 8)
    assert np.all(np.linalg.eigvalsh(sq) > 0)


def test_small_linear_helpers():
    rng = np.random.default_rng(1)
    A = rng.standard_normal((5, 5))
    V = A @ A.T + 5 * np.eye(5)
    B = rng.standard_normal((5, 3))
    np.testing.assert_allclose(V @ solveV(V, B), B, rtol=1e-12, atol=1e-12)
    z = rng.standard_normal(5)
    mu = rng.standard_normal(5)
    np.testing.assert_allclose(norm_sim(z, mu, V), np.linalg.cholesky(V) @ z + mu, rtol=1e-13)
    icond = np.array([True, True, False, False, False])
    Ac, b, Vc = mvncond(mu, V, icond)
    np.testing.assert_allclose(Ac, V[2:, :2] @ np.linalg.inv(V[:2, :2]), rtol=1e-10)
    np.testing.assert_allclose(Vc, V[2:, 2:] - Ac @ V[:2, 2:], rtol=1e-12)
    np.testing.assert_allclose(b, mu[2:] - Ac @ mu[:2], rtol=1e-12)


def test_thin_index_matches_reference_thinning():
    """rodeo_legacy_b200.inference.thin_index against the indices the reference's own
    inference.thinning picked (tests/golden/make_golden.py)."""
    from rodeo_legacy_b200.inference import thin_index, inference
    g = load_golden("fitz_n400")
    np.testing.assert_array_equal(thin_index(g["data_t"], np.linspace(0, 40, 401)), g["thin_ind"])
    t = load_golden("thinning_n130")
    np.testing.assert_array_equal(thin_index(t["data_t"], t["ode_t"]), t["ind"])
    inf = inference([0, 3], 0, 40, None)
    X = np.arange(401.0)[:, None] * np.ones((1, 6))
    np.testing.assert_array_equal(inf.thinning(g["data_t"], np.linspace(0, 40, 401), X)[:, 0], g["thin_ind"])
    # loglike = the reference's: scale passed as `var`
    np.testing.assert_allclose(inf.loglike(np.array([0.3]), np.array([0.1]), 0.2),
                               -0.5 * np.log(2 * np.pi) - np.log(0.2) - 0.5)
