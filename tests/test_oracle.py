"""The oracle (oracle/) against golden vectors produced by the live reference
(tests/golden/make_golden.py) and the reference's one visibility known answer."""
import numpy as np
import pytest

from oracle import dynamics as od
from oracle import ukf as ou
from oracle import visibility as ov
from oracle.constants import RE


def test_propagate_single_and_chain(golden):
    g = golden("propagate")
    np.testing.assert_array_equal(od.propagate_satellite(g["single_x0"], 0, 5), g["single_out"])
    ts, x = g["chain_t"], g["chain_x"]
    y = x[:, 0]
    for k in range(1, 9):
        y = od.propagate_satellite(y, ts[k - 1], ts[k])
        np.testing.assert_array_equal(y, x[:, k])


def test_propagate_batch_shapes_and_errors(golden):
    g = golden("propagate")
    x0 = g["cat_x0"]
    np.testing.assert_array_equal(od.propagate_satellite(x0[:, ::7], 0.0, 60.0), g["cat_dt60"][:, ::7])
    np.testing.assert_array_equal(od.propagate_satellite(x0[:, 5:8], 100.0, 40.0), g["cat_back"][:, 5:8])
    out = od.propagate_satellite(x0[:, :1], 0.0, 60.0)
    assert out.shape == (6,)
    np.testing.assert_array_equal(out, g["col61_out"])
    with pytest.raises(ValueError):
        od.propagate_satellite(x0, 3.0, 3.0)


def test_transforms(golden):
    g = golden("transforms")
    for i in range(10):
        np.testing.assert_array_equal(od.lla2ecef(g["lla"][:, i]), g["lla_ecef"][:, i])
    for jd in (0, 1234, 259200):
        jdf = {0: 0.0, 1234: 1234.5, 259200: 259200.0}[jd]
        np.testing.assert_allclose(od.ecef2eci(g["x"], jdf), g[f"ecef2eci_{jd}"], rtol=0, atol=1e-11)
        np.testing.assert_allclose(od.eci2ecef(g["x"], jdf), g[f"eci2ecef_{jd}"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(od.propagate_terrestrial(g["ter_x0"], 0.0, 100.0), g["ter_dt100"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(od.propagate_terrestrial(g["ter_x0"][:, 0], 50.0, 110.0), g["ter_single"], rtol=0, atol=1e-11)
    for i in range(16):
        np.testing.assert_allclose(od.coe2eci(*g["coe"][:, i]), g["coe_eci"][:, i], rtol=1e-15, atol=1e-11)
    with pytest.raises(ValueError):
        od.ecef2eci(np.zeros((5, 3)))
    with pytest.raises(ValueError):
        od.coe2eci(RE - 1, 0, 0, 0, 0, 0)


def test_ukf_weights_and_ez(golden):
    g = golden("ukf")
    gamma, wm, wc = ou.ukf_weights()
    assert gamma == float(g["w_gamma"])
    np.testing.assert_array_equal(wm, g["w_mean"])
    np.testing.assert_array_equal(wc, g["w_cvr_diag"])
    x = np.array([8000.0, 0, 0, 0, 8, 0])
    f = ou.ez_ukf(x, 0.1, 0.1, 0.1)
    f.predict_and_update(5, x)
    np.testing.assert_array_equal(f.est_x, g["ez_est_x"])
    np.testing.assert_array_equal(f.est_p, g["ez_est_p"])
    np.testing.assert_array_equal(f.pred_p, g["ez_pred_p"])
    assert float(f.nis) == float(g["ez_nis"])


def test_ukf_orbit_rollout(golden):
    g = golden("ukf")
    f = ou.OracleUKF(0, g["orbit_x0"], g["orbit_P0"], g["orbit_Q"], g["orbit_R"])
    for k, t in enumerate(g["orbit_times"][:10]):
        z = g["orbit_z"][k]
        f.predict_and_update(t, None if np.isnan(z[0]) else z)
        np.testing.assert_array_equal(f.est_x, g["orbit_est_x"][k])
        np.testing.assert_array_equal(f.est_p, g["orbit_est_p"][k])
        np.testing.assert_array_equal(f.pred_p, g["orbit_pred_p"][k])
    with pytest.raises(ValueError):
        f.update(np.zeros((6, 1)))


def test_ukf_nearest_pd_fallback(golden):
    g = golden("ukf")
    f = ou.OracleUKF(0, g["cat_truth0"][:, 0], g["cat_P0"], g["cat_Q"], g["cat_R"])
    s = f.sigma_points_of(g["cat_truth0"][:, 0], g["npd_cov"])
    assert f.used_nearest_pd
    np.testing.assert_allclose(s, g["npd_sigma"], rtol=1e-13, atol=1e-12)


def test_visibility_known_answer():
    """tests/environment/test_env.py:31-41,132-136 of the reference: the single
    geometric known answer available for the satvis boundary."""
    sens = np.zeros((6, 3))
    sens[0] = [RE, RE + 1, RE + 2]
    targ = np.zeros((6, 4))
    targ[0] = [RE + 400, -RE - 500, RE + 600, RE + 700]
    vm = ov.calc_vis_map(sens, targ)
    assert vm.shape == (4, 3) and vm.dtype == np.int64
    np.testing.assert_array_equal(vm, [[1, 1, 1], [0, 0, 0], [1, 1, 1], [1, 1, 1]])


def test_visibility_analytic_cases():
    # antipodal -> hidden; grazing geometry: phi == a1 + a2 at the limb
    r1 = np.array([RE + 500.0, 0, 0])
    assert not ov.is_vis(r1, -r1, RE)
    a1 = np.arccos(RE / (RE + 500.0))
    a2 = np.arccos(RE / (RE + 800.0))
    for eps, want in ((-1e-6, True), (1e-6, False)):
        phi = a1 + a2 + eps
        r2 = (RE + 800.0) * np.array([np.cos(phi), np.sin(phi), 0])
        assert ov.is_vis(r1, r2, RE) is want
    v = ov.calc_vis_map(np.r_[r1, 0, 0, 0].reshape(6, 1), np.r_[r1 * 2, 0, 0, 0].reshape(6, 1), binary=False)
    assert v.dtype == np.float64 and v[0, 0] == pytest.approx(a1 + np.arccos(RE / (2 * (RE + 500.0))))
    below = np.array([RE - 1.0, 0, 0])
    assert not ov.is_vis(below, r1, RE)
    assert np.isnan(ov.visibility_func(below, r1, RE)[0])
    with pytest.raises(ValueError):
        ov.calc_vis_map(np.zeros((5, 2)), np.zeros((6, 2)))


def test_radial_rate_and_vis_derivative(golden):
    """Oracle restatements behind calcVisMapAndDerivative: radial rate pinned against the live
    reference (orbits.py:39-69), derivative checked against a central difference of the visibility
    function along two-body motion (the satvis formula itself is unpinned)."""
    from oracle.dynamics import propagate_satellite
    from oracle.visibility import calc_vis_and_der_vis, calc_vis_map_and_derivative, radial_rate, visibility_func
    g = golden("orbits")
    X = g["x"]
    got = np.array([radial_rate(X[:3, i], X[3:, i]) for i in range(X.shape[1])])
    np.testing.assert_allclose(got, g["rdot"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(got, (X[:3] * X[3:]).sum(0) / np.linalg.norm(X[:3], axis=0), rtol=0, atol=1e-13)
    RE, h = 6378.1363, 0.5
    for a, b in ((3, 17), (5, 40), (11, 29)):
        xa, xb = X[:, a], X[:, b]
        v0, d0 = calc_vis_and_der_vis(xa[:3], xa[3:], radial_rate(xa[:3], xa[3:]), xb[:3], xb[3:],
                                      radial_rate(xb[:3], xb[3:]), RE)
        vp = visibility_func(propagate_satellite(xa, 0, h)[:3], propagate_satellite(xb, 0, h)[:3], RE)[0]
        vm = visibility_func(propagate_satellite(xa, 0, -h)[:3], propagate_satellite(xb, 0, -h)[:3], RE)[0]
        assert d0 == pytest.approx((vp - vm) / (2 * h), rel=1e-6, abs=1e-12)
        assert v0 == visibility_func(xa[:3], xb[:3], RE, 0.0, 1e-6)[0]
    vm, dm = calc_vis_map_and_derivative(X[:, 48:52], X[:, :6])
    assert vm.shape == dm.shape == (6, 4) and set(np.unique(vm)) <= {0, 1}


def test_ez_ukf_parameter_forms_without_gpu():
    """ezUKF / getRandomParams (ez_ukf.py:17-119): float -> v I6, (6,6) ndarray as given, {dist, params, seed}
    -> diag of numpy's default_rng draw; time defaults to 0; weights as in ukf_v2.py:58-128.  Building a
    filter touches no device."""
    import numpy as np
    import pytest
    from numpy.random import default_rng
    from clockpunch_b200.dynamics.dynamics_classes import SatDynamicsModel, StaticTerrestrial
    from clockpunch_b200.estimation.ez_ukf import ezUKF, getRandomParams
    x = np.array([8000.0, 0, 0, 0, 8, 0])
    spec = {"dist": "uniform", "params": [[0.1] * 6, [1.0] * 6], "seed": 2}
    R = np.diag([1, 1, 1, 0.01, 0.01, 0.01])
    f = ezUKF({"x_init": x, "p_init": 0.1, "dynamics_type": "satellite", "Q": spec, "R": R})
    np.testing.assert_array_equal(f.est_p, 0.1 * np.eye(6))
    np.testing.assert_array_equal(f.q_matrix, np.diag(default_rng(2).uniform([0.1] * 6, [1.0] * 6)))
    np.testing.assert_array_equal(f.r_matrix, R)
    assert f.time == 0 and f.x_dim == 6 and f.y_dim == 6 and isinstance(f.dynamics, SatDynamicsModel)
    n, alpha, beta, kappa = 6, 1e-3, 2.0, 3 - 6
    lam = alpha**2 * (n + kappa) - n
    assert f.gamma == np.sqrt(n + lam)
    assert f.mean_weight[0] == lam / (n + lam) and f.mean_weight[1] == 1 / (2 * (n + lam))
    assert f.cvr_weight[0, 0] == lam / (n + lam) + 1 - alpha**2 + beta
    g = ezUKF({"time": 7, "x_init": x, "p_init": np.eye(6), "dynamics_type": "satellite", "Q": 2, "R": 0.5})
    assert g.time == 7
    np.testing.assert_array_equal(g.q_matrix, 2 * np.eye(6))
    np.testing.assert_array_equal(g.r_matrix, 0.5 * np.eye(6))
    # a filter on Earth-fixed dynamics is not on the hot path: it fails loudly instead of running on the CPU
    with pytest.raises(NotImplementedError):
        ezUKF({"x_init": x, "p_init": 1, "dynamics_type": "terrestrial", "Q": 2, "R": 0.5})
    assert StaticTerrestrial is not None
    np.testing.assert_array_equal(getRandomParams("normal", [[0, 0], [1, 2]], seed=5), default_rng(5).normal([0, 0], [1, 2]))
    with pytest.raises(AssertionError):
        ezUKF({"x_init": x, "p_init": 1, "dynamics_type": "satellite", "Q": {"dist": "uniform", "params": [[0] * 5, [1] * 5]}, "R": 1})
