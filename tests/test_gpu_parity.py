"""Parity of the CUDA path (through the C-ABI) against the oracle and the golden vectors
from the live reference.  Tolerances are the ones SURVEY.md section 8d / BASELINE.md state:
truth states <= 1e-9 km, vis bits identical outside |V| <= 1e-9 rad, est_p / pred_p <= 1e-8
max-relative and est_x <= 1e-5 km after one step with the same injected measurement."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

POS_TOL = 1e-9      # km
VEL_TOL = 1e-11     # km/s
COV_RTOL = 1e-8
ESTX_TOL = 1e-5     # km (and km/s, looser than needed for velocity) for LEO; see estx_tol()
GRAZE = 1e-9        # rad


def estx_tol(x):
    """The reference forms pred_x = sum_i w_i X_i with w_0 = -2e6: its own float64 round-off is
    ~ulp(2e6 |x|) = 4.4e-10 |x|, i.e. 3e-6 km in LEO and 2e-5 km at GEO (measured spread between
    two runs of the reference itself).  Parity on est_x after a measurement update can
    only be asked for to that level: max(1e-5 km, 1e-9 |r|)."""
    return np.maximum(ESTX_TOL, 1e-9 * np.linalg.norm(np.asarray(x)[:3], axis=0))


def relmax(a, b):
    return np.max(np.abs(a - b)) / np.max(np.abs(b))


def cov_close(a, b, rtol=COV_RTOL):
    scale = np.sqrt(np.abs(np.einsum("...ii,...jj->...ij", b, b)))     # sqrt(b_ii b_jj)
    return np.max(np.abs(a - b) / scale) <= rtol


@pytest.fixture(scope="module")
def dyn():
    from clockpunch_b200.dynamics.dynamics_classes import SatDynamicsModel, StaticTerrestrial
    return SatDynamicsModel(), StaticTerrestrial()


def test_propagate_golden(golden, dyn):
    sat, _ = dyn
    g = golden("propagate")
    out = sat.propagate(g["single_x0"], 0, 5)
    assert out.shape == (6,)
    assert np.abs(out - g["single_out"])[:3].max() <= POS_TOL and np.abs(out - g["single_out"])[3:].max() <= VEL_TOL
    x0 = g["cat_x0"]
    for key, t0, tf in (("cat_dt60", 0.0, 60.0), ("cat_dt100", 0.0, 100.0), ("cat_back", 100.0, 40.0)):
        out = sat.propagate(x0, t0, tf)
        assert out.shape == x0.shape
        err = np.abs(out - g[key])
        assert err[:3].max() <= POS_TOL, (key, err[:3].max())
        assert err[3:].max() <= VEL_TOL, (key, err[3:].max())
    # dt = 600 s: the reference's adaptive DOP853 (rtol = atol = 1e-10) is itself 1e-8..1e-7 km off
    # the 30-digit truth for LEO/HEO, so parity is its error band; against the truth the fixed-step
    # kernel must sit at the FP64 round-off floor and be at least as close as the reference.
    out = sat.propagate(x0, 0.0, 600.0)
    ref_err = np.abs(g["cat_dt600"] - g["cat_dt600_truth"])[:3].max()
    assert np.abs(out - g["cat_dt600"])[:3].max() <= 2 * ref_err + POS_TOL
    assert np.abs(out - g["cat_dt600_truth"])[:3].max() <= POS_TOL
    assert np.abs(out - g["cat_dt600_truth"])[:3].max() <= ref_err
    out = sat.propagate(x0, 0.0, 100.0)
    assert np.abs(out - g["cat_dt100_truth"])[:3].max() <= 1e-10
    assert np.abs(out - g["cat_dt100_truth"])[3:].max() <= 1e-12
    assert sat.propagate(x0[:, :1], 0.0, 60.0).shape == (6,)
    with pytest.raises(ValueError):
        sat.propagate(x0, 5.0, 5.0)
    # 49-step chain of the reference's test_dynamics_classes.py:143-149
    ts, xs = g["chain_t"], g["chain_x"]
    y = xs[:, 0]
    for k in range(1, len(ts)):
        y = sat.propagate(y, ts[k - 1], ts[k])
    # 48 chained calls over 1.7 orbits: the reference accumulates its own 1e-10-relative step errors,
    # so parity is its drift band; against the closed-form truth the kernel is far closer.
    ref_drift = np.abs(xs[:, -1] - g["chain_truth_end"])[:3].max()
    assert np.abs(y - xs[:, -1])[:3].max() <= 2 * ref_drift + POS_TOL
    assert np.abs(y - g["chain_truth_end"])[:3].max() <= 1e-8
    assert np.abs(y - g["chain_truth_end"])[:3].max() <= ref_drift


def test_propagate_substep_override_and_j2(dyn):
    from clockpunch_b200.dynamics.propagator import propagate_batch
    from clockpunch_b200 import _lib
    from oracle.dynamics import propagate_satellite
    x = np.array([[7000.0, 100.0, 50.0, 0.1, 7.4, 1.0]]).T
    ref = propagate_satellite(x[:, 0], 0.0, 100.0)
    outs = [propagate_batch(x, 0.0, 100.0, substeps=ns)[:, 0] for ns in (0, 1, 3, 8)]
    for out in outs:
        assert np.abs(out - ref)[:3].max() <= POS_TOL and np.abs(out - ref)[3:].max() <= VEL_TOL
    assert np.abs(outs[2] - outs[3])[:3].max() <= 1e-11       # 3 vs 8 substeps: both converged
    assert np.abs(outs[0] - outs[1])[:3].max() <= 1e-10       # auto rule vs one forced substep
    # J2 is an opt-in extension with no reference counterpart: check against scipy DOP853
    from scipy.integrate import solve_ivp
    MU, RE, J2 = 398600.0, 6378.1363, 1.08262668e-3

    def rhs(t, s):
        r = s[:3]; rm = np.linalg.norm(r); z2 = 5 * r[2] ** 2 / rm**2
        k = 1.5 * J2 * (RE / rm) ** 2
        a = -MU * r / rm**3 * (1 + k * np.array([1 - z2, 1 - z2, 3 - z2]))
        return np.r_[s[3:], a]
    sol = solve_ivp(rhs, [0, 300.0], x[:, 0], method="DOP853", rtol=1e-13, atol=1e-13)
    out = propagate_batch(x, 0.0, 300.0, force_model=_lib.FORCE_TWOBODY_J2)[:, 0]
    assert np.abs(out - sol.y[:, -1])[:3].max() <= 1e-8
    tb = propagate_batch(x, 0.0, 300.0)[:, 0]
    assert np.abs(out - tb)[:3].max() > 1e-3            # J2 really is on


def test_step_with_j2_and_space_sensors():
    """The bench's `c2_j2_space_sensors` line (BASELINE.json configs[1] as worded): the FUSED step with the J2 force
    model and sensors in orbit.  No reference counterpart, so the step is checked against the library's own
    stand-alone kernels, which test_propagate_substep_override_and_j2 pins to scipy: truth states and sensors equal
    propagate_batch(J2) bit for bit, an unmeasured target's estimate is its J2-propagated centre point, the true
    visibility map is the map of those states, and a measured target moves towards its measurement."""
    from clockpunch_b200 import _lib
    from clockpunch_b200.common.visibility import calcVisMap
    from clockpunch_b200.dynamics.propagator import propagate_batch
    from clockpunch_b200.engine import CatalogEngine
    from clockpunch_b200.simulation.catalog import synth_catalog
    cat = synth_catalog(3000, 40, seed=99, space_sensor_frac=0.25)
    assert (cat.sensor_kinds == _lib.KIND_SATELLITE).sum() == 10
    eng = CatalogEngine(cat.sensors, cat.sensor_kinds, cat.targets, cat.est_x, cat.est_p, cat.Q, cat.R, seed=3,
                        force_model=_lib.FORCE_TWOBODY_J2)
    two = CatalogEngine(cat.sensors, cat.sensor_kinds, cat.targets, cat.est_x, cat.est_p, cat.Q, cat.R, seed=3)
    tasked = np.zeros(cat.N, np.uint8); tasked[::50] = 1
    z = propagate_batch(cat.targets, 0.0, 100.0, force_model=_lib.FORCE_TWOBODY_J2)
    eng.step(100.0, tasked=tasked, z=z)
    two.step(100.0, tasked=tasked, z=z)
    x1, s1 = eng.host("truth_x"), eng.host("sens_x")
    np.testing.assert_array_equal(x1, z)
    np.testing.assert_array_equal(s1, propagate_batch(cat.sensors, 0.0, 100.0, kind=cat.sensor_kinds, force_model=_lib.FORCE_TWOBODY_J2))
    assert np.abs(x1 - two.host("truth_x"))[:3].max() > 1e-3          # J2 really is on in the step
    un = tasked == 0
    ex = eng.host("est_x")
    np.testing.assert_array_equal(ex[:, un], propagate_batch(cat.est_x, 0.0, 100.0, force_model=_lib.FORCE_TWOBODY_J2)[:, un])
    vt = eng.host("vis_truth")
    np.testing.assert_array_equal(vt, calcVisMap(s1, x1))
    # measured targets: pulled from the (noisy) prediction to the exact measurement
    d0 = np.linalg.norm((propagate_batch(cat.est_x, 0.0, 100.0, force_model=_lib.FORCE_TWOBODY_J2) - z)[:3, ~un], axis=0)
    d1 = np.linalg.norm((ex - z)[:3, ~un], axis=0)
    assert (d1 < d0).all() and d1.mean() < 0.5 * d0.mean()
    assert (eng.host("flags") & ~np.uint32(_lib.FLAG_SUBSTEP_CLAMP) == 0).all()


def test_transforms_golden(golden, dyn):
    from clockpunch_b200.common import transforms as T
    _, ter = dyn
    g = golden("transforms")
    for i in range(10):
        np.testing.assert_allclose(T.lla2ecef(g["lla"][:, i]), g["lla_ecef"][:, i], rtol=0, atol=1e-11)
    for tag, jd in ((0, 0.0), (1234, 1234.5), (259200, 259200.0)):
        np.testing.assert_allclose(T.ecef2eci(g["x"], jd), g[f"ecef2eci_{tag}"], rtol=0, atol=2e-11)
        np.testing.assert_allclose(T.eci2ecef(g["x"], jd), g[f"eci2ecef_{tag}"], rtol=0, atol=2e-11)
    assert T.ecef2eci(g["x"][:, :1], 3.0).shape == (6,)
    np.testing.assert_allclose(ter.propagate(g["ter_x0"], 0.0, 100.0), g["ter_dt100"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(ter.propagate(ter.propagate(g["ter_x0"], 0.0, 100.0), 100.0, 4000.0),
                               g["ter_chain"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(ter.propagate(g["ter_x0"][:, 0], 50.0, 110.0), g["ter_single"], rtol=0, atol=1e-11)
    for i in range(16):
        np.testing.assert_allclose(T.coe2eci(*g["coe"][:, i]), g["coe_eci"][:, i], rtol=1e-14, atol=1e-10)
    np.testing.assert_allclose(T.coe2eci_batch(g["coe"]), g["coe_eci"], rtol=1e-14, atol=1e-10)


def _v_values(sens, targ, RE=6378.1363):
    from oracle.visibility import visibility_func
    return np.array([[visibility_func(sens[:3, j], targ[:3, i], RE)[0] for j in range(sens.shape[1])]
                     for i in range(targ.shape[1])])


def test_vismap_known_answer_and_oracle():
    from clockpunch_b200.common.visibility import calcVisMap
    from clockpunch_b200.simulation.catalog import synth_catalog
    from oracle.visibility import calc_vis_map
    RE = 6378.1363
    sens = np.zeros((6, 3)); sens[0] = [RE, RE + 1, RE + 2]
    targ = np.zeros((6, 4)); targ[0] = [RE + 400, -RE - 500, RE + 600, RE + 700]
    vm = calcVisMap(sens, targ)
    assert vm.shape == (4, 3) and vm.dtype == np.int64
    np.testing.assert_array_equal(vm, [[1, 1, 1], [0, 0, 0], [1, 1, 1], [1, 1, 1]])   # test_env.py:132-136
    # ragged sizes around the 32-sensor word boundary and the 128-thread CTA boundary
    for (n, m, seed) in ((1, 1, 1), (5, 31, 2), (7, 32, 3), (33, 33, 4), (130, 65, 5), (257, 100, 6)):
        cat = synth_catalog(n, m, seed=seed, space_sensor_frac=0.3)
        got = calcVisMap(cat.sensors, cat.targets)
        want = calc_vis_map(cat.sensors, cat.targets)
        V = _v_values(cat.sensors, cat.targets)
        bad = (got != want) & (np.abs(V) > GRAZE)
        assert not bad.any(), (n, m, int(bad.sum()))
        cont = calcVisMap(cat.sensors, cat.targets, binary=False)
        assert cont.dtype == np.float64 and cont.shape == (n, m)
        np.testing.assert_allclose(cont, calc_vis_map(cat.sensors, cat.targets, binary=False), rtol=0, atol=1e-12)
    # 1-D inputs and below-surface agents
    one = calcVisMap(sens[:, 0], targ[:, 0])
    assert one.shape == (1, 1) and one[0, 0] == 1
    under = targ.copy(); under[0, 0] = RE - 5.0
    assert calcVisMap(sens, under)[0].sum() == 0
    assert np.isnan(calcVisMap(sens, under, binary=False)[0]).all()


def test_vismap_analytic_cases_on_device():
    """The analytic cases of SURVEY.md section 8c (the satvis boundary is unpinned, so geometry is the
    anchor), through the standalone kernel AND through the visibility rows the fused step writes:
    co-linear on the same side => V = a1 + a2 > 0 (visible); antipodal => hidden; at the limb
    phi = a1 + a2 exactly: visible just inside, hidden just outside, for offsets from 1e-6 down to the
    stated grazing band of 1e-9 rad; below the surface => never visible, NaN in continuous mode."""
    from clockpunch_b200.common.visibility import calcVisMap
    from clockpunch_b200.engine import CatalogEngine
    RE = 6378.1363
    r1 = RE + 500.0
    a1 = np.arccos(RE / r1)
    sens = np.zeros((6, 1)); sens[0, 0] = r1
    rows, want, cont = [], [], []
    for r2 in (RE + 0.001, RE + 800.0, 42164.0):
        a2 = np.arccos(RE / r2)
        rows.append([2 * r2, 0.0, 0.0]); want.append(1); cont.append(a1 + np.arccos(RE / (2 * r2)))   # co-linear, same side
        rows.append([-r2, 0.0, 0.0]); want.append(0); cont.append(a1 + a2 - np.pi)                     # antipodal
        for eps in (1e-6, 1e-7, 1e-8, 2e-9):
            for sign in (-1.0, 1.0):
                phi = a1 + a2 + sign * eps
                rows.append([r2 * np.cos(phi), r2 * np.sin(phi), 0.0]); want.append(int(sign < 0)); cont.append(-sign * eps)
    targ = np.zeros((6, len(rows))); targ[:3] = np.array(rows).T
    got = calcVisMap(sens, targ)
    np.testing.assert_array_equal(got[:, 0], want)
    val = calcVisMap(sens, targ, binary=False)[:, 0]
    np.testing.assert_allclose(val, cont, rtol=0, atol=1e-11)
    # the same pairs through the fused step: circular-orbit velocities, one step, rows recomputed from the
    # states the step itself produced must equal what it wrote (both maps), and the analytic answers hold
    # for a zero-length-equivalent check via compute_vis() on the initial states
    v = np.sqrt(398600.0 / np.linalg.norm(targ[:3], axis=0))
    tx = targ.copy(); tx[5] = v                                   # velocity along +z: any bound orbit will do
    sx = sens.copy(); sx[4, 0] = np.sqrt(398600.0 / r1)
    n = tx.shape[1]
    P0 = np.diag([1.0, 1, 1, 1e-2, 1e-2, 1e-2])
    eng = CatalogEngine(sx, np.zeros(1, np.uint8), tx, tx.copy(), np.broadcast_to(P0, (n, 6, 6)).copy(), 0.1 * P0, P0)
    np.testing.assert_array_equal(eng.host("vis_truth")[:, 0], want)          # reset path (pc_vismap_packed)
    np.testing.assert_array_equal(eng.host("vis_est")[:, 0], want)
    eng.step(dt=30.0)
    np.testing.assert_array_equal(eng.host("vis_truth"), calcVisMap(eng.host("sens_x"), eng.host("truth_x")))
    np.testing.assert_array_equal(eng.host("vis_est"), calcVisMap(eng.host("sens_x"), eng.host("est_x")))
    # below the surface
    under = targ[:, :1].copy(); under[:3, 0] = [RE - 1.0, 0, 0]
    assert calcVisMap(sens, under)[0, 0] == 0 and np.isnan(calcVisMap(sens, under, binary=False)[0, 0])


def _golden_step_inputs(g, mode, dt):
    key = f"cat_{mode}_dt{dt}"
    est_x, est_p, pred_p = g[key + "_est_x"], g[key + "_est_p"], g[key + "_pred_p"]
    z, nis = g[key + "_z"], g[key + "_nis"]
    return est_x, est_p, pred_p, z, nis


def test_ukf_one_step_golden(golden):
    """Every stored step of the reference rollouts, restarted from the reference's own
    previous state: 16 targets (LEO/MEO/GEO/HEO) x 4 steps x {tasked, untasked, mixed}."""
    from clockpunch_b200 import _lib
    from clockpunch_b200.estimation.ukf_v2 import ukf_batch_host
    g = golden("ukf")
    params = _lib.make_ukf_params(g["cat_Q"], g["cat_R"])
    worst = dict(est_x=0.0, est_p=0.0, pred_p=0.0)
    for mode in ("tasked", "untasked", "mixed"):
        for dt in (60, 100):
            est_x, est_p, pred_p, z, nis = _golden_step_inputs(g, mode, dt)
            nt, ns = est_x.shape[:2]
            for s in range(ns):
                x_in = g["cat_est0"] if s == 0 else est_x[:, s - 1].T
                p_in = np.broadcast_to(g["cat_P0"], (nt, 6, 6)) if s == 0 else est_p[:, s - 1]
                tasked = ~np.isnan(z[:, s, 0])
                zz = np.where(np.isnan(z[:, s]), 0.0, z[:, s]).T
                out = ukf_batch_host(x_in, p_in, zz, tasked, s * float(dt), (s + 1) * float(dt), params)
                assert (out["flags"] == 0).all()
                assert cov_close(out["pred_p"], pred_p[:, s]), (mode, dt, s)
                assert cov_close(out["est_p"], est_p[:, s]), (mode, dt, s)
                dev = np.abs(out["est_x"] - est_x[:, s].T).max(axis=0)
                assert (dev <= estx_tol(est_x[:, s].T)).all(), (mode, dt, s, dev.max())
                assert (dev[~tasked] <= 1e-9).all()        # no measurement: est_x is a plain propagation
                ex = dev.max()
                np.testing.assert_allclose(out["nis"][tasked], nis[tasked, s], rtol=1e-4)  # innovation carries the 1e-5 km pred_x noise
                assert np.isnan(out["nis"][~tasked]).all()
                worst["est_x"] = max(worst["est_x"], ex)
    print("worst est_x deviation [km]:", worst["est_x"])


def test_ukf_object_api_golden(golden):
    """The reference's own scenarios through the mirrored class API
    (tests/estimation/test_ez_ukf.py:25-40, test_ukf_v2.py:254-313,408-478)."""
    from clockpunch_b200.dynamics.dynamics_classes import SatDynamicsModel
    from clockpunch_b200.estimation.ez_ukf import ezUKF
    from clockpunch_b200.estimation.ukf_v2 import UnscentedKalmanFilter
    g = golden("ukf")
    x = np.array([8000.0, 0, 0, 0, 8, 0])
    f = ezUKF({"x_init": x, "p_init": 0.1, "dynamics_type": "satellite", "Q": 0.1, "R": 0.1})
    assert f.gamma == float(g["w_gamma"]) and np.array_equal(f.mean_weight, g["w_mean"])
    assert np.array_equal(np.diag(f.cvr_weight), g["w_cvr_diag"])
    f.predictAndUpdate(5, x)
    assert np.abs(f.est_x - g["ez_est_x"]).max() <= ESTX_TOL
    assert cov_close(f.est_p, g["ez_est_p"]) and cov_close(f.pred_p, g["ez_pred_p"])
    assert f.nis == pytest.approx(float(g["ez_nis"]), rel=1e-6)
    assert f.time == 5 and f.last_measurement_time == 5
    f = UnscentedKalmanFilter(0, g["orbit_x0"], g["orbit_P0"], SatDynamicsModel(), lambda s: s,
                              g["orbit_Q"], g["orbit_R"])
    for k, t in enumerate(g["orbit_times"]):
        z = g["orbit_z"][k]
        f.predictAndUpdate(t, None if np.isnan(z[0]) else z)
        if k < 4:                                           # before the first measurement: tight
            assert np.abs(f.est_x - g["orbit_est_x"][k]).max() <= 1e-8
            assert cov_close(f.est_p, g["orbit_est_p"][k], 1e-8 * (k + 1))
    # 25-step rollout drift is reported against the reference's own noise floor, loosely gated
    assert np.abs(f.est_x - g["orbit_est_x"][-1])[:3].max() <= 1e-3
    assert cov_close(f.est_p, g["orbit_est_p"][-1], 1e-6)
    # split predict() / update() equals predictAndUpdate(); reset() restores
    f.reset()
    f.predict(100.0)
    assert cov_close(f.pred_p, g["orbit_pred_p"][0]) and f.time == 100.0
    np.testing.assert_array_equal(f.est_x, g["orbit_x0"])
    f.update(None)
    assert np.abs(f.est_x - g["orbit_est_x"][0]).max() <= 1e-8
    with pytest.raises(ValueError):
        f.predictAndUpdate(200.0, np.zeros((6, 1)))
    with pytest.raises(NotImplementedError):
        UnscentedKalmanFilter(0, np.zeros(2), np.eye(2), SatDynamicsModel(), lambda s: s, np.eye(2), np.eye(2))


def test_ukf_nonpd_is_flagged_and_repaired(golden):
    """A non-PD covariance is DATA (flag), and -- like the reference's nearest-PD fallback
    (ukf_v2.py:170-176, ukf_utils.py:155-198) -- the filter keeps going on the repaired factor."""
    import ctypes as C
    import torch
    from clockpunch_b200 import _lib, _buffers as B
    from clockpunch_b200.estimation.ukf_v2 import ukf_batch_host
    from oracle.ukf import OracleUKF
    g = golden("ukf")
    params = _lib.make_ukf_params(g["cat_Q"], g["cat_R"])
    x = g["cat_truth0"][:, :3]
    p = np.broadcast_to(g["cat_P0"], (3, 6, 6)).copy()
    p[1] = g["npd_cov"]
    # sigma points of the repaired factor against the LIVE reference's own (generateSigmaPoints on
    # the same matrix, tests/golden/make_golden.py): the repaired direction agrees to ~1e-10 km
    ctx = _lib.get_ctx()
    dev = B.device_of(ctx)
    dx = torch.from_numpy(np.ascontiguousarray(g["cat_truth0"][:, :1])).to(dev)
    dp = torch.from_numpy(g["npd_cov"].reshape(1, 36).copy()).to(dev)
    ws = torch.zeros((14, 6, 1), dtype=torch.float64, device=dev)
    fl = torch.zeros((2, 1), dtype=torch.int32, device=dev)
    ctx.check(ctx.lib.pc_sigma_points(ctx.handle, dx.data_ptr(), dp.data_ptr(), 1, 1, params.gamma, ws.data_ptr(),
                                      fl.data_ptr(), None))
    sig = ws[:13, :, 0].cpu().numpy().T
    assert fl[0, 0].item() & _lib.FLAG_NONPD_EST
    np.testing.assert_allclose(sig, g["npd_sigma"], rtol=0, atol=1e-8)
    # one full predict step: flagged, finite, and equal to the oracle (which took its nearest-PD path)
    out = ukf_batch_host(x, p, None, None, 0.0, 60.0, params)
    assert out["flags"][1] & _lib.FLAG_NONPD_EST and not out["flags"][1] & _lib.FLAG_NONFINITE
    assert out["flags"][0] == 0 and out["flags"][2] == 0
    assert np.isfinite(out["est_p"]).all() and np.isfinite(out["est_x"]).all()
    f = OracleUKF(0.0, x[:, 1], p[1], g["cat_Q"], g["cat_R"])
    f.predict_and_update(60.0, None)
    assert f.used_nearest_pd
    assert np.abs(out["est_x"][:, 1] - f.est_x).max() <= 1e-8
    assert cov_close(out["pred_p"][1], f.pred_p, 1e-7)
    # NaN input: flagged non-finite, neighbours untouched
    p2 = p.copy(); p2[1, 2, 2] = np.nan
    out = ukf_batch_host(x, p2, None, None, 0.0, 60.0, params)
    assert out["flags"][1] & _lib.FLAG_NONFINITE and out["flags"][0] == 0 and out["flags"][2] == 0


def _engine_and_oracle(n, m, seed, space_frac=0.25):
    from clockpunch_b200.engine import CatalogEngine
    from clockpunch_b200.simulation.catalog import synth_catalog
    from oracle.step import OracleCatalog
    cat = synth_catalog(n, m, seed=seed, space_sensor_frac=space_frac)
    eng = CatalogEngine(cat.sensors, cat.sensor_kinds, cat.targets, cat.est_x, cat.est_p, cat.Q, cat.R)
    kinds = ["terrestrial" if k else "satellite" for k in cat.sensor_kinds]
    orc = OracleCatalog(cat.sensors, kinds, cat.targets, cat.est_x, cat.est_p, cat.Q, cat.R)
    return cat, eng, orc


def test_engine_steps_match_oracle():
    """Full fused step (sensors, truth, UKF, both vis maps, tasking bookkeeping) vs the oracle's
    restatement of SSAScheduler.step on the same inputs and injected measurements."""
    cat, eng, orc = _engine_and_oracle(21, 5, seed=11)
    rng = np.random.default_rng(5)
    vt0, ve0 = orc.vis_maps()
    np.testing.assert_array_equal(eng.host("vis_truth"), vt0)
    for s in range(4):
        t1 = 100.0 * (s + 1)
        actions = rng.integers(0, cat.N + 1, size=cat.M)
        ve_prev = orc.vis_maps()[1]
        tasked = np.zeros(cat.N, bool)
        for j, a in enumerate(actions):
            if a < cat.N and ve_prev[a, j]:
                tasked[a] = True
        got_mask = eng.task_from_actions(actions).cpu().numpy().astype(bool)
        np.testing.assert_array_equal(got_mask, tasked)
        truth_next = np.stack([__import__("oracle.dynamics", fromlist=["x"]).propagate_satellite(orc.targets[:, i], orc.time, t1)
                               for i in range(cat.N)], axis=1)
        z = truth_next + rng.normal(size=(6, cat.N)) * np.sqrt(np.diag(cat.R))[:, None]
        vt, ve = orc.step(t1, tasked, z)
        eng.step(t1, tasked=eng.tasked, z=z)
        assert np.abs(eng.host("truth_x") - orc.targets)[:3].max() <= POS_TOL * (s + 1)
        assert np.abs(eng.host("sens_x") - orc.sensors)[:3].max() <= POS_TOL * (s + 1)
        assert (np.abs(eng.host("est_x") - orc.est_x).max(axis=0) <= estx_tol(orc.est_x) * (s + 1)).all()
        assert cov_close(eng.host("est_p"), orc.est_p, COV_RTOL * 10 * (s + 1))
        assert cov_close(eng.host("pred_p"), orc.pred_p, COV_RTOL * 10 * (s + 1))
        Vt = _v_values(orc.sensors, orc.targets)
        Ve = _v_values(orc.sensors, orc.est_x)
        assert not ((eng.host("vis_truth") != vt) & (np.abs(Vt) > 1e-7)).any()
        assert not ((eng.host("vis_est") != ve) & (np.abs(Ve) > 1e-7)).any()
        np.testing.assert_array_equal(eng.host("num_tasked"), orc.num_tasked)
        np.testing.assert_array_equal(eng.host("last_time_tasked"), orc.last_time_tasked)
    assert (eng.host("flags") == 0).all()
    # the same steps through a captured CUDA graph (device clock, actions buffer) give the same results
    eng.reset(rewind_rng=True)
    rng = np.random.default_rng(5)
    direct = {}
    acts = [rng.integers(0, cat.N + 1, size=cat.M) for _ in range(3)]
    for s, a in enumerate(acts):
        eng.task_from_actions(a)
        eng.step(100.0 * (s + 1), tasked=eng.tasked, z=None)
    for k in ("truth_x", "est_x", "est_p", "vis_est", "num_tasked", "last_time_tasked"):
        direct[k] = eng.host(k)
    eng.reset(rewind_rng=True)
    gr = eng.build_graph(100.0)
    for a in acts:
        eng.actions[: cat.M].copy_(__import__("torch").from_numpy(a))
        eng.replay(gr)
    for k, v in direct.items():
        if k in ("est_x", "est_p"):
            # with an actions buffer the few tasked targets are updated by dedicated warps (pc_step_args.task_of):
            # same equations as the in-line update of the mask path, sums in a different order
            np.testing.assert_allclose(eng.host(k), v, rtol=1e-11, atol=1e-13, err_msg=k)
        else:
            np.testing.assert_array_equal(eng.host(k), v, err_msg=k)
    assert eng.time == 300.0
    # reset restores the construction-time snapshot exactly
    eng.reset()
    np.testing.assert_array_equal(eng.host("truth_x"), cat.targets)
    np.testing.assert_array_equal(eng.host("vis_truth"), vt0)
    assert eng.time == 0.0 and eng.host("num_tasked").sum() == 0


def test_engine_single_steps_restarted_from_oracle_state_meet_stated_gates():
    """The multi-step test above lets both sides run freely, so their differences compound (the +-2e6
    sigma weights amplify 1e-12 km of integrator noise) and its tolerances grow
    with the step count.  Here every step is RESTARTED from the oracle's state -- a fresh engine on the
    oracle's (truth, est_x, est_p) -- and must meet the one-step gates exactly as stated in SURVEY.md
    section 8d: truth 1e-9 km, est_p / pred_p 1e-8 relative, est_x max(1e-5 km, 1e-9 |r|) after an update
    and 1e-9 km without, visibility bits identical outside 1e-9 rad of grazing."""
    from clockpunch_b200.engine import CatalogEngine
    from oracle.dynamics import propagate_satellite
    cat, _, orc = _engine_and_oracle(21, 5, seed=11)
    rng = np.random.default_rng(6)
    for s in range(4):
        t0, t1 = orc.time, orc.time + 100.0
        eng = CatalogEngine(orc.sensors, cat.sensor_kinds, orc.targets, orc.est_x, orc.est_p, cat.Q, cat.R, time=t0)
        np.testing.assert_array_equal(eng.host("vis_est"), orc.vis_maps()[1])
        tasked = rng.random(cat.N) < 0.4
        truth_next = np.stack([propagate_satellite(orc.targets[:, i], t0, t1) for i in range(cat.N)], axis=1)
        z = truth_next + rng.normal(size=(6, cat.N)) * np.sqrt(np.diag(cat.R))[:, None]
        vt, ve = orc.step(t1, tasked, z)
        eng.step(t1, tasked=tasked.astype(np.uint8), z=z)
        assert np.abs(eng.host("truth_x") - orc.targets)[:3].max() <= POS_TOL
        assert np.abs(eng.host("truth_x") - orc.targets)[3:].max() <= VEL_TOL
        assert np.abs(eng.host("sens_x") - orc.sensors)[:3].max() <= POS_TOL
        dx = np.abs(eng.host("est_x") - orc.est_x).max(axis=0)
        assert (dx[tasked] <= estx_tol(orc.est_x[:, tasked])).all(), dx[tasked].max()
        assert (dx[~tasked] <= POS_TOL).all(), dx[~tasked].max()
        assert cov_close(eng.host("est_p"), orc.est_p, COV_RTOL)
        assert cov_close(eng.host("pred_p"), orc.pred_p, COV_RTOL)
        Vt, Ve = _v_values(orc.sensors, orc.targets), _v_values(orc.sensors, orc.est_x)
        assert not ((eng.host("vis_truth") != vt) & (np.abs(Vt) > GRAZE)).any()
        # the estimated map is evaluated at est_x, which the two sides hold to estx_tol (1e-5 km of a
        # ~1e4 km range = 1e-9 rad): the band is the stated one plus that angle
        assert not ((eng.host("vis_est") != ve) & (np.abs(Ve) > GRAZE + 2e-9)).any()


def test_device_measurement_draw_is_distributionally_right():
    """agents.py:173-183 draws z ~ N(truth, R) from numpy's global RNG; the device path draws
    with Philox: check moments, determinism per (seed, step) and that tasked targets move."""
    from clockpunch_b200 import _lib, _buffers as B
    from clockpunch_b200.simulation.catalog import synth_catalog
    import ctypes as C, torch
    cat = synth_catalog(20000, 1, seed=21)
    ctx = _lib.get_ctx()
    params = _lib.make_ukf_params(cat.Q, cat.R)
    dev = B.device_of(ctx)
    n = cat.N

    def run(seed, step):
        ex = torch.from_numpy(cat.est_x).to(dev); ep = torch.from_numpy(cat.est_p.reshape(n, 36)).to(dev)
        tx = torch.from_numpy(cat.targets).to(dev); zo = torch.zeros((6, n), dtype=torch.float64, device=dev)
        tk = torch.ones(n, dtype=torch.uint8, device=dev)
        ctx.check(ctx.lib.pc_ukf_predict_update(ctx.handle, ex.data_ptr(), ep.data_ptr(), tx.data_ptr(), None,
                                                tk.data_ptr(), n, n, 0.0, 60.0, 0, 0, C.byref(params), seed, step,
                                                None, None, None, None, zo.data_ptr(), None))
        return (zo - tx).cpu().numpy(), ex.cpu().numpy()
    d1, ex1 = run(7, 3)
    d2, _ = run(7, 3)
    d3, _ = run(7, 4)
    np.testing.assert_array_equal(d1, d2)
    assert np.abs(d1 - d3).max() > 0.1
    sd = np.sqrt(np.diag(cat.R))
    assert np.abs(d1.mean(axis=1) / sd).max() < 0.05
    np.testing.assert_allclose(d1.std(axis=1), sd, rtol=0.03)
    assert np.abs(np.corrcoef(d1)[np.triu_indices(6, 1)]).max() < 0.05
    assert np.abs(ex1 - cat.est_x)[:3].max() > 1.0


def test_full_size_properties():
    """BASELINE.json configs[1] size (100 sensors x 10k mixed targets): size-independent
    invariants instead of an oracle that would take hours."""
    from clockpunch_b200.engine import CatalogEngine
    from clockpunch_b200.dynamics.propagator import propagate_batch
    from clockpunch_b200.simulation.catalog import synth_catalog, MU
    cat = synth_catalog(10000, 100, seed=20240627)
    eng = CatalogEngine(cat.sensors, cat.sensor_kinds, cat.targets, cat.est_x, cat.est_p, cat.Q, cat.R, seed=9)
    x0 = cat.targets
    tasked = np.zeros(cat.N, np.uint8); tasked[::97] = 1
    eng.step(100.0, tasked=tasked, z=None)
    x1 = eng.host("truth_x")
    # two-body invariants: specific energy and angular momentum
    def energy(x): return 0.5 * (x[3:] ** 2).sum(0) - MU / np.linalg.norm(x[:3], axis=0)
    def hvec(x): return np.cross(x[:3].T, x[3:].T)
    assert np.abs(energy(x1) / energy(x0) - 1).max() < 1e-13
    assert np.abs(hvec(x1) - hvec(x0)).max() / np.abs(hvec(x0)).max() < 1e-13
    # the fused kernel's truth lane equals the standalone propagator bit for bit
    np.testing.assert_array_equal(x1, propagate_batch(x0, 0.0, 100.0))
    # forward then backward returns to the start
    back = propagate_batch(x1, 100.0, 0.0)
    assert np.abs(back - x0)[:3].max() < 1e-8
    # untasked targets: est_x is the propagated centre point; covariances symmetric PD
    ex = eng.host("est_x"); ep = eng.host("est_p"); pp = eng.host("pred_p")
    un = tasked == 0
    np.testing.assert_array_equal(ex[:, un], propagate_batch(cat.est_x, 0.0, 100.0)[:, un])
    np.testing.assert_array_equal(ep[un], pp[un])
    np.testing.assert_array_equal(ep, ep.transpose(0, 2, 1))
    assert np.linalg.eigvalsh(ep).min() > 0
    # a measurement can only shrink the covariance: pred_p - est_p is PSD for tasked targets
    diff = pp[~un] - ep[~un]
    assert np.linalg.eigvalsh(diff).min() > -1e-9 and np.trace(diff, axis1=1, axis2=2).min() > 0
    assert (eng.host("flags") == 0).all()
    assert eng.host("num_tasked").sum() == tasked.sum()
    # packed map: padding bits beyond M are zero, every row consistent with a 2nd standalone call
    words = eng.vis_truth.cpu().numpy().view(np.uint32)
    assert (words[:, -1] >> (cat.M - 32 * (eng.wpr - 1))).max() == 0
    from clockpunch_b200.common.visibility import calcVisMap
    np.testing.assert_array_equal(eng.host("vis_truth"), calcVisMap(eng.host("sens_x"), x1))
    np.testing.assert_array_equal(eng.host("vis_est"), calcVisMap(eng.host("sens_x"), ex))
    # GEO targets see ~40% of ground sites, LEO far fewer; antipodal check
    vt = eng.host("vis_truth")
    assert vt[cat.regime == 2].mean() > 3 * vt[cat.regime == 0].mean() > 0


def test_step_host_matches_graph_replay():
    """pc_step_host (host actions in, float32 observation + packed vis map + num_tasked out, graph
    cached inside the library) produces exactly what the device-resident replay path produces,
    and its observation is the float32 cast of the device state (env.py:261-270, 419-452, 508-530)."""
    import torch
    cat, eng, _ = _engine_and_oracle(37, 6, seed=13)
    rng = np.random.default_rng(3)
    acts = [rng.integers(0, cat.N + 1, size=cat.M) for _ in range(4)]
    gr = eng.build_graph(100.0, pack_obs=True)
    ref = []
    for a in acts:
        eng.actions[: cat.M].copy_(torch.from_numpy(a))
        eng.replay(gr)
        ref.append((eng.obs_f32.cpu().numpy().copy(), eng.vis_est.cpu().numpy().copy(), eng.host("num_tasked").copy()))
    eng.reset(rewind_rng=True)
    eng._graph = None
    for pinned in (True, False):
        eng.reset(rewind_rng=True)
        hb = eng.host_buffers(pinned=pinned)
        for k, a in enumerate(acts):
            hb["actions"][: cat.M].copy_(torch.from_numpy(a))
            eng.step_host(hb, 100.0)
            np.testing.assert_array_equal(hb["obs"].numpy(), ref[k][0])
            np.testing.assert_array_equal(hb["vis_est"].numpy(), ref[k][1])
            np.testing.assert_array_equal(hb["num_tasked"].numpy(), ref[k][2])
        assert eng.time == 400.0
    # the observation block is the float32 cast of the float64 device state
    A = cat.M + cat.N
    obs = hb["obs"].numpy()
    eci = obs[: 6 * A].reshape(6, A)
    np.testing.assert_array_equal(eci[:, : cat.M], eng.host("sens_x").astype(np.float32))
    np.testing.assert_array_equal(eci[:, cat.M:], eng.host("est_x").astype(np.float32))
    np.testing.assert_array_equal(obs[6 * A: 6 * A + 36 * cat.N].reshape(cat.N, 6, 6), eng.host("est_p").astype(np.float32))
    stale = obs[6 * A + 36 * cat.N:]
    np.testing.assert_array_equal(stale, (eng.time - eng.host("last_time_tasked").astype(np.float64)).astype(np.float32))
    assert eng.host("num_tasked").sum() > 0


def test_vismap_derivative_matches_oracle(golden):
    """calcVisMapAndDerivative (visibility.py:119-180) through pc_vismap_der_host against the oracle
    restatement (radial rate pinned to the live reference; satvis derivative formula unpinned)."""
    from clockpunch_b200.common.visibility import calcVisMapAndDerivative
    from oracle.visibility import calc_vis_map_and_derivative
    g = golden("orbits")
    X = g["x"]
    sens, targ = np.concatenate([X[:, 48:], X[:, :5]], axis=1), X[:, 5:48]     # ground + space sensors
    vis, der = calcVisMapAndDerivative(sens, targ)
    vis_o, der_o = calc_vis_map_and_derivative(sens, targ)
    assert vis.shape == der.shape == (43, 13)
    np.testing.assert_array_equal(vis, vis_o)
    # d/dt of the angle sum: scale is the orbital angular rate (~1e-3 rad/s); ground sites sit
    # 1e-6 km above the surface, where da/dt divides a 1e-16 radial rate by sqrt(3e-10)
    np.testing.assert_allclose(der, der_o, rtol=1e-9, atol=1e-11)
    cont, _ = calcVisMapAndDerivative(sens, targ, binary=False)
    vo, _ = calc_vis_map_and_derivative(sens, targ, binary=False)
    np.testing.assert_allclose(cont, vo, rtol=0, atol=1e-12)
    # aligned position vectors: infinite (or NaN) derivative, as the reference documents
    a = np.array([[7000.0, 0, 0, 0, 7.5, 0]]).T
    b = np.array([[14000.0, 0, 0, 0.1, 5.0, 0]]).T
    _, d = calcVisMapAndDerivative(a, b)
    assert not np.isfinite(d[0, 0])
    with pytest.raises(ValueError):
        calcVisMapAndDerivative(np.zeros((5, 2)), targ)


def test_c5_every_target_updated_every_step():
    """BASELINE.json configs[4] shape: 100k targets, measurement update on EVERY target EVERY step
    (thread-per-target kernel, update path on all lanes).  Size-independent properties plus an
    oracle check on a random sample of targets with the measurements the device drew."""
    from clockpunch_b200.engine import CatalogEngine
    from clockpunch_b200.simulation.catalog import synth_catalog
    from oracle.ukf import OracleUKF
    N = 100_000
    cat = synth_catalog(N, 16, seed=20240629)
    eng = CatalogEngine(cat.sensors, cat.sensor_kinds, cat.targets, cat.est_x, cat.est_p, cat.Q, cat.R, seed=3)
    ones = np.ones(N, np.uint8)
    z = cat.targets * 0
    import torch
    for s in range(3):
        x_prev, p_prev = eng.host("est_x"), eng.host("est_p")
        t0 = eng.time
        # draw z on the host from the device's own truth so that the oracle can replay it
        eng_truth_next = None
        eng.step(dt=60.0, tasked=ones, z=None)
        assert (eng.host("flags") == 0).all()
        assert (eng.host("num_tasked") == s + 1).all()
        np.testing.assert_array_equal(eng.host("last_time_tasked"), np.float32(eng.time))
        ep, pp = eng.host("est_p"), eng.host("pred_p")
        np.testing.assert_array_equal(ep, ep.transpose(0, 2, 1))
        idx = np.random.default_rng(s).choice(N, 2000, replace=False)
        assert np.linalg.eigvalsh(ep[idx]).min() > 0
        assert np.linalg.eigvalsh(pp[idx] - ep[idx]).min() > -1e-9          # a measurement only shrinks P
        nis = eng.host("nis")
        assert np.isfinite(nis).all() and (nis >= 0).all()
    # after three updates every estimate is within a few measurement sigmas (1 km, 0.1 km/s) of the truth
    err = np.abs(eng.host("est_x") - eng.host("truth_x"))
    assert np.percentile(err[:3].max(axis=0), 99) < 8.0 and np.percentile(err[3:].max(axis=0), 99) < 0.6
    # oracle replay of one step on a sample, with the measurement the device drew (out_z)
    from clockpunch_b200 import _lib, _buffers as B
    import ctypes as C
    ctx = _lib.get_ctx()
    dev = B.device_of(ctx)
    k = 24
    ex = torch.from_numpy(cat.est_x[:, :k].copy()).to(dev); epd = torch.from_numpy(cat.est_p[:k].reshape(k, 36).copy()).to(dev)
    tx = torch.from_numpy(cat.targets[:, :k].copy()).to(dev); zo = torch.zeros((6, k), dtype=torch.float64, device=dev)
    tk = torch.ones(k, dtype=torch.uint8, device=dev)
    params = _lib.make_ukf_params(cat.Q, cat.R)
    ctx.check(ctx.lib.pc_ukf_predict_update(ctx.handle, ex.data_ptr(), epd.data_ptr(), tx.data_ptr(), None, tk.data_ptr(),
                                            k, k, 0.0, 60.0, 0, 0, C.byref(params), 11, 0, None, None, None, None,
                                            zo.data_ptr(), None))
    zh, exh, eph = zo.cpu().numpy(), ex.cpu().numpy(), epd.cpu().numpy().reshape(k, 6, 6)
    for i in range(k):
        f = OracleUKF(0.0, cat.est_x[:, i], cat.est_p[i], cat.Q, cat.R)
        f.predict_and_update(60.0, zh[:, i])
        assert np.abs(exh[:, i] - f.est_x).max() <= estx_tol(f.est_x)
        assert cov_close(eph[i], f.est_p)


def test_c4_million_targets_on_one_gpu():
    """BASELINE.json configs[3] catalog size (1M objects x 100 sensors) on ONE device: invariants that do
    not need an oracle, the thread-per-target kernel with 128-thread CTAs, and the north-star step time."""
    import torch
    from clockpunch_b200.engine import CatalogEngine
    from clockpunch_b200.simulation.catalog import synth_catalog, MU
    N = 1_000_000
    cat = synth_catalog(N, 100, seed=20240630)
    eng = CatalogEngine(cat.sensors, cat.sensor_kinds, cat.targets, cat.est_x, cat.est_p, cat.Q, cat.R, seed=4)
    g = eng.build_graph(60.0, policy_probes=64, policy_seed=1)
    for _ in range(3):
        eng.replay(g)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        eng.replay(g)
    e1.record()
    torch.cuda.synchronize()
    assert e0.elapsed_time(e1) / 5 < 10.0           # north star: < 10 ms per step (on 8 GPUs; here on one)
    x0, x1 = cat.targets, eng.host("truth_x")
    energy = lambda x: 0.5 * (x[3:] ** 2).sum(0) - MU / np.linalg.norm(x[:3], axis=0)
    assert np.abs(energy(x1) / energy(x0) - 1).max() < 5e-13
    h0, h1 = np.cross(x0[:3].T, x0[3:].T), np.cross(x1[:3].T, x1[3:].T)
    assert np.abs(h1 - h0).max() / np.abs(h0).max() < 5e-13
    assert eng.time == 8 * 60.0 and 0 < eng.host("num_tasked").sum() <= 8 * 100
    fl = eng.host("flags")
    assert (fl & ~np.uint32(16)).max() == 0 and (fl != 0).mean() < 1e-3       # at most rate-clamp flags
    ep = eng.host("est_p")
    idx = np.random.default_rng(0).choice(N, 4000, replace=False)
    np.testing.assert_array_equal(ep[idx], ep[idx].transpose(0, 2, 1))
    assert np.linalg.eigvalsh(ep[idx]).min() > 0
    # rows of both maps equal the standalone kernel on the same states (bit for bit), on a slice
    from clockpunch_b200.common.visibility import calcVisMap
    sl = slice(500_000, 500_512)
    np.testing.assert_array_equal(eng.host("vis_truth")[sl], calcVisMap(eng.host("sens_x"), x1[:, sl]))
    np.testing.assert_array_equal(eng.host("vis_est")[sl], calcVisMap(eng.host("sens_x"), eng.host("est_x")[:, sl]))


def test_fused_visibility_with_many_sensors_matches_standalone():
    """More sensors than one shared-memory tile holds (M = 300 -> 10 words -> two bulk-copy chunks,
    ragged last row) in both per-target kernels: the fused maps equal the standalone kernel's."""
    from clockpunch_b200.common.visibility import calcVisMap
    from clockpunch_b200.engine import CatalogEngine
    from clockpunch_b200.simulation.catalog import synth_catalog
    for n in (77, 40_000):                       # quad kernel / thread-per-target kernel
        cat = synth_catalog(n, 300, seed=31 + n, space_sensor_frac=0.2)
        eng = CatalogEngine(cat.sensors, cat.sensor_kinds, cat.targets, cat.est_x, cat.est_p, cat.Q, cat.R, seed=2)
        tasked = np.zeros(n, np.uint8); tasked[::5] = 1
        eng.step(dt=100.0, tasked=tasked)
        sl = slice(0, n) if n < 1000 else slice(n - 700, n)
        sx = eng.host("sens_x")
        np.testing.assert_array_equal(eng.host("vis_truth")[sl], calcVisMap(sx, eng.host("truth_x")[:, sl]))
        np.testing.assert_array_equal(eng.host("vis_est")[sl], calcVisMap(sx, eng.host("est_x")[:, sl]))
        words = eng.vis_truth.cpu().numpy().view(np.uint32)
        assert (words[:, -1] >> (300 - 32 * 9)).max() == 0          # padding bits of the ragged last word stay 0


def test_peer_push_single_rank_matches_summary():
    """The fused exchange on one GPU (world = 1): what the per-target kernel pushes into the peer-mapped
    gathered buffer is byte-identical to the shard's own summary buffer, the flag / wait pair
    completes, parity slots alternate, and an engine reset does not rewind the flags."""
    import os
    import torch
    import torch.distributed as dist
    from clockpunch_b200.engine import CatalogEngine
    from clockpunch_b200.simulation.catalog import synth_catalog
    if not dist.is_initialized():
        import socket
        with socket.socket() as sock:
            sock.bind(("127.0.0.1", 0))
            port = sock.getsockname()[1]
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        try:
            dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
        except Exception as exc:
            pytest.skip(f"cannot start a single-rank NCCL group here: {exc!r}")
    for n in (300, 40_000):                       # quad kernel / thread-per-target kernel
        cat = synth_catalog(n, 70, seed=50 + n)
        eng = CatalogEngine(cat.sensors, cat.sensor_kinds, cat.targets, cat.est_x, cat.est_p, cat.Q, cat.R, seed=6)
        gathered = eng.enable_peer_push()         # pc_comm_init / pc_comm_attach (CUDA IPC); one per context
        g = eng.build_graph(100.0, policy_probes=32, policy_seed=3)
        seen = []
        last = eng.peer_status()[1]               # the push counter belongs to the library context
        for s in range(5):
            if s == 3:
                eng.reset()
            eng.replay(g)                             # its sensor kernel pushes the PREVIOUS step's summaries
            err, pushes = eng.peer_status()           # exchanges the current ones, then reports
            assert err == 0 and pushes == last + 2    # one push per step, one per explicit exchange
            last = pushes
            torch.testing.assert_close(gathered[pushes & 1, 0], eng.summary, rtol=0, atol=0)
            f = eng.peer_fields(pushes & 1, 0)
            np.testing.assert_array_equal(f["vis_est"].cpu().numpy(), eng.vis_est.cpu().numpy())
            np.testing.assert_array_equal(f["num_tasked"].cpu().numpy(), eng.host("num_tasked"))
            seen.append(gathered[pushes & 1, 0].clone())
        assert not torch.equal(seen[0], seen[1])          # the two parity slots hold different steps
        assert eng.host("num_tasked").sum() > 0


def test_guard_bands_no_out_of_bounds_writes():
    """compute-sanitizer is closed on this pool, so out-of-bounds writes are hunted the old way: every
    buffer the step kernels write is re-homed inside a larger allocation whose margins hold a
    sentinel, the step runs (all kernels: sensors, propagation, per-target stage with update and
    visibility, observation pack, standalone maps) at sizes that do not divide any tile, and the
    margins must come back untouched."""
    import torch
    from clockpunch_b200.engine import CatalogEngine
    from clockpunch_b200.simulation.catalog import synth_catalog
    G = 96
    for n, m in ((77, 45), (33_001, 70), (130, 300)):
        cat = synth_catalog(n, m, seed=7 * n + m, space_sensor_frac=0.3)
        eng = CatalogEngine(cat.sensors, cat.sensor_kinds, cat.targets, cat.est_x, cat.est_p, cat.Q, cat.R, seed=1)
        bands = {}

        def guard(name):
            t = getattr(eng, name)
            flat = t.reshape(-1)
            fill = 12345.678 if t.dtype.is_floating_point else 0x5A
            big = torch.full((flat.numel() + 2 * G,), fill, dtype=t.dtype, device=t.device)
            big[G: G + flat.numel()] = flat
            setattr(eng, name, big[G: G + flat.numel()].view(t.shape))
            bands[name] = (big, fill)

        for name in ("sens_x", "sens_q", "sigma", "sig_flags", "est_p", "pred_p", "nis", "flags", "vis_truth",
                     "out_block", "tasked", "z", "actions"):
            guard(name)
        eng.est_x, eng.truth_x = eng.sigma[0], eng.sigma[13]
        eng.obs_f32 = eng.out_block[: 4 * sum(eng._obs_sizes)].view(torch.float32)
        eng.summary = eng.out_block[eng._obs_bytes:]
        eng._bind_summary_views()
        eng.snapshot()
        rng = np.random.default_rng(n)
        eng.compute_vis()
        for s in range(3):
            eng.step(dt=100.0, actions=rng.integers(0, n + 1, size=m))
            eng.pack_obs()
        hb = eng.host_buffers()
        hb["actions"][:m].copy_(torch.from_numpy(rng.integers(0, n + 1, size=m)))
        eng.step_host(hb, 100.0)
        eng.step_host(hb, 100.0)
        torch.cuda.synchronize()
        assert eng.host("num_tasked").sum() > 0 and (eng.host("flags") & ~np.uint32(16)).max() == 0
        for name, (big, fill) in bands.items():
            lo, hi = big[:G].cpu().numpy(), big[-G:].cpu().numpy()
            ref = np.full(G, fill, dtype=lo.dtype)
            np.testing.assert_array_equal(lo, ref, err_msg=f"{name}: wrote before the buffer (n={n}, m={m})")
            np.testing.assert_array_equal(hi, ref, err_msg=f"{name}: wrote past the buffer (n={n}, m={m})")


def test_empty_and_degenerate_inputs():
    """Zero-length batches, t0 == tf, NaN states: the entry points return / raise like the reference
    instead of launching on nothing."""
    from clockpunch_b200 import _lib
    from clockpunch_b200.common.visibility import calcVisMap, calcVisMapAndDerivative
    from clockpunch_b200.dynamics.dynamics_classes import SatDynamicsModel, StaticTerrestrial
    from clockpunch_b200.dynamics.propagator import propagate_batch
    from clockpunch_b200.estimation.ukf_v2 import ukf_batch_host
    assert propagate_batch(np.zeros((6, 0)), 0.0, 10.0).shape == (6, 0)
    sens = np.array([[7000.0, 0, 0, 0, 7.5, 0]]).T
    assert calcVisMap(sens, np.zeros((6, 0))).shape == (0, 1)
    assert calcVisMap(np.zeros((6, 0)), sens).shape == (1, 0)
    v, d = calcVisMapAndDerivative(sens, np.zeros((6, 0)))
    assert v.shape == d.shape == (0, 1)
    params = _lib.make_ukf_params(np.eye(6), np.eye(6))
    out = ukf_batch_host(np.zeros((6, 0)), np.zeros((0, 6, 6)), None, None, 0.0, 60.0, params)
    assert out["est_x"].shape == (6, 0) and out["flags"].shape == (0,)
    with pytest.raises(ValueError):
        SatDynamicsModel().propagate(sens, 3.0, 3.0)                      # propagator.py:40-41
    np.testing.assert_array_equal(StaticTerrestrial().propagate(sens[:, 0], 3.0, 3.0), sens[:, 0])   # zero rotation
    # a NaN state propagates to NaN (no hang in the substep rule) and is invisible
    bad = sens.copy(); bad[1, 0] = np.nan
    assert np.isnan(SatDynamicsModel().propagate(bad, 0.0, 100.0)).any()
    assert calcVisMap(sens, bad)[0, 0] == 0
    # a state at the origin-crossing extreme takes the clamped substep path and returns
    fall = np.array([[6500.0, 0, 0, 0, 0.01, 0]]).T               # nearly radial plunge: orbit rate clamp
    assert np.isfinite(SatDynamicsModel().propagate(fall, 0.0, 10.0)).all()


@pytest.mark.parametrize("given_z", [False, True])
def test_update_warps_equal_inline_update(given_z):
    """With an actions buffer the tasked targets are updated by dedicated warps of the per-target kernel
    (pc_step_args.task_of); with a tasking mask by the in-line update of the target's own lanes.  Same
    equations: results agree to round-off, bookkeeping exactly -- also when several sensors task the
    same target (one update, num_tasked + 1) and when the measurement is drawn on the device (both
    paths key the draw by (seed, step, target))."""
    from clockpunch_b200.engine import CatalogEngine
    from clockpunch_b200.simulation.catalog import synth_catalog
    cat = synth_catalog(150, 24, seed=77, space_sensor_frac=0.3)
    mk = lambda: CatalogEngine(cat.sensors, cat.sensor_kinds, cat.targets, cat.est_x, cat.est_p, cat.Q, cat.R, seed=9)
    a, b = mk(), mk()
    rng = np.random.default_rng(3)
    n_upd = 0
    for s in range(4):
        vis = a.host("vis_est")                                       # (N, M) estimated map of the previous step
        acts = np.full(cat.M, cat.N, np.int64)
        for j in range(cat.M):
            seen = np.flatnonzero(vis[:, j])
            if len(seen):
                acts[j] = seen[rng.integers(0, min(3, len(seen)))]   # few distinct picks: duplicates on purpose
        acts[::5] = rng.integers(0, cat.N + 1, size=len(acts[::5]))  # and some actions on invisible targets / inaction
        z = (cat.targets + rng.normal(size=cat.targets.shape)) if given_z else None
        a.step(dt=100.0, actions=acts, z=z)
        b.task_from_actions(acts)
        b.step(dt=100.0, tasked=b.tasked, z=z)
        for k in ("truth_x", "vis_truth", "vis_est", "num_tasked", "last_time_tasked", "flags"):
            np.testing.assert_array_equal(a.host(k), b.host(k), err_msg=f"{k} step {s}")
        # (the transform's weights are -2e6 / +1.7e5: a different summation order moves pred_p by ~1e-10
        # relative, inside the bar of SURVEY.md section 8d -- covariances 1e-8, est_x 1e-5 km per step)
        assert (np.abs(a.host("est_x") - b.host("est_x")).max(axis=0) <= estx_tol(b.host("est_x")) * (s + 1)).all(), f"est_x step {s}"
        assert cov_close(a.host("est_p"), b.host("est_p"), 1e-8 * (s + 1)), f"est_p step {s}"
        assert cov_close(a.host("pred_p"), b.host("pred_p"), 1e-8 * (s + 1)), f"pred_p step {s}"
        np.testing.assert_allclose(a.host("nis"), b.host("nis"), rtol=1e-6, equal_nan=True, err_msg=f"nis step {s}")
        n_upd = int(a.host("num_tasked").sum())
    assert n_upd > 8 and a.host("num_tasked").max() <= 4      # updates happened; never two in one step


def test_host_policy_equals_device_policy():
    """The stand-in policy has a device form (pc_policy_random_visible, also inside the step's sensor kernel)
    and a host form (pc_policy_random_visible_host, bench.py's end-to-end loop): same picks."""
    import ctypes as C
    import torch
    from clockpunch_b200 import _lib
    ctx = _lib.get_ctx(0)
    rng = np.random.default_rng(8)
    N, M, probes = 700, 45, 32
    wpr = (M + 31) // 32
    packed = rng.integers(0, 2**32, size=(N, wpr), dtype=np.uint64).astype(np.uint32) \
        & rng.integers(0, 2**32, size=(N, wpr), dtype=np.uint64).astype(np.uint32) \
        & rng.integers(0, 2**32, size=(N, wpr), dtype=np.uint64).astype(np.uint32)       # ~12 % visible
    d_vis = torch.from_numpy(packed.view(np.int32)).cuda()
    d_act = torch.zeros(M, dtype=torch.int64, device="cuda")
    host = np.zeros(M, np.int64)
    for step in (0, 5, 123456789):
        ctx.check(ctx.lib.pc_policy_random_visible(ctx.handle, d_vis.data_ptr(), N, M, wpr, 99, step, None, probes,
                                                   d_act.data_ptr(), torch.cuda.current_stream().cuda_stream))
        assert ctx.lib.pc_policy_random_visible_host(packed.ctypes.data_as(C.c_void_p), N, M, wpr, 99, step, probes,
                                                     host.ctypes.data_as(C.c_void_p)) == 0
        np.testing.assert_array_equal(d_act.cpu().numpy(), host)
        assert (host < N).sum() > M // 2


def test_folded_propagation_equals_single_state_kernel():
    """Above 16 k targets the sigma points are advanced two per thread out of a sorted 256-target window
    (propagate_sigma_kernel), below one per thread (propagate_sigma_single_kernel): the same arithmetic per
    state, so the first targets of a large batch must come out bit-identical to the same targets run
    alone -- including states whose substep count exceeds the 8-bit field of the window sort (the rule is
    then evaluated again) and the clamp flag."""
    from clockpunch_b200 import _lib
    from clockpunch_b200.estimation.ukf_v2 import ukf_batch_host
    from clockpunch_b200.simulation.catalog import synth_catalog
    cat = synth_catalog(17_000, 4, seed=31)
    x, p = cat.est_x.copy(), cat.est_p.copy()
    mu = 398600.0
    for k in range(24):                                   # eccentric orbits at perigee: orbit rates up to the clamp
        rp, e = 6600.0 + 40.0 * k, 0.05 + 0.035 * k
        vp = np.sqrt(mu * (1 + e) / rp)
        x[:, 7 * k] = [rp, 0.0, 0.0, 0.0, vp * np.cos(0.3), vp * np.sin(0.3)]
    params = _lib.make_ukf_params(cat.Q, cat.R)
    dt = 30_000.0
    big = ukf_batch_host(x, p, None, None, 0.0, dt, params)
    m = 300
    small = ukf_batch_host(x[:, :m], p[:m], None, None, 0.0, dt, params)
    for k in ("est_x", "pred_x"):
        np.testing.assert_array_equal(big[k][:, :m], small[k], err_msg=k)
    for k in ("est_p", "pred_p", "flags"):
        np.testing.assert_array_equal(big[k][:m], small[k], err_msg=k)
    assert np.isfinite(small["pred_x"]).all()



def test_propagation_mappings_agree_bit_for_bit():
    """The two thread mappings of the sigma-point propagation (pc_option PC_OPT_PROPAGATE_KERNEL: one state
    per thread with a CTA-wide sort of a 128-target window, two states per thread out of a sorted 256-target
    window) integrate every state with the same substep count and the same arithmetic: whole engine steps -- truth, estimates, covariances, maps, flags -- must come out identical,
    including eccentric estimates (several substeps, mixed within a window) and a ragged last window."""
    from clockpunch_b200 import _lib
    from clockpunch_b200.engine import CatalogEngine
    from clockpunch_b200.simulation.catalog import synth_catalog
    n = 3000 + 77
    cat = synth_catalog(n, 40, seed=123, space_sensor_frac=0.25)
    mu = 398600.0
    for k in range(60):                                   # eccentric estimates at perigee, sprinkled over the windows
        rp, e = 6700.0 + 30.0 * k, 0.03 + 0.011 * k
        vp = np.sqrt(mu * (1 + e) / rp)
        cat.est_x[:, 50 * k + 3] = [rp, 0.0, 0.0, 0.0, vp * np.cos(0.4), vp * np.sin(0.4)]
    ctx = _lib.get_ctx()
    out = {}
    rng = np.random.default_rng(4)
    tasked = (rng.random(n) < 0.05).astype(np.uint8)
    z = cat.targets + rng.normal(size=cat.targets.shape)
    for mapping in (1, 2):
        old = ctx.set_option(_lib.OPT_PROPAGATE_KERNEL, mapping)
        try:
            eng = CatalogEngine(cat.sensors, cat.sensor_kinds, cat.targets, cat.est_x, cat.est_p, cat.Q, cat.R, seed=2)
            for s in range(3):
                eng.step(dt=100.0 if s < 2 else 900.0, tasked=tasked, z=z)      # the last step: dt above the 128 s horizon
            out[mapping] = {k: eng.host(k) for k in ("truth_x", "est_x", "est_p", "pred_p", "vis_truth", "vis_est", "flags", "nis")}
        finally:
            ctx.set_option(_lib.OPT_PROPAGATE_KERNEL, old)
    for mapping in (2,):
        for k, v in out[1].items():
            np.testing.assert_array_equal(out[mapping][k], v, err_msg=f"mapping {mapping}: {k}")
    assert np.isfinite(out[1]["est_x"]).all()
