"""Access-window forecasting against the live reference's
AccessWindowCalculator (tests/golden/windows.npz) and forecastVisMap against the oracle."""
import numpy as np
import pytest


def _calc(g, tag):
    from clockpunch_b200.schedule_tree.access_windows import AccessWindowCalculator
    horizon, dt, fixed, merge, t0 = g[f"{tag}_cfg"]
    kinds = ["terrestrial" if k else "satellite" for k in g[f"{tag}_kinds"]]
    xs, xt = g[f"{tag}_xs"], g[f"{tag}_xt"]
    calc = AccessWindowCalculator(num_sensors=xs.shape[1], num_targets=xt.shape[1], dynamics_sensors=kinds,
                                  dynamics_targets="satellite", horizon=int(horizon), dt=float(dt),
                                  fixed_horizon=bool(fixed), merge_windows=bool(merge))
    return calc, xs, xt, float(t0)


def test_window_bookkeeping_matches_reference(golden):
    """Host-side window logic on the reference's own (T, N, M) histories."""
    g = golden("windows")
    for tag in ("a", "b"):
        calc, xs, xt, t0 = _calc(g, tag)
        vh = g[f"{tag}_vis_hist"]
        merged = calc._mergeWindows(vh)
        np.testing.assert_array_equal(merged, g[f"{tag}_vis_hist_targets"])
        np.testing.assert_array_equal(calc.sumWindows(vh, merged, calc.merge_windows), g[f"{tag}_num_windows"])
        np.testing.assert_array_equal(calc._calcTimeToWindowHist(t0, merged), g[f"{tag}_ttw"])
        calc.t_now = t0
        np.testing.assert_array_equal(calc._genTime(), g[f"{tag}_time"])
    np.testing.assert_array_equal(calc.calcStepsToWindow(np.array([1, 0, 0, 1, 0, 1, 0, 0])),
                                  [0, 2, 1, 0, 1, 0, np.inf, np.inf])
    from clockpunch_b200.schedule_tree.access_windows import AccessWindowCalculator
    with pytest.raises(AssertionError):
        AccessWindowCalculator(2, 2, "satellite", "satellite", horizon=0)


@pytest.mark.gpu
def test_vis_history_matches_live_reference(golden):
    from oracle.visibility import visibility_func
    g = golden("windows")
    for tag in ("a", "b"):
        calc, xs, xt, t0 = _calc(g, tag)
        nw, vh, vht, th, ttw = calc.calcNumWindows(xs, xt, t0, return_vis_hist=True)
        ref = g[f"{tag}_vis_hist"]
        assert vh.shape == ref.shape and vh.dtype == ref.dtype
        np.testing.assert_array_equal(th, g[f"{tag}_time"])
        diff = vh != ref
        assert diff.sum() <= 1, (tag, int(diff.sum()))       # at most a grazing pair among T N M
        if not diff.any():
            np.testing.assert_array_equal(nw, g[f"{tag}_num_windows"])
            np.testing.assert_array_equal(vht, g[f"{tag}_vis_hist_targets"])
            np.testing.assert_array_equal(ttw, g[f"{tag}_ttw"])
    # the bit-packed words the library returned hold the same history
    bits = np.unpackbits(calc._words.view(np.uint8), axis=2, bitorder="little")[:, :, : xs.shape[1]]
    np.testing.assert_array_equal(bits, vh)


@pytest.mark.gpu
def test_forecast_vis_map_matches_oracle(golden):
    """forecastVisMap (env_utils.py:130-180), truth and estimate flavours, against the oracle's
    propagate / UKF predict / visibility on the env golden's agents."""
    from test_env import _agents_from_golden
    from clockpunch_b200.environment.env_utils import forecastVisMap, getVisMapEstOrTruth
    from oracle.dynamics import propagate_satellite, propagate_terrestrial
    from oracle.ukf import OracleUKF
    from oracle.visibility import calc_vis_map
    g = golden("env")
    agents = _agents_from_golden(g, "default")
    M = 3
    S, T, X0 = g["default_sensors"], g["default_targets"], g["default_est_x0"]
    dt = 250.0
    s1 = np.stack([propagate_terrestrial(S[:, j], 0.0, dt) for j in range(M)], axis=1)
    t1 = np.stack([propagate_satellite(T[:, i], 0.0, dt) for i in range(T.shape[1])], axis=1)
    np.testing.assert_array_equal(forecastVisMap(agents, dt), calc_vis_map(s1, t1))
    px = []
    for i in range(T.shape[1]):
        f = OracleUKF(0.0, X0[:, i], g["default_P0"], g["default_Q"], g["default_R"])
        f.predict(dt)
        px.append(f.pred_x)
    np.testing.assert_array_equal(forecastVisMap(agents, dt, estimate=True), calc_vis_map(s1, np.stack(px, axis=1)))
    # the agents were not touched
    assert agents[0].time == 0 and agents[M].target_filter.time == 0
    np.testing.assert_array_equal(getVisMapEstOrTruth(agents, truth_flag=True), g["default_reset_vis_truth"])
    np.testing.assert_array_equal(getVisMapEstOrTruth(agents, truth_flag=False), g["default_reset_vis_est"])
