"""Vectorised envs (SURVEY.md row N4 / BASELINE.json configs[2]): E environments stepped as one
catalog must equal E independent single-environment engines bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _batch(E, M, N, seed):
    from clockpunch_b200.simulation.catalog import synth_catalog
    cats = [synth_catalog(N, M, seed=seed + e, space_sensor_frac=0.25) for e in range(E)]
    st = lambda k: np.stack([getattr(c, k) for c in cats])
    return cats, st("sensors"), st("sensor_kinds"), st("targets"), st("est_x"), st("est_p")


@pytest.mark.parametrize("E,M,N", [(3, 4, 9), (2, 40, 37), (5, 10, 100)])
def test_batched_envs_equal_independent_engines(E, M, N):
    from clockpunch_b200.engine import CatalogEngine
    from clockpunch_b200.environment.vector_env import VectorSSAScheduler
    cats, S, K, T, X, P = _batch(E, M, N, seed=100 * E + N)
    venv = VectorSSAScheduler(S, K, T, X, P, cats[0].Q, cats[0].R, time_step=100.0, horizon=50)
    singles = [CatalogEngine(c.sensors, c.sensor_kinds, c.targets, c.est_x, c.est_p, c.Q, c.R) for c in cats]
    obs, _ = venv.reset()
    for e, eng in enumerate(singles):
        np.testing.assert_array_equal(obs["vis_map_est"][e], eng.host("vis_est"))
    rng = np.random.default_rng(7)
    for s in range(3):
        actions = rng.integers(0, N + 1, size=(E, M))
        z = T + rng.normal(size=T.shape)                 # any finite measurement will do: same on both sides
        obs, rew, done, trunc, info = venv.step(actions, measurements=z)
        for e, eng in enumerate(singles):
            eng.step(dt=100.0, z=z[e], actions=actions[e])
            eng.pack_obs()
            o = eng.obs_f32.cpu().numpy()
            A = M + N
            np.testing.assert_array_equal(obs["eci_state"][e], o[: 6 * A].reshape(6, A))
            np.testing.assert_array_equal(obs["est_cov"][e], o[6 * A: 6 * A + 36 * N].reshape(N, 6, 6))
            np.testing.assert_array_equal(obs["obs_staleness"][e, 0], o[6 * A + 36 * N:])
            np.testing.assert_array_equal(obs["vis_map_est"][e], eng.host("vis_est"))
            np.testing.assert_array_equal(obs["num_tasked"][e, 0], eng.host("num_tasked"))
            sl = slice(e * N, (e + 1) * N)
            np.testing.assert_array_equal(venv._engine.host("truth_x")[:, sl], eng.host("truth_x"))
            np.testing.assert_array_equal(venv._engine.host("vis_truth")[sl], eng.host("vis_truth"))
        assert obs["num_tasked"].sum() > 0 or s == 0
    assert not done.any() and (rew == 0).all()


def test_c3_shape_rollout_stepping():
    """configs[2]: 1024 envs x 100 targets x 10 sensors, stepped through pc_step_host (graph replay,
    measurements drawn on device): shapes, dtypes, horizon reset and spot checks of the maps."""
    from clockpunch_b200.common.visibility import calcVisMap
    from clockpunch_b200.environment.vector_env import VectorSSAScheduler
    from clockpunch_b200.simulation.catalog import synth_catalog
    E, M, N = 1024, 10, 100
    cat = synth_catalog(E * N, E * M, seed=20240628, mix="LEO")
    S = cat.sensors.reshape(6, E, M).transpose(1, 0, 2)
    T = cat.targets.reshape(6, E, N).transpose(1, 0, 2)
    X = cat.est_x.reshape(6, E, N).transpose(1, 0, 2)
    venv = VectorSSAScheduler(S, cat.sensor_kinds.reshape(E, M), T, X, cat.est_p[0], cat.Q, cat.R, 100.0, horizon=4, seed=1)
    # the same batch returning VIEWS of the pinned host block instead of copies (copy=False, as gymnasium's vector envs)
    views = VectorSSAScheduler(S, cat.sensor_kinds.reshape(E, M), T, X, cat.est_p[0], cat.Q, cat.R, 100.0, horizon=4, seed=1,
                               copy=False)
    obs0, _ = venv.reset()
    views.reset()
    rng = np.random.default_rng(0)
    for s in range(4):
        vis = obs0["vis_map_est"] if s == 0 else obs["vis_map_est"]
        # greedy: every sensor tasks its first estimated-visible target, else inaction
        has = vis.any(axis=1)                                   # (E, M)
        actions = np.where(has, vis.argmax(axis=1), N)
        obs, rew, done, trunc, info = venv.step(actions)
        obs_v, _, _, _, info_v = views.step(actions)
        for k in obs:
            np.testing.assert_array_equal(obs_v[k], obs[k], err_msg=k)
        # (on the horizon step the env resets itself, which rewrites the block: that step returns copies)
        assert np.shares_memory(obs_v["est_cov"], views._hb["obs"].numpy()) == (s < 3)
        assert not np.shares_memory(obs["est_cov"], venv._hb["obs"].numpy())
        np.testing.assert_array_equal(info_v["mean_pos_var"], info["mean_pos_var"])
        # the kernel's float32 traces are the traces of the float32 observation (env.py:599-609)
        np.testing.assert_allclose(info["mean_pos_var"], np.trace(obs["est_cov"][:, :, :3, :3], axis1=2, axis2=3).mean(axis=1), rtol=2e-6)
        np.testing.assert_allclose(info["mean_vel_var"], np.trace(obs["est_cov"][:, :, 3:, 3:], axis1=2, axis2=3).mean(axis=1), rtol=2e-6)
        assert obs["eci_state"].shape == (E, 6, M + N) and obs["eci_state"].dtype == np.float32
        assert obs["est_cov"].shape == (E, N, 6, 6) and obs["vis_map_est"].shape == (E, N, M)
        assert obs["obs_staleness"].shape == (E, 1, N) and obs["num_tasked"].dtype == np.int64
        assert info["mean_pos_var"].shape == (E,)
        assert done.all() == (s == 3)
    eng = venv._engine
    assert venv.num_steps == 0 and eng.time == 0.0                 # horizon reached: reset inside step
    np.testing.assert_array_equal(venv._obs()["eci_state"], obs0["eci_state"])
    # tasked targets were really updated in the last step before the reset
    assert obs["num_tasked"].sum() >= E          # at least one tasking per env over the rollout
    for e in (0, 511, 1023):
        sl_t, sl_s = slice(e * N, (e + 1) * N), slice(e * M, (e + 1) * M)
        np.testing.assert_array_equal(eng.host("vis_truth")[sl_t],
                                      calcVisMap(eng.host("sens_x")[:, sl_s], eng.host("truth_x")[:, sl_t]))
