"""Golden SSAScheduler rollouts from the LIVE reference env (run in the build container only).

    python tests/golden/make_golden_env.py      # writes tests/golden/env.npz

The reference package needs gymnasium / ray / satvis / intervaltree / matplotlib, none of which
is installed here.  They are replaced by TEST-ONLY stand-ins:
  * a meta-path finder fabricating permissive modules and classes for everything missing
    (gymnasium.Env, the spaces, ray.*, ...): none of them does arithmetic;
  * `satvis.visibility_func` -> the oracle's restatement (oracle/visibility.py, parity unpinned);
  * numpy.Inf (removed in numpy 2), which two reference modules import.
The environment, agents, filters, propagator and metrics tracker are the reference's own code.
Measurement noise comes from numpy's legacy global RNG (agents.py:181); it is seeded and every
drawn z is recorded so that the CUDA env can be fed the same measurements.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_stubs  # noqa: E402,F401  (registers the stand-ins and the bare `punchclock` package)

from punchclock.common import agents as ref_agents  # noqa: E402
from punchclock.common.transforms import ecef2eci  # noqa: E402
from punchclock.environment.env import SSAScheduler  # noqa: E402
from punchclock.environment.env_parameters import SSASchedulerParams  # noqa: E402

RE = 6378.1363
MU = 398600.0


def scenario(kind):
    if kind == "test_env":            # tests/environment/test_env.py:35-79
        sens_ecef = np.zeros([3, 6]); sens_ecef[:, 0] = [RE, RE + 1, RE + 2]
        sensors = ecef2eci(sens_ecef.T, JD=0).T
        targets = np.zeros([4, 6]); targets[:, 0] = [RE + 400, -RE - 500, RE + 600, RE + 700]
        for i in range(4):
            targets[i, 4] = np.sqrt(MU / np.linalg.norm(targets[i, :3]))
        D = np.diag([1, 1, 1, 0.01, 0.01, 0.01])
        return sensors, targets, {"Q": 0.001 * D, "R": 0.01 * D, "p_init": 1 * D}, 1.2, 10
    # default-like env: 3 ground sensors x 10 LEO targets, reference default filter, dt = 100 s
    rng = np.random.default_rng(20240625)
    lat, lon = rng.uniform(-1.2, 1.2, 3), rng.uniform(-np.pi, np.pi, 3)
    sens_ecef = np.zeros([3, 6])
    sens_ecef[:, 0] = (RE + 1e-6) * np.cos(lat) * np.cos(lon)
    sens_ecef[:, 1] = (RE + 1e-6) * np.cos(lat) * np.sin(lon)
    sens_ecef[:, 2] = (RE + 1e-6) * np.sin(lat)
    sensors = ecef2eci(sens_ecef.T, JD=0).T
    from punchclock.common.transforms import coe2eci
    targets = np.array([coe2eci(rng.uniform(RE + 400, RE + 1000), rng.uniform(0, 0.01), rng.uniform(0, np.pi),
                                *rng.uniform(0, 2 * np.pi, 3)) for _ in range(10)])
    D = np.diag([1, 1, 1, 0.01, 0.01, 0.01])
    return sensors, targets, {"Q": 0.1 * D, "R": 1.0 * D, "p_init": 10 * D}, 100.0, 12


out = {}
for kind in ("test_env", "default"):
    sensors, targets, fp, dt, steps = scenario(kind)
    M, N = len(sensors), len(targets)
    agent_params = {"num_sensors": M, "num_targets": N, "sensor_starting_num": 1000, "target_starting_num": 9000,
                    "sensor_dynamics": "terrestrial", "target_dynamics": "satellite", "sensor_dist": None,
                    "target_dist": None, "sensor_dist_frame": None, "target_dist_frame": None,
                    "sensor_dist_params": None, "target_dist_params": None, "fixed_sensors": sensors,
                    "fixed_targets": targets, "init_num_tasked": None, "init_last_time_tasked": None}
    np.random.seed(1234)                      # initial-estimate noise (env_parameters.py:372-377) + measurements
    params = SSASchedulerParams(time_step=dt, horizon=steps + 5, agent_params=agent_params, filter_params=fp)
    env = SSAScheduler(params)
    zlog = []
    orig = ref_agents.Target.getMeasurement

    def patched(self, _orig=orig):
        z = _orig(self)
        zlog.append((self.agent_id, z.copy()))
        return z
    ref_agents.Target.getMeasurement = patched
    obs, info = env.reset()
    tg = [a for a in env.agents if type(a) is ref_agents.Target]
    out[f"{kind}_sensors"] = np.array([a.eci_state.squeeze() for a in env.agents if type(a) is ref_agents.Sensor]).T
    out[f"{kind}_targets"] = np.array([a.eci_state.squeeze() for a in tg]).T
    out[f"{kind}_est_x0"] = np.array([np.asarray(a.target_filter.est_x).squeeze() for a in tg]).T
    out[f"{kind}_P0"], out[f"{kind}_Q"], out[f"{kind}_R"] = fp["p_init"], fp["Q"], fp["R"]
    out[f"{kind}_dt"], out[f"{kind}_ids"] = dt, np.array([a.agent_id for a in env.agents])
    out[f"{kind}_reset_vis_est"], out[f"{kind}_reset_vis_truth"] = info["vis_map_est"], info["vis_map_truth"]
    out[f"{kind}_reset_eci"], out[f"{kind}_reset_cov"] = obs["eci_state"], obs["est_cov"]
    arng = np.random.default_rng(99)
    rec = {k: [] for k in ("action", "z", "eci", "cov", "vis_est", "vis_truth", "stale", "num_tasked", "true",
                           "pred_p", "mpv", "mvv", "uniq", "nv_est", "nv_truth", "multi", "nvs_est", "nvs_truth",
                           "time", "estx64", "estp64")}
    for s in range(steps):
        action = arng.integers(0, N + 1, size=M)
        if kind == "test_env" and s == 0:
            action = np.array([1, 0, 4])        # test_env.py:137
        zlog.clear()
        obs, rew, done, trunc, info = env.step(action)
        z = np.full((6, N), np.nan)
        for aid, zz in zlog:
            z[:, list(env.target_ids).index(aid)] = zz
        tg = [a for a in env.agents if type(a) is ref_agents.Target]
        rec["action"].append(action); rec["z"].append(z); rec["eci"].append(obs["eci_state"])
        rec["cov"].append(obs["est_cov"]); rec["vis_est"].append(obs["vis_map_est"])
        rec["vis_truth"].append(info["vis_map_truth"]); rec["stale"].append(obs["obs_staleness"])
        rec["num_tasked"].append(obs["num_tasked"]); rec["true"].append(info["true_states"])
        rec["pred_p"].append(info["filter_pred_p"]); rec["mpv"].append(info["mean_pos_var"])
        rec["mvv"].append(info["mean_vel_var"]); rec["uniq"].append(info["num_unique_targets_tasked"])
        rec["nv_est"].append(info["num_non_vis_taskings_est"]); rec["nv_truth"].append(info["num_non_vis_taskings_truth"])
        rec["multi"].append(info["num_multiple_taskings"]); rec["nvs_est"].append(np.array(info["non_vis_by_sensor_est"]))
        rec["nvs_truth"].append(np.array(info["non_vis_by_sensor_truth"])); rec["time"].append(info["time_now"])
        rec["estx64"].append(np.array([a.target_filter.est_x for a in tg]).T)
        rec["estp64"].append(np.array([a.target_filter.est_p for a in tg]))
        assert rew == 0 and not done
    ref_agents.Target.getMeasurement = orig
    for k, v in rec.items():
        out[f"{kind}_{k}"] = np.array(v)
    print(kind, "steps", steps, "tasked per step", [int((~np.isnan(z[0])).sum()) for z in rec["z"]])

np.savez_compressed(os.path.join(HERE, "env.npz"), **out)
print("wrote env.npz", {k: v.shape for k, v in out.items() if hasattr(v, "shape")})
