"""Golden rollout of a WRAPPED SSAScheduler built by the reference's own buildEnv (run in the build
container only).

    python tests/golden/make_golden_seam.py      # writes tests/golden/seam.npz

The environment configuration is the one the reference's tests generate for their simulation tests
(tests/simulation/data/gen_config_env.py: 2 fixed ground sensors, 4 targets drawn in COE, and the wrapper
stack FilterObservation -> CopyObsInfoItem -> VisMap2ActionMask -> CopyObsInfoItem -> NestObsItems ->
IdentityWrapper -> FlatDict), with a seed added so that the run is reproducible, a longer horizon, and
8 steps of a visibility-aware action sequence.  Everything that runs is the reference's code
(ray/build_env.py:33-165, environment/*_wrappers.py, env.py, agents, filters, propagator) under the
test-only stand-ins of _ref_stubs.py (functional mini-gymnasium, satvis -> oracle restatement).

Recorded: the wrapped observation / info of every step (what a policy sees), and -- in the key
format of env.npz -- the base environment's states, actions and drawn measurements, so that
  * tests/test_seam.py can rebuild the SAME wrapped env through `install_as_punchclock()` (reference
    buildEnv + wrappers on top of the mirrored SSAScheduler) and compare, and
  * tests/test_env.py can replay the base env on the GPU with the measurements injected.
"""
import os
import sys
from copy import deepcopy

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

RE = 6378.1363
NUM_SENSORS, NUM_TARGETS, STEPS = 2, 4, 8


def env_config():
    """tests/simulation/data/gen_config_env.py:20-118, with a seed and horizon 20, and with agents that can
    see each other: the reference's file puts the sensors at |r| = RE exactly and draws the targets'
    semi-major axes around 5000 km (inside the Earth), so nothing is ever visible or tasked there; here
    the sensors sit 10 m above the surface and the targets at 9000 +- 500 km."""
    from numpy import array, diag, pi
    from punchclock.common.transforms import ecef2eci
    sensor_init_ecef = array([[RE + 0.01, 0, 0, 0, 0, 0], [0, RE + 0.01, 0, 0, 0, 0]])
    sensor_init_eci = ecef2eci(sensor_init_ecef.transpose(), 0).transpose()
    agent_params = {
        "num_sensors": NUM_SENSORS, "num_targets": NUM_TARGETS, "sensor_dynamics": "terrestrial",
        "target_dynamics": "satellite", "sensor_dist": None, "target_dist": "normal", "sensor_dist_frame": None,
        "target_dist_frame": "COE", "sensor_dist_params": None,
        "target_dist_params": [[9000, 500], [0, 0], [0, pi], [0, 2 * pi], [0, 2 * pi], [0, 2 * pi]],
        "fixed_sensors": sensor_init_eci, "fixed_targets": None,
    }
    temp_matrix = diag([1, 1, 1, 0.01, 0.01, 0.01])
    filter_params = {"Q": 0.001 * temp_matrix, "R": 0.1 * temp_matrix, "p_init": 10 * temp_matrix}
    constructor_params = {"wrappers": [
        {"wrapper": "FilterObservation", "wrapper_config": {"filter_keys": ["vis_map_est", "est_cov"]}},
        {"wrapper": "CopyObsInfoItem", "wrapper_config": {"copy_from": "obs", "copy_to": "info",
                                                          "from_key": "vis_map_est", "to_key": "vm_copy"}},
        {"wrapper": "VisMap2ActionMask", "wrapper_config": {"obs_info": "info", "vis_map_key": "vm_copy",
                                                            "new_key": "action_mask"}},
        {"wrapper": "CopyObsInfoItem", "wrapper_config": {
            "copy_from": "info", "copy_to": "obs", "from_key": "action_mask", "to_key": "action_mask",
            "info_space_config": {"space": "MultiBinary", "n": [NUM_TARGETS + 1, NUM_SENSORS]}}},
        {"wrapper": "NestObsItems", "wrapper_config": {"new_key": "observations",
                                                       "keys_to_nest": ["vis_map_est", "est_cov"]}},
        {"wrapper": "IdentityWrapper"},
        {"wrapper": "FlatDict"},
    ]}
    return {"horizon": 20, "agent_params": agent_params, "filter_params": filter_params, "time_step": 100,
            "seed": 315, "constructor_params": constructor_params}


NP_SEED = 2718      # numpy's legacy global RNG: initial-estimate noise (env_parameters.py:372-377) + measurements


def actions_for(step, action_mask):
    """Deterministic visibility-aware policy: sensor j takes the (step + j)-th allowed row of its column
    of the wrapped observation's action mask ((N + 1, M), last row = inaction)."""
    mask = np.asarray(action_mask).reshape(NUM_TARGETS + 1, NUM_SENSORS)
    out = []
    for j in range(NUM_SENSORS):
        allowed = np.flatnonzero(mask[:, j])
        out.append(int(allowed[(step + j) % len(allowed)]))
    return np.array(out)


def flat_items(prefix, obj, out):
    """Nested dict of arrays / scalars -> flat {prefix/key: array} (wrapped obs and info)."""
    if isinstance(obj, dict):
        for k, v in obj.items():
            flat_items(f"{prefix}/{k}", v, out)
    elif isinstance(obj, (list, tuple)) and obj and isinstance(obj[0], dict):
        return                                   # agent_map and friends: not numeric
    else:
        try:
            a = np.asarray(obj)
        except Exception:
            return
        if a.dtype != object and a.dtype.kind in "fiub":
            out[prefix] = a


def rollout(build_env, target_class, record_base=None):
    """Build, reset and step the wrapped env; returns {key: stacked arrays}."""
    np.random.seed(NP_SEED)
    env = build_env(deepcopy(env_config()))
    out = {}
    obs, info = env.reset()
    flat_items("reset_obs", obs, out)
    flat_items("reset_info", info, out)
    rec = {}
    acts = []
    for s in range(STEPS):
        a = actions_for(s, obs["action_mask"])
        acts.append(a)
        if record_base is not None:
            record_base("pre", env, s, a)
        obs, rew, done, trunc, info = env.step(a)
        if record_base is not None:
            record_base("post", env, s, a)
        step_items = {}
        flat_items("obs", obs, step_items)
        flat_items("info", info, step_items)
        step_items["reward"] = np.asarray(rew, dtype=float)
        for k, v in step_items.items():
            rec.setdefault(k, []).append(v)
        assert not done
    for k, v in rec.items():
        if len(v) == STEPS and all(x.shape == v[0].shape for x in v):
            out[f"step_{k}"] = np.stack(v)
    out["actions"] = np.stack(acts)
    return out, env


if __name__ == "__main__":
    import _ref_stubs  # noqa: F401  (registers the stand-ins and the bare `punchclock` package -> reference)
    from punchclock.common import agents as ref_agents
    from punchclock.ray.build_env import buildEnv

    zlog, base = [], {k: [] for k in ("action", "z", "eci", "cov", "vis_est", "vis_truth", "stale", "num_tasked",
                                       "true", "pred_p", "mpv", "mvv", "uniq", "nv_est", "nv_truth", "multi",
                                       "nvs_est", "nvs_truth", "time", "estx64", "estp64")}
    orig = ref_agents.Target.getMeasurement

    def patched(self, _orig=orig):
        z = _orig(self)
        zlog.append((self.agent_id, z.copy()))
        return z
    ref_agents.Target.getMeasurement = patched

    def record_base(when, env, s, action):
        b = env.unwrapped
        if when == "pre":
            zlog.clear()
            return
        N = NUM_TARGETS
        obs, info = b._getObs(), b._getInfo()
        z = np.full((6, N), np.nan)
        for aid, zz in zlog:
            z[:, list(b.target_ids).index(aid)] = zz
        tg = [a for a in b.agents if type(a) is ref_agents.Target]
        for k, v in (("action", action), ("z", z), ("eci", obs["eci_state"]), ("cov", obs["est_cov"]),
                     ("vis_est", obs["vis_map_est"]), ("vis_truth", info["vis_map_truth"]), ("stale", obs["obs_staleness"]),
                     ("num_tasked", obs["num_tasked"]), ("true", info["true_states"]), ("pred_p", info["filter_pred_p"]),
                     ("mpv", info["mean_pos_var"]), ("mvv", info["mean_vel_var"]), ("uniq", info["num_unique_targets_tasked"]),
                     ("nv_est", info["num_non_vis_taskings_est"]), ("nv_truth", info["num_non_vis_taskings_truth"]),
                     ("multi", info["num_multiple_taskings"]), ("nvs_est", np.array(info["non_vis_by_sensor_est"])),
                     ("nvs_truth", np.array(info["non_vis_by_sensor_truth"])), ("time", info["time_now"]),
                     ("estx64", np.array([a.target_filter.est_x for a in tg]).T),
                     ("estp64", np.array([a.target_filter.est_p for a in tg]))):
            base[k].append(np.array(v))

    # initial conditions of the base env (before the first step), as env.npz stores them
    np.random.seed(NP_SEED)
    probe = buildEnv(deepcopy(env_config()))
    pobs, pinfo = probe.unwrapped.reset()
    b = probe.unwrapped
    tg = [a for a in b.agents if type(a) is ref_agents.Target]
    kind = "seam"
    out = {f"{kind}_sensors": np.array([a.eci_state.squeeze() for a in b.agents if type(a) is ref_agents.Sensor]).T,
           f"{kind}_targets": np.array([a.eci_state.squeeze() for a in tg]).T,
           f"{kind}_est_x0": np.array([np.asarray(a.target_filter.est_x).squeeze() for a in tg]).T,
           f"{kind}_P0": env_config()["filter_params"]["p_init"], f"{kind}_Q": env_config()["filter_params"]["Q"],
           f"{kind}_R": env_config()["filter_params"]["R"], f"{kind}_dt": 100.0,
           f"{kind}_ids": np.array([a.agent_id for a in b.agents]),
           f"{kind}_reset_vis_est": pinfo["vis_map_est"], f"{kind}_reset_vis_truth": pinfo["vis_map_truth"],
           f"{kind}_reset_eci": pobs["eci_state"], f"{kind}_reset_cov": pobs["est_cov"]}
    wrapped, env = rollout(buildEnv, ref_agents.Target, record_base)
    ref_agents.Target.getMeasurement = orig
    for k, v in base.items():
        out[f"{kind}_{k}"] = np.array(v)
    for k, v in wrapped.items():
        out[f"wrapped/{k}"] = v
    out["wrapper_chain"] = np.array(str(env))
    np.savez_compressed(os.path.join(HERE, "seam.npz"), **out)
    print("wrote seam.npz:", str(env))
    print("tasked per step:", [int((~np.isnan(z[0])).sum()) for z in base["z"]])
    print("wrapped keys:", sorted(k for k in wrapped if k.startswith("step_obs") or k.startswith("reset_obs")))
