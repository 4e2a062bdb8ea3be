"""Golden rollout of an SSAScheduler wrapped with the reference's INFO wrappers that call the hot path themselves
(run in the build container only).

    python tests/golden/make_golden_infowrap.py      # writes tests/golden/infowrap.npz

environment/info_wrappers.py: `NumWindows` (:163-637; an AccessWindowCalculator forecast -- (T, N, M) visibility
history from chained propagations, recalculated on every step) and `VisMap` (:1572-1771; continuous visibility
value and its time derivative of every pair from the ESTIMATED states, `calcVisMapAndDerivative`), stacked by the
reference's buildEnv on the scenario of tests/golden/make_golden_seam.py.  Everything that runs is the reference's
code under the stand-ins of _ref_stubs.py (satvis arithmetic: the oracle's restatement, parity unpinned).
tests/test_seam.py repeats the rollout through `install_as_punchclock()`: the same reference wrappers, on top of the
mirrored env / AccessWindowCalculator / calcVisMapAndDerivative.
"""
import os
import sys
from copy import deepcopy

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_seam as seam  # noqa: E402

STEPS = 5
KEYS = ("num_windows_left", "vis_forecast", "time_to_next_window", "vis_map", "vis_map_dot", "est_x", "num_tasked")


def env_config():
    cfg = seam.env_config()
    cfg["horizon"] = 8
    cfg["constructor_params"] = {"wrappers": [
        {"wrapper": "NumWindows", "wrapper_config": {"use_estimates": True, "open_loop": False}},
        {"wrapper": "VisMap", "wrapper_config": {"binary": False, "state_key": "est_x"}},
    ]}
    return cfg


def rollout(build_env, der_fn=None):
    """der_fn (the golden run passes the reference's calcVisMapAndDerivative): also record the derivative map
    evaluated by that function on float64 copies of the float32 info["est_x"] the VisMap wrapper uses -- see the
    note on `vis_map_dot64` in tests/test_seam.py."""
    np.random.seed(seam.NP_SEED)
    env = build_env(deepcopy(env_config()))
    obs, info = env.reset()
    out = {f"reset_{k}": np.asarray(info[k]) for k in KEYS}
    rec = {k: [] for k in KEYS}
    acts = []
    M = seam.NUM_SENSORS

    def dot64(info):
        x = np.asarray(info["est_x"]).astype(np.float64)
        return np.nan_to_num(der_fn(sensor_states=x[:, :M], target_states=x[:, M:], binary=False)[1], nan=0).astype(np.float32)

    if der_fn is not None:
        out["reset_vis_map_dot64"] = dot64(info)
        rec["vis_map_dot64"] = []
    for s in range(STEPS):
        vm = np.asarray(obs["vis_map_est"])                     # (N, M) base observation
        a = []
        for j in range(seam.NUM_SENSORS):
            seen = np.flatnonzero(vm[:, j])
            a.append(int(seen[(s + j) % len(seen)]) if len(seen) else seam.NUM_TARGETS)
        acts.append(a)
        obs, rew, done, trunc, info = env.step(np.array(a))
        for k in KEYS:
            rec[k].append(np.asarray(info[k]))
        if der_fn is not None:
            rec["vis_map_dot64"].append(dot64(info))
    out["actions"] = np.array(acts)
    for k, v in rec.items():
        shapes = {x.shape for x in v}
        if len(shapes) == 1:
            out[f"step_{k}"] = np.stack(v)
        else:                                                    # the forecast shrinks towards the fixed horizon
            for s, x in enumerate(v):
                out[f"step{s}_{k}"] = x
    return out, env


if __name__ == "__main__":
    import _ref_stubs  # noqa: F401
    from punchclock.common.visibility import calcVisMapAndDerivative
    from punchclock.ray.build_env import buildEnv
    out, env = rollout(buildEnv, calcVisMapAndDerivative)
    print("float32 vs float64 evaluation of the reference's own derivative map, max |difference| per step:",
          np.abs(out["step_vis_map_dot"] - out["step_vis_map_dot64"]).reshape(STEPS, -1).max(axis=1))
    np.savez_compressed(os.path.join(HERE, "infowrap.npz"), **out)
    print("wrote infowrap.npz:", str(env)[:90])
    for k, v in out.items():
        print(" ", k, v.shape, v.dtype, np.asarray(v).ravel()[:4])
