"""Golden SimRunner run of the all-reference stack (run in the build container only).

    python tests/golden/make_golden_simrunner.py      # writes tests/golden/simrunner.npz

The scenario of the reference's own tests/simulation/test_sim_runner.py:86-345: 2 Earth-fixed sensors,
3 satellites at fixed initial conditions, Q = R = 0.1, P0 = 1, ten instants 30 s apart, the wrapper stack
FilterObservation -> CopyObsInfoItem -> VisMap2ActionMask -> NestObsItems -> IdentityWrapper built by
`buildEnv`, the reference's GreedyCovariance policy (epsilon = 0: deterministic) and
`SimRunner(env, policy, max_steps).runSim()` (simulation/sim_runner.py:63-350), which deep-copies the wrapped
env every step.  Everything that runs is the reference's code under the stand-ins of _ref_stubs.py.
tests/test_seam.py repeats the run through `install_as_punchclock()` and compares the SimResults.
"""
import os
import sys
from collections import OrderedDict
from copy import deepcopy

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
HORIZON = 10
NP_SEED = 1618


def config():
    sensor_ICs = np.array([[6500, 0, 0, 0, 0, 0], [0, 6500, 0, 0, 0, 0]])
    target_ICs = np.array([[7000, 0, 0, 0, 8, 0], [7000, 1000, 0, 0, 8, 0], [-7500, 1000, 0, 0, 8, 0]])
    agent_params = {
        "num_sensors": 2, "num_targets": 3, "sensor_starting_num": 101, "target_starting_num": 1,
        "sensor_dynamics": "terrestrial", "target_dynamics": "satellite", "sensor_dist_frame": None,
        "target_dist_frame": None, "sensor_dist": None, "target_dist": None, "sensor_dist_params": None,
        "target_dist_params": None, "fixed_sensors": sensor_ICs, "fixed_targets": target_ICs,
        "init_num_tasked": None, "init_last_time_tasked": None,
    }
    keys = ["eci_state", "est_cov", "obs_staleness", "vis_map_est"]
    return OrderedDict({
        "time_step": 30.0, "horizon": HORIZON, "agent_params": agent_params,
        "filter_params": {"Q": 0.1, "R": 0.1, "p_init": 1},
        "constructor_params": {"wrappers": [
            {"wrapper": "FilterObservation", "wrapper_config": {"filter_keys": list(keys)}},
            {"wrapper": "CopyObsInfoItem", "wrapper_config": {"copy_from": "obs", "copy_to": "obs",
                                                              "from_key": "vis_map_est", "to_key": "action_mask"}},
            {"wrapper": "VisMap2ActionMask", "wrapper_config": {"obs_info": "obs", "vis_map_key": "action_mask"}},
            {"wrapper": "NestObsItems", "wrapper_config": {"new_key": "observations", "keys_to_nest": list(keys)}},
            {"wrapper": "IdentityWrapper"},
        ]},
    })


def run(buildEnv, GreedyCovariance, SimRunner):
    """Build env + policy + runner from the given (reference or swapped) classes; returns flat arrays."""
    np.random.seed(NP_SEED)
    env = buildEnv(deepcopy(config()))
    env.reset()
    policy = GreedyCovariance(observation_space=env.observation_space, action_space=env.action_space, epsilon=0)
    res = SimRunner(env=env, policy=policy, max_steps=HORIZON).runSim()
    out = {"actions": np.array([np.asarray(a) for a in res.actions]),
           "reward": np.array(res.reward, dtype=float), "terminated": np.array(res.terminated, dtype=bool)}
    for key in ("eci_state", "est_cov", "obs_staleness", "vis_map_est"):
        out[f"obs_{key}"] = np.array([np.asarray(o["observations"][key]) for o in res.obs])
    out["obs_action_mask"] = np.array([np.asarray(o["action_mask"]) for o in res.obs])
    for key in ("true_states", "est_x", "num_tasked", "vis_map_truth", "mean_pos_var", "time_now", "num_steps",
                "num_unique_targets_tasked", "num_non_vis_taskings_est"):
        out[f"info_{key}"] = np.array([np.asarray(i[key]) for i in res.info])
    return out, env


if __name__ == "__main__":
    import _ref_stubs  # noqa: F401
    from punchclock.policies.greedy_cov_v2 import GreedyCovariance
    from punchclock.ray.build_env import buildEnv
    from punchclock.simulation.sim_runner import SimRunner
    out, env = run(buildEnv, GreedyCovariance, SimRunner)
    np.savez_compressed(os.path.join(HERE, "simrunner.npz"), **out)
    print("wrote simrunner.npz:", str(env)[:120])
    print("actions:\n", out["actions"].T, "\nnum_tasked final:", out["info_num_tasked"][-1], "steps", out["info_num_steps"])
