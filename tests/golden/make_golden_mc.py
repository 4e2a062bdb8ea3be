"""Golden Monte-Carlo run of the all-reference stack (run in the build container only).

    python tests/golden/make_golden_mc.py      # writes tests/golden/mc.npz

`MonteCarloRunner` (simulation/mc.py:32-660) in its serial mode: two episodes of one trial, environment built from
a JSON-able config by the reference's buildEnv (seed from the config: the same initial conditions every episode),
policy built from a config by `buildCustomPolicy` (GreedyCovariance, epsilon = 0), each episode run by SimRunner, the
per-step results turned into a pandas DataFrame with the post-processed covariance columns and saved / re-read by the
runner itself.  tests/test_seam.py repeats it through `install_as_punchclock()` and compares the numeric columns.
"""
import os
import sys
import tempfile
from copy import deepcopy

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_simrunner as sr  # noqa: E402

COLUMNS = ("step", "episode", "action", "reward", "obs_action_mask", "obs_observations.eci_state", "obs_observations.est_cov",
           "obs_observations.obs_staleness", "obs_observations.vis_map_est", "info_num_tasked", "info_est_x",
           "info_filter_pred_p", "info_true_states", "info_vis_map_truth", "info_mean_pos_var", "info_time_now",
           "info_num_unique_targets_tasked", "cov_tr", "pos_cov_tr", "vel_cov_tr", "cov_mean", "cov_max")


def run(buildEnv, buildSpaceConfig, MonteCarloRunner):
    cfg = dict(sr.config())
    cfg["seed"], cfg["horizon"] = 11, 5
    cfg["agent_params"]["fixed_sensors"] = np.asarray(cfg["agent_params"]["fixed_sensors"]).tolist()
    cfg["agent_params"]["fixed_targets"] = np.asarray(cfg["agent_params"]["fixed_targets"]).tolist()
    np.random.seed(5)
    env = buildEnv(deepcopy(cfg))
    policy_config = {"policy": "GreedyCovariance", "observation_space": buildSpaceConfig(env.observation_space).toDict(),
                     "action_space": buildSpaceConfig(env.action_space).toDict(), "epsilon": 0}
    mcr = MonteCarloRunner(num_episodes=2, policy_configs=[policy_config], results_dir=tempfile.mkdtemp(), env_configs=[cfg],
                           print_status=False, multiprocess=False, static_initial_conditions=None, save_format="pkl")
    mcr.runMC()
    df = mcr.getDataFrame()
    out = {"n_rows": np.array(len(df)), "columns": np.array(sorted(map(str, df.columns)))}
    for c in COLUMNS:
        out[c] = np.stack([np.asarray(v, dtype=float) for v in df[c]])
    return out


if __name__ == "__main__":
    import _ref_stubs  # noqa: F401
    from punchclock.policies.policy_builder import buildSpaceConfig
    from punchclock.ray.build_env import buildEnv
    from punchclock.simulation.mc import MonteCarloRunner
    out = run(buildEnv, buildSpaceConfig, MonteCarloRunner)
    np.savez_compressed(os.path.join(HERE, "mc.npz"), **out)
    print("wrote mc.npz:", int(out["n_rows"]), "rows;", {k: v.shape for k, v in out.items() if k not in ("columns",)})
