"""Oracle: one SSAScheduler physics step over catalog arrays (test infrastructure).

Restates the per-step call stack of punchclock/environment/env.py:533-597
 on plain arrays, keeping the reference's loop structure
(one solve per agent, one filter per target, one Python call per pair):

  _propagateAgents   env.py:503-506  -> agents.py:59-71
  _taskAgents        env.py:464-501  -> agents.py:144-171 (measurement injected)
  vis maps           env.py:455-462  -> env_utils.py:22-98 -> visibility.py:56-116
"""
from __future__ import annotations

import numpy as np

from .dynamics import propagate_satellite, propagate_terrestrial
from .ukf import OracleUKF
from .visibility import calc_vis_map


class OracleCatalog:
    """M sensors + N targets with one OracleUKF per target."""

    def __init__(self, sensor_states, sensor_kinds, target_states, est_x, est_p, Q, R, time=0.0):
        self.sensors = np.array(sensor_states, dtype=float)       # (6,M)
        self.sensor_kinds = list(sensor_kinds)                     # "terrestrial" | "satellite"
        self.targets = np.array(target_states, dtype=float)        # (6,N)
        self.time = float(time)
        n = self.targets.shape[1]
        self.filters = [OracleUKF(time, est_x[:, i], est_p[i], Q, R) for i in range(n)]
        self.num_tasked = np.zeros(n, dtype=np.int64)
        self.last_time_tasked = np.zeros(n, dtype=np.float32)

    def step(self, t_next, tasked, measurements):
        """tasked: (N,) bool; measurements: (6,N) (columns used where tasked)."""
        for j, kind in enumerate(self.sensor_kinds):
            f = propagate_terrestrial if kind == "terrestrial" else propagate_satellite
            self.sensors[:, j] = f(self.sensors[:, j], self.time, t_next)
        for i in range(self.targets.shape[1]):
            self.targets[:, i] = propagate_satellite(self.targets[:, i], self.time, t_next)
        self.time = t_next
        for i, f in enumerate(self.filters):
            if tasked[i]:
                self.num_tasked[i] += 1
                f.predict_and_update(t_next, measurements[:, i])
            else:
                f.predict_and_update(t_next, None)
            if f.last_measurement_time is not None:
                self.last_time_tasked[i] = np.float32(f.last_measurement_time)
        return self.vis_maps()

    @property
    def est_x(self):
        return np.stack([f.est_x for f in self.filters], axis=1)

    @property
    def est_p(self):
        return np.stack([f.est_p for f in self.filters], axis=0)

    @property
    def pred_p(self):
        return np.stack([f.pred_p for f in self.filters], axis=0)

    def vis_maps(self):
        truth = calc_vis_map(self.sensors, self.targets)
        est = calc_vis_map(self.sensors, self.est_x)
        return truth, est


def target_step(x_truth, est_x, est_p, Q, R, t0, tf, z, sensors):
    """One *target-step* (SURVEY.md section 8d): truth propagate, UKF
    predictAndUpdate, and the target's row in both vis maps.  Used by the CPU
    baseline timer; returns (truth', est_x', est_p', pred_p, vis_truth_row, vis_est_row)."""
    xt = propagate_satellite(x_truth, t0, tf)
    f = OracleUKF(t0, est_x, est_p, Q, R)
    f.predict_and_update(tf, z)
    vt = calc_vis_map(sensors, xt.reshape(6, 1))[0]
    ve = calc_vis_map(sensors, f.est_x.reshape(6, 1))[0]
    return xt, f.est_x, f.est_p, f.pred_p, vt, ve
