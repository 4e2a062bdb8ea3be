"""Access-window forecasting (mirror of punchclock/schedule_tree/access_windows.py:21-537).

Same class, constructor arguments, methods and return shapes as the reference's
``AccessWindowCalculator``.  ``calcVisHist`` -- in the reference T x (N + M) solve_ivp calls and
T N M Python pair tests per call, repeated every env step by the NumWindows wrapper -- is one call
into the CUDA library (pc_vis_history_host: 3 launches per epoch on device-resident states).
The window bookkeeping on the resulting (T, N, M) array is the reference's own numpy logic,
vectorised.
"""
from __future__ import annotations

from copy import deepcopy
from typing import Tuple

import numpy as np
from numpy import arange, ndarray

from .. import _lib
from ..common.constants import getConstants
from ..common.visibility import BINARY_TOL
from ..dynamics.dynamics_classes import DynamicsModel, SatDynamicsModel, StaticTerrestrial


def _kinds(dynamics, n: int) -> ndarray:
    if isinstance(dynamics, (str, DynamicsModel)):
        dynamics = [dynamics] * n
    out = np.empty(n, dtype=np.uint8)
    for i, d in enumerate(dynamics):
        if isinstance(d, str):
            assert d in ["terrestrial", "satellite"]
            out[i] = _lib.KIND_TERRESTRIAL if d == "terrestrial" else _lib.KIND_SATELLITE
        elif isinstance(d, StaticTerrestrial):
            out[i] = _lib.KIND_TERRESTRIAL
        elif isinstance(d, SatDynamicsModel):
            out[i] = _lib.KIND_SATELLITE
        else:
            raise NotImplementedError("only the built-in satellite / terrestrial dynamics have a GPU path")
    return out


class AccessWindowCalculator:
    """Calculates access windows between sensors and targets (access_windows.py:21-146)."""

    def __init__(self, num_sensors: int, num_targets: int, dynamics_sensors, dynamics_targets, horizon: int = 1,
                 dt: float = 100.0, fixed_horizon: bool = True, merge_windows: bool = True):
        assert isinstance(dynamics_sensors, (list, str, DynamicsModel))
        assert isinstance(dynamics_targets, (list, str, DynamicsModel))
        if isinstance(dynamics_sensors, list):
            assert len(dynamics_sensors) == num_sensors
        if isinstance(dynamics_targets, list):
            assert len(dynamics_targets) == num_targets
        assert horizon >= 1, "horizon not >= 1."
        self.dynamics_sensors, self.dynamics_targets = dynamics_sensors, dynamics_targets
        self._sens_kind, self._targ_kind = _kinds(dynamics_sensors, num_sensors), _kinds(dynamics_targets, num_targets)
        self.merge_windows, self.fixed_horizon = merge_windows, fixed_horizon
        self.RE = getConstants()["earth_radius"]
        self.num_sensors, self.num_targets = num_sensors, num_targets
        self.dt = min(dt, (horizon * dt) / 5)                       # access_windows.py:139
        self.horizon = horizon
        self.fixed_horizon_time = (horizon + 1) * self.dt if fixed_horizon is True else None

    def _genTime(self) -> ndarray:
        """access_windows.py:479-506."""
        if self.fixed_horizon is True:
            assert self.t_now < self.fixed_horizon_time, "Current time exceeds fixed horizon."
            return arange(start=self.t_now + self.dt, stop=self.t_now + (self.fixed_horizon_time - self.t_now),
                          step=self.dt)
        return arange(start=self.t_now + self.dt, stop=self.t_now + (self.horizon + 1) * self.dt, step=self.dt)

    def calcVisHist(self, x_sensors: ndarray, x_targets: ndarray, t: float) -> ndarray:
        """(T, N, M) binary visibility history (access_windows.py:148-195).  As in the reference,
        frame 0 is evaluated at the CURRENT states and frame i >= 1 at time_vec[i] (the first
        propagation therefore spans time_vec[1] - t)."""
        self.t_now = t
        self.time_vec = self._genTime()
        T, M, N = len(self.time_vec), self.num_sensors, self.num_targets
        xs = np.ascontiguousarray(np.asarray(x_sensors, dtype=np.float64).reshape(6, M))
        xt = np.ascontiguousarray(np.asarray(x_targets, dtype=np.float64).reshape(6, N))
        epochs = np.ascontiguousarray(np.concatenate([[float(t)], self.time_vec[1:]]), dtype=np.float64)
        wpr = (M + 31) // 32
        words = np.zeros((T, N, wpr), dtype=np.uint32)
        if T:
            ctx = _lib.get_ctx()
            ctx.check(ctx.lib.pc_vis_history_host(ctx.handle, xs.ctypes.data, self._sens_kind.ctypes.data, M,
                                                  xt.ctypes.data, self._targ_kind.ctypes.data, N, float(t),
                                                  epochs.ctypes.data, T, 0, _lib.FORCE_TWOBODY, float(self.RE),
                                                  BINARY_TOL, words.ctypes.data))
        self._words = words
        bits = np.unpackbits(words.view(np.uint8), axis=2, bitorder="little")
        return bits[:, :, :M].astype(int)

    def calcNumWindows(self, x_sensors: ndarray, x_targets: ndarray, t: float, return_vis_hist: bool = False):
        """access_windows.py:197-257."""
        vis_hist = self.calcVisHist(x_sensors=x_sensors, x_targets=x_targets, t=t)
        vis_hist_targets = self._mergeWindows(vis_hist)
        num_windows = self.sumWindows(vis_hist=vis_hist, vis_hist_targets=vis_hist_targets,
                                      merge_windows=self.merge_windows)
        time_to_window_hist = self._calcTimeToWindowHist(t_now=t, vis_hist_targets=vis_hist_targets)
        if return_vis_hist is False:
            return num_windows
        return num_windows, vis_hist, vis_hist_targets, deepcopy(self.time_vec), time_to_window_hist

    def _calcTimeToWindowHist(self, t_now: float, vis_hist_targets: ndarray) -> ndarray:
        steps = np.zeros(vis_hist_targets.shape, dtype=float)
        for i, col in enumerate(vis_hist_targets.T):
            steps[:, i] = self.calcStepsToWindow(col)
        return self.dt * steps

    def calcStepsToWindow(self, A: ndarray) -> ndarray:
        """[1,0,0,1,0,1,0,0] -> [0,2,1,0,1,0,Inf,Inf] (access_windows.py:281-305), O(len)."""
        A = np.asarray(A)
        out = np.full(len(A), np.inf)
        nxt = np.inf
        for i in range(len(A) - 1, -1, -1):
            if A[i] == 1:
                nxt = i
            out[i] = nxt - i
        return out

    # -- helpers of the reference class that its own tests call (tests/schedule_tree/test_access_windows.py);
    #    calcVisHist above does not need them: the whole history is one library call --------------------------
    def _setup(self, x_sensors: ndarray, x_targets: ndarray, t: float):
        """Time vector and surrogate agents (access_windows.py:337-366)."""
        self.t_now = t
        self.time_vec = self._genTime()
        self.sensors = self._buildAgents(x=x_sensors, time=self.t_now, dynamics=self.dynamics_sensors)
        self.targets = self._buildAgents(x=x_targets, time=self.t_now, dynamics=self.dynamics_targets)

    def _getStates(self) -> ndarray:
        """(6, M + N) current states of the surrogate agents (access_windows.py:368-380)."""
        return np.asarray([agent.eci_state for agent in self.sensors + self.targets]).squeeze().transpose()

    def _getVis(self, x_sensors: ndarray, x_targets: ndarray) -> ndarray:
        """(N, M) binary visibility of the given states (access_windows.py:382-402)."""
        from ..common.visibility import calcVisMap
        return np.asarray(calcVisMap(np.asarray(x_sensors, dtype=np.float64).reshape(6, -1),
                                     np.asarray(x_targets, dtype=np.float64).reshape(6, -1), body_radius=self.RE), dtype=int)

    def _buildAgents(self, x: ndarray, time: float, dynamics) -> list:
        """Generic agents at the columns of x (access_windows.py:431-477)."""
        from ..common.agents import Agent
        from ..dynamics.dynamics_classes import SatDynamicsModel, StaticTerrestrial
        dynamics_model_map = {"terrestrial": StaticTerrestrial(), "satellite": SatDynamicsModel()}
        if isinstance(dynamics, (str, DynamicsModel)):
            dynamics = [dynamics for _ in range(x.shape[1])]
        if isinstance(dynamics[0], str):
            dynamics = [dynamics_model_map[d] for d in dynamics]
        return [Agent(dynamics_model=dyn, init_eci_state=x[:, i], time=time) for i, dyn in zip(range(x.shape[1]), dynamics)]

    def sumWindows(self, vis_hist: ndarray, vis_hist_targets: ndarray, merge_windows: bool) -> ndarray:
        if merge_windows is False:
            return np.sum(np.sum(vis_hist, axis=2), axis=0)
        return np.sum(vis_hist_targets, axis=0)

    def _mergeWindows(self, vis_hist: ndarray) -> ndarray:
        return (np.asarray(vis_hist).sum(axis=2) > 0).astype(int)
