"""E SSAScheduler environments stepped as ONE catalog (BASELINE.json configs[2]:
1024 envs x 100 targets x 10 sensors for RLlib rollout stepping).

The reference scales rollouts by running one Python env per Ray worker process
(ray/build_tuner.py:169-213).  Here the E environments are laid out environment-major in one
CatalogEngine (`num_envs = E`): one pc_step_host call advances all of them -- every target is
tested only against its own environment's sensors, actions are local to each environment -- and
the float32 observation comes back as E reference observations back to back, so
`obs["eci_state"][e]` is exactly what env e's `_getObs()` would return (env.py:220-270).
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

from ..common.visibility import unpack_vis_words
from ..engine import CatalogEngine


class VectorSSAScheduler:
    """Batched env.  `sensors` (E, 6, M), `targets` (E, 6, N), `est_x` (E, 6, N), `est_p` (E, N, 6, 6)
    (or one (6, 6) P0 for all), shared Q, R, time_step and horizon."""

    def __init__(self, sensors, sensor_kinds, targets, est_x, est_p, Q, R, time_step: float, horizon: int,
                 device: int | None = None, seed: int | None = None, copy: bool = True):
        """copy (as gymnasium.vector.SyncVectorEnv's): True returns fresh observation arrays every step; False
        returns views of the pinned host block the step writes (valid until the next step / reset) and saves the
        16 MB of copies a 1024-environment observation costs."""
        self.copy = bool(copy)
        if seed is None:
            from .env import fresh_seed
            seed = fresh_seed()
        sensors, targets, est_x = (np.asarray(a, dtype=np.float64) for a in (sensors, targets, est_x))
        self.num_envs, _, self.num_sensors = sensors.shape
        self.num_targets = targets.shape[2]
        E, M, N = self.num_envs, self.num_sensors, self.num_targets
        est_p = np.asarray(est_p, dtype=np.float64)
        if est_p.shape == (6, 6):
            est_p = np.broadcast_to(est_p, (E, N, 6, 6))
        kinds = np.asarray(sensor_kinds, dtype=np.uint8)
        kinds = np.broadcast_to(kinds, (E, M)) if kinds.ndim == 1 else kinds
        flat = lambda a: np.ascontiguousarray(a.transpose(1, 0, 2).reshape(6, -1))     # (E,6,K) -> (6, E K) env-major
        self.time_step, self.horizon = float(time_step), int(horizon)
        self._engine = CatalogEngine(flat(sensors), kinds.reshape(-1), flat(targets), flat(est_x),
                                     est_p.reshape(E * N, 6, 6), Q, R, device=device, seed=seed, num_envs=E)
        self._hb = self._engine.host_buffers(pinned=True)
        self.num_steps = 0

    # ------------------------------------------------------------------
    def _obs(self) -> OrderedDict:
        E, M, N = self.num_envs, self.num_sensors, self.num_targets
        A = M + N
        o = self._hb["obs"].numpy()
        n0, n1 = 6 * E * A, 36 * E * N
        words = self._hb["vis_est"].numpy().view(np.uint32)
        own = (lambda a: a.copy()) if self.copy else (lambda a: a)
        return OrderedDict({
            "eci_state": own(o[:n0].reshape(E, 6, A)),
            "est_cov": own(o[n0: n0 + n1].reshape(E, N, 6, 6)),
            "vis_map_est": unpack_vis_words(words, M).reshape(E, N, M),
            "obs_staleness": own(o[n0 + n1:].reshape(E, 1, N)),
            "num_tasked": own(self._hb["num_tasked"].numpy().reshape(E, 1, N))})

    def reset(self, seed: int | None = None):
        """New episode; the measurement-noise stream keeps running unless `seed` restarts it."""
        e, hb = self._engine, self._hb
        e.reset(seed=seed)
        self.num_steps = 0
        e.pack_obs()
        hb["obs"].copy_(e.obs_f32)
        hb["vis_est"].copy_(e.vis_est)
        hb["num_tasked"].copy_(e.num_tasked)
        return self._obs(), {"time_now": e.time, "num_steps": 0}

    def step(self, actions, measurements=None):
        """actions (E, M), each entry a LOCAL target index 0..N (N = inaction).  Returns
        (obs, reward (E,), done (E,), truncated (E,), info); all environments share the clock, so
        they reach the horizon together and are reset together (env.py:588-591)."""
        E, M, N = self.num_envs, self.num_sensors, self.num_targets
        e, hb = self._engine, self._hb
        a = np.ascontiguousarray(np.asarray(actions, dtype=np.int64).reshape(E * M))
        hb["actions"][: E * M].copy_(e.torch.from_numpy(a))
        traces = None
        if measurements is None:
            e.step_host(hb, self.time_step)
            # float32 traces of the float32 covariance blocks, formed by the step kernel (env.py:599-609)
            traces = (hb["tr_pos"].numpy().reshape(E, N), hb["tr_vel"].numpy().reshape(E, N))
        else:
            z = np.asarray(measurements, dtype=np.float64)               # (E, 6, N)
            z = np.where(np.isnan(z), 0.0, z).transpose(1, 0, 2).reshape(6, E * N)
            e.step(dt=self.time_step, z=z, actions=a)
            e.pack_obs()
            hb["obs"].copy_(e.obs_f32)
            hb["vis_est"].copy_(e.vis_est)
            hb["num_tasked"].copy_(e.num_tasked)
        self.num_steps += 1
        obs = self._obs()
        if traces is None:
            traces = (np.trace(obs["est_cov"][:, :, :3, :3], axis1=2, axis2=3), np.trace(obs["est_cov"][:, :, 3:, 3:], axis1=2, axis2=3))
        info = {"time_now": e.time, "num_steps": self.num_steps,
                "mean_pos_var": traces[0].mean(axis=1), "mean_vel_var": traces[1].mean(axis=1)}
        done = np.full(E, self.num_steps == self.horizon)
        if done[0]:
            if not self.copy:      # the reset below rewrites the host block these views point into
                obs = OrderedDict((k, v.copy()) for k, v in obs.items())
            self.reset()
        return obs, np.zeros(E), done, np.zeros(E, dtype=bool), info
