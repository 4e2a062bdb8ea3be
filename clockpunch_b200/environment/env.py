"""SSAScheduler on the device-resident catalog engine (mirror of punchclock/environment/env.py:39-609).

Same constructor argument (an object with ``time_step``, ``horizon``, ``agents``), the same
``reset() / step(action) / _getObs() / _getInfo()`` contract, observation and info keys, shapes
and dtypes (the asserts of env.py:279-323 are kept), the same step semantics: metrics tracker
updated with the PREVIOUS maps (env.py:386-390), actions masked by the PREVIOUS estimated
visibility (env.py:571), horizon auto-reset inside step (env.py:588-591).  What differs is
where the arithmetic runs: the whole step -- sensor + truth propagation, N filter
predict/updates, both visibility maps, float32 observation packing -- is ONE call into
libpunch_b200 (``CatalogEngine.step_host`` -> ``pc_step_host``), instead of 14 N + M scipy solves
and 2 N M Python pair tests.

``env.agents`` stays a list of real ``Sensor`` / ``Target`` objects (reference wrappers test
``type(agent) is Target``); their states are refreshed from the device lazily, when somebody
reads the list after a step.  gymnasium is optional: when importable the class is a
``gymnasium.Env`` with the reference's spaces, otherwise a plain class with the same methods.

Measurement noise: the reference draws from numpy's legacy global RNG (agents.py:181), which is
never reseeded by the env: every episode and every worker process sees different noise.  Here it
is drawn on device (Philox keyed by (seed, step counter, target)); the step counter keeps running
across ``reset()`` (also the horizon reset inside ``step``), ``seed=None`` (default) takes a fresh
seed per env instance from the OS, and ``reset(seed=s)`` restarts the stream under ``s`` (the
Gymnasium convention).  ``step(action, measurements=z)`` injects a (6, N) array instead -- that is
how tests/test_env.py replays rollouts of the live reference.
"""
from __future__ import annotations

from collections import OrderedDict
from copy import deepcopy

import numpy as np
from numpy import float32, inf, int64

from .. import _lib
from ..common.agents import Sensor, Target
from ..common.metrics import TaskingMetricTracker
from ..common.visibility import unpack_vis_words
from ..dynamics.dynamics_classes import StaticTerrestrial
from ..engine import CatalogEngine
from .env_parameters import SSASchedulerParams  # noqa: F401  (ray/build_env.py:24 imports it from this module)
from .env_utils import buildAgentInfoMap

try:  # gymnasium is not part of this image; the env does not need it to run
    import gymnasium as _gym
    from gymnasium.spaces import Box, Dict, MultiBinary, MultiDiscrete
    _EnvBase = _gym.Env
except ModuleNotFoundError:  # pragma: no cover - exercised in this image
    _gym = None
    _EnvBase = object


class _Space:
    """Shape/dtype record standing in for a gymnasium space when gymnasium is absent."""

    def __init__(self, kind, shape, dtype, low=None, high=None, nvec=None):
        self.kind, self.shape, self.dtype, self.low, self.high, self.nvec = kind, tuple(shape), dtype, low, high, nvec

    def sample(self, rng=None):
        rng = rng or np.random.default_rng()
        if self.kind == "multidiscrete":
            return rng.integers(0, self.nvec)
        if self.kind == "multibinary":
            return rng.integers(0, 2, size=self.shape)
        return rng.normal(size=self.shape).astype(self.dtype)


def fresh_seed() -> int:
    """63 bits from the OS: distinct per env instance and per (Ray) worker process."""
    import os
    return int.from_bytes(os.urandom(8), "little") >> 1


_IMMUTABLE = (int, float, str, bool, bytes, type(None), np.generic)


def _independent_copy(v):
    """What deepcopy(v) returns for the kinds of values the info dict holds, without visiting every scalar:
    arrays are copied, lists of immutables are copied shallowly (their items cannot be changed in place), lists of
    flat dicts (the agent map) entry by entry; anything else falls back to deepcopy."""
    if isinstance(v, np.ndarray):
        return v.copy()
    if isinstance(v, _IMMUTABLE):
        return v
    if isinstance(v, list):
        kinds = set(map(type, v))                       # (the distinct item types: a handful, found at C speed)
        if all(issubclass(t, _IMMUTABLE) for t in kinds):
            return list(v)
        if kinds == {dict}:
            inner = set()
            for x in v:
                inner.update(map(type, x.values()))
            if all(issubclass(t, _IMMUTABLE) for t in inner):
                return [dict(x) for x in v]
    return deepcopy(v)


class SSAScheduler(_EnvBase):
    metadata = {"render_modes": [None]}

    def __init__(self, scenario_params, device: int | None = None, seed: int | None = None):
        if seed is None:
            seed = fresh_seed()
        self.time_step = scenario_params.time_step
        self.horizon = scenario_params.horizon
        self._agents = deepcopy(list(scenario_params.agents))
        self._sens_idx = [k for k, a in enumerate(self._agents) if isinstance(a, Sensor)]
        self._targ_idx = [k for k, a in enumerate(self._agents) if isinstance(a, Target)]
        self.num_sensors, self.num_targets = len(self._sens_idx), len(self._targ_idx)
        self.num_agents = len(self._agents)
        assert self.num_sensors + self.num_targets == self.num_agents, "agents must be Sensor or Target"
        self.info = {"num_targets": self.num_targets, "num_sensors": self.num_sensors,
                     "num_agents": self.num_agents, "agent_map": buildAgentInfoMap(self._agents)}
        self.reset_params = deepcopy(scenario_params)
        self.sensor_ids = [self._agents[k].agent_id for k in self._sens_idx]
        self.target_ids = [self._agents[k].agent_id for k in self._targ_idx]
        self.tracker = TaskingMetricTracker(self.sensor_ids, self.target_ids)
        N, M, A = self.num_targets, self.num_sensors, self.num_agents
        if _gym is not None:
            self.observation_space = Dict({
                "eci_state": Box(low=-inf, high=inf, shape=(6, A)),
                "est_cov": Box(low=-inf, high=inf, shape=(N, 6, 6)),
                "vis_map_est": MultiBinary((N, M)),
                "obs_staleness": Box(low=0, high=inf, shape=(1, N)),
                "num_tasked": Box(low=0, high=inf, shape=(1, N), dtype=int)})
            self.action_space = MultiDiscrete((N + 1) * np.ones([M]))
        else:
            self.observation_space = {
                "eci_state": _Space("box", (6, A), float32, -inf, inf), "est_cov": _Space("box", (N, 6, 6), float32, -inf, inf),
                "vis_map_est": _Space("multibinary", (N, M), int64), "obs_staleness": _Space("box", (1, N), float32, 0, inf),
                "num_tasked": _Space("box", (1, N), int64, 0, inf)}
            self.action_space = _Space("multidiscrete", (M,), int64, nvec=(N + 1) * np.ones(M, dtype=int64))
        self._engine = self._build_engine(device, seed)
        self._hb = self._engine.host_buffers(pinned=True)
        # column k of (6, M+N) arrays is agent k: where each engine row goes
        self._perm = np.array(self._sens_idx + self._targ_idx)
        self._agents_stale = False
        self._vis_truth_words = None

    # ------------------------------------------------------------------ construction
    def _build_engine(self, device, seed):
        sens = [self._agents[k] for k in self._sens_idx]
        targ = [self._agents[k] for k in self._targ_idx]
        t0 = {float(a.time) for a in self._agents}
        assert len(t0) == 1, "all agents must start at the same time"
        filters = [t.target_filter for t in targ]
        q, r = np.asarray(filters[0].q_matrix, float), np.asarray(filters[0].r_matrix, float)
        for f in filters:
            if not (np.array_equal(f.q_matrix, q) and np.array_equal(f.r_matrix, r)):
                raise NotImplementedError("the batched step shares Q and R across targets (as genFilters, "
                                          "env_parameters.py:326-393, builds them)")
        kinds = np.array([_lib.KIND_TERRESTRIAL if isinstance(a.dynamics, StaticTerrestrial) else _lib.KIND_SATELLITE
                          for a in sens], dtype=np.uint8)
        col = lambda xs: np.stack([np.asarray(x, float).reshape(6) for x in xs], axis=1) if xs else np.zeros((6, 0))
        return CatalogEngine(
            col([a.eci_state for a in sens]), kinds, col([a.eci_state for a in targ]),
            col([f.est_x for f in filters]), np.stack([np.asarray(f.est_p, float) for f in filters]), q, r,
            time=t0.pop(), device=device, seed=seed,
            init_num_tasked=[a.num_tasked for a in targ], init_last_time_tasked=[a.last_time_tasked for a in targ])

    # ------------------------------------------------------------------ agents view
    @property
    def agents(self):
        """List of Sensor / Target objects, synchronised with the device state on demand."""
        if self._agents_stale:
            e = self._engine
            sx, tx, ex = e.host("sens_x"), e.host("truth_x"), e.host("est_x")
            ep, pp = e.host("est_p"), e.host("pred_p")
            nt, lt = e.host("num_tasked"), e.host("last_time_tasked")
            for j, k in enumerate(self._sens_idx):
                a = self._agents[k]
                a.eci_state, a.time = sx[:, j].copy(), e.time
            for i, k in enumerate(self._targ_idx):
                a = self._agents[k]
                a.eci_state, a.time = tx[:, i].copy(), e.time
                a.num_tasked, a.last_time_tasked = int(nt[i]), float32(lt[i])
                f = a.target_filter
                f.est_x, f.est_p, f.pred_p, f.time = ex[:, i].copy(), ep[i].copy(), pp[i].copy(), e.time
                if nt[i] > 0:
                    f.last_measurement_time = float(lt[i])
            self._agents_stale = False
        return self._agents

    # ------------------------------------------------------------------ gym API
    def reset(self, *, seed=None, options=None):
        if _gym is not None:
            super().reset(seed=seed)
        self.tracker.reset()
        self.info["num_steps"] = 0
        self.info["time_now"] = 0.0
        self.info["num_tasked"] = np.zeros(self.num_targets, dtype=int).tolist()
        self._copy_tracker_to_info()
        self._engine.reset(seed=seed)
        self._agents = deepcopy(list(self.reset_params.agents))
        self._agents_stale = False
        self._engine.pack_obs()
        hb = self._hb
        hb["obs"].copy_(self._engine.obs_f32)
        hb["vis_est"].copy_(self._engine.vis_est)
        hb["num_tasked"].copy_(self._engine.num_tasked)
        self.updateInfoPostTasking()
        return self._getObs(), self._getInfo()

    def _copy_tracker_to_info(self):
        t = self.tracker
        self.info["targets_tasked"] = t.targets_tasked
        self.info["num_unique_targets_tasked"] = t.unique_tasks
        self.info["num_non_vis_taskings_est"] = t.non_vis_tasked_est
        self.info["num_non_vis_taskings_truth"] = t.non_vis_tasked_truth
        self.info["non_vis_by_sensor_est"] = t.non_vis_by_sensor_est
        self.info["non_vis_by_sensor_truth"] = t.non_vis_by_sensor_truth
        self.info["num_multiple_taskings"] = t.multiple_taskings

    def updateInfoPreTasking(self, action):
        """Time, step count and the tasking metrics, computed with the maps of the END of the
        previous step (env.py:340-390), from the packed words (O(M))."""
        self.info["time_now"] += self.time_step
        self.info["num_steps"] += 1
        self.tracker.update_packed(np.asarray(action), self._hb["vis_est"].numpy(), self._vis_truth_words)
        self._copy_tracker_to_info()

    def updateInfoPostTasking(self):
        """Info entries that depend on the new states (env.py:392-462), from the host copies of the
        buffers the step kernels wrote."""
        e, M, N, A = self._engine, self.num_sensors, self.num_targets, self.num_agents
        obs = self._hb["obs"].numpy()
        inv = np.argsort(self._perm)                      # engine column order -> agent order
        self.info["est_x"] = obs[: 6 * A].reshape(6, A)[:, inv].copy()
        self.info["est_cov"] = obs[6 * A: 6 * A + 36 * N].reshape(N, 6, 6).copy()
        self.info["filter_pred_p"] = e.host("pred_p").astype(float32)
        tr_pos = np.trace(self.info["est_cov"][:, :3, :3], axis1=1, axis2=2)
        tr_vel = np.trace(self.info["est_cov"][:, 3:, 3:], axis1=1, axis2=2)
        self.info["mean_pos_var"], self.info["mean_vel_var"] = np.mean(tr_pos), np.mean(tr_vel)
        true = np.concatenate([e.host("sens_x"), e.host("truth_x")], axis=1)[:, inv]
        self.info["true_states"] = true.astype(float32)
        self.info["num_tasked"] = list(self._hb["num_tasked"].numpy().astype(int64))
        self.info["obs_staleness"] = [float(v) for v in obs[6 * A + 36 * N:].astype(np.float64)]
        self._vis_truth_words = e.vis_truth.cpu().numpy().view(np.uint32)
        self.info["vis_map_truth"] = unpack_vis_words(self._vis_truth_words, M)
        self.info["vis_map_est"] = unpack_vis_words(self._hb["vis_est"].numpy().view(np.uint32), M)

    def _getObs(self) -> OrderedDict:
        """env.py:261-270: five entries of (a copy of) the info."""
        self._check_info()
        info = self.info
        return OrderedDict({
            "eci_state": info["est_x"].copy(), "est_cov": info["est_cov"].copy(), "vis_map_est": info["vis_map_est"].copy(),
            "obs_staleness": np.asarray(info["obs_staleness"], dtype=float32).reshape((1, -1)),
            "num_tasked": np.asarray(info["num_tasked"]).reshape((1, -1))})

    def _getInfo(self) -> dict:
        """A deep copy of the info after the reference's shape / dtype checks (env.py:279-325).  `deepcopy` walks
        every Python scalar of the per-target lists and every entry of the agent map (60 ms per step at 10 000
        targets); the same independent copy is built per value kind instead."""
        self._check_info()
        return {k: _independent_copy(v) for k, v in self.info.items()}

    def _check_info(self) -> None:
        i, N, M, A = self.info, self.num_targets, self.num_sensors, self.num_agents
        assert i["est_x"].shape == (6, A), "est_x must have shape (6, N+M)"
        assert i["est_cov"].shape == (N, 6, 6), "est_cov must have shape (N, 6, 6)"
        assert isinstance(i["est_cov"][0, 0, 0], float32), "Entries of est_cov must be floats"
        assert i["true_states"].shape == (6, A), "true_states shape must be (6, N+M)"
        assert isinstance(i["true_states"][0, 0], float32), "true_states dtype must be float32"
        assert i["vis_map_est"].shape == (N, M), "vis_map_est shape must be (N, M)"
        assert isinstance(i["vis_map_est"][0, 0], int64), "vis_map_est entries must be ints"
        assert i["vis_map_truth"].shape == (N, M), "vis_map_truth shape must be (N, M)"
        assert isinstance(i["vis_map_truth"][0, 0], int64), "vis_map_truth entries must be ints"
        assert len(i["num_tasked"]) == N, "num_tasked must be N-long list"
        assert isinstance(i["num_tasked"][0], int64), "Entries of num_tasked must be ints"
        assert len(i["obs_staleness"]) == N, "obs_staleness must be N-long list"
        assert isinstance(i["obs_staleness"][0], float), "Entries of obs_staleness must be floats"

    def _earnReward(self, actions) -> float:
        return 0

    def step(self, action, measurements=None):
        """(observation, reward, done, truncated, info) -- env.py:533-597.  `measurements`:
        optional (6, N) array used instead of the device draw for the targets that end up tasked."""
        action = np.asarray(action)
        self.updateInfoPreTasking(action)
        e, hb = self._engine, self._hb
        hb["actions"][: self.num_sensors].copy_(e.torch.from_numpy(action.astype(np.int64)))
        # the reference propagates every agent TO info["time_now"] (env.py:560-562), which is one time step
        # ahead of the agents unless a caller has advanced the clock by hand (updateInfoPreTasking on its own,
        # as tests/environment/test_env.py:113-126 does): then the interval is longer
        dt = float(self.info["time_now"]) - float(e.time)
        if abs(dt - self.time_step) <= 1e-9 * max(1.0, abs(float(self.info["time_now"]))):
            dt = self.time_step
        if measurements is None:
            e.step_host(hb, dt)
        else:
            z = np.where(np.isnan(measurements), 0.0, np.asarray(measurements, float))
            e.step(dt=dt, z=z, actions=hb["actions"][: self.num_sensors].numpy())
            e.pack_obs()
            hb["obs"].copy_(e.obs_f32)
            hb["vis_est"].copy_(e.vis_est)
            hb["num_tasked"].copy_(e.num_tasked)
        self._agents_stale = True
        reward = self._earnReward(action)
        self.updateInfoPostTasking()
        observation, info = self._getObs(), self._getInfo()
        if self.info["num_steps"] == self.horizon:
            done = True
            print("Horizon reached, resetting...")
            self.reset()
        else:
            done = False
        return observation, reward, done, False, info

    # -- the two halves of the reference's step as separate calls (env.py:464-506; its own
    #    tests/environment/test_env.py:157-163 drives them by hand between updateInfoPreTasking and
    #    updateInfoPostTasking) -------------------------------------------------------------------------
    def _propagateAgents(self, time):
        """Propagate all agents to `time` (env.py:503-506).  The engine advances truth states and filters in ONE
        launch, so this only records the instant; the `_taskAgents` call that follows it in the reference's step
        (env.py:560-575) performs both.  `env.agents` read in between still shows the previous instant."""
        self._pending_time = float(time)

    def _taskAgents(self, actions_array, vis_map_est) -> None:
        """(N, M) binary assignments, masked by the estimated map: a target with a 1 left in its row gets a
        measurement, every other target's filter is advanced without one (env.py:464-501)."""
        masked = np.multiply(np.asarray(actions_array), np.asarray(vis_map_est))
        tasked = (masked == 1).any(axis=1).astype(np.uint8)
        t_next = getattr(self, "_pending_time", None)
        e, hb = self._engine, self._hb
        e.step(t_next=float(self.info["time_now"]) if t_next is None else t_next, tasked=tasked)
        self._pending_time = None
        e.pack_obs()
        hb["obs"].copy_(e.obs_f32)
        hb["vis_est"].copy_(e.vis_est)
        hb["num_tasked"].copy_(e.num_tasked)
        self._agents_stale = True

    def _getVarMeans(self):
        est_cov = self.info["est_cov"]
        return (np.mean(np.trace(est_cov[:, :3, :3], axis1=1, axis2=2)),
                np.mean(np.trace(est_cov[:, 3:, 3:], axis1=1, axis2=2)))

    def getEstStates(self):
        return self.info["est_x"].copy()

    def __deepcopy__(self, memo):
        """wrappers and SimRunner deep-copy the env every step (utilities.py:370-374,
        sim_runner.py:173,229): device buffers are cloned, the library context is shared."""
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k == "_engine":
                new._engine = v.clone()
            elif k == "_hb":
                continue
            else:
                setattr(new, k, deepcopy(v, memo))
        new._hb = new._engine.host_buffers(pinned=True)          # views of one pinned block: rebuild, then fill
        new._hb["block"].copy_(self._hb["block"])
        new._hb["actions"].copy_(self._hb["actions"])
        return new
