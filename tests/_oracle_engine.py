"""TEST-ONLY CPU double of clockpunch_b200.engine.CatalogEngine built on the oracle (oracle/step.py).

The product has no CPU path (clockpunch_b200 raises without the CUDA library and a device).  The
build container has no GPU, and the GPU box has no reference tree, so the one test that needs BOTH the
reference's own callers (buildEnv + the Gymnasium wrapper stack, imported through
install_as_punchclock) and a stepping environment -- tests/test_seam.py -- plugs this double under the
mirrored SSAScheduler: same constructor, buffers and methods as the engine, arithmetic by the oracle.
It checks the Python seam (module composition, class identity, attribute protocol between the reference's
wrappers and the mirrored env); that the real engine equals the oracle is what the `-m gpu` tests prove.

Measurement noise is drawn exactly as the reference does (numpy.random.multivariate_normal on the legacy
global RNG, one draw per tasked target in target order, agents.py:173-183, env.py:493-501), so that a
rollout under the same numpy seed reproduces the live reference's measurements.
"""
from __future__ import annotations

import copy

import numpy as np
import torch

from oracle.step import OracleCatalog


def pack_vis(vm: np.ndarray) -> np.ndarray:
    """(N, M) 0/1 map -> (N, ceil(M/32)) uint32 words, bit j % 32 of word j // 32 = sensor j."""
    N, M = vm.shape
    wpr = max(1, (M + 31) // 32)
    words = np.zeros((N, wpr), np.uint32)
    for j in range(M):
        words[:, j // 32] |= (vm[:, j].astype(np.uint32) & 1) << np.uint32(j % 32)
    return words


class OracleEngine:
    def __init__(self, sensor_states, sensor_kinds, target_states, est_x, est_p, Q, R, time: float = 0.0,
                 device=None, seed=0, init_num_tasked=None, init_last_time_tasked=None, **_unused):
        self.torch = torch
        sensor_states = np.asarray(sensor_states, float).reshape(6, -1)
        target_states = np.asarray(target_states, float).reshape(6, -1)
        self.M, self.N = sensor_states.shape[1], target_states.shape[1]
        self.wpr = max(1, (self.M + 31) // 32)
        self.seed = seed
        kinds = ["terrestrial" if int(k) == 1 else "satellite" for k in np.asarray(sensor_kinds)]
        self.R = np.asarray(R, float)
        self._orc = OracleCatalog(sensor_states, kinds, target_states, np.asarray(est_x, float).reshape(6, -1),
                                  np.asarray(est_p, float).reshape(self.N, 6, 6), Q, R, time=float(time))
        if init_num_tasked is not None:
            self._orc.num_tasked[:] = np.asarray(init_num_tasked, np.int64)
        if init_last_time_tasked is not None:
            self._orc.last_time_tasked[:] = np.asarray(init_last_time_tasked, np.float32)
        self._vt, self._ve = self._orc.vis_maps()
        A = self.M + self.N
        self._n_obs = 6 * A + 37 * self.N
        self.obs_f32 = torch.zeros(self._n_obs, dtype=torch.float32)
        self.vis_est = torch.zeros((self.N, self.wpr), dtype=torch.int32)
        self.vis_truth = torch.zeros((self.N, self.wpr), dtype=torch.int32)
        self.num_tasked = torch.zeros(self.N, dtype=torch.int64)
        self._publish()
        self._snap = copy.deepcopy(self._orc)

    # -- state ------------------------------------------------------------------------------------
    @property
    def time(self):
        return self._orc.time

    def _publish(self):
        self._vt, self._ve = self._orc.vis_maps()
        self.vis_est.copy_(torch.from_numpy(pack_vis(self._ve).view(np.int32)))
        self.vis_truth.copy_(torch.from_numpy(pack_vis(self._vt).view(np.int32)))
        self.num_tasked.copy_(torch.from_numpy(self._orc.num_tasked.copy()))

    def host(self, name):
        o = self._orc
        return {"sens_x": lambda: o.sensors.copy(), "truth_x": lambda: o.targets.copy(), "est_x": lambda: o.est_x,
                "est_p": lambda: o.est_p, "pred_p": lambda: o.pred_p, "num_tasked": lambda: o.num_tasked.copy(),
                "last_time_tasked": lambda: o.last_time_tasked.copy(), "vis_truth": lambda: self._vt.copy(),
                "vis_est": lambda: self._ve.copy()}[name]()

    def pack_obs(self):
        o = self._orc
        eci = np.concatenate([o.sensors, o.est_x], axis=1).astype(np.float32)
        stale = (o.time - o.last_time_tasked.astype(np.float64)).astype(np.float32)
        flat = np.concatenate([eci.ravel(), o.est_p.astype(np.float32).ravel(), stale])
        self.obs_f32.copy_(torch.from_numpy(flat))
        return self.obs_f32

    def host_buffers(self, pinned: bool = True):
        nbytes = 4 * self._n_obs
        pad = (nbytes + 255) // 256 * 256
        block = torch.zeros(pad + self.N * (8 + 4 * self.wpr), dtype=torch.uint8)
        return {"actions": torch.zeros(max(1, self.M), dtype=torch.int64), "block": block,
                "obs": block[:nbytes].view(torch.float32),
                "num_tasked": block[pad: pad + 8 * self.N].view(torch.int64),
                "vis_est": block[pad + 8 * self.N:].view(torch.int32).view(self.N, self.wpr)}

    # -- stepping -----------------------------------------------------------------------------------
    def _tasked_from(self, actions):
        tasked = np.zeros(self.N, bool)
        for j, a in enumerate(np.asarray(actions, np.int64)[: self.M]):
            if 0 <= a < self.N and self._ve[a, j]:
                tasked[a] = True
        return tasked

    def step(self, t_next=None, tasked=None, z=None, dt=None, actions=None):
        o = self._orc
        t_next = o.time + float(dt) if t_next is None else float(t_next)
        if actions is not None:
            tasked = self._tasked_from(actions)
        tasked = np.zeros(self.N, bool) if tasked is None else np.asarray(tasked).astype(bool)
        if z is None:
            # the reference propagates every agent first, then draws one measurement per tasked target in
            # target order from the global numpy RNG (env.py:503-506, 493-501; agents.py:173-183)
            from oracle.dynamics import propagate_satellite
            z = np.zeros((6, self.N))
            for i in np.flatnonzero(tasked):
                truth = propagate_satellite(o.targets[:, i], o.time, t_next)
                z[:, i] = np.random.multivariate_normal(truth, self.R)
        o.step(t_next, tasked, np.asarray(z, float))
        self._publish()

    def step_host(self, hb, dt, **_kw):
        self.step(dt=dt, actions=hb["actions"][: self.M].numpy())
        self.pack_obs()
        hb["obs"].copy_(self.obs_f32)
        hb["vis_est"].copy_(self.vis_est)
        hb["num_tasked"].copy_(self.num_tasked)

    def reset(self, seed=None, rewind_rng=False):
        self._orc = copy.deepcopy(self._snap)
        self._publish()

    def clone(self):
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        new = object.__new__(OracleEngine)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            setattr(new, k, v if k == "torch" else copy.deepcopy(v, memo))
        return new


def install_cpu_doubles():
    """Route the CUDA entry points that the mirrored modules reach to the oracle, for the CPU-only seam tests
    (see the module docstring): the engine, the frame / element conversions used for initial conditions as
    Python functions; the host-buffer entry points of the library (propagation, filter step,
    visibility map / value / derivative, visibility history) through an object with the library's own interface.
    Returns that context (its `lib.calls` counts the entry points used)."""
    import clockpunch_b200.common.transforms as tr
    import clockpunch_b200.environment.env as env
    import clockpunch_b200.common.agents  # noqa: F401  (binds lla2ecef by name)
    import clockpunch_b200.simulation.sim_utils  # noqa: F401  (binds the batch conversions by name)
    from oracle import dynamics as od

    def frame_change(x, JD, to_eci):
        x = tr._check_dimensions(x)
        return (od.ecef2eci if to_eci else od.eci2ecef)(x, JD).squeeze()

    def coe2eci_batch(coe, mu=tr.mu):
        coe = np.ascontiguousarray(coe, dtype=np.float64)
        tr._check_coe(*coe)
        return np.stack([od.coe2eci(*coe[:, i], mu=mu) for i in range(coe.shape[1])], axis=1)

    def lla2eci_batch(lla, JD=0.0):
        lla = np.asarray(lla, dtype=np.float64)[:3]
        return np.stack([od.ecef2eci(od.lla2ecef(lla[:, i]), JD) for i in range(lla.shape[1])], axis=1)

    # (modules that bound these names with `from ... import` get the doubles as well)
    import sys
    swaps = {tr._frame_change: frame_change, tr.coe2eci_batch: coe2eci_batch, tr.lla2eci_batch: lla2eci_batch,
             tr.lla2ecef: od.lla2ecef}
    for name, mod in list(sys.modules.items()):
        if name.startswith("clockpunch_b200") and mod is not None:
            for attr, val in list(vars(mod).items()):
                if callable(val) and val in swaps:
                    setattr(mod, attr, swaps[val])
    env.CatalogEngine = OracleEngine
    # every other mirrored function reaches the library through `_lib.get_ctx().lib.pc_*_host(...)` with host
    # buffers: those entry points are served by tests/_oracle_lib.OracleLib (same names, arguments and status
    # codes), so that the mirrors' own marshalling code runs
    import clockpunch_b200._lib as lib
    from _oracle_lib import OracleContext
    ctx = OracleContext()
    lib.get_ctx = lambda device=None: ctx
    return ctx
