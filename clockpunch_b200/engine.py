"""Device-resident catalog engine: the batched form of SSAScheduler's physics step.

One ``CatalogEngine`` holds M sensors and N targets (truth states, one 6-D UKF per
target) in HBM and advances them with ``pc_step_fused``:

    sensors propagate (K1/K2) -> targets truth + UKF (K4) -> both packed vis maps (K3)

which is the arithmetic of SSAScheduler.step (punchclock/environment/env.py:533-597):
_propagateAgents (:503), _taskAgents (:464), updateInfoPostTasking (:392).  PyTorch
tensors are used only as device buffers (allocation, stream, copies); every kernel is in
libpunch_b200.  Layouts are those of include/punch_b200.h: states (6, ld) SoA, covariances
(N, 6, 6), vis maps bit-packed (N, ceil(M/32)) uint32.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _buffers as B
from . import _lib
from .common.visibility import BINARY_TOL, unpack_vis_words

RE = 6378.1363


def _pad(n: int, mult: int = 16) -> int:
    return max(mult, (n + mult - 1) // mult * mult)


class CatalogEngine:
    def __init__(self, sensor_states, sensor_kinds, target_states, est_x, est_p, Q, R, time: float = 0.0,
                 device: int | None = None, substeps: int = 0, force_model: int = _lib.FORCE_TWOBODY,
                 seed: int = 0, alpha: float = 0.001, beta: float = 2.0, kappa=None,
                 body_radius: float = RE, vis_tol: float = BINARY_TOL,
                 init_num_tasked=None, init_last_time_tasked=None, num_envs: int = 1):
        torch = B.torch_mod()
        self.ctx = _lib.get_ctx(device)
        self.dev = B.device_of(self.ctx)
        self.torch = torch
        sensor_states = np.asarray(sensor_states, dtype=np.float64).reshape(6, -1)
        target_states = np.asarray(target_states, dtype=np.float64).reshape(6, -1)
        self.M, self.N = sensor_states.shape[1], target_states.shape[1]
        self.ld_s, self.ld = _pad(self.M), _pad(self.N)
        # a batch of `num_envs` independent environments stored environment-major: each target is
        # tested against its own environment's sensors only (vectorised SSAScheduler envs)
        self.E = int(num_envs)
        if self.E < 1 or self.M % self.E or self.N % self.E:
            raise ValueError("num_envs must divide both the sensor and the target count")
        self.Me, self.Ne = self.M // self.E, self.N // self.E
        self.wpr = max(1, (self.Me + 31) // 32)
        self.time = float(time)
        self.step_count = 0
        # counter of the measurement / stand-in-policy random streams (Philox and splitmix are keyed by
        # (seed, rng_step, target)).  It keeps running across reset(): the reference draws from numpy's
        # global RNG (agents.py:173-183), which does not rewind between episodes either.
        self.rng_step = 0
        self.substeps, self.force_model, self.seed = int(substeps), int(force_model), int(seed)
        self.body_radius, self.vis_tol = float(body_radius), float(vis_tol)
        self.params = _lib.make_ukf_params(Q, R, alpha, beta, kappa)
        f64, dev = torch.float64, self.dev

        def state_buf(a, ld):
            t = torch.zeros((6, ld), dtype=f64, device=dev)
            t[:, : a.shape[1]] = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
            return t

        self.sens_x = state_buf(sensor_states, self.ld_s)
        self.sens_kind = torch.from_numpy(np.ascontiguousarray(sensor_kinds, dtype=np.uint8)).to(dev)
        # sigma buffer: 14 blocks of (6, ld); est_x IS block 0 and the truth state IS block 13
        self.sigma = torch.zeros((14, 6, self.ld), dtype=f64, device=dev)
        self.est_x, self.truth_x = self.sigma[0], self.sigma[13]
        self.truth_x.copy_(state_buf(target_states, self.ld))
        self.est_x.copy_(state_buf(np.asarray(est_x, dtype=np.float64).reshape(6, -1), self.ld))
        self.sig_flags = torch.zeros((2, self.N), dtype=torch.int32, device=dev)
        self.sens_q = torch.zeros((self.ld_s, 4), dtype=f64, device=dev)   # double4 per sensor
        self.task_of = torch.full((max(self.ld_s, 1),), -1, dtype=torch.int64, device=dev)   # pc_step_args.task_of
        self._sigma_valid = False   # blocks 1..12 match (est_x, est_p)?
        self.est_p = torch.from_numpy(np.ascontiguousarray(est_p, dtype=np.float64).reshape(self.N, 36)).to(dev)
        self.pred_p = self.est_p.clone()
        self.nis = torch.full((self.N,), float("nan"), dtype=f64, device=dev)
        self.flags = torch.zeros((self.N,), dtype=torch.int32, device=dev)
        self.vis_truth = torch.zeros((self.N, self.wpr), dtype=torch.int32, device=dev)
        # what every other shard's scheduler needs from this one each step lives in
        # ONE buffer, so that the per-step exchange is a single all-gather: num_tasked (int64) |
        # packed estimated vis map (int32 words) | tr_pos | tr_vel | last_time_tasked (float32)
        # ... and that buffer sits right behind the float32 observation in ONE output block, so that
        # the per-step device-to-host transfer of step_host is a single copy
        A = self.M + self.N
        self._obs_sizes = (6 * A, 36 * self.N, self.N)
        self._obs_bytes = (4 * sum(self._obs_sizes) + 255) // 256 * 256
        self._sum_bytes = self.N * (8 + 4 * self.wpr + 12)
        self.out_block = torch.zeros((self._obs_bytes + self._sum_bytes,), dtype=torch.uint8, device=dev)
        self.obs_f32 = self.out_block[: 4 * sum(self._obs_sizes)].view(torch.float32)
        self.summary = self.out_block[self._obs_bytes:]
        self._bind_summary_views()
        self.tasked = torch.zeros((self.N,), dtype=torch.uint8, device=dev)
        self.z = torch.zeros((6, self.ld), dtype=f64, device=dev)
        nt = np.zeros(self.N, np.int64) if init_num_tasked is None else np.asarray(init_num_tasked, np.int64)
        lt = np.zeros(self.N, np.float32) if init_last_time_tasked is None else np.asarray(init_last_time_tasked, np.float32)
        self.num_tasked.copy_(torch.from_numpy(nt.copy()))
        self.last_time_tasked.copy_(torch.from_numpy(lt.copy()))
        self.actions = torch.zeros((max(1, self.M),), dtype=torch.int64, device=dev)
        self._args = _lib.StepArgs()
        self._snapshot = None
        # device-resident step clock (pc_clock: t_now, step, t_cur, step_cur) for graph replay
        self.clock = torch.zeros((32,), dtype=torch.uint8, device=dev)
        self._graph = None
        self._hs_dt = None
        self._hs_key = None
        self._peer = None
        self._hs_io = _lib.StepIO()
        # work enqueued on torch's stream since the last step_host (which runs on the library's own
        # stream when torch's is the legacy default stream, see _order_streams)
        self._torch_stream_dirty = True
        self._tasked_dirty = False     # self.tasked holds a caller's mask (steps with actions need it zeroed)
        self._tasked_all = False       # ... the all-ones mask of an all_tasked graph
        self._sync_clock()
        self.compute_vis()
        self.snapshot()

    def _bind_summary_views(self):
        torch, N, w = self.torch, self.N, self.wpr
        b, o = self.summary, 0
        self.num_tasked = b[o: o + 8 * N].view(torch.int64); o += 8 * N
        self.vis_est = b[o: o + 4 * w * N].view(torch.int32).view(N, w); o += 4 * w * N
        self.tr_pos = b[o: o + 4 * N].view(torch.float32); o += 4 * N
        self.tr_vel = b[o: o + 4 * N].view(torch.float32); o += 4 * N
        self.last_time_tasked = b[o: o + 4 * N].view(torch.float32)

    @staticmethod
    def summary_views(block, N: int, wpr: int) -> dict:
        """Field views of one shard's summary block (as received from the all-gather)."""
        import torch
        o, out = 0, {}
        out["num_tasked"] = block[o: o + 8 * N].view(torch.int64); o += 8 * N
        out["vis_est"] = block[o: o + 4 * wpr * N].view(torch.int32).view(N, wpr); o += 4 * wpr * N
        for k in ("tr_pos", "tr_vel", "last_time_tasked"):
            out[k] = block[o: o + 4 * N].view(torch.float32); o += 4 * N
        return out

    # -- helpers -----------------------------------------------------------
    def _stream(self):
        return self.torch.cuda.current_stream(self.dev).cuda_stream

    def _sync_clock(self):
        rec = np.zeros(1, dtype=[("t", "<f8"), ("s", "<u8"), ("tc", "<f8"), ("sc", "<u8")])
        rec["t"], rec["s"] = self.time, self.rng_step
        self.clock.copy_(self.torch.from_numpy(rec.view(np.uint8).copy()))
        self._torch_stream_dirty = True

    def _zero_tasked(self):
        self.tasked.zero_()
        self._tasked_dirty, self._tasked_all = False, False
        self._torch_stream_dirty = True

    def _order_streams(self):
        """pc_step_host captures and replays its graph on the library's own (non-blocking) stream when
        torch's current stream is the legacy default stream, which cannot be captured; that stream is
        not ordered against torch's.  Everything else the engine does (reset, step, sigma generation,
        clock writes, observation packs) is enqueued on torch's stream, so the first step_host after any
        of it waits for that work (a host-side wait, paid only on such transitions); step_host itself
        returns only when its stream is idle, so nothing is needed in the other direction."""
        if self._torch_stream_dirty:
            self.torch.cuda.current_stream(self.dev).synchronize()
            self._torch_stream_dirty = False

    def _fill_args(self, t0, tf, z, tasked, want_vis=True, use_clock=False, actions=None, policy_probes=0,
                   policy_seed=0):
        a = self._args
        self._hs_key = None          # the cached step_host call points at this block
        a.sens_x, a.sens_kind, a.M, a.ld_s = self.sens_x.data_ptr(), self.sens_kind.data_ptr(), self.M, self.ld_s
        a.truth_x, a.est_x, a.est_p = self.truth_x.data_ptr(), self.est_x.data_ptr(), self.est_p.data_ptr()
        a.z = None if z is None else z.data_ptr()
        a.tasked = None if tasked is None else tasked.data_ptr()
        a.N, a.ld, a.t0, a.tf = self.N, self.ld, float(t0), float(tf)
        a.substeps, a.force_model = self.substeps, self.force_model
        a.rng_seed, a.rng_step = self.seed, self.rng_step
        a.pred_p, a.nis, a.flags = self.pred_p.data_ptr(), self.nis.data_ptr(), self.flags.data_ptr()
        a.vis_truth = self.vis_truth.data_ptr() if want_vis else None
        a.vis_est = self.vis_est.data_ptr() if want_vis else None
        a.words_per_row, a.body_radius, a.vis_tol = self.wpr, self.body_radius, self.vis_tol
        a.num_tasked, a.last_time_tasked = self.num_tasked.data_ptr(), self.last_time_tasked.data_ptr()
        a.tr_pos, a.tr_vel = self.tr_pos.data_ptr(), self.tr_vel.data_ptr()
        a.clock = self.clock.data_ptr() if use_clock else None
        a.actions = None if actions is None else actions.data_ptr()
        a.policy_probes, a.policy_seed = int(policy_probes), int(policy_seed)
        a.sigma_ws, a.sig_flags = self.sigma.data_ptr(), self.sig_flags.data_ptr()
        a.sens_q = self.sens_q.data_ptr()
        a.task_of = self.task_of.data_ptr()
        a.env_targets, a.env_sensors = (self.Ne, self.Me) if self.E > 1 else (0, 0)
        if self._peer is not None:
            a.peer_table, a.peer_world, a.peer_rank = self._peer["table"], self._peer["world"], self._peer["rank"]
            a.summary_base, a.summary_bytes = self.summary.data_ptr(), self.summary.numel()
        else:
            a.peer_table, a.peer_world, a.peer_rank, a.summary_base, a.summary_bytes = None, 0, 0, None, 0
        a.sigma_valid = 1 if self._sigma_valid else 0
        return a

    # -- physics -------------------------------------------------------------
    def compute_vis(self):
        """Both packed maps from the current states (reset path, env.py:212)."""
        c = self.ctx
        self._torch_stream_dirty = True
        if self.E > 1:
            c.check(c.lib.pc_vismap_packed_envs(c.handle, self.sens_x.data_ptr(), self.ld_s, self.truth_x.data_ptr(),
                                                self.est_x.data_ptr(), self.ld, self.E, self.Me, self.Ne,
                                                self.body_radius, self.vis_tol, self.vis_truth.data_ptr(),
                                                self.vis_est.data_ptr(), self._stream()))
            return
        c.check(c.lib.pc_vismap_packed(c.handle, self.sens_x.data_ptr(), self.M, self.ld_s, None, 0.0,
                                       self.truth_x.data_ptr(), self.est_x.data_ptr(), self.N, self.ld,
                                       self.body_radius, self.vis_tol, self.vis_truth.data_ptr(),
                                       self.vis_est.data_ptr(), self.wpr, self._stream()))

    def task_from_actions(self, actions):
        """(M,) MultiDiscrete actions (host or device) -> device tasking mask, masked by the
        estimated visibility of the previous step (env.py:564-571)."""
        if isinstance(actions, np.ndarray) or not hasattr(actions, "data_ptr"):
            self.actions[: self.M].copy_(self.torch.from_numpy(np.ascontiguousarray(actions, dtype=np.int64)),
                                         non_blocking=True)
            actions = self.actions
        c = self.ctx
        self._torch_stream_dirty, self._tasked_dirty = True, True
        c.check(c.lib.pc_task_from_actions(c.handle, actions.data_ptr(), self.M, self.vis_est.data_ptr(),
                                           self.N, self.wpr, self.tasked.data_ptr(), self._stream()))
        return self.tasked

    def step(self, t_next: float | None = None, tasked=None, z=None, dt: float | None = None, actions=None):
        """Advance every agent to t_next.  tasked: (N,) uint8 device tensor / numpy bool or None;
        actions: (M,) MultiDiscrete actions (numpy or device int64) -- then the step derives the
        tasking mask itself from the previous estimated vis map (env.py:564-571);
        z: (6,N) measurements (numpy or (6,ld) device tensor) or None -> drawn on device."""
        if t_next is None:
            t_next = self.time + float(dt)
        if actions is not None:
            if not hasattr(actions, "data_ptr"):
                self.actions[: self.M].copy_(self.torch.from_numpy(np.ascontiguousarray(actions, dtype=np.int64)))
                actions = self.actions
            self._zero_tasked()
            tasked = self.tasked
        elif tasked is not None:
            if not hasattr(tasked, "data_ptr"):
                self.tasked.copy_(self.torch.from_numpy(np.ascontiguousarray(tasked, dtype=np.uint8)))
                tasked = self.tasked
            # a caller's mask is left in place by the step: steps that build the mask from actions must
            # start from zeros (their update warps own exactly the targets pc_step_args.task_of names)
            self._tasked_dirty = self._tasked_dirty or tasked is self.tasked
        self._torch_stream_dirty = True
        if z is not None and not hasattr(z, "data_ptr"):
            self.z[:, : self.N].copy_(self.torch.from_numpy(np.ascontiguousarray(z, dtype=np.float64)))
            z = self.z
        a = self._fill_args(self.time, t_next, z, tasked, actions=actions)
        c = self.ctx
        c.check(c.lib.pc_step_fused(c.handle, C.byref(a), C.byref(self.params), self._stream()))
        self._sigma_valid = True
        self.time = float(t_next)
        self.step_count += 1
        self.rng_step += 1
        if self._graph is not None or self._hs_dt is not None:
            self._sync_clock()

    # -- CUDA-graph replay of the whole step ------------------------------------------
    def build_graph(self, dt: float, policy_probes: int = 0, policy_seed: int = 0, pack_obs: bool = False,
                    post=None, all_tasked: bool = False):
        """Capture pc_step_fused (tasking from self.actions -- or the stand-in policy when
        policy_probes > 0 -- masked by the previous estimated vis map, sensors, truth + UKF, both
        vis maps) [+ obs pack] into one CUDA graph.  Times and RNG counters come from the device
        clock, so replay() needs no argument updates.  Measurements are drawn on device.
        all_tasked: every target receives a measurement every step (BASELINE.json configs[4]: a fixed
        all-ones tasking mask instead of actions)."""
        self._zero_tasked()
        torch, c = self.torch, self.ctx
        if all_tasked:
            self.tasked.fill_(1)
            self._tasked_all = True
        self._sync_clock()
        self.ensure_sigma()
        before = c.launch_count
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                a = self._fill_args(0.0, float(dt), None, self.tasked, use_clock=True,
                                    actions=None if all_tasked else self.actions,
                                    policy_probes=0 if all_tasked else policy_probes, policy_seed=policy_seed)
                c.check(c.lib.pc_step_fused(c.handle, C.byref(a), C.byref(self.params), self._stream()))
                if pack_obs:
                    self.pack_obs(use_clock=True)
                if post is not None:
                    post()          # e.g. the shard all-gather, captured into the same graph
        torch.cuda.current_stream(self.dev).wait_stream(side)
        g.pc_dt, g.pc_launches, g.pc_all_tasked = float(dt), c.launch_count - before, bool(all_tasked)
        self._graph = g
        return g

    def step_policy_eager(self, dt: float, policy_probes: int, policy_seed: int = 0):
        """The step of build_graph(policy_probes > 0) launched kernel by kernel instead of replayed
        (device clock, stand-in policy, measurements drawn on device).  Used by bench.py's roofline
        pass: plain cudaEventRecord calls around one kernel measure it without the latency of the
        event-record NODES a captured graph needs."""
        self.ensure_sigma()
        if self._tasked_dirty or self._tasked_all:
            self._zero_tasked()
        if self._hs_dt is None and self._graph is None:
            self._sync_clock()
            self._hs_dt = float(dt)
        a = self._fill_args(0.0, float(dt), None, self.tasked, use_clock=True, actions=self.actions,
                            policy_probes=policy_probes, policy_seed=policy_seed)
        c = self.ctx
        self._torch_stream_dirty = True
        c.check(c.lib.pc_step_fused(c.handle, C.byref(a), C.byref(self.params), self._stream()))
        self.time += float(dt)
        self.step_count += 1
        self.rng_step += 1

    # -- multi-GPU: summaries pushed into every peer's gathered buffer by the step itself ----------
    def enable_peer_push(self, group=None):
        """Set up the peer exchange of the step (pc_comm_init / pc_comm_attach: every rank allocates its
        gathered buffer, the 64-byte CUDA IPC handles are all-gathered through `group`, every rank maps
        every peer's buffer over NVLink) and make every following step push its per-target summaries
        straight into all of them (from the next step's sensor kernel, beside the propagation), replacing
        the per-step all-gather.  All ranks must call this
        together, with equal shard sizes.  torch.distributed is only the transport of the handles;
        a non-torch caller does the same three C calls with its own transport.
        Returns the local gathered tensor, shape (2, world, summary_bytes) uint8: [step parity][source rank]."""
        import weakref
        import torch.distributed as dist
        torch, c = self.torch, self.ctx
        group = group or dist.group.WORLD
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        nbytes = self.summary.numel()
        # a library context owns ONE exchange: an engine that had it before loses it (its buffers are freed below)
        prev = getattr(c, "_peer_engine", None)
        prev = prev() if prev is not None else None
        if prev is not None and prev is not self:
            prev._peer, prev._hs_key, prev._graph = None, None, None
        c._peer_engine = weakref.ref(self)
        mine = np.zeros(_lib.COMM_HANDLE_BYTES, np.uint8)
        # an exchange left on this context by an earlier engine: every rank unmaps its peers before
        # anybody frees (pc_comm_init frees the old buffer)
        torch.cuda.synchronize(self.dev)
        c.check(c.lib.pc_comm_detach(c.handle))
        dist.barrier(group)
        c.check(c.lib.pc_comm_init(c.handle, world, rank, nbytes, mine.ctypes.data))
        on_gpu = dist.get_backend(group) == "nccl"
        send = torch.from_numpy(mine).to(self.dev) if on_gpu else torch.from_numpy(mine)
        recv = torch.empty((world, _lib.COMM_HANDLE_BYTES), dtype=torch.uint8, device=send.device)
        dist.all_gather_into_tensor(recv, send, group=group)
        handles = np.ascontiguousarray(recv.cpu().numpy())
        c.check(c.lib.pc_comm_attach(c.handle, handles.ctypes.data))
        table, gathered = C.c_void_p(), C.c_void_p()
        c.check(c.lib.pc_comm_table(c.handle, C.byref(table), C.byref(gathered), None, None, None))
        view = B.device_view(gathered.value, 2 * world * nbytes, self.dev).view(2, world, nbytes)
        torch.cuda.synchronize(self.dev)
        dist.barrier(group)              # everybody attached before anybody's first push
        self._peer = {"table": table.value, "world": world, "rank": rank, "gathered": view}
        self._hs_key = None
        return view

    def peer_sync(self):
        """Exchange the CURRENT summaries now (pc_peer_exchange: push + wait as two kernels on torch's
        stream): what follows on the stream (or the host, after a synchronise) may read `gathered`
        [push count & 1].  Steps do not end with an exchange -- the NEXT step's sensor kernel pushes this
        step's summaries and waits for the peers', beside its propagation -- so a scheduler that needs
        every shard's summaries between two steps calls this.  All ranks must call it together."""
        p, c = self._peer, self.ctx
        self._torch_stream_dirty = True
        c.check(c.lib.pc_peer_exchange(c.handle, p["table"], p["world"], p["rank"], self.summary.data_ptr(),
                                       self.summary.numel(), self._stream()))

    def peer_fields(self, parity: int, src_rank: int) -> dict:
        """Typed views of what `src_rank` pushed for steps of the given parity."""
        return self.summary_views(self._peer["gathered"][parity, src_rank], self.N, self.wpr)

    def peer_status(self) -> tuple[int, int]:
        """(error mask: bit r = peer r's flag never arrived within the bounded wait; pushes done so
        far -- its parity is the slot of `gathered` holding the latest summaries).  Exchanges the current
        summaries first (peer_sync) and synchronises."""
        err, n = C.c_uint32(0), C.c_uint64(0)
        self.peer_sync()
        self.torch.cuda.synchronize(self.dev)
        self.ctx.check(self.ctx.lib.pc_peer_status(self.ctx.handle, C.byref(err), C.byref(n)))
        return err.value, n.value

    def ensure_sigma(self):
        """(Re)generate sigma-point blocks 1..12 from the current (est_x, est_p) if stale."""
        if not self._sigma_valid:
            c = self.ctx
            self._torch_stream_dirty = True
            c.check(c.lib.pc_sigma_points(c.handle, self.est_x.data_ptr(), self.est_p.data_ptr(), self.N,
                                          self.ld, self.params.gamma, self.sigma.data_ptr(),
                                          self.sig_flags.data_ptr(), self._stream()))
            self._sigma_valid = True

    def invalidate_sigma(self):
        """Call after writing est_x / est_p from outside the step kernels."""
        self._sigma_valid = False

    def replay(self, graph=None):
        """One step through a captured graph (actions must already be in self.actions unless the
        graph contains the stand-in policy)."""
        g = graph or self._graph
        self.ensure_sigma()
        if g.pc_all_tasked:
            if not self._tasked_all:              # a reset zeroed the workspace
                self.tasked.fill_(1)
                self._tasked_all, self._tasked_dirty = True, True
        elif self._tasked_dirty or self._tasked_all:
            self._zero_tasked()
        self._torch_stream_dirty = True
        g.replay()
        self.time += g.pc_dt
        self.step_count += 1
        self.rng_step += 1

    # -- the reference-facing step: host actions in, host observation out ------------------
    def host_buffers(self, pinned: bool = True):
        """Host-side (pinned) buffers for step_host: actions (M,) int64 in; obs float32
        [eci_state (6, M+N) | est_cov (N, 6, 6) | obs_staleness (N)], packed estimated vis map
        (N, wpr) and num_tasked (N,) out."""
        torch = self.torch
        mk = (lambda *a, **k: torch.empty(*a, **k).pin_memory()) if pinned else torch.empty
        block = mk((self.out_block.numel(),), dtype=torch.uint8)        # mirror of self.out_block
        v = self.summary_views(block[self._obs_bytes:], self.N, self.wpr)
        return {"actions": mk((max(1, self.M),), dtype=torch.int64), "block": block,
                "obs": block[: 4 * sum(self._obs_sizes)].view(torch.float32),
                "vis_est": v["vis_est"], "num_tasked": v["num_tasked"], "tr_pos": v["tr_pos"],
                "tr_vel": v["tr_vel"], "last_time_tasked": v["last_time_tasked"]}

    def step_host(self, hb: dict, dt: float, policy_probes: int = 0, policy_seed: int = 0, zero_copy: bool = True):
        """One SSAScheduler.step for the whole catalog through pc_step_host: copies hb["actions"]
        to the device, runs the step (a CUDA graph inside the library after the first call),
        packs the float32 observation and copies it, the packed estimated vis map and num_tasked
        into hb.  One C call per step; returns when the host buffers are filled."""
        key = (float(dt), int(policy_probes), int(policy_seed), id(hb), self._sigma_valid, bool(zero_copy))
        if self._hs_key != key:
            # (re)build the argument blocks once; steady-state steps below are one C call
            self.ensure_sigma()
            if self._hs_dt != dt:
                self._sync_clock()
                self._hs_dt = dt
            self._fill_args(0.0, float(dt), None, self.tasked, use_clock=True, actions=self.actions,
                            policy_probes=policy_probes, policy_seed=policy_seed)
            io = self._hs_io
            io.h_actions = hb["actions"].data_ptr() if policy_probes == 0 else None
            io.d_obs, io.h_obs = self.obs_f32.data_ptr(), hb["obs"].data_ptr()
            io.h_vis_est, io.h_num_tasked = hb["vis_est"].data_ptr(), hb["num_tasked"].data_ptr()
            if "block" in hb:      # obs | summary in one device block -> one D2H copy
                io.d_block, io.h_block, io.block_bytes = (self.out_block.data_ptr(), hb["block"].data_ptr(),
                                                          self.out_block.numel())
                # pinned block: the kernels store the outputs into it themselves while they run (no copy);
                # the library falls back to the single copy when the catalog does not qualify.  The device
                # observation (self.obs_f32) is not refreshed on that path: pack_obs() does that.
                io.zero_copy = 1 if (zero_copy and hb["block"].is_pinned()) else 0
            self._hs_call = (self.ctx.lib.pc_step_host, self.ctx.handle, C.byref(self._args), C.byref(self.params),
                             C.byref(io), self._stream())
            self._hs_key = (float(dt), int(policy_probes), int(policy_seed), id(hb), True, bool(zero_copy))
        if self._tasked_dirty or self._tasked_all:
            self._zero_tasked()
        if self._torch_stream_dirty:
            self._order_streams()
        f, *cargs = self._hs_call
        rc = f(*cargs)
        if rc:
            self.ctx.check(rc)
        self.time += float(dt)
        self.step_count += 1
        self.rng_step += 1

    def pack_obs(self, use_clock: bool = False):
        """float32 eci_state / est_cov / obs_staleness into self.obs_f32 (device)."""
        c = self.ctx
        self._torch_stream_dirty = True
        n0, n1, _ = self._obs_sizes
        base = self.obs_f32.data_ptr()
        if self.E > 1:
            c.check(c.lib.pc_obs_pack_envs(c.handle, self.sens_x.data_ptr(), self.ld_s, self.est_x.data_ptr(),
                                           self.est_p.data_ptr(), self.ld, self.last_time_tasked.data_ptr(),
                                           self.time, self.clock.data_ptr() if use_clock else None, self.E,
                                           self.Me, self.Ne, base, base + 4 * n0, base + 4 * (n0 + n1),
                                           self._stream()))
            return self.obs_f32
        c.check(c.lib.pc_obs_pack(c.handle, self.sens_x.data_ptr(), self.M, self.ld_s, self.est_x.data_ptr(),
                                  self.est_p.data_ptr(), self.N, self.ld, self.last_time_tasked.data_ptr(),
                                  self.time, self.clock.data_ptr() if use_clock else None,
                                  base, base + 4 * n0, base + 4 * (n0 + n1), self._stream()))
        return self.obs_f32

    # -- reset snapshot (env.py:142,208: deepcopy of the agent list) ---------------
    _STATE = ("sens_x", "sigma", "sig_flags", "est_p", "pred_p", "nis", "flags", "vis_truth", "summary")

    def snapshot(self):
        self._snapshot = ({k: getattr(self, k).clone() for k in self._STATE}, self.time, self.step_count,
                          self._sigma_valid)
        self._snapshot_rng_step = self.rng_step

    def reset(self, seed: int | None = None, rewind_rng: bool = False):
        """Back to the snapshot (env.py:208).  The random streams of the measurement noise and of the
        stand-in policy keep running -- a new episode sees new noise, as with the reference's global numpy
        RNG -- unless a new `seed` is given (the streams then restart under it) or `rewind_rng` asks for
        a bit-identical replay of the snapshot's episode."""
        if seed is not None and int(seed) != self.seed and self._graph is not None:
            raise ValueError("the seed is part of the captured step: drop the graph (build_graph again) to reseed")
        bufs, t, sc, sv = self._snapshot
        for k, v in bufs.items():
            getattr(self, k).copy_(v)
        self.time, self.step_count, self._sigma_valid = t, sc, sv
        if seed is not None:
            self.seed, self.rng_step = int(seed), 0
            self._hs_key = None          # the seed is part of the cached argument block
        elif rewind_rng:
            self.rng_step = self._snapshot_rng_step
        self._zero_tasked()
        self._sync_clock()

    def clone(self) -> "CatalogEngine":
        """Independent copy: device buffers cloned, library context shared (deepcopy of an env)."""
        torch = self.torch
        new = object.__new__(CatalogEngine)
        for k, v in self.__dict__.items():
            if isinstance(v, torch.Tensor):
                setattr(new, k, v.clone())
            elif k in ("_args", "_hs_io", "_graph", "_snapshot", "est_x", "truth_x"):
                continue
            else:
                setattr(new, k, v)
        new.est_x, new.truth_x = new.sigma[0], new.sigma[13]      # views into the cloned sigma buffer
        new.obs_f32 = new.out_block[: 4 * sum(new._obs_sizes)].view(torch.float32)
        new.summary = new.out_block[new._obs_bytes:]
        new._bind_summary_views()
        new._args, new._hs_io = _lib.StepArgs(), _lib.StepIO()
        new._graph, new._hs_dt, new._hs_key, new._peer = None, None, None, None
        bufs, t, sc, sv = self._snapshot
        new._snapshot = ({k: v.clone() for k, v in bufs.items()}, t, sc, sv)
        new._sync_clock()
        return new

    def __deepcopy__(self, memo):
        return self.clone()

    # -- host views --------------------------------------------------------------
    def host(self, name: str) -> np.ndarray:
        t = getattr(self, name)
        if name in ("sens_x",):
            return t[:, : self.M].cpu().numpy()
        if name in ("truth_x", "est_x", "z"):
            return t[:, : self.N].cpu().numpy()
        if name in ("est_p", "pred_p"):
            return t.cpu().numpy().reshape(self.N, 6, 6)
        if name in ("vis_truth", "vis_est"):
            return unpack_vis_words(t.cpu().numpy().view(np.uint32), self.Me)      # (N, Me): own env's sensors
        return t.cpu().numpy()
