#!/usr/bin/env python
"""bench.py -- target-steps/s of the punchclock hot path (propagate + vis + UKF).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): 100 ground sensors x 10 000 mixed LEO/MEO/GEO targets,
two-body, dt = 100 s, reference-default filter matrices, ~M targets tasked per step by a
random-visible-target policy.  One "step" = one SSAScheduler physics step over the whole
catalog: every sensor propagated, every target's truth propagated, every target's UKF
predictAndUpdate (13 sigma-point propagations), both (N, M) visibility maps.  One
target-step = that work for one target.

Own arm: catalog resident in HBM (`value`), and end to end through CatalogEngine with host
actions in / host observation out every step (`e2e`).  At N > 1 every rank owns its own
catalog of the same size (weak scaling, targets partitioned, sensors replicated) and the
ranks all-gather the packed estimated vis map + per-target summaries each step.

Reference arm (--impl reference): the CPU oracle port of the reference algorithm
(oracle/, one scipy DOP853 solve per state, Python loops per pair -- the reference's own
structure) on all host cores, each step a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 20240625 + 2
N_TARGETS, N_SENSORS, DT = 10_000, 100, 100.0
EPISODE = 50        # steps per episode: the catalog is reset to its initial state every EPISODE steps (untimed), as
                    # SSAScheduler does at its horizon -- untasked reference-default filters drift, so a
                    # longer run would no longer be measuring the stated workload
METRIC = "target-steps/sec (propagate+vis+UKF)"
UNIT = "target-steps/s"
# Algorithmic work per unit, SURVEY.md section 8(d) (FMA = 2, add/mul/sqrt/div = 1):
FLOP_DOP853_STEP = 1150.0          # one fixed DOP853 step: 12 RHS + stage sums + final combination
FLOP_UKF_NOMEAS, FLOP_UKF_MEAS = 1500.0, 5600.0
FLOP_VIS_PAIR = 8.0
FLOP_NEXT_CHOL = 110.0             # Cholesky of the new est_p (next step's sigma points)
FP64_INSTR_DOP853_STEP = 435.0     # FP64 instructions the kernel executes per DOP853 step (320 DFMA + 109 DMUL + 6 DADD;
                                   # tools/sass_mix.py on the built library)
BYTES_TARGET_STEP = 1130.0         # truth/est_x/est_p r+w, pred_p w, z r (+ M/4 vis bits)
BYTES_FINISH_TARGET = 1824.0       # per-target kernel: 13 states read (624), est_p + pred_p written (576), next sigma
                                   # points written (624) (+ M/4 vis bits)


def workload_config(n_targets, n_sensors):
    which = ("BASELINE.json configs[1]" if (n_targets, n_sensors) == (N_TARGETS, N_SENSORS) else
             "BASELINE.json configs[3] shard (1 M targets over 8 GPUs)" if (n_targets, n_sensors) == (125_000, 100) else
             "other size of the configs[1] workload")
    return {"workload": f"{n_sensors} ground sensors x {n_targets} mixed LEO/MEO/GEO targets, two-body, "
                        f"dt={DT:.0f}s, UKF per target ({which})",
            "targets_per_gpu": n_targets, "sensors": n_sensors, "dt_s": DT,
            "tasking": "each sensor tasks a random estimated-visible target (<= M per step)",
            "episode": f"catalog reset to its initial state every {EPISODE} steps (outside the timed events)",
            "l2": "256 MiB flush between timed steps (state fits the 126 MB L2)",
            "physics": "the reference's: two-body, its default ground sensors (dynamics.py:16-40; the reference has no J2 "
                       "and this line is compared with the reference's CPU path); BASELINE.json's wording of configs[1] "
                       "(ground/space sensors, two-body + J2) is timed beside it as extra.configs.c2_j2_space_sensors"}


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and clock-event (throttle) reasons sampled DURING the timed region.  The timed
    region of this workload is milliseconds long, so the sampler polls NVML every 2 ms from a thread
    (nvidia-smi -lms cannot go below 100 ms); nvidia-smi is the fallback when NVML is unavailable."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.thread, self.run = index, [], None, None, False
        self.sm, self.mx, self.reasons, self.power = [], [], set(), []

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap"}
            self.run = True

            def poll():
                while self.run:
                    try:
                        self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                        self.mx.append(float(mx))
                        r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                        self.reasons.update(n for bit, n in names.items() if r & bit)
                    except nv.NVMLError:
                        pass
                    time.sleep(0.002)
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.run = False
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.thread is not None:
            self.run = False
            self.thread.join(timeout=1.0)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None,
                    "sm_max_mhz": max(self.mx) if self.mx else None, "reasons": sorted(self.reasons),
                    "samples": len(self.sm), "source": "NVML polled every 2 ms during the timed passes"}
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for name, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


# ----------------------------------------------------------------------------- CPU reference arm
def _cpu_chunk(args):
    """Oracle target-steps for a slice of targets (runs in a worker process)."""
    from oracle.step import target_step
    truth, est_x, est_p, Q, R, z, tasked, sensors, t0, tf = args
    out = 0.0
    for i in range(truth.shape[1]):
        r = target_step(truth[:, i], est_x[:, i], est_p[i], Q, R, t0, tf, z[:, i] if tasked[i] else None, sensors)
        out += float(r[0][0])
    return out


def cpu_reference_run(cat, sample, steps, warmup, cores):
    """Times `steps` passes over a `sample`-target slice of the workload on `cores` processes."""
    import multiprocessing as mp
    rng = np.random.default_rng(SEED)
    sample = min(sample, cat.N)
    tasked = np.zeros(sample, bool)
    tasked[rng.choice(sample, size=max(1, round(sample * cat.M / cat.N)), replace=False)] = True
    z = cat.targets[:, :sample] + rng.normal(size=(6, sample)) * np.sqrt(np.diag(cat.R))[:, None]
    bounds = np.linspace(0, sample, cores + 1).astype(int)
    chunks = [(cat.targets[:, a:b], cat.est_x[:, a:b], cat.est_p[a:b], cat.Q, cat.R, z[:, a:b], tasked[a:b],
               cat.sensors, 0.0, DT) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for _ in range(warmup):
            pool.map(_cpu_chunk, chunks)
        t0 = time.perf_counter()
        for _ in range(steps):
            pool.map(_cpu_chunk, chunks)
        el = time.perf_counter() - t0
    return sample * steps / el, el / steps, sample


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from clockpunch_b200.simulation.catalog import synth_catalog
    cat = synth_catalog(args.targets, args.sensors, seed=SEED)
    cores = host_cores()
    # bounded sample per step: the whole --steps/--warmup run is sized to ~2 minutes of CPU work
    # (the oracle port does ~40 target-steps/s per core)
    per_core = args.cpu_sample_per_core or int(np.clip(120.0 / max(1, args.steps + args.warmup) * 40.0, 4, 640))
    sample = min(cat.N, max(cores, per_core * cores))
    value, s_per_step, sample = cpu_reference_run(cat, sample, args.steps, args.warmup, cores)
    desc = (f"{sample} of {cat.N} targets x {cat.M} sensors per step, oracle port (scipy DOP853 per state, "
            f"python loop per pair), {cores} processes")
    line = {"metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": s_per_step * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.targets, args.sensors),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- own arm
def kernel_sources_sha256():
    """Content hash of the CUDA sources: profile summaries under profiles/ are only trusted for the exact
    kernels they were captured from."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "clockpunch_b200", "csrc")
    for f in sorted(os.listdir(d)):
        # device code and its launchers: kernels_*.cu and the headers; api.cu is host-only glue (argument checks,
        # graph cache, host policy) and holds neither a kernel nor a launch
        if f.endswith(".cuh") or (f.startswith("kernels_") and f.endswith(".cu")):
            h.update(f.encode())
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()


def load_profile_summary(n_targets, n_sensors):
    """profiles/r2_kernel_summary_<N>.json (tools/kernel_profile_summary.py: per-kernel ncu durations and
    DRAM traffic of this workload) -- or None when absent or captured from other kernel sources."""
    path = os.path.join(ROOT, "profiles", f"r2_kernel_summary_{n_targets}.json")
    try:
        d = json.load(open(path))
    except (OSError, ValueError):
        return None
    if d.get("sources_sha256") != kernel_sources_sha256() or d.get("sensors") != n_sensors:
        return None
    d["file"] = os.path.relpath(path, ROOT)
    return d


def pin_rank_to_cores(local, world_local):
    """Give every rank of a multi-rank run its own slice of the host cores, on the NUMA node of its GPU
    when sysfs tells (eight processes sharing one core set made the end-to-end step 2x slower at N = 8)."""
    try:
        cores = sorted(os.sched_getaffinity(0))
    except AttributeError:
        return None
    if world_local <= 1 or len(cores) < 2 * world_local:
        return None
    node_cores = None
    try:
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[local]) if vis and vis.split(",")[local].isdigit() else local
        bus = nv.nvmlDeviceGetPciInfo(nv.nvmlDeviceGetHandleByIndex(phys)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        bus = bus[4:] if len(bus) > 12 else bus
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node >= 0:
            txt = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
            nc = set()
            for part in txt.split(","):
                lo, _, hi = part.partition("-")
                nc.update(range(int(lo), int(hi or lo) + 1))
            node_cores = [c for c in cores if c in nc]
    except Exception:
        node_cores = None
    per = max(1, len(cores) // world_local)
    mine = cores[local * per:(local + 1) * per]
    if node_cores and len(node_cores) >= per:
        k = local % max(1, len(node_cores) // per)
        mine = node_cores[k * per:(k + 1) * per] or mine
    try:
        os.sched_setaffinity(0, mine)
    except OSError:
        return None
    return mine


def run_own(args):
    # stdout carries exactly one JSON line: anything libraries print while we work (NCCL prints its
    # version banner to fd 1) is routed to stderr, and fd 1 is restored for the final print
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line))
        sys.stdout.flush()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    pinned_cores = pin_rank_to_cores(local, int(os.environ.get("LOCAL_WORLD_SIZE", str(world))))

    import torch
    import torch.distributed as dist
    from clockpunch_b200 import _lib
    from clockpunch_b200.engine import CatalogEngine
    from clockpunch_b200.simulation.catalog import synth_catalog
    import ctypes as C

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ctx = _lib.get_ctx(local)
    lib = ctx.lib
    ctx.set_option(_lib.OPT_PROPAGATE_KERNEL, args.opt_propagate)
    ctx.set_option(_lib.OPT_FINISH_KERNEL, args.opt_finish)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    peak_tf = C.c_double()
    ctx.check(lib.pc_fp64_peak(ctx.handle, 5, C.byref(peak_tf)))
    peak_tf = peak_tf.value

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ------------------------------------------------------------------ one workload on this rank
    class Workload:
        """A catalog shard on this rank + its exchange + its step graph."""

        def __init__(self, n, m, dt, seed, all_tasked=False, num_envs=1, mix="mixed", space_sensor_frac=0.0,
                     force_model=_lib.FORCE_TWOBODY):
            self.n, self.m, self.dt, self.all_tasked, self.E = n, m, dt, all_tasked, num_envs
            # every rank owns its own shard of targets; the sensor network is replicated
            self.cat = synth_catalog(n, m, seed=seed + 1000 * rank, sensor_seed=seed, mix=mix,
                                     space_sensor_frac=space_sensor_frac)
            c = self.cat
            self.eng = CatalogEngine(c.sensors, c.sensor_kinds, c.targets, c.est_x, c.est_p, c.Q, c.R, device=local,
                                     seed=seed + rank, num_envs=num_envs, force_model=force_model)
            self.sg, self.exchange, self.use_nccl, self.post = None, "none", False, None
            # per-step summaries every rank needs from every other: packed est vis map,
            # trace(P_pos), trace(P_vel), last_time_tasked, num_tasked.  Default: the NEXT step's sensor kernel copies
            # them into every peer's gathered buffer over NVLink (pc_comm_*: CUDA IPC) and waits for the peers,
            # beside the target propagation; --exchange nccl: one all-gather of the
            # summary buffer captured into the step's graph.  Environments (num_envs > 1) are independent
            # replicas: nothing is exchanged.
            if world > 1 and num_envs == 1 and args.exchange == "none":
                # attribution runs only (profiles/README.md): every rank steps its shard and nobody learns the others'
                # summaries -- the difference to the default run is what the exchange costs
                self.exchange = "NONE (--exchange none: independent shards, for attributing the cost of the exchange)"
            elif world > 1 and num_envs == 1:
                from clockpunch_b200.sharding import PackedGather
                self.sg = PackedGather(world, rank, self.eng.summary, n, self.eng.wpr)
                self.exchange = "nccl all_gather_into_tensor" + ("" if args.eager_gather else " (captured in the step graph)")
                if args.exchange in ("peer", "auto"):
                    try:
                        self.eng.enable_peer_push()
                        self.exchange = ("the NEXT step's sensor kernel copies this step's summary buffer into every peer's "
                                         "gathered buffer (pc_comm_*: CUDA IPC over NVLink) and waits for the peers' flags, "
                                         "beside the target propagation")
                    except Exception as exc:        # no peer access on this box: keep the collective
                        print(f"peer push unavailable ({exc!r}); using NCCL", file=sys.stderr)
                self.use_nccl = self.eng._peer is None
                if self.use_nccl and not args.eager_gather:
                    self.post = self.sg.gather
            self.graph = None

        def build(self):
            if self.sg is not None:
                self.sg.gather()                 # first NCCL call (communicator set-up) outside any capture
                torch.cuda.synchronize(dev)
            self.graph = self.eng.build_graph(self.dt, policy_probes=0 if self.all_tasked else 64, policy_seed=SEED,
                                              post=self.post, all_tasked=self.all_tasked)

        def step(self):
            self.eng.replay(self.graph)
            if self.use_nccl and self.post is None:
                self.sg.gather()

        def verify_exchange(self):
            """The pushed copies once against a plain all-gather of the same buffers (untimed); if any rank
            disagrees, every rank falls back to the NCCL exchange."""
            if self.eng._peer is None:
                return
            ref = self.sg.gather().clone()
            err, pushes = self.eng.peer_status()
            bad = torch.tensor([int(bool(err) or not torch.equal(self.eng._peer["gathered"][pushes & 1], ref))], device=dev)
            dist.all_reduce(bad, op=dist.ReduceOp.MAX)
            if int(bad.item()):
                print(f"rank {rank}: peer-pushed summaries differ from the all-gather (error mask {err:#x}); "
                      "falling back to NCCL", file=sys.stderr)
                self.eng._peer, self.eng._hs_key = None, None
                self.use_nccl = True
                self.post = None if args.eager_gather else self.sg.gather
                self.exchange = "nccl all_gather_into_tensor (fallback: peer push failed its check)"
                self.eng.reset()
                self.build()

        def timed(self, steps, warmup, episode=EPISODE):
            """Device-resident pass: graph replay, CUDA events per step, L2 flushed between steps; max over
            ranks.  The last step's exchange completes inside the timed region (the deferred wait of the
            peer push is enqueued before the final event)."""
            for _ in range(warmup):
                self.step()
            self.verify_exchange()
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            barrier()
            wall0 = time.perf_counter()
            for k in range(steps):
                if k and k % episode == 0:         # a new episode, as the env does at its horizon (untimed)
                    self.eng.reset()
                    self.eng.ensure_sigma()
                if args.l2 == "flush":
                    flush.zero_()                  # L2 flush, outside the timed events
                evs[k][0].record()
                self.step()
                if self.eng._peer is not None and k == steps - 1:
                    self.eng.peer_sync()           # the exchange a following step would have done
                evs[k][1].record()
            barrier()
            wall = time.perf_counter() - wall0
            ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in evs)) / steps
            return ms, wall, self.graph.pc_launches * steps

    def back_to_back(w, steps, episode=EPISODE):
        """The same graph replayed back to back, no L2 flush between steps (what a rollout does; the headline keeps
        the flush the timing rules ask for): one event pair around each episode."""
        tot = 0.0
        for k0 in range(0, steps, episode):
            w.eng.reset()
            w.eng.ensure_sigma()
            for _ in range(3):
                w.step()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            a.record()
            n = min(episode, steps - k0)
            for _ in range(n):
                w.step()
            if w.eng._peer is not None:
                w.eng.peer_sync()
            b.record()
            torch.cuda.synchronize(dev)
            tot += a.elapsed_time(b)
        return max_over_ranks(tot) / steps

    def roofline_kernels(w, steps_done_before, warmup, n_roof):
        """Per-kernel rooflines of one workload (a separate pass): the step's kernels launched one by one (not the
        graph), three CUDA events recorded by the library on the launch stream -- before the sigma-point
        propagation, between it and the per-target kernel, after the per-target kernel.  Inside a captured
        graph the events would be extra event-record NODES whose latency (~4 us each) would be charged to
        kernels of 12-15 us.  Algorithmic work: SURVEY.md section 8(d)."""
        eng, N, M, dt = w.eng, w.n, w.m, w.dt
        eng.reset()
        eng._graph = None
        ctx.check(lib.pc_set_timing(ctx.handle, 1))

        def timed_step():
            if w.all_tasked:
                eng.step(dt=dt, tasked=ones)
            else:
                eng.step_policy_eager(dt, 64, SEED)       # same kernels as the graph, launched one by one
            if w.use_nccl:
                w.sg.gather()

        ones = torch.ones(N, dtype=torch.uint8, device=dev)
        for _ in range(warmup):
            timed_step()
        inv_theta = np.float32(1.0 / 0.12)

        def step_rate(x):
            """The substep rule's rate for a step of dt from state(s) x (csrc/propagate.cuh orbit_rates):
            min(perigee rate with margin, angular rate at the smallest radius reachable within the step)."""
            mu_ = 398600.0
            r, v = x[:3], x[3:]
            hv = np.cross(r.T, v.T)
            h2, r2, v2 = (hv * hv).sum(1), (r * r).sum(0), (v * v).sum(0)
            rn = np.sqrt(r2)
            e1 = 1.0 + np.sqrt(np.maximum(0.0, 1.0 + 2.0 * (0.5 * v2 - mu_ / rn) * h2 / mu_ ** 2))
            wp = mu_ ** 2 * e1 ** 3.5 / h2 ** 1.5
            rmin = rn - np.abs((r * v).sum(0) / rn) * abs(dt) - 0.5 * mu_ / r2 * dt * dt
            wl = np.sqrt(h2) / np.maximum(rmin, 1e-9) ** 2 * np.sqrt(e1)
            return np.where((rmin > h2 / (mu_ * e1)) & (wl < wp), wl, wp)

        truth_sub = np.maximum(1.0, np.ceil(abs(dt) * step_rate(eng.host("truth_x")) / 0.12))

        def substeps_now():
            # the rates the kernels themselves stored with the sigma points (sig_flags row 1: two bfloat16,
            # low half = bound for steps up to 128 s, high half = perigee rate)
            w = eng.sig_flags[1].cpu().numpy().view(np.uint32)
            half = (w & 0xFFFF) if abs(dt) <= 128.0 else (w >> 16)
            om = (half.astype(np.uint32) << 16).view(np.float32)
            est = np.maximum(1.0, np.ceil(np.float32(abs(dt)) * om * inv_theta)).astype(np.float64)
            return 13.0 * est.sum() + truth_sub.sum(), est.mean()

        p_ms, f_ms, state_steps_sum, est_sub_sum = 0.0, 0.0, 0.0, 0.0
        a_ms, b_ms = C.c_double(), C.c_double()
        nt0 = int(eng.num_tasked.sum().item())
        for k in range(n_roof + 2):
            ss, es = substeps_now()
            if args.l2 == "flush":
                flush.zero_()
            timed_step()
            torch.cuda.synchronize(dev)                                   # the events are this step's
            ctx.check(lib.pc_get_kernel_timing(ctx.handle, C.byref(a_ms), C.byref(b_ms)))
            if k >= 2:
                p_ms += a_ms.value
                f_ms += b_ms.value
                state_steps_sum += ss
                est_sub_sum += es
        ctx.check(lib.pc_set_timing(ctx.handle, 0))
        tasked_frac = (int(eng.num_tasked.sum().item()) - nt0) / max(1, N * (n_roof + 2))
        flags_bad = int((eng.flags != 0).sum().item())
        p_ms, f_ms = p_ms / n_roof, f_ms / n_roof
        state_steps = state_steps_sum / n_roof                  # DOP853 steps per launch, summed over 14 N states
        prof = load_profile_summary(N, M)

        def from_profile(family, key):
            return (prof or {}).get("kernels", {}).get(family, {}).get(key)

        # -- sigma-point propagation: FP64 bound
        prop_name = "propagate_sigma_kernel" if N > 16 * 1024 else "propagate_sigma_single_kernel"
        flop_prop = FLOP_DOP853_STEP * state_steps
        ach = flop_prop / (p_ms * 1e-3) * 1e-12
        instr = FP64_INSTR_DOP853_STEP * state_steps
        prop_ncu_ms = from_profile("propagate_sigma", "ms_mean")
        prop = {"kernel": prop_name, "bound": "fp64", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": ach / peak_tf, "kernel_ms": p_ms, "flop_per_launch": flop_prop,
                "dop853_steps_per_launch": state_steps, "mean_substeps_per_sigma_propagate": est_sub_sum / n_roof,
                "algorithmic_flop_per_unit": FLOP_DOP853_STEP, "unit_of_work": "one DOP853 step of one state",
                # what the FP64 pipe actually executes: 435 FP64 instructions per DOP853 step (320 DFMA + 109 DMUL +
                # 6 DADD, tools/sass_mix.py), against the pipe's issue rate = measured DFMA peak / 2 flop
                "fp64_instr_per_launch": instr, "fp64_issue_frac": instr / (p_ms * 1e-3) / (peak_tf * 1e12 / 2.0),
                "traffic": from_profile("propagate_sigma", "dram_bytes"), "kernel_ms_ncu": prop_ncu_ms,
                "frac_on_ncu_duration": (flop_prop / (prop_ncu_ms * 1e-3) * 1e-12 / peak_tf) if prop_ncu_ms else None,
                "hbm": {"achieved": 14 * 96.0 * N / (p_ms * 1e-3) * 1e-9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": 14 * 96.0 * N / (p_ms * 1e-3) * 1e-9 / hbm_peak, "bytes_per_launch": 14 * 96.0 * N}}
        # -- per-target kernel: HBM bound at large catalogs, latency bound at small ones -- both fractions
        quad_max = 8192 if w.all_tasked else 32768     # the library's size rule (a tasking mask may be dense)
        heavy = w.all_tasked or M * 16 >= N             # many measured targets per step: the 255-register build
        fin_name = ("ukf_finish_quad_kernel" if N <= quad_max else "ukf_finish_kernel<64, 4>" if heavy else
                    "ukf_finish_kernel<64, 6>" if N <= 131072 else "ukf_finish_kernel<128, 3>")
        m_seen = M // w.E                               # sensors one target is tested against (its own environment's)
        vis_bytes = 8.0 * ((m_seen + 31) // 32)         # two packed rows of 32-sensor words
        bytes_fin = (BYTES_FINISH_TARGET + vis_bytes) * N
        flop_fin = ((1 - tasked_frac) * FLOP_UKF_NOMEAS + tasked_frac * FLOP_UKF_MEAS + FLOP_NEXT_CHOL + 2 * m_seen * FLOP_VIS_PAIR) * N
        gbs = bytes_fin / (f_ms * 1e-3) * 1e-9
        tfs = flop_fin / (f_ms * 1e-3) * 1e-12
        fin_ncu_ms = from_profile("ukf_finish", "ms_mean")
        fin = {"kernel": fin_name, "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
               "kernel_ms": f_ms, "bytes_per_launch": bytes_fin, "algorithmic_bytes_per_unit": BYTES_FINISH_TARGET + vis_bytes,
               "sensors_per_target": m_seen,
               "unit_of_work": "one target: 13 sigma points read, est_p / pred_p / next sigma points written, 2 vis rows",
               "fp64": {"achieved": tfs, "peak": peak_tf, "unit": "TFLOP/s", "frac": tfs / peak_tf, "flop_per_launch": flop_fin},
               "traffic": from_profile("ukf_finish", "dram_bytes"), "kernel_ms_ncu": fin_ncu_ms,
               "frac_on_ncu_duration": (bytes_fin / (fin_ncu_ms * 1e-3) * 1e-9 / hbm_peak) if fin_ncu_ms else None,
               "tasked_fraction": tasked_frac}
        note = ("kernel_ms: CUDA events recorded by the library around each kernel of an eagerly launched step "
                "(the events cost a few us, which matters for kernels of 12-15 us at 10 000 targets and not at 125 000+); "
                "kernel_ms_ncu / traffic / frac_on_ncu_duration come from " +
                (prof["file"] + " (ncu capture of these very kernel sources, hash checked)" if prof else
                 "no profile summary matching the current kernel sources: null"))
        return prop, fin, note, tasked_frac, flags_bad, flop_prop

    # ------------------------------------------------------------------ headline: BASELINE.json configs[1]
    N, M = args.targets, args.sensors
    head = Workload(N, M, DT, SEED)
    head.build()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_per_step, wall, launches = head.timed(args.steps, args.warmup)
    value = world * N / (ms_per_step * 1e-3)

    b2b_ms = back_to_back(head, min(args.steps, 200))
    n_roof = max(3, min(args.steps, 50))
    prop, fin, roof_note, tasked_frac, flags_bad, flop_prop = roofline_kernels(head, 0, args.warmup, n_roof)
    flop_per_target_step = flop_prop / N + (1 - tasked_frac) * FLOP_UKF_NOMEAS + tasked_frac * FLOP_UKF_MEAS \
        + FLOP_NEXT_CHOL + 2 * M * FLOP_VIS_PAIR
    # `roofline`: the dominant kernel of the step (the longer of the two, by its live duration), with every kernel
    # of the step in `kernels`
    roofline = dict(prop if prop["kernel_ms"] >= fin["kernel_ms"] else fin)
    roofline.update({
        "kernels": [prop, fin], "note": roof_note,
        "peak_source": "FP64: DFMA probe measured live in this run (pc_fp64_peak; MEASURED_PEAKS.json has no FP64 entry, "
                       "nominal 37 TFLOP/s); HBM: MEASURED_PEAKS.json" + ("" if peaks else " missing -> fallback"),
        "kernel_share_of_step": {"propagate": prop["kernel_ms"] / ms_per_step, "finish": fin["kernel_ms"] / ms_per_step,
                                 "note": "event-timed kernels of the eager pass over the graph-replayed step: shares can "
                                         "exceed 1 in sum (events serialise what the graph overlaps)"},
        "whole_step_flop_per_target_step": flop_per_target_step,
        "whole_step_achieved_tflops": flop_per_target_step * N / (ms_per_step * 1e-3) * 1e-12,
        "whole_step_bytes_per_target_step": BYTES_TARGET_STEP + M / 4.0})
    clocks = sampler.stop() if rank == 0 else None

    # ------------------------------------------------------------------ end to end through the public API (`e2e`)
    # CatalogEngine.step_host -> pc_step_host: pinned host actions -> device, the step + observation
    # pack (CUDA graph inside the library), observation + packed vis map + num_tasked -> pinned host,
    # stream synchronise.  The policy runs on the host from the previous step's observation.
    def e2e_pass(w, steps, warmup):
        eng = w.eng
        eng.reset()
        eng._graph = None
        hb = eng.host_buffers()
        h_vis_ptr = C.c_void_p(hb["vis_est"].data_ptr())
        h_act_ptr = C.c_void_p(hb["actions"].data_ptr())
        host_policy_fn = lib.pc_policy_random_visible_envs_host
        Ne, Me, E = w.n // w.E, w.m // w.E, w.E

        def host_policy(k):
            # the stand-in policy of the `value` pass in its host form, on the packed estimated map the previous
            # pc_step_host wrote
            rc = host_policy_fn(h_vis_ptr, E, Ne, Me, eng.wpr, SEED + rank, k, 64, h_act_ptr)
            if rc:
                raise RuntimeError(f"pc_policy_random_visible_envs_host status {rc}")

        def one(k):
            host_policy(k)
            eng.step_host(hb, w.dt)
            if w.use_nccl:
                w.sg.gather()
                torch.cuda.current_stream(dev).synchronize()

        hb["vis_est"].copy_(eng.vis_est)
        for k in range(warmup):
            one(k)
        barrier()
        wall_s = 0.0
        for k0 in range(0, steps, EPISODE):            # episodes of EPISODE steps; the reset between them is untimed
            if k0:
                eng.reset()
                eng.ensure_sigma()
                hb["vis_est"].copy_(eng.vis_est)
                torch.cuda.synchronize(dev)
            tw = time.perf_counter()
            for k in range(k0, min(steps, k0 + EPISODE)):
                one(warmup + k)
            if eng._peer is not None:                   # the last step's exchange completes inside the timed region
                eng.peer_sync()
                torch.cuda.current_stream(dev).synchronize()
            wall_s += time.perf_counter() - tw          # every step ends in a stream synchronise
        barrier()
        ms = max_over_ranks(wall_s * 1e3) / steps
        return ms, hb

    e2e_ms, hb = e2e_pass(head, args.steps, args.warmup)
    e2e = {"value": world * N / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(M * 8),
           "d2h_bytes_per_step": int(hb["block"].numel()),     # float32 obs | num_tasked | packed vis | traces, one block
           "ms_per_step": e2e_ms, "host_cores_of_this_rank": pinned_cores,
           "path": "host policy (pc_policy_random_visible_host on the previous step's packed map) -> CatalogEngine.step_host -> "
                   "pc_step_host: pinned host actions -> [CUDA graph: H2D actions, pc_step_fused + pc_obs_pack] -> output block "
                   "(float32 obs | num_tasked | packed vis map | covariance traces | last_time_tasked) in pinned host memory, "
                   "stream sync; timed by the host clock around the loop.  Up to 32768 targets x 256 sensors per GPU the kernels "
                   "store the block into the (mapped) host buffer themselves while they run (zero-copy form of pc_step_host, "
                   "every byte rewritten every step; also beside a peer exchange); otherwise ONE device-to-host copy node "
                   "at the end of the graph"}
    head_exchange = head.exchange
    del head, hb
    torch.cuda.empty_cache()

    # ------------------------------------------------------------------ the other BASELINE.json configurations
    # (driver-visible numbers for the north-star sizes; fewer steps, same timing rules)
    extra_cfgs = {}
    if not args.no_extra_configs:
        xs = max(20, min(100, args.steps))

        def run_extra(key, desc, n, m, dt, total_units, unit_name, **kw):
            w = Workload(n, m, dt, SEED + 7, **kw)
            w.build()
            ms, _, nl = w.timed(xs, args.warmup)
            p2, f2, note2, tf2, bad2, _ = roofline_kernels(w, 0, args.warmup, min(20, xs))
            e_ms, hb2 = e2e_pass(w, min(xs, 50), args.warmup)
            out = {"workload": desc, "targets_per_gpu": n, "sensors": m, "dt_s": dt, "steps": xs, "ms_per_step": ms,
                   "target_steps_per_s": world * n / (ms * 1e-3), "scaling": "weak",
                   "launches_per_step": nl / xs, "exchange": w.exchange, "tasked_fraction": tf2, "flagged_targets": bad2,
                   "e2e_ms_per_step": e_ms, "e2e_target_steps_per_s": world * n / (e_ms * 1e-3),
                   "e2e_d2h_bytes_per_step": int(hb2["block"].numel()),
                   "roofline": [p2, f2]}
            if unit_name:
                out[unit_name] = total_units / (ms * 1e-3)
            if kw.get("force_model", _lib.FORCE_TWOBODY) != _lib.FORCE_TWOBODY:
                out["roofline_note"] = ("flop and FP64-instruction counts of the propagation kernel are the TWO-BODY figures "
                                        "(1150 flop / 435 instructions per DOP853 step); the J2 term adds work that is not counted")
            extra_cfgs[key] = out
            del w, hb2
            torch.cuda.empty_cache()

        # BASELINE.json words configs[1] as "100 ground/space sensors ... two-body+J2".  The reference's dynamics are
        # two-body only (dynamics.py:16-40), so the headline -- which is compared with the
        # reference's CPU path -- runs two-body with the reference's default ground sensors; this line is the same
        # catalog with 20 of the 100 sensors in LEO and the J2 term switched on (no reference counterpart: J2 is
        # checked against scipy in tests/test_gpu_parity.py)
        if args.targets == 10_000 and args.sensors == 100:
            run_extra("c2_j2_space_sensors", "BASELINE.json configs[1] as worded: 80 ground + 20 LEO space sensors x 10 000 mixed "
                      "LEO/MEO/GEO targets, two-body + J2, dt = 100 s", 10_000, 100, DT, 0, None, space_sensor_frac=0.2,
                      force_model=_lib.FORCE_TWOBODY_J2)
        run_extra("c4_shard", "BASELINE.json configs[3] shard: 125 000 debris targets per GPU x 100 sensors, 60 s steps "
                  "(8 GPUs = the 1 M-object catalog)", 125_000, 100, 60.0, 0, None)
        n5 = 100_000 // world
        run_extra("c5", f"BASELINE.json configs[4]: 100 000 targets in total ({n5} per GPU), full 6x6 predict + measurement "
                  "UPDATE on every target every step, 60 s steps", n5, 100, 60.0, 0, None, all_tasked=True)
        extra_cfgs["c5"]["scaling"] = "strong"
        extra_cfgs["c5"]["targets_total"] = n5 * world
        e3 = 1024 // world
        run_extra("c3", f"BASELINE.json configs[2]: 1024 vectorised SSAScheduler envs x 100 targets x 10 sensors in total "
                  f"({e3} envs per GPU, independent replicas), dt = 100 s", e3 * 100, e3 * 10, 100.0, 1024, "env_steps_per_s",
                  num_envs=e3, mix="LEO")
        extra_cfgs["c3"]["scaling"] = "strong"
        extra_cfgs["c3"]["e2e_env_steps_per_s"] = 1024 / (extra_cfgs["c3"]["e2e_ms_per_step"] * 1e-3)

    # BASELINE.json configs[0]: the reference's DEFAULT SSAScheduler env (3 ground sensors, 10 LEO targets, a UKF per
    # target), one 100-step episode through the Gymnasium API of the drop-in env -- dict observation and info built
    # every step, one library call per step.  The reference runs this episode on the CPU in 17.5 s.
    if not args.no_extra_configs and rank == 0:
        try:
            from clockpunch_b200.environment.env import SSAScheduler
            from clockpunch_b200.environment.env_parameters import SSASchedulerParams
            env1 = SSAScheduler(SSASchedulerParams(horizon=100, agent_params={"num_sensors": 3, "num_targets": 10}, seed=3), seed=3)
            assert env1.num_sensors == 3 and env1.num_targets == 10
            obs1, _ = env1.reset()
            rng1 = np.random.default_rng(0)
            e1 = env1._engine
            c1_state = dict(sens=e1.host("sens_x"), targ=e1.host("truth_x"), est_x=e1.host("est_x"), est_p=e1.host("est_p"),
                            Q=np.array(e1.params.Q[:]).reshape(6, 6), R=np.array(e1.params.R[:]).reshape(6, 6))

            def episode():
                o = obs1
                t0e = time.perf_counter()
                for _ in range(100):
                    vis = o["vis_map_est"]
                    act = np.where(vis.any(axis=0), (vis * (1 + rng1.random(vis.shape))).argmax(axis=0), env1.num_targets)
                    o, _, done1, _, info1 = env1.step(act)
                return time.perf_counter() - t0e, done1, info1

            episode()                                    # warm-up episode (graph capture); the env resets itself at the horizon
            secs, done1, info1 = episode()
            extra_cfgs["c1_default_env_episode"] = {
                "workload": "BASELINE.json configs[0]: default SSAScheduler env (3 ground sensors x 10 LEO targets, UKF per "
                            "target), one 100-step episode through env.step (dict observation + info every step)",
                "episode_s": secs, "ms_per_env_step": 10.0 * secs, "env_steps_per_s": 100.0 / secs,
                "target_steps_per_s": 1000.0 / secs, "episode_done": bool(done1), "sensors": 3, "targets": 10,
                "reference_note": "the reference's own profile of this config: 17.5 s per episode (SURVEY.md section 8d, measured in "
                                  "the build container); cpu_episode_s below is the oracle port of the same episode on this box"}
            if world == 1 and not args.no_cpu_baseline:
                # the same episode on one host core through the oracle port of the reference's step (one scipy DOP853
                # solve per state, one filter object per target, one Python call per pair): SURVEY.md section 8(d)
                from oracle.step import OracleCatalog
                st = c1_state
                orc = OracleCatalog(st["sens"], ["terrestrial"] * st["sens"].shape[1], st["targ"], st["est_x"], st["est_p"],
                                    st["Q"], st["R"])
                _, ve = orc.vis_maps()
                rng2 = np.random.default_rng(0)
                t0c = time.perf_counter()
                for k in range(100):
                    act = np.where(ve.any(axis=0), (ve * (1 + rng2.random(ve.shape))).argmax(axis=0), ve.shape[0])
                    tk = np.zeros(ve.shape[0], bool)
                    for j, a_ in enumerate(act):
                        if a_ < ve.shape[0] and ve[a_, j]:
                            tk[a_] = True
                    z = orc.targets + rng2.normal(size=orc.targets.shape) * np.sqrt(np.diag(st["R"]))[:, None]
                    # (the measurement of a tasked target is its propagated truth plus noise: drawn around the
                    # pre-step truth here, the cost is the same)
                    _, ve = orc.step(100.0 * (k + 1), tk, z)
                extra_cfgs["c1_default_env_episode"]["cpu_episode_s"] = time.perf_counter() - t0c
                extra_cfgs["c1_default_env_episode"]["cpu_kind"] = "oracle port, one core"
            del env1
        except Exception as exc:                         # informational line: never fail the bench on it
            extra_cfgs["c1_default_env_episode"] = {"error": repr(exc)}

    def leave():
        # Ranks leave together and without tearing NCCL down: destroying a process group whose
        # collectives live inside instantiated CUDA graphs can block at exit.
        if world > 1:
            sys.stdout.flush()
            torch.cuda.synchronize(dev)
            dist.barrier()
            torch.cuda.synchronize(dev)
            os._exit(0)

    if rank != 0:
        leave()
        return

    # ------------------------------------------------------------------ CPU baseline beside it (rank 0, N = 1 only)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cat = synth_catalog(N, M, seed=SEED, sensor_seed=SEED)
        cores = host_cores()
        sample = min(N, max(cores, (args.cpu_sample_per_core or 640) * cores))    # ~15 s of CPU work
        v, s_step, sample = cpu_reference_run(cat, sample, 1, 0, cores)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{sample} of {N} targets x {M} sensors, one pass ({s_step:.1f} s), oracle port on "
                         f"{cores} processes (scipy DOP853 per state, python loop per pair)"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(N, M),
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "extra": {"tasked_per_step": tasked_frac * N, "flagged_targets": flags_bad,
                      "wall_s_timed_region": wall, "launches_per_step": launches / args.steps, "cuda_graph": True,
                      "exchange": head_exchange, "target_propagation_steps_per_s": 14 * value,
                      "back_to_back": {"ms_per_step": b2b_ms, "target_steps_per_s": world * N / (b2b_ms * 1e-3),
                                       "note": "the same step graph replayed back to back WITHOUT the L2 flush between steps "
                                               "(a rollout); informational, the headline keeps the flush"},
                      "configs": extra_cfgs}}
    emit(line)
    leave()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--targets", type=int, default=N_TARGETS)
    ap.add_argument("--sensors", type=int, default=N_SENSORS)
    ap.add_argument("--substeps", type=int, default=0)
    ap.add_argument("--cpu-sample-per-core", type=int, default=0, help="0 = sized for the time budget")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true",
                    help="skip the BASELINE.json configs[2..4] passes reported under extra.configs")
    ap.add_argument("--l2", default="flush", choices=["flush", "none"])
    ap.add_argument("--eager-gather", action="store_true", help="N > 1: all-gather outside the step's CUDA graph")
    ap.add_argument("--opt-propagate", type=int, default=0, help="tuning: pc_option PC_OPT_PROPAGATE_KERNEL (0 = size rule)")
    ap.add_argument("--opt-finish", type=int, default=0, help="tuning: pc_option PC_OPT_FINISH_KERNEL (0 = size rule)")
    ap.add_argument("--exchange", default="auto", choices=["auto", "peer", "nccl", "none"],
                    help="N > 1: summaries pushed into peers' memory by the step kernel, or one NCCL all-gather "
                         "(none: no exchange at all -- attribution runs only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "own" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
