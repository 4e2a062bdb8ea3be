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
# DRAM traffic per launch of propagate_sigma_kernel from the committed ncu captures: (targets, sensors) -> bytes
TRAFFIC_NCU = {(10_000, 100): 6.784e6 + 0.0, (1_000_000, 100): 680.5e6 + 614.8e6}
# the same kernel's gpu__time_duration.sum (ms, mean of 26 launches) in the committed launch lists
# profiles/r1d_launches_bench_{10k,1M}.csv: what the kernel itself takes, without the event pair's ~8 us
NCU_KERNEL_MS = {(10_000, 100): 0.013758, (1_000_000, 100): 0.724994}
BYTES_TARGET_STEP = 1130.0         # truth/est_x/est_p r+w, pred_p w, z r (+ M/4 vis bits)


def workload_config(n_targets, n_sensors):
    which = ("BASELINE.json configs[1]" if (n_targets, n_sensors) == (N_TARGETS, N_SENSORS) else
             "BASELINE.json configs[3] shard (1 M targets over 8 GPUs)" if (n_targets, n_sensors) == (125_000, 100) else
             "other size of the configs[1] workload")
    return {"workload": f"{n_sensors} ground sensors x {n_targets} mixed LEO/MEO/GEO targets, two-body, "
                        f"dt={DT:.0f}s, UKF per target ({which})",
            "targets_per_gpu": n_targets, "sensors": n_sensors, "dt_s": DT,
            "tasking": "each sensor tasks a random estimated-visible target (<= M per step)",
            "episode": f"catalog reset to its initial state every {EPISODE} steps (outside the timed events)",
            "l2": "256 MiB flush between timed steps (state fits the 126 MB L2)"}


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

    import torch
    import torch.distributed as dist
    from clockpunch_b200 import _lib
    from clockpunch_b200.engine import CatalogEngine
    from clockpunch_b200.simulation.catalog import synth_catalog
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    N, M = args.targets, args.sensors
    # every rank owns its own shard of targets; the sensor network is replicated
    cat = synth_catalog(N, M, seed=SEED + 1000 * rank, sensor_seed=SEED)
    eng = CatalogEngine(cat.sensors, cat.sensor_kinds, cat.targets, cat.est_x, cat.est_p, cat.Q, cat.R,
                        device=local, seed=SEED + rank)
    ctx, lib = eng.ctx, eng.ctx.lib
    stream = torch.cuda.current_stream(dev).cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    # per-step summaries every rank needs from every other: packed est vis map,
    # trace(P_pos), trace(P_vel), last_time_tasked, num_tasked -- all-gathered straight out of the
    # buffers the kernels write (equal shards: no staging copy)
    # Exchange: by default the per-target kernel itself stores every summary into all peers'
    # gathered buffers over NVLink (peer-mapped symmetric memory) and the step ends with a flag wait;
    # --exchange nccl uses one all-gather of the summary buffer captured into the step's graph.
    sg, exchange = None, "none"
    if world > 1:
        from clockpunch_b200.sharding import PackedGather
        sg = PackedGather(world, rank, eng.summary, N, eng.wpr)
        exchange = "nccl all_gather_into_tensor" + ("" if args.eager_gather else " (captured in the step graph)")
        # measured on 8 GPUs: 10 000 targets per GPU 61 us per step (stores from the per-target kernel)
        # vs 79 us with the all-gather; 125 000 per GPU 0.264 ms (copy kernel behind it) vs 0.287 ms
        if args.exchange in ("peer", "auto"):
            try:
                eng.enable_peer_push()
                mode = {"auto": 0, "store": 1, "copy": 2}[args.peer_push]
                ctx_opt = eng.ctx.set_option(_lib.OPT_PEER_PUSH, mode)
                by_copy = mode == 2 or (mode == 0 and (N > 32768 or world >= 8))
                exchange = ("peer stores from the per-target kernel" if not by_copy else
                            "summary pushed into every peer by a copy kernel behind the per-target kernel") + \
                    " (symmetric memory over NVLink) + flag wait"
            except Exception as exc:        # no symmetric memory on this box: keep the collective
                print(f"peer push unavailable ({exc!r}); using NCCL", file=sys.stderr)
    use_nccl = sg is not None and eng._peer is None
    gather_in_graph = use_nccl and not args.eager_gather
    post = sg.gather if gather_in_graph else None

    # policy -> tasking mask -> sensors -> truth + UKF -> vis maps [-> all-gather], captured once,
    # replayed per step
    g_value = None

    def device_step():
        eng.replay(g_value)
        if use_nccl and not gather_in_graph:
            sg.gather()

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---------------- device-resident throughput (`value`) ----------------
    if sg is not None:
        sg.gather()                      # first NCCL call (communicator set-up) outside any capture
        torch.cuda.synchronize(dev)
    g_value = eng.build_graph(DT, policy_probes=64, policy_seed=SEED, post=post)
    for k in range(args.warmup):
        device_step()
    if eng._peer is not None:
        # check the pushed copies once against a plain all-gather of the same buffers (untimed); if any
        # rank disagrees, every rank falls back to the NCCL exchange
        ref = sg.gather().clone()
        err, pushes = eng.peer_status()
        bad = torch.tensor([int(bool(err) or not torch.equal(eng._peer["gathered"][pushes & 1], ref))], device=dev)
        dist.all_reduce(bad, op=dist.ReduceOp.MAX)
        if int(bad.item()):
            print(f"rank {rank}: peer-pushed summaries differ from the all-gather (error mask {err:#x}); "
                  "falling back to NCCL", file=sys.stderr)
            eng._peer, eng._hs_key = None, None
            use_nccl, gather_in_graph = True, not args.eager_gather
            post = sg.gather if gather_in_graph else None
            exchange = "nccl all_gather_into_tensor (fallback: peer push failed its check)"
            eng.reset()
            g_value = eng.build_graph(DT, policy_probes=64, policy_seed=SEED, post=post)
            for k in range(args.warmup):
                device_step()
    peak_tf = C.c_double()
    ctx.check(lib.pc_fp64_peak(ctx.handle, 5, C.byref(peak_tf)))
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    wall0 = time.perf_counter()
    for k in range(args.steps):
        if k and k % EPISODE == 0:         # a new episode, as the env does at its horizon (untimed)
            eng.reset()
            eng.ensure_sigma()
        if args.l2 == "flush":
            flush.zero_()                  # L2 flush, outside the timed events
        evs[k][0].record()
        device_step()
        evs[k][1].record()
    barrier()
    wall = time.perf_counter() - wall0
    ms = sum(a.elapsed_time(b) for a, b in evs)
    launches = g_value.pc_launches * args.steps
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(t_ms.item()) / args.steps
    value = world * N / (ms_per_step * 1e-3)

    # ---------------- roofline of the dominant kernel (separate pass) ----------------
    # propagate_sigma_kernel (two states per thread out of a sorted 256-target window; catalogs up to
    # 16 k targets: propagate_sigma_single_kernel, one state per thread): all 13 sigma points + the
    # truth state of every target.  Its launch is bracketed by a CUDA-event pair recorded by the library on the launch
    # stream.  This pass launches the step's kernels one by one (not the graph `value` replays):
    # inside a captured graph the pair would be two extra event-record NODES, whose latency (~4 us
    # each) would be charged to a 14 us kernel.
    # Algorithmic work: SURVEY.md section 8(d), 1150 flop per fixed DOP853 step x the substeps each state
    # actually takes (read back from the per-target orbit rates the step itself uses); algorithmic
    # bytes: each 6-vector read and written once.
    # Every pass (value, roofline, e2e) starts from the same catalog state: reset + warm-up.
    eng.reset()
    ctx.check(lib.pc_set_timing(ctx.handle, 1))
    eng._graph = None

    def timed_step():
        eng.step_policy_eager(DT, 64, SEED)         # same kernels, launched one by one: plain event records
        if use_nccl:
            sg.gather()

    for k in range(args.warmup):
        timed_step()
    inv_theta = np.float32(1.0 / 0.12)

    def substeps_now():
        om = eng.sig_flags[1].cpu().numpy().view(np.float32)
        est = np.maximum(1.0, np.ceil(np.float32(DT) * om * inv_theta)).astype(np.float64)
        return 13.0 * est.sum() + truth_sub.sum(), est.mean()

    tx = eng.host("truth_x")
    hv = np.cross(tx[:3].T, tx[3:].T)
    h2, r2, v2 = (hv * hv).sum(1), (tx[:3] ** 2).sum(0), (tx[3:] ** 2).sum(0)
    e1 = 1.0 + np.sqrt(np.maximum(0.0, 1.0 + 2.0 * (0.5 * v2 - 398600.0 / np.sqrt(r2)) * h2 / 398600.0 ** 2))
    truth_sub = np.maximum(1.0, np.ceil(DT * (398600.0 ** 2 * e1 ** 3.5 / h2 ** 1.5) / 0.12))
    ukf_ms_sum, state_steps_sum, est_sub_sum, ukf_last = 0.0, 0.0, 0.0, C.c_double()
    n_roof = max(3, min(args.steps, 50))
    for k in range(n_roof + 2):
        ss, es = substeps_now()
        if args.l2 == "flush":
            flush.zero_()
        timed_step()
        torch.cuda.synchronize(dev)                                   # both events are this step's
        ctx.check(lib.pc_get_timing(ctx.handle, C.byref(ukf_last)))
        if k >= 2:
            ukf_ms_sum += ukf_last.value
            state_steps_sum += ss
            est_sub_sum += es
    ctx.check(lib.pc_set_timing(ctx.handle, 0))
    tasked_total = int(eng.num_tasked.sum().item())
    flags_bad = int((eng.flags != 0).sum().item())
    total_steps = args.warmup + n_roof + 2
    tasked_frac = tasked_total / max(1, N * total_steps)
    state_steps = state_steps_sum / n_roof                  # DOP853 steps per launch, summed over 14 N states
    substeps_used = est_sub_sum / n_roof
    flop_per_launch = FLOP_DOP853_STEP * state_steps
    flop_per_target_step = flop_per_launch / N + (1 - tasked_frac) * FLOP_UKF_NOMEAS + tasked_frac * FLOP_UKF_MEAS \
        + 2 * M * FLOP_VIS_PAIR
    ukf_ms_avg = ukf_ms_sum / n_roof
    achieved_tf = flop_per_launch / (ukf_ms_avg * 1e-3) * 1e-12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_ach = 14 * 96.0 * N / (ukf_ms_avg * 1e-3) * 1e-9
    roofline = {"bound": "fp64", "kernel": "propagate_sigma_kernel" if N > 16 * 1024 else "propagate_sigma_single_kernel",
                "achieved": achieved_tf, "peak": peak_tf.value,
                "unit": "TFLOP/s", "frac": achieved_tf / peak_tf.value if peak_tf.value else None,
                "traffic": TRAFFIC_NCU.get((N, M)), "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of this "
                "kernel in the committed ncu --set full capture of this workload (profiles/r1d_ncu_full_step_*_metrics.csv); "
                "null for other sizes", "peak_source": "FP64 DFMA probe measured live in this run (pc_fp64_peak); "
                "MEASURED_PEAKS.json has no FP64 entry (HBM/bf16 only); nominal 37 TFLOP/s",
                "flop_per_launch": flop_per_launch, "dop853_steps_per_launch": state_steps,
                "mean_substeps_per_sigma_propagate": substeps_used,
                "kernel_ms": ukf_ms_avg, "kernel_share_of_step": ukf_ms_avg / ms_per_step,
                "kernel_ms_ncu": NCU_KERNEL_MS.get((N, M)),
                "frac_on_ncu_duration": (flop_per_launch / (NCU_KERNEL_MS[(N, M)] * 1e-3) * 1e-12 / peak_tf.value
                                         if (N, M) in NCU_KERNEL_MS and peak_tf.value else None),
                "frac_note": "frac uses kernel_ms, a CUDA-event pair around ONE launch per step (the pair itself "
                "costs ~8 us, which matters for a 14 us kernel at 10 000 targets and not at 1 M); frac_on_ncu_duration "
                "divides the same live flop count by the kernel's committed ncu duration (profiles/r1d_launches_*.csv)",
                "whole_step_flop_per_target_step": flop_per_target_step,
                "whole_step_achieved": flop_per_target_step * N / (ms_per_step * 1e-3) * 1e-12,
                "hbm": {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                        "bytes_per_launch": 14 * 96.0 * N,
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}}
    clocks = sampler.stop() if rank == 0 else None

    # ---------------- end to end through the public API (`e2e`) ----------------
    # CatalogEngine.step_host -> pc_step_host: pinned host actions -> device, the step + observation
    # pack (CUDA graph inside the library), observation + packed vis map + num_tasked -> pinned host,
    # stream synchronise.  The policy runs on the host from the previous step's observation.
    eng.reset()
    eng._graph = None
    hb = eng.host_buffers()
    # Host-side scheduler: the stand-in policy of the `value` pass (first estimated-visible target among 64
    # hashed probes per sensor, else inaction) in its host form, reading the packed estimated map that the
    # previous pc_step_host copied out and writing the pinned actions buffer of the next one.
    import ctypes as C
    h_vis_ptr = C.c_void_p(hb["vis_est"].data_ptr())
    h_act_ptr = C.c_void_p(hb["actions"].data_ptr())
    host_policy_fn = eng.ctx.lib.pc_policy_random_visible_host

    def host_policy(k):
        rc = host_policy_fn(h_vis_ptr, N, M, eng.wpr, SEED + rank, k, 64, h_act_ptr)
        if rc:
            raise RuntimeError(f"pc_policy_random_visible_host status {rc}")

    def e2e_step(k):
        host_policy(k)
        eng.step_host(hb, DT)
        if use_nccl:
            sg.gather()
            torch.cuda.current_stream(dev).synchronize()

    hb["vis_est"].copy_(eng.vis_est)
    for k in range(args.warmup):
        e2e_step(k)
    barrier()
    e2e_wall = 0.0
    for k0 in range(0, args.steps, EPISODE):            # episodes of EPISODE steps; the reset between them is untimed
        if k0:
            eng.reset()
            eng.ensure_sigma()
            hb["vis_est"].copy_(eng.vis_est)
            torch.cuda.synchronize(dev)
        tw = time.perf_counter()
        for k in range(k0, min(args.steps, k0 + EPISODE)):
            e2e_step(args.warmup + k)
        e2e_wall += time.perf_counter() - tw            # every step ends in a stream synchronise
    barrier()
    e2e_ms = torch.tensor([e2e_wall * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * N * args.steps / (float(e2e_ms.item()) * 1e-3)
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(M * 8),
           "d2h_bytes_per_step": int(hb["block"].numel()),     # float32 obs | num_tasked | packed vis | traces, one copy
           "ms_per_step": float(e2e_ms.item()) / args.steps,
           "path": "host policy (pc_policy_random_visible_host on the previous step's packed map) -> CatalogEngine.step_host -> pc_step_host: pinned host actions "
                   "-> [CUDA graph: H2D actions, pc_step_fused + pc_obs_pack] -> output block (float32 obs | num_tasked | packed vis "
                   "map | covariance traces | last_time_tasked) in pinned host memory, stream sync; timed by the host clock "
                   "around the loop.  Up to 32768 targets x 256 sensors per GPU and without a peer exchange the kernels store "
                   "the block into the (mapped) host buffer themselves while they run (zero-copy form of pc_step_host, every "
                   "byte rewritten every step); otherwise ONE device-to-host copy node at the end of the graph"}

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

    # ---------------- CPU baseline beside it (rank 0, N = 1 only) ----------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
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
            "extra": {"tasked_per_step": tasked_total / total_steps, "flagged_targets": flags_bad,
                      "wall_s_timed_region": wall, "launches_per_step": launches / args.steps, "cuda_graph": True,
                      "exchange": exchange,
                      "target_propagation_steps_per_s": 14 * value}}
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
    ap.add_argument("--l2", default="flush", choices=["flush", "none"])
    ap.add_argument("--eager-gather", action="store_true", help="N > 1: all-gather outside the step's CUDA graph")
    ap.add_argument("--peer-push", default="auto", choices=["auto", "store", "copy"],
                    help="N > 1, peer exchange: stores from the per-target kernel or a copy kernel behind it (auto: size rule)")
    ap.add_argument("--exchange", default="auto", choices=["auto", "peer", "nccl"],
                    help="N > 1: summaries pushed into peers' memory by the step kernel, or one NCCL all-gather")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "own" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
