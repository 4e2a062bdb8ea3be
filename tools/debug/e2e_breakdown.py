"""Where the end-to-end step goes (host clock, no L2 flush -- as bench.py's e2e pass):
python tools/debug/e2e_breakdown.py [targets sensors envs dt]      (default: configs[1]; configs[2]: 102400 10240 1024 100)"""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
# optional: run on the cores of the GPU's NUMA node (E2E_NODE=local) or on the others (E2E_NODE=remote)
if os.environ.get("E2E_NODE"):
    import pynvml as nv
    nv.nvmlInit()
    bus = nv.nvmlDeviceGetPciInfo(nv.nvmlDeviceGetHandleByIndex(0)).busId
    bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
    bus = bus[4:] if len(bus) > 12 else bus
    node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
    nodes = sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
    def cpus(nd):
        out = set()
        for part in open(f"/sys/devices/system/node/node{nd}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            out.update(range(int(lo), int(hi or lo) + 1))
        return out
    allowed = set(os.sched_getaffinity(0))
    local = cpus(node) & allowed if node >= 0 else set()
    remote = allowed - local
    pick = local if os.environ["E2E_NODE"] == "local" else remote
    print(f"GPU 0 on NUMA node {node} of {nodes}; allowed cores {len(allowed)}, local {len(local)}, remote {len(remote)}; using {os.environ['E2E_NODE']}")
    if pick:
        os.sched_setaffinity(0, pick)
from clockpunch_b200.engine import CatalogEngine
from clockpunch_b200.simulation.catalog import synth_catalog
SEED = 20240627
N, M, E, dt = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])) if len(sys.argv) > 4 else (10000, 100, 1, 100.0)
K = 400 if N <= 20000 else 100
cat = synth_catalog(N, M, seed=SEED, sensor_seed=SEED, mix="mixed" if E == 1 else "LEO")
eng = CatalogEngine(cat.sensors, cat.sensor_kinds, cat.targets, cat.est_x, cat.est_p, cat.Q, cat.R, device=0, seed=SEED, num_envs=E)
lib = eng.ctx.lib
hb = eng.host_buffers()
hv, ha = C.c_void_p(hb["vis_est"].data_ptr()), C.c_void_p(hb["actions"].data_ptr())
hb["vis_est"].copy_(eng.vis_est)
def policy(k): lib.pc_policy_random_visible_envs_host(hv, E, N // E, M // E, eng.wpr, SEED, k, 64, ha)
def loop(fn, n=K, episode=50):
    """Episodes of 50 steps from the initial catalog, like bench.py (the reset is untimed)."""
    tot = 0.0
    for k0 in range(0, n, episode):
        eng.reset(); eng.ensure_sigma(); hb["vis_est"].copy_(eng.vis_est); torch.cuda.synchronize()
        for k in range(5): fn(k)
        torch.cuda.synchronize(); t = time.perf_counter()
        for k in range(episode): fn(5 + k0 + k)
        torch.cuda.synchronize(); tot += time.perf_counter() - t
    return tot / n * 1e6
print("host policy alone             %.1f us" % loop(policy))
def full(k): policy(k); eng.step_host(hb, dt)
print("policy + step_host (zero copy) %.1f us" % loop(full))
print("step_host (zero copy)          %.1f us" % loop(lambda k: eng.step_host(hb, dt)))
eng._hs_key = None
print("step_host (one D2H copy node)  %.1f us" % loop(lambda k: eng.step_host(hb, dt, zero_copy=False)))
g = eng.build_graph(dt)                      # tasking from eng.actions (device), no host block
s = torch.cuda.current_stream()
def dev(k): eng.replay(g); s.synchronize()
print("graph replay + stream sync     %.1f us" % loop(dev))
def dev2(k): eng.replay(g)
print("graph replay back to back      %.1f us" % loop(dev2))
