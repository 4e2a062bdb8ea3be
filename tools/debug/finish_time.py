"""Step time at N targets with nobody / ~M / everybody tasked (CUDA events over graph replays)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from clockpunch_b200.engine import CatalogEngine
from clockpunch_b200.simulation.catalog import synth_catalog
SEED = 20240627
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
cat = synth_catalog(N, 100, seed=SEED, sensor_seed=SEED)
eng = CatalogEngine(cat.sensors, cat.sensor_kinds, cat.targets, cat.est_x, cat.est_p, cat.Q, cat.R, device=0, seed=SEED)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run(label, graph, steps=30, pre=None):
    eng.reset()
    for _ in range(5):
        if pre: pre()
        eng.replay(graph)
    ms = 0.0
    for _ in range(steps):
        if pre: pre()
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.replay(graph); e1.record(); torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    print(f"{label:28s} {ms/steps*1e3:8.1f} us/step   tasked/step {eng.host('num_tasked').sum()/(steps+5):.0f}")
g_pol = eng.build_graph(100.0, policy_probes=64, policy_seed=SEED)
run("policy (~M tasked)", g_pol)
g_act = eng.build_graph(100.0, policy_probes=0)
eng.actions.fill_(N)
run("nobody tasked", g_act)
# everybody tasked: a graph whose tasked mask is all ones (set before each replay)
import ctypes as C
def all_tasked():
    eng.tasked.fill_(1)
eng.reset(); eng.ensure_sigma()
g = torch.cuda.CUDAGraph()
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
eng._sync_clock()
with torch.cuda.stream(side):
    with torch.cuda.graph(g, stream=side):
        a = eng._fill_args(0.0, 100.0, None, eng.tasked, use_clock=True)
        eng.ctx.check(eng.ctx.lib.pc_step_fused(eng.ctx.handle, C.byref(a), C.byref(eng.params), eng._stream()))
torch.cuda.current_stream().wait_stream(side)
g.pc_dt, g.pc_launches = 100.0, 3
run("everybody tasked", g, pre=all_tasked)
