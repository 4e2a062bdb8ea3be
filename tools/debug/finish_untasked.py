"""Step time of the bench catalog with and without any tasking (are the update warps or the regular path the
longer chain of the per-target kernel?): python tools/debug/finish_untasked.py [targets]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from clockpunch_b200.engine import CatalogEngine
from clockpunch_b200.simulation.catalog import synth_catalog
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
SEED = 20240627
cat = synth_catalog(N, 100, seed=SEED, sensor_seed=SEED)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(eng, g, steps=300, episode=50):
    evs = []
    for k in range(steps):
        if k % episode == 0:
            eng.reset(); eng.ensure_sigma()
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.replay(g); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in evs) / steps * 1e3


for label, probes in (("policy (about M targets measured per step)", 64), ("no tasking at all", 0)):
    eng = CatalogEngine(cat.sensors, cat.sensor_kinds, cat.targets, cat.est_x, cat.est_p, cat.Q, cat.R, device=0, seed=SEED)
    eng.actions.fill_(N)                     # inaction unless the policy overwrites it
    g = eng.build_graph(100.0, policy_probes=probes, policy_seed=SEED)
    for _ in range(5):
        eng.replay(g)
    print(f"{label}: {timed(eng, g):.2f} us / step, measured in total {int(eng.num_tasked.sum().item())}")
