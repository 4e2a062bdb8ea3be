"""Which per-target kernel mapping suits a catalog whose EVERY target is measured every step (BASELINE.json configs[4]
shards): python tools/debug/all_tasked_mapping.py [targets ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from clockpunch_b200 import _lib
from clockpunch_b200.engine import CatalogEngine
from clockpunch_b200.simulation.catalog import synth_catalog
sizes = [int(a) for a in sys.argv[1:]] or [12500, 25000, 32768]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for n in sizes:
    cat = synth_catalog(n, 100, seed=7)
    for variant in (0, 2, 3):
        ctx = _lib.get_ctx()
        ctx.set_option(_lib.OPT_FINISH_KERNEL, variant)
        eng = CatalogEngine(cat.sensors, cat.sensor_kinds, cat.targets, cat.est_x, cat.est_p, cat.Q, cat.R, seed=1)
        g = eng.build_graph(60.0, all_tasked=True)
        for _ in range(5):
            eng.replay(g)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(50)]
        for a, b in evs:
            flush.zero_()
            a.record(); eng.replay(g); b.record()
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in evs) / len(evs)
        print(f"n={n} all tasked, finish variant {variant}: {ms*1e3:.1f} us/step  flags {(eng.flags != 0).sum().item()}")
        ctx.set_option(_lib.OPT_FINISH_KERNEL, 0)
