"""Graph-replayed step of the large configurations (what bench.py's extra.configs time), for A/B runs of variant
libraries: PUNCH_B200_LIB=variants/libpunch_X.so python tools/debug/large_configs.py [finish variant]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from clockpunch_b200 import _lib
from clockpunch_b200.engine import CatalogEngine
from clockpunch_b200.simulation.catalog import synth_catalog
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
cases = [("c3", 102400, 10240, 1024, 100.0), ("c4", 125000, 100, 1, 60.0), ("c5", 100000, 100, 1, 60.0), ("200k", 200000, 100, 1, 60.0), ("1M", 1000000, 100, 1, 60.0)]
for name, n, m, E, dt in cases:
    cat = synth_catalog(n, m, seed=7, space_sensor_frac=0.0) if E > 1 else synth_catalog(n, m, seed=7)
    ctx = _lib.get_ctx()
    ctx.set_option(_lib.OPT_FINISH_KERNEL, variant)
    eng = CatalogEngine(cat.sensors, cat.sensor_kinds, cat.targets, cat.est_x, cat.est_p, cat.Q, cat.R, seed=1, num_envs=E)
    g = eng.build_graph(dt, all_tasked=True) if name == "c5" else eng.build_graph(dt, policy_probes=64, policy_seed=3)
    for _ in range(5):
        eng.replay(g)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(40)]
    for a, b in evs:
        flush.zero_()
        a.record(); eng.replay(g); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    print(f"{name} n={n}: mean {sum(ts)/len(ts)*1e3:.1f} median {ts[len(ts)//2]*1e3:.1f} min {ts[0]*1e3:.1f} us/step flags {(eng.flags != 0).sum().item()}")
    del eng, g
