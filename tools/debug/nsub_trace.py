"""Trace the per-target substep counts of the bench workload over many steps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from clockpunch_b200.engine import CatalogEngine
from clockpunch_b200.simulation.catalog import synth_catalog
SEED = 20240627
cat = synth_catalog(10000, 100, seed=SEED, sensor_seed=SEED)
eng = CatalogEngine(cat.sensors, cat.sensor_kinds, cat.targets, cat.est_x, cat.est_p, cat.Q, cat.R, device=0, seed=SEED)
g = eng.build_graph(100.0, policy_probes=64, policy_seed=SEED)
for k in range(300):
    eng.replay(g)
    if k % 20 == 19 or k < 3:
        w = eng.sig_flags[1].cpu().numpy().view(np.uint32)      # packed rates: low half = bound for steps <= 128 s
        om = ((w & 0xFFFF).astype(np.uint32) << 16).view(np.float32)
        ns = np.maximum(1, np.ceil(np.float32(100.0) * om * np.float32(1 / 0.12)))
        top = np.argsort(ns)[-3:]
        nt = eng.host("num_tasked")
        ex, tx = eng.host("est_x"), eng.host("truth_x")
        err = np.linalg.norm(ex[:3] - tx[:3], axis=0)
        print(k, "mean nsub %.3f max %d" % (ns.mean(), ns.max()), "top", top, ns[top], "regime", cat.regime[top],
              "num_tasked", nt[top], "pos err km", err[top].round(1), "flags", eng.host("flags")[top],
              "median err %.2f p99 %.1f" % (np.median(err), np.percentile(err, 99)), "tasked total", nt.sum())
