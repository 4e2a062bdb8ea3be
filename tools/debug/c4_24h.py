"""BASELINE.json configs[3] in full on one shard: 125 000 debris targets x 100 sensors, 24 h at 60 s steps = 1440 steps
of the device-resident step graph (stand-in policy, measurements drawn on device), no reset in between.
python tools/debug/c4_24h.py [targets] [steps]"""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from clockpunch_b200 import _lib
from clockpunch_b200.engine import CatalogEngine
from clockpunch_b200.simulation.catalog import synth_catalog, MU
N = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 1440
cat = synth_catalog(N, 100, seed=20240628)
eng = CatalogEngine(cat.sensors, cat.sensor_kinds, cat.targets, cat.est_x, cat.est_p, cat.Q, cat.R, seed=1)
g = eng.build_graph(60.0, policy_probes=64, policy_seed=3)
for _ in range(3):
    eng.replay(g)
eng.reset(); eng.ensure_sigma()
torch.cuda.synchronize()
x0 = eng.host("truth_x")
marks = {}
t0 = time.perf_counter()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for k in range(STEPS):
    eng.replay(g)
    if k + 1 in (1, 60, 360, 720, 1440):
        torch.cuda.synchronize()
        marks[k + 1] = time.perf_counter() - t0
b.record(); torch.cuda.synchronize()
wall = time.perf_counter() - t0
x1 = eng.host("truth_x")
en = lambda x: 0.5 * (x[3:] ** 2).sum(0) - MU / np.linalg.norm(x[:3], axis=0)
fl = eng.host("flags")
ep = eng.host("est_p")
nt = eng.host("num_tasked")
out = {"targets": N, "sensors": 100, "steps": STEPS, "dt_s": 60.0, "simulated_hours": STEPS * 60.0 / 3600.0,
       "device_s": a.elapsed_time(b) * 1e-3, "wall_s": wall, "ms_per_step_mean": a.elapsed_time(b) / STEPS,
       "wall_s_at_step": marks,
       "truth_energy_rel_drift_max": float(np.abs(en(x1) / en(x0) - 1).max()),
       "est_p_finite": bool(np.isfinite(ep).all()), "est_x_finite": bool(np.isfinite(eng.host("est_x")).all()),
       "targets_flagged_last_step": {name: int(((fl & getattr(_lib, name)) != 0).sum()) for name in
                                     ("FLAG_NONPD_EST", "FLAG_NONPD_PRED", "FLAG_NONPD_INNOV", "FLAG_NONFINITE", "FLAG_SUBSTEP_CLAMP")},
       "measurements_total": int(nt.sum()), "targets_measured_at_least_once": int((nt > 0).sum()),
       "median_position_sigma_km": float(np.median(np.sqrt(ep[:, 0, 0] + ep[:, 1, 1] + ep[:, 2, 2])))}
print(json.dumps(out, indent=1))
