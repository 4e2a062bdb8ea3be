"""Rollout-stepping rate of BASELINE.json configs[2] (1024 envs x 100 targets x 10 sensors):
python tools/bench_vector_env.py [E] [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from clockpunch_b200.environment.vector_env import VectorSSAScheduler
from clockpunch_b200.simulation.catalog import synth_catalog
E = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
M, N = 10, 100
cat = synth_catalog(E * N, E * M, seed=20240628, mix="LEO")
S = cat.sensors.reshape(6, E, M).transpose(1, 0, 2)
T = cat.targets.reshape(6, E, N).transpose(1, 0, 2)
X = cat.est_x.reshape(6, E, N).transpose(1, 0, 2)
venv = VectorSSAScheduler(S, cat.sensor_kinds.reshape(E, M), T, X, cat.est_p[0], cat.Q, cat.R, 100.0, horizon=10**9, seed=1)
obs, _ = venv.reset()
rng = np.random.default_rng(0)
acts = rng.integers(0, N + 1, size=(steps + 5, E, M))
for k in range(5):
    venv.step(acts[k])
e, hb = venv._engine, venv._hb
t0 = time.perf_counter()
for k in range(steps):                     # engine-level: actions in, packed obs out (no dict assembly)
    hb["actions"][: E * M].copy_(e.torch.from_numpy(acts[5 + k].reshape(-1)))
    e.step_host(hb, 100.0)
t1 = time.perf_counter()
for k in range(steps):                     # env-level: dict observation with the (E, N, M) int64 map
    venv.step(acts[5 + k])
t2 = time.perf_counter()
views = VectorSSAScheduler(S, cat.sensor_kinds.reshape(E, M), T, X, cat.est_p[0], cat.Q, cat.R, 100.0, horizon=10**9, seed=1, copy=False)
views.reset()
for k in range(5):
    views.step(acts[k])
t3 = time.perf_counter()
for k in range(steps):                     # env-level, observation arrays as views of the pinned block
    views.step(acts[5 + k])
t4 = time.perf_counter()
print(f"VectorSSAScheduler(copy=False).step {1e3*(t4-t3)/steps:.3f} ms/step = {E*steps/(t4-t3):.3g} env-steps/s")
print(f"E={E} envs x {N} targets x {M} sensors: step_host {1e3*(t1-t0)/steps:.3f} ms/step = {E*steps/(t1-t0):.3g} env-steps/s "
      f"({E*N*steps/(t1-t0):.3g} target-steps/s); VectorSSAScheduler.step {1e3*(t2-t1)/steps:.3f} ms/step = {E*steps/(t2-t1):.3g} env-steps/s")
