"""Cost of one SSAScheduler.step (the Gymnasium env of the reference, dict observation + info) at a given size, next
to the engine-level call it wraps: python tools/bench_env_step.py [targets] [sensors] [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from clockpunch_b200.common.agents import Sensor, Target
from clockpunch_b200.dynamics.dynamics_classes import SatDynamicsModel, StaticTerrestrial
from clockpunch_b200.environment.env import SSAScheduler
from clockpunch_b200.estimation.ez_ukf import ezUKF
from clockpunch_b200.simulation.catalog import synth_catalog
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
M = int(sys.argv[2]) if len(sys.argv) > 2 else 100
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
cat = synth_catalog(N, M, seed=20240627)


class P:
    time_step, horizon = 100.0, 10**9


t0 = time.perf_counter()
agents = [Sensor(StaticTerrestrial(), cat.sensors[:, j].copy(), agent_id=f"s{j}") for j in range(M)]
for i in range(N):
    f = ezUKF({"x_init": cat.est_x[:, i].copy(), "p_init": cat.est_p[i], "dynamics_type": "satellite", "Q": cat.Q, "R": cat.R})
    agents.append(Target(SatDynamicsModel(), cat.targets[:, i].copy(), f, agent_id=f"t{i}"))
P.agents = agents
env = SSAScheduler(P, seed=1)
obs, info = env.reset()
print(f"build + reset: {time.perf_counter() - t0:.1f} s")
rng = np.random.default_rng(0)
for k in range(3):
    env.step(rng.integers(0, N + 1, size=M))
import cProfile, pstats
t1 = time.perf_counter()
for k in range(steps):
    env.step(rng.integers(0, N + 1, size=M))
t2 = time.perf_counter()
print(f"SSAScheduler.step at {N} targets x {M} sensors: {1e3 * (t2 - t1) / steps:.2f} ms/step")
pr = cProfile.Profile(); pr.enable()
for k in range(5):
    env.step(rng.integers(0, N + 1, size=M))
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
