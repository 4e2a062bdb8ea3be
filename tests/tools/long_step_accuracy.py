"""Who is off at long steps: the oracle (scipy DOP853 at the reference's rtol = atol = 1e-10) or the kernel (fixed-step DOP853,
substep rule)?  Both against a 3e-14-tolerance solution: python tests/tools/long_step_accuracy.py"""
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from scipy.integrate import solve_ivp
from clockpunch_b200.dynamics.propagator import propagate_batch
from clockpunch_b200.simulation.catalog import synth_catalog
from oracle.dynamics import propagate_satellite, two_body_rhs
cat = synth_catalog(257, 32, seed=556035638)
x = np.concatenate([cat.targets, cat.est_x], axis=1)
for dt in (100.0, 300.0, 600.0):
    gpu = propagate_batch(x, 0.0, dt)
    eo, eg = [], []
    for i in range(0, x.shape[1], 8):
        tight = solve_ivp(two_body_rhs, [0, dt], x[:, i], method="DOP853", rtol=3e-14, atol=1e-14).y[:, -1]
        orc = propagate_satellite(x[:, i], 0.0, dt)
        eo.append(np.abs(orc - tight)[:3].max()); eg.append(np.abs(gpu[:, i] - tight)[:3].max())
    print(f"dt={dt}: oracle (rtol=atol=1e-10) max err {max(eo):.2e} km, kernel max err {max(eg):.2e} km (vs rtol 3e-14 solution, {len(eo)} states)")
