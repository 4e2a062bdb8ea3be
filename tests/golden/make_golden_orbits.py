"""Golden radial rates from the LIVE reference (punchclock/common/orbits.py:39-69), used by
calcVisMapAndDerivative (common/visibility.py:119-180).  Build container only:

    python tests/golden/make_golden_orbits.py      # writes tests/golden/orbits.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_stubs  # noqa: E402,F401

from punchclock.common.orbits import getRadialRate  # noqa: E402
from punchclock.common.transforms import coe2eci, ecef2eci, lla2ecef  # noqa: E402

RE = 6378.1363
rng = np.random.default_rng(20240701)
states = []
for _ in range(48):
    a = rng.uniform(RE + 400, 43000)
    e = rng.uniform(0, 0.7)
    e = min(e, 1 - (RE + 300) / a)
    states.append(coe2eci(a, e, rng.uniform(0, np.pi), *rng.uniform(0, 2 * np.pi, 3)))
for _ in range(8):      # ground sites (Earth-fixed: r . v = 0)
    lla = np.array([rng.uniform(-1.4, 1.4), rng.uniform(-np.pi, np.pi), 1e-6])
    states.append(ecef2eci(lla2ecef(lla), rng.uniform(0, 86400)))
X = np.array(states).T
rdot = np.array([getRadialRate(X[:3, i], X[3:, i]) for i in range(X.shape[1])])
np.savez_compressed(os.path.join(HERE, "orbits.npz"), x=X, rdot=rdot)
print("wrote orbits.npz", X.shape, "max |rdot - r.v/|r|| =",
      np.abs(rdot - (X[:3] * X[3:]).sum(0) / np.linalg.norm(X[:3], axis=0)).max())
