"""Oracle: Earth-occlusion visibility (test infrastructure).  PARITY UNPINNED.

The arithmetic lives in the third-party package ``satvis==0.2.4``
(punchclock pyproject.toml:21), which is not vendored under /root/reference and
is not installable offline.  This file restates its published algorithm
(Lawton 1987 visibility function) as used at the reference's call sites:
  common/visibility.py:106      isVis(r_sensor, r_target, RE)
  common/visibility.py:108-110  visibilityFunc(r1=, r2=, RE=, hg=0, tol=1e-6)[0]
and the loop structure / output conventions of calcVisMap
(common/visibility.py:56-116).  The only reference-held known answer is
tests/environment/test_env.py:31-41,132-136 (checked in tests/test_oracle.py).
"""
from __future__ import annotations

import numpy as np

from .constants import RE

DEFAULT_TOL = 1e-13


def visibility_func(r1, r2, RE_body, hg=0.0, tol=DEFAULT_TOL):
    """V = a1 + a2 - phi; a_i = acos(RE'/|r_i|); phi = angle(r1, r2).

    |r| below RE' by less than ``tol`` is clamped to RE'; farther below gives
    acos(>1) = NaN (so the pair is "not visible" in binary mode).
    """
    rp = RE_body + hg
    m1 = float(np.linalg.norm(r1))
    m2 = float(np.linalg.norm(r2))
    if m1 < rp and rp - m1 <= tol:
        m1 = rp
    if m2 < rp and rp - m2 <= tol:
        m2 = rp
    with np.errstate(invalid="ignore"):
        a1 = np.arccos(rp / m1)
        a2 = np.arccos(rp / m2)
        c = float(np.dot(r1, r2)) / (m1 * m2)
        phi = np.arccos(min(1.0, max(-1.0, c)))
    return a1 + a2 - phi, phi, a1, a2


def is_vis(r1, r2, RE_body, hg=0.0, tol=DEFAULT_TOL):
    v = visibility_func(r1, r2, RE_body, hg, tol)[0]
    return bool(v > 0)


def calc_vis_map(sensor_states, target_states, body_radius=None, binary=True):
    """(6,M) sensors, (6,N) targets -> (N,M) int64 / float64 (visibility.py:56-116)."""
    if body_radius is None:
        body_radius = RE
    sensor_states = np.asarray(sensor_states, dtype=float)
    target_states = np.asarray(target_states, dtype=float)
    if sensor_states.shape[0] != 6:
        raise ValueError("Bad input: sensor_states must be (6, M)")
    if target_states.shape[0] != 6:
        raise ValueError("Bad input: target_states must be (6, M)")
    if sensor_states.ndim == 1:
        sensor_states = sensor_states.reshape(6, 1)
    if target_states.ndim == 1:
        target_states = target_states.reshape(6, 1)
    M, N = sensor_states.shape[1], target_states.shape[1]
    vm = np.zeros((N, M))
    for col in range(M):
        rs = sensor_states[:3, col]
        for row in range(N):
            rt = target_states[:3, row]
            if binary:
                vm[row, col] = is_vis(rs, rt, body_radius)
            else:
                vm[row, col] = visibility_func(rs, rt, body_radius, hg=0.0, tol=1e-6)[0]
    return vm.astype("int") if binary else vm
