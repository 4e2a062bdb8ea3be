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


def radial_rate(r_vec, v_vec, mu=398600.0):
    """d|r|/dt through the orbit-element chain of the reference (common/orbits.py:39-69:
    eccentricity vector -> true anomaly -> sma -> mean motion -> true-anomaly rate).  Pinned against
    the live reference in tests/golden/orbits.npz; equals r.v/|r| to 5e-14 km/s on that set."""
    r_vec, v_vec = np.asarray(r_vec, float), np.asarray(v_vec, float)
    r, v = np.linalg.norm(r_vec), np.linalg.norm(v_vec)
    e_vec = ((v**2 - mu / r) * r_vec - np.vdot(r_vec, v_vec) * v_vec) / mu      # orbits.py:128-156
    e = np.linalg.norm(e_vec)
    if not abs(e) < 1e-12:
        e_vec = e_vec / e
    c = np.dot(e_vec, r_vec) / r
    ta = np.arccos(min(1.0, max(-1.0, c)))                                         # safeArccos
    if np.dot(r_vec, v_vec) < 0.0:
        ta = 2 * np.pi - ta                                                        # fixAngleQuadrant
    sma = -0.5 * mu / (0.5 * v**2 - mu / r)
    n = np.sqrt(mu / sma**3)
    ta_dot = (n * sma**2 / r**2) * np.sqrt(1 - e**2)
    return r * ta_dot * e * np.sin(ta) / (1 + e * np.cos(ta))


def calc_vis_and_der_vis(r1, r1dot, r1mag_dot, r2, r2dot, r2mag_dot, RE_body, hg=0.0, tol=1e-6):
    """(V, dV/dt) of satvis.visibility_func.calcVisAndDerVis as called at
    common/visibility.py:166-174 (PARITY UNPINNED: restated from the published formula):
        V = a1 + a2 - phi,   a_i = acos(RE'/|r_i|),   phi = acos(r1.r2 / (|r1||r2|))
        da_i/dt = RE' |r_i|' / (|r_i|^2 sqrt(1 - (RE'/|r_i|)^2))
        dphi/dt = -[ (r1'.r2 + r1.r2')/(|r1||r2|) - cos(phi) (|r1|'/|r1| + |r2|'/|r2|) ] / sin(phi)
    Aligned position vectors (sin(phi) = 0) give +-Inf (or NaN for 0/0), as the reference documents
    (visibility.py:141-143)."""
    r1, r2, r1dot, r2dot = (np.asarray(a, float) for a in (r1, r2, r1dot, r2dot))
    v, phi, a1, a2 = visibility_func(r1, r2, RE_body, hg, tol)
    rp = RE_body + hg
    m1, m2 = np.linalg.norm(r1), np.linalg.norm(r2)
    with np.errstate(divide="ignore", invalid="ignore"):
        a1dot = rp * r1mag_dot / (m1**2 * np.sqrt(1 - (rp / m1) ** 2))
        a2dot = rp * r2mag_dot / (m2**2 * np.sqrt(1 - (rp / m2) ** 2))
        c = np.dot(r1, r2) / (m1 * m2)
        cdot = (np.dot(r1dot, r2) + np.dot(r1, r2dot)) / (m1 * m2) - c * (r1mag_dot / m1 + r2mag_dot / m2)
        phidot = -np.float64(cdot) / np.sqrt(np.float64(1 - c**2))
    return v, a1dot + a2dot - phidot


def calc_vis_map_and_derivative(sensor_states, target_states, body_radius=None, binary=True):
    """(N, M) visibility map and its time derivative (common/visibility.py:119-180)."""
    if body_radius is None:
        body_radius = RE
    S = np.asarray(sensor_states, float).reshape(6, -1)
    T = np.asarray(target_states, float).reshape(6, -1)
    vm, dm = np.zeros((T.shape[1], S.shape[1])), np.zeros((T.shape[1], S.shape[1]))
    for col in range(S.shape[1]):
        for row in range(T.shape[1]):
            s, t = S[:, col], T[:, row]
            vm[row, col], dm[row, col] = calc_vis_and_der_vis(
                s[:3], s[3:], radial_rate(s[:3], s[3:]), t[:3], t[3:], radial_rate(t[:3], t[3:]), body_radius)
    if binary is True:
        vm = np.where(vm > 0, 1, 0)
    return vm, dm
