"""Oracle: two-body propagation and frame transforms (test infrastructure).

Follows punchclock/dynamics/dynamics.py:16-86, dynamics/propagator.py:12-81,
dynamics/dynamics_classes.py:51-121 and common/transforms.py:34-292.
"""
from __future__ import annotations

import numpy as np
from scipy.integrate import solve_ivp

from .constants import MU, OMEGA_EARTH, RE


def two_body_rhs(t, x):
    """xdot = [v, -mu r/|r|^3]   (dynamics.py:16-40 + a2body :71-86)."""
    r = x[:3]
    rmag = np.linalg.norm(r)
    return np.concatenate((x[3:6], -MU * r / rmag**3))


def propagate_satellite(x0, t0, tf):
    """One DOP853 solve per column, dense output at tf (propagator.py:12-81).

    Accepts (6,), (6,1) or (6,K); returns the squeezed shape like the reference.
    """
    if t0 == tf:
        raise ValueError("Woah buddy, t0 and tf are equal.")  # propagator.py:40-41
    x0 = np.asarray(x0, dtype=float).squeeze()
    cols = x0.reshape(6, -1)
    out = np.zeros_like(cols)
    for k in range(cols.shape[1]):
        sol = solve_ivp(two_body_rhs, [t0, tf], cols[:, k], t_eval=[tf],
                        method="DOP853", atol=1e-10, rtol=1e-10)  # propagator.py:60-68
        if sol.y.size == 0:
            raise ValueError("solve_ivp did not converge, likely due to a bad initial condition.")
        out[:, k] = sol.y[:, 0]
    return out.reshape(x0.shape)


def _rot3(theta):
    c, s = np.cos(theta), np.sin(theta)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])  # transforms.py:226-237


def _frame_change(x, R, omega):
    """r1 = R r0 ; v1 = R v0 + omega x (R r0)  (transforms.py:143-178)."""
    r1 = R @ x[:3]
    v1 = R @ x[3:] + np.cross(omega, r1)
    return np.concatenate((r1, v1))


def _as_6n(x):
    x = np.asarray(x, dtype=float)
    if x.ndim == 1:
        x = x.reshape(6, 1)
    if x.shape[0] != 6:
        raise ValueError("Argument must be a (6,N) array.")  # transforms.py:138-139
    return x


def ecef2eci(x_ecef, JD=0.0):
    """transforms.py:55-90."""
    x = _as_6n(x_ecef)
    R = _rot3(OMEGA_EARTH * JD)
    w = np.array([0.0, 0.0, OMEGA_EARTH])
    out = np.stack([_frame_change(x[:, i], R, w) for i in range(x.shape[1])], axis=1)
    return out.squeeze()


def eci2ecef(x_eci, JD):
    """transforms.py:93-129."""
    x = _as_6n(x_eci)
    R = _rot3(-OMEGA_EARTH * JD)
    w = np.array([0.0, 0.0, -OMEGA_EARTH])
    out = np.stack([_frame_change(x[:, i], R, w) for i in range(x.shape[1])], axis=1)
    return out.squeeze()


def propagate_terrestrial(x0, t0, tf):
    """Earth-fixed agent: eci2ecef at t0 then ecef2eci at tf, column by column
    (dynamics_classes.py:87-121, dynamics.py:43-68)."""
    x0 = np.asarray(x0, dtype=float)
    cols = x0.reshape(6, -1) if x0.ndim == 1 else x0
    out = np.zeros(cols.shape)
    for i in range(cols.shape[1]):
        out[:, i] = ecef2eci(eci2ecef(cols[:, i], t0), tf)
    return out.squeeze()


def lla2ecef(x_lla):
    """Spherical Earth, zero ECEF velocity (transforms.py:34-52)."""
    lat, lon, alt = x_lla[0], x_lla[1], x_lla[2]
    rd = (RE + alt) * np.cos(lat)
    return np.array([rd * np.cos(lon), rd * np.sin(lon), (RE + alt) * np.sin(lat), 0.0, 0.0, 0.0])


def coe2eci(sma, ecc, inc, raan, argp, true_anom, mu=MU):
    """COE -> PQW -> ECI with the reference's range checks (transforms.py:240-292)."""
    two_pi = 2 * np.pi
    if sma < RE:
        raise ValueError("Semi-major axis is less than Earth radius.")
    if ecc < 0 or ecc >= 1:
        raise ValueError("Eccentricity is out of range 0<=ecc<1.")
    if inc < 0 or inc > np.pi:
        raise ValueError("Inclination is out of range 0<=inc<=pi.")
    if raan < 0 or raan >= two_pi:
        raise ValueError("RAAN is out of range 0<=raan<2*pi.")
    if argp < 0 or argp >= two_pi:
        raise ValueError("Argument of perigee is out of range 0<=argp<2*pi.")
    if true_anom < 0 or true_anom >= two_pi:
        raise ValueError("True anomaly is out of range 0<=true_anom<2*pi.")
    ca, sa = np.cos(true_anom), np.sin(true_anom)
    p = sma * (1.0 - ecc**2)
    r_pqw = p / (1.0 + ecc * ca) * np.array([ca, sa, 0.0])
    v_pqw = np.sqrt(mu / p) * np.array([-sa, ecc + ca, 0.0])

    def rot1(th):
        c, s = np.cos(th), np.sin(th)
        return np.array([[1.0, 0.0, 0.0], [0.0, c, -s], [0.0, s, c]])

    T = _rot3(-raan).dot(rot1(-inc).dot(_rot3(-argp)))
    return np.concatenate([T.dot(r_pqw), T.dot(v_pqw)])
