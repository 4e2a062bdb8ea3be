"""Coordinate transforms (mirror of punchclock/common/transforms.py:34-292).

Same names, argument meaning, shapes and exceptions as the reference.  The batched
arithmetic (frame rotation + transport theorem, COE -> ECI, LLA -> ECI) runs in
libpunch_b200 kernels (pc_frame_change / pc_coe2eci / pc_lla2eci); the 3x3 rotation
matrix builders are plain array constructors exactly as in the reference.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
from numpy import array, cos, ndarray, pi, sin

from .. import _buffers as B
from .. import _lib
from .constants import getConstants

mu = getConstants()["mu"]
RE = getConstants()["earth_radius"]
OMEGA_EARTH = getConstants()["earth_rotation_rate"]


def _check_dimensions(x: ndarray) -> ndarray:
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x.reshape(6, 1) if x.size == 6 else x.reshape(x.shape[0], 1)
    if x.shape[0] != 6:
        raise ValueError("Argument must be a (6,N) array.")      # transforms.py:138-139
    return x


def _frame_change(x: ndarray, JD: float, to_eci: bool) -> ndarray:
    x = _check_dimensions(x)
    ctx = _lib.get_ctx()
    n = x.shape[1]
    d = B.to_device(x, ctx)
    ctx.check(ctx.lib.pc_frame_change(ctx.handle, d.data_ptr(), d.data_ptr(), n, n, float(JD),
                                      1 if to_eci else 0, B.stream_ptr(ctx)))
    return d.cpu().numpy().squeeze()


def ecef2eci(x_ecef: ndarray, JD: float = 0) -> ndarray:
    """(6,) or (6,N) ECEF -> ECI at JD seconds since the frames were aligned
    (transforms.py:55-90).  (6,1) input returns (6,)."""
    return _frame_change(x_ecef, JD, True)


def eci2ecef(x_eci: ndarray, JD: float) -> ndarray:
    """(6,) or (6,N) ECI -> ECEF (transforms.py:93-129)."""
    return _frame_change(x_eci, JD, False)


def lla2ecef(x_lla: ndarray) -> ndarray:
    """[lat, lon, alt] -> ECEF state with zero velocity (transforms.py:34-52)."""
    ctx = _lib.get_ctx()
    lla = np.asarray(x_lla, dtype=np.float64)[:3].reshape(3, 1)
    d_in = B.to_device(lla, ctx)
    d_out = B.empty((6, 1), ctx)
    # JD = 0: ECI and ECEF coincide in position; ECEF velocity is zero by definition
    ctx.check(ctx.lib.pc_lla2eci(ctx.handle, d_in.data_ptr(), d_out.data_ptr(), 1, 1, 0.0,
                                 B.stream_ptr(ctx)))
    out = d_out.cpu().numpy().reshape(6)
    out[3:] = 0.0
    return out


def lla2eci_batch(lla: ndarray, JD: float = 0.0) -> ndarray:
    """(3,N) [lat; lon; alt] -> (6,N) ECI = ecef2eci(lla2ecef(.), JD), the ground-agent
    initial-condition path of genInitStates (sim_utils.py:96-101), in one launch."""
    lla = np.ascontiguousarray(np.asarray(lla, dtype=np.float64)[:3])
    ctx = _lib.get_ctx()
    n = lla.shape[1]
    d_in = B.to_device(lla, ctx)
    d_out = B.empty((6, n), ctx)
    ctx.check(ctx.lib.pc_lla2eci(ctx.handle, d_in.data_ptr(), d_out.data_ptr(), n, n, float(JD),
                                 B.stream_ptr(ctx)))
    return d_out.cpu().numpy()


def rot1(theta: float) -> ndarray:
    return array([[1, 0, 0], [0, cos(theta), -sin(theta)], [0, sin(theta), cos(theta)]])


def rot2(theta: float) -> ndarray:
    return array([[cos(theta), 0, sin(theta)], [0, 1, 0], [-sin(theta), 0, cos(theta)]])


def rot3(theta: float) -> ndarray:
    return array([[cos(theta), -sin(theta), 0], [sin(theta), cos(theta), 0], [0, 0, 1]])


def _check_coe(sma, ecc, inc, raan, argp, true_anom):
    """Range checks of the reference (transforms.py:268-279), vectorised."""
    sma, ecc, inc, raan, argp, true_anom = (np.asarray(v, dtype=np.float64) for v in
                                            (sma, ecc, inc, raan, argp, true_anom))
    if np.any(sma < RE):
        raise ValueError("Semi-major axis is less than Earth radius.")
    if np.any(ecc < 0) or np.any(ecc >= 1):
        raise ValueError("Eccentricity is out of range 0<=ecc<1.")
    if np.any(inc < 0) or np.any(inc > pi):
        raise ValueError("Inclination is out of range 0<=inc<=pi.")
    if np.any(raan < 0) or np.any(raan >= 2 * pi):
        raise ValueError("RAAN is out of range 0<=raan<2*pi.")
    if np.any(argp < 0) or np.any(argp >= 2 * pi):
        raise ValueError("Argument of perigee is out of range 0<=argp<2*pi.")
    if np.any(true_anom < 0) or np.any(true_anom >= 2 * pi):
        raise ValueError("True anomaly is out of range 0<=true_anom<2*pi.")


def coe2eci_batch(coe: ndarray, mu: float = mu) -> ndarray:
    """(6,N) rows [sma, ecc, inc, raan, argp, true_anom] -> (6,N) ECI in one launch."""
    coe = np.ascontiguousarray(coe, dtype=np.float64)
    if coe.ndim != 2 or coe.shape[0] != 6:
        raise ValueError("coe must be (6, N)")
    _check_coe(*coe)
    ctx = _lib.get_ctx()
    n = coe.shape[1]
    d_in = B.to_device(coe, ctx)
    d_out = B.empty((6, n), ctx)
    ctx.check(ctx.lib.pc_coe2eci(ctx.handle, d_in.data_ptr(), d_out.data_ptr(), n, n, float(mu),
                                 B.stream_ptr(ctx)))
    return d_out.cpu().numpy()


def coe2eci(sma: float, ecc: float, inc: float, raan: float, argp: float, true_anom: float,
            mu: Optional[float] = mu) -> ndarray:
    """Classical orbital elements -> (6,) ECI state (transforms.py:240-292)."""
    coe = np.array([[sma], [ecc], [inc], [raan], [argp], [true_anom]], dtype=np.float64)
    return coe2eci_batch(coe, mu)[:, 0]
