"""Dynamics functions (mirror of punchclock/dynamics/dynamics.py:16-86).

`a2body` and `satelliteDynamics` are the reference's right-hand side as plain numpy, kept because callers
import them by name (`simplePropagate(satelliteDynamics, ...)` uses the function OBJECT as the tag that
selects the CUDA propagator).  They are never evaluated on the step path: the integrator in
csrc/propagate.cuh computes the same expression inside its stages, and nothing in this package calls them."""
from __future__ import annotations

import numpy as np
from numpy import ndarray

from .. import _lib
from ..common.constants import getConstants


def a2body(r_eci: ndarray, mu: float) -> ndarray:
    """Two-body acceleration -mu r / |r|^3 (dynamics.py:71-86).  Kept for API parity:
    the GPU integrator evaluates the same expression inside its stages."""
    r_eci = np.asarray(r_eci, dtype=np.float64)
    return -mu * r_eci / (np.linalg.norm(r_eci) ** 3)


def satelliteDynamics(t, x0: ndarray) -> ndarray:
    """State derivative [v, a2body(r)] (dynamics.py:16-40)."""
    x0 = np.asarray(x0, dtype=np.float64)
    return np.concatenate((x0[3:6], a2body(x0[:3], getConstants()["mu"])))


def terrestrialDynamics(t, x0: ndarray, JD: float) -> ndarray:
    """[6 x T] ECI history of an Earth-fixed point (dynamics.py:48-68): the point is
    taken to ECEF at JD and back to ECI at each t -- one pc_rotate_terrestrial launch each."""
    if type(t) is not list:
        t = [t]
    x0 = np.ascontiguousarray(np.asarray(x0, dtype=np.float64).reshape(6, 1))
    ctx = _lib.get_ctx()
    hist = np.zeros((6, len(t)))
    out = np.empty((6, 1))
    kind = np.array([_lib.KIND_TERRESTRIAL], dtype=np.uint8)
    for i, ti in enumerate(t):
        if ti == JD:
            hist[:, i] = x0[:, 0]
            continue
        ctx.check(ctx.lib.pc_propagate_agents_host(ctx.handle, x0.ctypes.data, out.ctypes.data,
                                                   kind.ctypes.data, 1, 1, float(JD), float(ti), 0,
                                                   _lib.FORCE_TWOBODY))
        hist[:, i] = out[:, 0]
    return hist


def terrestrialDynamicsAlt(x0: ndarray, t0: float, tf: float) -> ndarray:
    """Argument-order shim (dynamics.py:43-45)."""
    return terrestrialDynamics(t=tf, x0=x0, JD=t0)
