"""Visibility map (mirror of punchclock/common/visibility.py:17-116)."""
from __future__ import annotations

import numpy as np
from numpy import ndarray

from .. import _lib
from .constants import getConstants

RE = getConstants()["earth_radius"]
# |r| may fall below the body radius by this much and still count as "on the surface";
# farther below is never visible (NaN in continuous mode).  satvis==0.2.4's own default is
# not verifiable offline (parity unpinned); the continuous path uses the 1e-6
# the reference passes explicitly (visibility.py:108-110).
BINARY_TOL = 1e-13
CONTINUOUS_TOL = 1e-6


def _prepVisMapInputs(sensor_states: ndarray, target_states: ndarray, body_radius: float = None):
    """Shape checks and defaults (visibility.py:17-53)."""
    if body_radius is None:
        body_radius = RE
    sensor_states = np.asarray(sensor_states)
    target_states = np.asarray(target_states)
    if sensor_states.shape[0] != 6:
        raise ValueError("Bad input: sensor_states must be (6, M)")
    if target_states.shape[0] != 6:
        raise ValueError("Bad input: target_states must be (6, M)")
    if sensor_states.ndim == 1:
        sensor_states = sensor_states.reshape((6, 1))
    if target_states.ndim == 1:
        target_states = target_states.reshape((6, 1))
    return body_radius, sensor_states.shape[1], target_states.shape[1], sensor_states, target_states


def calcVisMap(sensor_states: ndarray, target_states: ndarray, body_radius: float = None,
               binary: bool = True) -> ndarray:
    """(6,M) sensors x (6,N) targets -> (N,M) visibility map (visibility.py:56-116).

    binary=True  -> int64 ones/zeros;  binary=False -> float64 visibility-function values
    (> 0 means visible; NaN when an agent is below the surface).
    """
    body_radius, M, N, sens, targ = _prepVisMapInputs(sensor_states, target_states, body_radius)
    sens = np.ascontiguousarray(sens, dtype=np.float64)
    targ = np.ascontiguousarray(targ, dtype=np.float64)
    ctx = _lib.get_ctx()
    if binary is True:
        out = np.zeros((N, M), dtype=np.int64)
        ctx.check(ctx.lib.pc_vismap_host(ctx.handle, sens.ctypes.data, M, targ.ctypes.data, N,
                                         float(body_radius), BINARY_TOL, 1, out.ctypes.data, None))
        return out.astype("int")
    out = np.zeros((N, M), dtype=np.float64)
    ctx.check(ctx.lib.pc_vismap_host(ctx.handle, sens.ctypes.data, M, targ.ctypes.data, N,
                                     float(body_radius), CONTINUOUS_TOL, 0, None, out.ctypes.data))
    return out


def calcVisMapAndDerivative(sensor_states: ndarray, target_states: ndarray, body_radius: float = None,
                            binary=True) -> tuple[ndarray, ndarray]:
    """(N, M) visibility map and (N, M) time derivative of the visibility function
    (visibility.py:119-180).  binary=True turns the map (not the derivative) into 0/1.  Aligned
    position vectors give +-Inf / NaN derivatives, as in the reference."""
    body_radius, M, N, sens, targ = _prepVisMapInputs(sensor_states, target_states, body_radius)
    sens = np.ascontiguousarray(sens, dtype=np.float64)
    targ = np.ascontiguousarray(targ, dtype=np.float64)
    vis, der = np.zeros((N, M)), np.zeros((N, M))
    ctx = _lib.get_ctx()
    ctx.check(ctx.lib.pc_vismap_der_host(ctx.handle, sens.ctypes.data, M, targ.ctypes.data, N,
                                         float(body_radius), CONTINUOUS_TOL, vis.ctypes.data, der.ctypes.data))
    if binary is True:
        vis = np.where(vis > 0, 1, 0)
    return vis, der


def unpack_vis_words(words: ndarray, M: int) -> ndarray:
    """(N, ceil(M/32)) uint32 packed map -> (N, M) int64, host-side view helper."""
    words = np.ascontiguousarray(words, dtype=np.uint32)
    out = np.empty((words.shape[0], int(M)), dtype=np.int64)
    rc = _lib.load_library().pc_unpack_vis_host(words.ctypes.data, words.shape[0], words.shape[1], int(M), out.ctypes.data)
    if rc:
        raise ValueError(f"packed map of shape {words.shape} does not hold {M} sensors per row")
    return out
