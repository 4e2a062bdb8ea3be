"""Propagator (mirror of punchclock/dynamics/propagator.py:12-81)."""
from __future__ import annotations

from typing import Callable

import numpy as np
from numpy import ndarray

from .. import _lib
from .dynamics import satelliteDynamics


def propagate_batch(x0: ndarray, t0: float, tf: float, kind=None, substeps: int = 0,
                    force_model: int = _lib.FORCE_TWOBODY) -> ndarray:
    """(6,K) float64 -> (6,K) through pc_propagate_agents_host (one launch for all K)."""
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    k = x0.shape[1]
    out = np.empty_like(x0)
    kptr = None
    if kind is not None:
        kind = np.ascontiguousarray(kind, dtype=np.uint8)
        kptr = kind.ctypes.data
    ctx = _lib.get_ctx()
    ctx.check(ctx.lib.pc_propagate_agents_host(ctx.handle, x0.ctypes.data, out.ctypes.data, kptr, k, k,
                                               float(t0), float(tf), int(substeps), int(force_model)))
    return out


def simplePropagate(func: Callable, x0: ndarray, t0: float, tf: float) -> ndarray:
    """Propagate (N,) | (N,M) state(s) from t0 to tf; returns the input's (squeezed) shape.

    The built-in two-body ``satelliteDynamics`` is the path the CUDA propagator implements
    (fixed-step DOP853, see csrc/propagate.cuh).  An arbitrary Python right-hand side
    (tests/dynamics/test_propagator.py:10-68 integrates a toy linear ODE) is a Python callback
    per RHS evaluation, i.e. host work by nature: it is handed to the reference's own
    simplePropagate when the reference tree is available (clockpunch_b200._reference), and
    raises NotImplementedError otherwise.  The two-body path has no CPU fallback.

    Raises ValueError if t0 == tf (propagator.py:40-41).
    """
    if t0 == tf:
        raise ValueError("Woah buddy, t0 and tf are equal.")
    if func is not satelliteDynamics:
        from .. import _reference
        return _reference.load("dynamics.propagator").simplePropagate(func, x0, t0, tf)
    x0 = np.asarray(x0, dtype=np.float64).squeeze()
    if x0.shape[0] != 6:
        raise ValueError("state vectors must have 6 components")
    return propagate_batch(x0.reshape(6, -1), t0, tf).reshape(x0.shape)
