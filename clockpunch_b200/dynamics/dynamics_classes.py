"""Dynamics model classes (mirror of punchclock/dynamics/dynamics_classes.py:20-121)."""
from __future__ import annotations

from abc import ABC, abstractmethod
from functools import partial
from typing import Callable

import numpy as np
from numpy import ndarray

from .. import _lib
from .dynamics import satelliteDynamics, terrestrialDynamicsAlt
from .propagator import propagate_batch, simplePropagate


class DynamicsModel(ABC):
    """Abstract base class for dynamics models (dynamics_classes.py:20-48)."""

    def __init__(self, dynamics_func: Callable[[ndarray, float, float], ndarray]):
        self._dynamics_func = dynamics_func

    @abstractmethod
    def propagate(self, start_state: ndarray, start_time: float, end_time: float, **kwargs) -> ndarray:
        return self._dynamics_func(start_state, start_time, end_time, **kwargs)


class SatDynamicsModel(DynamicsModel):
    """Earth satellite two-body dynamics (dynamics_classes.py:51-76) on the GPU."""

    kind = _lib.KIND_SATELLITE

    def __init__(self):
        super().__init__(partial(simplePropagate, satelliteDynamics))

    def propagate(self, start_state: ndarray, start_time: float, end_time: float, **kwargs) -> ndarray:
        return super().propagate(start_state, start_time, end_time, **kwargs)


class StaticTerrestrial(DynamicsModel):
    """Earth-fixed agent (dynamics_classes.py:79-121): all columns in one launch."""

    kind = _lib.KIND_TERRESTRIAL

    def __init__(self):
        super().__init__(terrestrialDynamicsAlt)

    def propagate(self, start_state: ndarray, start_time: float, end_time: float, **kwargs) -> ndarray:
        x = np.asarray(start_state, dtype=np.float64)
        if x.ndim == 1:
            x = x.reshape(6, 1)
        if start_time == end_time:      # the reference's rotation by zero angle
            return x.copy().squeeze()
        kinds = np.full(x.shape[1], _lib.KIND_TERRESTRIAL, dtype=np.uint8)
        return propagate_batch(x, start_time, end_time, kind=kinds).squeeze()
