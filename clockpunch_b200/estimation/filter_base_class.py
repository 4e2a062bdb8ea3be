"""Filter base class (mirror of punchclock/estimation/filter_base_class.py:13-37)."""
from __future__ import annotations

from abc import ABCMeta, abstractmethod

from numpy import ndarray


class Filter(metaclass=ABCMeta):
    """Abstract base class for a generic filter object."""

    def __init__(self, time: float, est_x: ndarray, est_p: ndarray):
        self.time = time
        self.est_x = est_x
        self.est_p = est_p

    @abstractmethod
    def reset():
        """Resets filter."""
