"""Agents (mirror of punchclock/common/agents.py:31-322).

Same classes, constructor arguments and attributes as the reference: an Agent owns its ECI
state, time and dynamics model; a Target additionally owns its filter, tasking counter and
last-tasked time.  Every physics call (propagate, predictAndUpdate) goes through the CUDA
library; these per-object classes therefore cost one launch per call and exist for drop-in
compatibility with reference code that walks `env.agents`.  The batched path is
clockpunch_b200.engine.CatalogEngine, and clockpunch_b200.environment builds one from a
list of these objects.
"""
from __future__ import annotations

from copy import copy
from typing import Any

from numpy import concatenate, cross, eye
from numpy import float32 as npfloat32
from numpy import int64 as npint64
from numpy import ndarray, pi, sqrt
from numpy.linalg import norm
from numpy.random import default_rng, multivariate_normal

from ..dynamics.dynamics_classes import DynamicsModel, SatDynamicsModel, StaticTerrestrial
from ..estimation.ez_ukf import ezUKF
from ..estimation.filter_base_class import Filter
from .constants import getConstants
from .transforms import ecef2eci, lla2ecef


class Agent:
    """Generic agent (agents.py:31-96)."""

    def __init__(self, dynamics_model: DynamicsModel, init_eci_state: ndarray, agent_id: Any = None,
                 time: float | int = 0):
        assert isinstance(init_eci_state, ndarray), "init_eci_state must be a ndarray"
        self.dynamics = dynamics_model
        self.agent_id = agent_id
        self.time = time
        self.eci_state = init_eci_state.reshape([init_eci_state.size, 1])

    def propagate(self, time_next: float):
        """Update ECI state and time (agents.py:59-71).  State is (6, 1) before the first call
        and (6,) afterwards, as in the reference."""
        self.eci_state = self.dynamics.propagate(start_state=self.eci_state, start_time=self.time,
                                                 end_time=time_next)
        self.time = time_next

    def toDict(self) -> dict:
        json_dict = {}
        for attr, attr_val in self.__dict__.items():
            if attr in ["dynamics", "target_filter"]:
                json_dict[attr] = str(attr_val.__class__)
            elif isinstance(attr_val, ndarray):
                json_dict[attr] = attr_val.tolist()
            elif isinstance(attr_val, npfloat32):
                json_dict[attr] = float(attr_val)
            elif isinstance(attr_val, npint64):
                json_dict[attr] = int(attr_val)
            else:
                json_dict[attr] = attr_val
        return json_dict


class Target(Agent):
    """Agent with a filter (agents.py:99-183)."""

    def __init__(self, dynamics_model: DynamicsModel, init_eci_state: ndarray, target_filter: Filter,
                 agent_id: Any = None, time: float | int = 0, init_num_tasked: int = 0,
                 init_last_time_tasked: float = 0):
        super().__init__(dynamics_model=dynamics_model, init_eci_state=init_eci_state, agent_id=agent_id,
                         time=time)
        assert self.time == target_filter.time, "Target time != filter time."
        self.target_filter = target_filter
        self.num_tasked = init_num_tasked
        self.last_time_tasked = npfloat32(init_last_time_tasked)

    def updateNonPhysical(self, task: bool):
        """agents.py:144-171: measurement if tasked, then predictAndUpdate to self.time."""
        if task is True:
            self.num_tasked = self.num_tasked + 1
            measurement = self.getMeasurement()
        elif task is False:
            measurement = None
        self.target_filter.predictAndUpdate(final_time=self.time, measurement=measurement)
        if self.target_filter.last_measurement_time is not None:
            self.last_time_tasked = npfloat32(self.target_filter.last_measurement_time)

    def getMeasurement(self) -> ndarray:
        """z ~ N(truth, R) from numpy's legacy global RNG, exactly as agents.py:173-183."""
        true_state = self.eci_state.squeeze()
        return multivariate_normal(true_state, self.target_filter.r_matrix)


class Sensor(Agent):
    """agents.py:189-216."""


def buildRandomAgent(agent_type: str = "agent", dynamics_model: str = "satellite",
                     init_eci_state: ndarray = None, id: str = None, time: float | int = 0):  # noqa: A002
    """Quick-use random agent (agents.py:219-286)."""
    rng = default_rng()
    dynamics_model_tag = copy(dynamics_model)
    if dynamics_model == "satellite":
        dynamics_model = SatDynamicsModel()
    elif dynamics_model == "terrestrial":
        dynamics_model = StaticTerrestrial()
    if init_eci_state is None:
        init_eci_state = getRandomIC(dynamics_model_tag, time)
    if agent_type == "agent":
        return Agent(dynamics_model=dynamics_model, init_eci_state=init_eci_state, agent_id=id, time=time)
    if agent_type == "sensor":
        return Sensor(dynamics_model=dynamics_model, init_eci_state=init_eci_state, agent_id=id, time=time)
    if agent_type == "target":
        ukf = ezUKF({"time": time, "x_init": init_eci_state, "p_init": rng.normal() * eye(6),
                     "dynamics_type": dynamics_model_tag, "Q": rng.normal() * eye(6), "R": rng.normal() * eye(6)})
        return Target(dynamics_model=dynamics_model, init_eci_state=init_eci_state, target_filter=ukf,
                      agent_id=id, time=time)
    raise ValueError("agent_type must be 'agent', 'sensor' or 'target'")


def getRandomIC(satellite_terrestrial: str = "satellite", time: float = 0) -> ndarray:
    """Random circular-orbit or ground IC (agents.py:289-322)."""
    rng = default_rng()
    RE = getConstants()["earth_radius"]
    if satellite_terrestrial == "satellite":
        lla = rng.uniform([-pi / 2, -pi, RE + 400], [pi / 2, pi, RE + 36000])
        r = lla2ecef(lla)[:3]
        v_mag = sqrt(getConstants()["mu"] / norm(r))          # getCircOrbitVel, orbits.py:26-36
        a_hat = r / norm(r)                                     # normalVec, math.py:192-210 (its own generator)
        rng_b = default_rng()
        b = rng_b.uniform(size=3)
        b_hat = b / norm(b)
        while a_hat @ b_hat == 1:
            b = rng_b.uniform(size=3)
            b_hat = b / norm(b)
        return concatenate([r, v_mag * cross(a_hat, b_hat)])
    if satellite_terrestrial == "terrestrial":
        lla = rng.uniform([-pi / 2, -pi, 0], [pi / 2, pi, 0])
        return ecef2eci(lla2ecef(lla), time)
    raise ValueError("satellite_terrestrial must be 'satellite' or 'terrestrial'")
