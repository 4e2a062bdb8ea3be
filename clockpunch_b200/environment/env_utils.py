"""Environment utilities (mirror of punchclock/environment/env_utils.py:22-127)."""
from __future__ import annotations

from numpy import asarray, ndarray

from ..common.agents import Agent, Sensor, Target
from ..common.constants import getConstants
from ..common.visibility import calcVisMap


def _getTrueStates(list_of_agents: list[Agent]) -> ndarray:
    return asarray([agent.eci_state for agent in list_of_agents]).squeeze().T


def _getEstimatedStates(list_of_agents: list[Agent]) -> ndarray:
    return asarray([agent.target_filter.est_x for agent in list_of_agents]).squeeze().T


def getVisMapEstOrTruth(list_of_agents: list[Agent], truth_flag: bool, binary: bool = True) -> ndarray:
    """(N, M) visibility map from true sensor states and true (truth_flag) or estimated target
    states (env_utils.py:22-98); one CUDA launch for the whole map."""
    sensors = [a for a in list_of_agents if isinstance(a, Sensor)]
    targets = [a for a in list_of_agents if isinstance(a, Target)]
    times = [a.time for a in sensors] + [a.time if truth_flag is True else a.target_filter.time for a in targets]
    assert all(t == times[0] for t in times), "Sensor and target times are not equal. If truth_flag == False, \
        target target_filter times are not equal to sensor times."
    sensor_states = _getTrueStates(sensors)
    target_states = _getTrueStates(targets) if truth_flag is True else _getEstimatedStates(targets)
    return calcVisMap(sensor_states=sensor_states, target_states=target_states,
                      body_radius=getConstants()["earth_radius"], binary=binary)


def buildAgentInfoMap(agents: list[Agent]) -> list[dict]:
    """[{"agent_type": "Sensor" | "Target", "id": agent_id}, ...] in agent order
    (env_utils.py:183-208)."""
    agent_map = []
    for agent in agents:
        entry = {}
        if isinstance(agent, Sensor):
            entry["agent_type"] = "Sensor"
        elif isinstance(agent, Target):
            entry["agent_type"] = "Target"
        entry["id"] = agent.agent_id
        agent_map.append(entry)
    return agent_map


def forecastVisMap(agents: list[Agent], time_step: float, estimate: bool = False) -> ndarray:
    """Binary vis map `time_step` ahead of the agents' current time (env_utils.py:130-180): sensors
    and (estimate=False) targets propagated, or (estimate=True) the filters' predicted means.
    All sensors go through one propagation launch, all targets through one propagation or one
    batched UKF predict, and the map is one visibility launch; the agents are not modified."""
    import numpy as np
    from .. import _lib
    from ..dynamics.dynamics_classes import StaticTerrestrial
    from ..dynamics.propagator import propagate_batch
    from ..estimation.ukf_v2 import ukf_batch_host
    times = [ag.time for ag in agents]
    assert all(t == times[0] for t in times)
    t0, t1 = times[0], times[0] + time_step
    sensors = [a for a in agents if isinstance(a, Sensor)]
    targets = [a for a in agents if isinstance(a, Target)]
    col = lambda xs: np.stack([np.asarray(x, dtype=np.float64).reshape(6) for x in xs], axis=1)
    kinds = np.array([_lib.KIND_TERRESTRIAL if isinstance(a.dynamics, StaticTerrestrial) else _lib.KIND_SATELLITE
                      for a in sensors], dtype=np.uint8)
    x_sensors = propagate_batch(col([a.eci_state for a in sensors]), t0, t1, kind=kinds)
    if estimate is True:
        f0 = targets[0].target_filter
        out = ukf_batch_host(col([a.target_filter.est_x for a in targets]),
                             np.stack([np.asarray(a.target_filter.est_p, dtype=np.float64) for a in targets]),
                             None, None, t0, t1, _lib.make_ukf_params(f0.q_matrix, f0.r_matrix))
        x_targets = out["pred_x"]
    else:
        x_targets = propagate_batch(col([a.eci_state for a in targets]), t0, t1)
    return calcVisMap(sensor_states=x_sensors, target_states=x_targets,
                      body_radius=getConstants()["earth_radius"])
