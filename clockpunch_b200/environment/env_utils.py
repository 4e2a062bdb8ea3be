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
