"""SSASchedulerParams (mirror of punchclock/environment/env_parameters.py:29-579).

Turns the reference's primitive scenario description -- `horizon`, `time_step`, an `agent_params`
dict (counts, dynamics, fixed or randomly drawn initial conditions) and a `filter_params` dict (Q, R,
p_init) -- into the `.agents` list an `SSAScheduler` is built from.  Same keys, defaults,
validation errors, ID scheme ("S1000", ..., "5000", ...) and random-draw order as the reference:
initial conditions come from `genInitStates(seed)`, the initial estimate noise from numpy's legacy
GLOBAL generator (env_parameters.py:372-377), so `numpy.random.seed(k)` before construction
reproduces the reference's scenario exactly.  Frame conversions of drawn ICs run on the GPU in one
launch per population.
"""
from __future__ import annotations

from copy import deepcopy

from numpy import arange, asarray, diag, diagonal, eye, ndarray, pi, sqrt, zeros
from numpy.random import normal

from ..common.agents import Sensor, Target
from ..common.constants import getConstants
from ..dynamics.dynamics_classes import SatDynamicsModel, StaticTerrestrial
from ..estimation.ez_ukf import ezUKF, getRandomParams
from ..simulation.sim_utils import genInitStates

RE = getConstants()["earth_radius"]

_AGENT_DEFAULTS = {"num_sensors": 1, "num_targets": 1, "sensor_starting_num": 1000, "target_starting_num": 5000,
                   "sensor_dist": None, "target_dist": None, "sensor_dist_frame": None, "target_dist_frame": None,
                   "sensor_dist_params": None, "target_dist_params": None, "fixed_sensors": None,
                   "fixed_targets": None, "init_num_tasked": None, "init_last_time_tasked": None,
                   "sensor_dynamics": "terrestrial", "target_dynamics": "satellite"}

_DEFAULT_DISTS = {       # env_parameters.py:505-531
    "terrestrial": ("uniform", "LLA", [[-pi / 2, pi / 2], [-pi, pi], [1e-6, 1e-6], [0, 0], [0, 0], [0, 0]]),
    "satellite": ("uniform", "COE", [[RE + 400, RE + 1000], [0, 0], [-pi, pi], [0, 2 * pi], [0, 2 * pi], [0, 2 * pi]]),
}


class SSASchedulerParams:
    """Attributes: agent_params, filter_params, horizon, time_step, agents (sensors first)."""

    def __init__(self, horizon: int, agent_params: dict = None, filter_params: dict = None,
                 time_step: float = 100, seed: int = None):
        agent_params = {} if agent_params is None else agent_params
        filter_params = {} if filter_params is None else filter_params
        for key, default in _AGENT_DEFAULTS.items():
            agent_params[key] = agent_params.get(key, default)
        agent_params.update(self.getDefaultAgentDists(agent_params))
        filter_params.update(self.getDefaultFilterConfig(filter_params))
        for who in ("sensor", "target"):
            fixed = agent_params[f"fixed_{who}s"]
            if isinstance(fixed, list) and len(fixed) != agent_params[f"num_{who}s"]:
                print(f"Error: len(agent_params['fixed_{who}s']) != ", f"agent_params['num_{who}s']")
                raise ValueError(f"Input dimensions of 'fixed_{who}s'!='num_{who}s'. If you",
                                 "are using fixed agents, ensure they match.")
            drawn = any(agent_params[f"{who}_{k}"] is not None for k in ("dist", "dist_frame", "dist_params"))
            if drawn and isinstance(fixed, list):
                print(f"Error: Conflicting values set for '{who}_dist*' and 'fixed_{who}s'.")
                raise ValueError(f"Argument(s) was/were provided for stochastically-generated {who}s, but values "
                                 f"were also provided for agent_params['fixed_{who}s'].")
        self.time_step, self.horizon = time_step, horizon
        self.agent_params, self.filter_params = agent_params, filter_params

        ics = {}
        for who in ("sensor", "target"):          # sensors are drawn first, both with the same seed
            if agent_params[f"{who}_dist"] is None:
                ics[who] = [a.reshape((6, 1)) for a in asarray(agent_params[f"fixed_{who}s"])]
            else:
                ics[who] = genInitStates(num_initial_conditions=agent_params[f"num_{who}s"],
                                         dist=agent_params[f"{who}_dist"],
                                         dist_params=agent_params[f"{who}_dist_params"],
                                         frame=agent_params[f"{who}_dist_frame"], seed=seed)
        n = agent_params["num_targets"]
        if agent_params["init_num_tasked"] is None:
            agent_params["init_num_tasked"] = zeros(shape=[n], dtype=int)
        if agent_params["init_last_time_tasked"] is None:
            agent_params["init_last_time_tasked"] = zeros(shape=[n], dtype=float)
        filters = self.genFilters(ics["target"], agent_params["target_dynamics"], filter_params)
        sensors = [Sensor(**p) for p in self.getAgentParams(ics["sensor"], agent_params["sensor_dynamics"],
                                                            agent_params["sensor_starting_num"], "sensor")]
        targets = [Target(**p) for p in self.getAgentParams(
            agent_initial_conditions=ics["target"], dynamics_type=agent_params["target_dynamics"],
            id_start=agent_params["target_starting_num"], sensor_target="target", list_of_filters=filters,
            num_tasked=agent_params["init_num_tasked"], last_time_tasked=agent_params["init_last_time_tasked"])]
        self.agents = sensors[: agent_params["num_sensors"]] + targets[:n]

    def genFilters(self, initial_conditions: list[ndarray], dynamics_type: str, filter_params: dict) -> list:
        """One ezUKF per target, identical Q / R / p_init, estimate = IC + N(0, diag(p_init)) drawn
        from numpy's legacy global RNG, one call per target (env_parameters.py:326-393)."""
        for k in ("Q", "R", "p_init"):
            if isinstance(filter_params[k], list):
                filter_params[k] = asarray(filter_params[k])
        p = filter_params["p_init"]
        if isinstance(p, (float, int)):
            p_mat = p * eye(6)
        elif isinstance(p, dict):
            p_mat = diag(getRandomParams(**p))
        else:
            p_mat = p
        std = sqrt(diagonal(p_mat))
        noisy = [(normal(ic.transpose(), std)).transpose() for ic in initial_conditions]
        return [ezUKF({"x_init": x, "dynamics_type": dynamics_type, "Q": filter_params["Q"],
                       "R": filter_params["R"], "p_init": filter_params["p_init"]}) for x in noisy]

    def getAgentParams(self, agent_initial_conditions: list[ndarray], dynamics_type: str, id_start: int,
                       sensor_target: str, list_of_filters: list = None, num_tasked: list = None,
                       last_time_tasked: list = None) -> list[dict]:
        """Keyword dicts for Sensor(...) / Target(...) (env_parameters.py:395-481)."""
        k = len(agent_initial_conditions)
        model = StaticTerrestrial if dynamics_type == "terrestrial" else SatDynamicsModel
        ids = list(map(str, arange(id_start, id_start + k, 1)))
        if sensor_target == "sensor":
            return [{"dynamics_model": model(), "agent_id": "S" + i, "init_eci_state": x}
                    for i, x in zip(ids, agent_initial_conditions)]
        return [{"dynamics_model": model(), "agent_id": i, "init_eci_state": x, "target_filter": f,
                 "init_num_tasked": nt, "init_last_time_tasked": lt}
                for i, x, f, nt, lt in zip(ids, agent_initial_conditions, list_of_filters, num_tasked,
                                           last_time_tasked)]

    def getDefaultAgentDists(self, config: dict) -> dict:
        """Terrestrial agents default to uniform LLA, satellites to uniform circular LEO COEs; a
        population with a distribution or fixed states is left alone (env_parameters.py:483-559)."""
        new = deepcopy(config)
        for who in ("sensor", "target"):
            if config[f"{who}_dist"] is None and config[f"fixed_{who}s"] is None:
                dist, frame, params = _DEFAULT_DISTS[config[f"{who}_dynamics"]]
                new[f"{who}_dist"], new[f"{who}_dist_frame"], new[f"{who}_dist_params"] = dist, frame, params
        return new

    def getDefaultFilterConfig(self, config: dict) -> dict:
        """Q = 0.1 D, R = D, p_init = 10 D with D = diag(1, 1, 1, 1e-2, 1e-2, 1e-2) where not given
        (env_parameters.py:561-579)."""
        new = deepcopy(config)
        D = diag([1, 1, 1, 1e-2, 1e-2, 1e-2])
        new["Q"] = new.get("Q", 0.1 * D)
        new["R"] = new.get("R", 1 * D)
        new["p_init"] = new.get("p_init", 10 * D)
        return new
