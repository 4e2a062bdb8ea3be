"""Shortcut for building a UKF (mirror of punchclock/estimation/ez_ukf.py:17-119)."""
from __future__ import annotations

from numpy import diag, eye, ndarray
from numpy.random import default_rng

from ..dynamics.dynamics_classes import SatDynamicsModel, StaticTerrestrial
from .ukf_v2 import UnscentedKalmanFilter


def ezUKF(params: dict) -> UnscentedKalmanFilter:
    """6-D fully observable UKF from {"x_init","p_init","dynamics_type","Q","R"[, "time"]}.

    "Q", "R", "p_init": float/int -> v*I6; (6,6) ndarray as is; dict {"dist","params","seed"}
    -> diag of random draws (ez_ukf.py:62-80).
    """
    time = params.get("time", 0)
    if params["dynamics_type"] == "terrestrial":
        dynamics_model = StaticTerrestrial()
    elif params["dynamics_type"] == "satellite":
        dynamics_model = SatDynamicsModel()

    def fullObservable6D(state: ndarray):
        return state

    derived = {"Q": None, "R": None, "p_init": None}
    for k, v in params.items():
        if k not in derived:
            continue
        if isinstance(v, dict):
            assert "dist" in v and "params" in v
            assert len(v["params"]) == 2
            assert len(v["params"][0]) == 6 and len(v["params"][1]) == 6
            derived[k] = diag(getRandomParams(**v))
        elif isinstance(v, (float, int)):
            derived[k] = v * eye(6)
        elif isinstance(v, ndarray):
            assert v.shape == (6, 6)
            derived[k] = v
    return UnscentedKalmanFilter(time=time, est_x=params["x_init"], est_p=derived["p_init"],
                                 dynamics_model=dynamics_model, measurement_model=fullObservable6D,
                                 q_matrix=derived["Q"], r_matrix=derived["R"])


def getRandomParams(dist: str, params: list, seed: int = None) -> ndarray:
    """Random diagonal entries (ez_ukf.py:98-119)."""
    rng = default_rng(seed=seed)
    if dist == "uniform":
        f = rng.uniform
    elif dist == "normal":
        f = rng.normal
    return f(*params)
