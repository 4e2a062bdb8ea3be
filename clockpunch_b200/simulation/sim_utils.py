"""Initial-condition generation (mirror of punchclock/simulation/sim_utils.py:19-199).

Same functions, argument meaning and return shapes as the reference.  The random draws are
the reference's own (numpy ``default_rng(seed)`` with identical call order, so a given seed
yields the same COE / LLA / ECI samples); the frame conversion of the whole batch runs in one
CUDA launch (pc_coe2eci / pc_lla2eci) instead of one Python call per agent.  The policy builder
shim of the reference (sim_utils.py:202-231) drags Ray and the torch nets into the import chain
at module level; here it resolves its two builders when CALLED (from the reference tree, through
install_as_punchclock), so building an environment does not import Ray.
"""
from __future__ import annotations

import numpy as np
from numpy import asarray, ndarray, pi, zeros
from numpy.random import default_rng

from ..common.constants import getConstants
from ..common.transforms import coe2eci_batch, lla2eci_batch


def genInitStatesArray(num_initial_conditions: int, dist: str, dist_params: list[list], frame: str,
                       seed: int = None) -> ndarray:
    """As genInitStates, returning one (6, K) array (the layout the engine consumes)."""
    if seed is None:
        seed = default_rng().integers(9999999)
    print(f"RNG seed = {seed}")                                    # sim_utils.py:76
    rng = default_rng(seed)
    dist_params = asarray(dist_params)
    if dist == "normal":
        ic_array = rng.normal(loc=dist_params[:, 0], scale=dist_params[:, 1], size=[num_initial_conditions, 6])
    elif dist == "uniform":
        ic_array = rng.uniform(low=dist_params[:, 0], high=dist_params[:, 1], size=[num_initial_conditions, 6])
    else:
        raise ValueError("dist must be 'normal' or 'uniform'")
    if num_initial_conditions == 0:
        return zeros((6, 0))
    if frame == "ECI":
        return np.ascontiguousarray(ic_array.T)
    if frame == "LLA":
        # velocities ignored: zero ECEF velocity, non-zero ECI velocity (sim_utils.py:96-101)
        return lla2eci_batch(np.ascontiguousarray(ic_array[:, :3].T), 0.0)
    if frame == "COE":
        coe = saturateCOEs(ic_array[:, 0], ic_array[:, 1], ic_array[:, 2], ic_array[:, 3], ic_array[:, 4],
                           ic_array[:, 5], return_array=True)
        return coe2eci_batch(np.ascontiguousarray(coe.T))
    raise ValueError("frame must be 'ECI', 'COE' or 'LLA'")


def genInitStates(num_initial_conditions: int, dist: str, dist_params: list[list], frame: str,
                  seed: int = None) -> list[ndarray]:
    """Stochastically generate agent initial states in `frame`, output in ECI as a
    num_initial_conditions-long list of (6, 1) arrays (sim_utils.py:19-120)."""
    arr = genInitStatesArray(num_initial_conditions, dist, dist_params, frame, seed)
    return [arr[:, i].reshape((6, 1)) for i in range(arr.shape[1])]


def saturateCOEs(sma, ecc, inc, raan, argp, ta, RE: float = None, return_array: bool = False):
    """Wrap angles and saturate SMA / eccentricity (sim_utils.py:123-199): SMA >= RE,
    0 <= ecc <= 0.99, inc mod pi, RAAN / ARGP / TA mod 2 pi."""
    cols = [np.asarray(v, dtype=float) for v in (sma, ecc, inc, raan, argp, ta)]
    if not all(len(c) == len(cols[0]) for c in cols):
        raise ValueError("not all lists have same length!")
    if RE is None:
        RE = getConstants()["earth_radius"]
    out = zeros([len(cols[0]), 6])
    out[:, 0] = np.where(cols[0] < RE, RE, cols[0])
    out[:, 1] = np.where(cols[1] < 0, 0.0, np.where(cols[1] > 0.99, 0.99, cols[1]))
    out[:, 2] = cols[2] % pi
    for j in (3, 4, 5):
        out[:, j] = cols[j] % (2 * pi)
    if return_array:
        return out
    return {k: list(out[:, i]) for i, k in enumerate(("sma", "ecc", "inc", "raan", "argp", "ta"))}


def buildCustomOrRayPolicy(config_or_path):
    """A CustomPolicy from a config dict, or a Ray policy from a checkpoint path
    (sim_utils.py:208-231; used by simulation/mc.py).  Not on the hot path: the builders are the
    reference's own, imported on first use."""
    assert isinstance(config_or_path, (str, dict)), "config_or_path must be a dict or str"
    if isinstance(config_or_path, str):
        from punchclock.ray.build_ray_policy import buildCustomRayPolicy
        return buildCustomRayPolicy(config_or_path)
    from punchclock.policies.policy_builder import buildCustomPolicy
    return buildCustomPolicy(config_or_path)
