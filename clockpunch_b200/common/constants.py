"""Constants (mirror of punchclock/common/constants.py:9-41)."""
from numpy import pi


def getConstants() -> dict:
    """Same keys and (deliberately low-precision) values as the reference."""
    params = {
        "earth_radius": 6378.1363,          # constants.py:21
        "mu": 398600,                       # constants.py:24
        "earth_rotation_rate": 7.27e-5,     # constants.py:27
        "sidereal_day": 86164.090517,
        "solar_day": 86400,
    }
    params["gso_radius"] = ((params["sidereal_day"] ** 2) / (4 * (pi**2)) * params["mu"]) ** (1 / 3)
    params["gso_altitude"] = params["gso_radius"] - params["earth_radius"]
    return params
