"""Constants of the reference (punchclock/common/constants.py:21-27)."""
RE = 6378.1363          # constants.py:21  spherical Earth radius, km
MU = 398600.0           # constants.py:24  km^3/s^2 (reference uses the integer 398600)
OMEGA_EARTH = 7.27e-5   # constants.py:27  rad/s
