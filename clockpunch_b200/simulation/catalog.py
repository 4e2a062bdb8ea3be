"""Synthetic catalogs for benchmarks and tests.

Pure-numpy INPUT synthesis (no physics stepping): seeded LEO/MEO/GEO targets drawn in
classical elements with the reference's generator semantics (genInitStates(frame="COE"),
simulation/sim_utils.py:19-120; default LEO band env_parameters.py:523), ground sensors in
LLA at 1e-6 km altitude (env_parameters.py:509-517) and the reference's default filter
matrices (env_parameters.py:572-577).  Both bench arms consume the arrays made here.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

RE = 6378.1363
MU = 398600.0
OMEGA_EARTH = 7.27e-5

REGIMES = {"LEO": (RE + 400.0, RE + 1000.0), "MEO": (20000.0, 30000.0), "GEO": (42000.0, 43000.0)}
MIXES = {"mixed": (("LEO", 0.6), ("MEO", 0.1), ("GEO", 0.3)), "LEO": (("LEO", 1.0),),
         "MEO": (("MEO", 1.0),), "GEO": (("GEO", 1.0),)}


def coe_to_eci(coe: np.ndarray, mu: float = MU) -> np.ndarray:
    """Vectorised COE -> ECI, (6,N) -> (6,N); same formulas as transforms.py:240-292."""
    sma, ecc, inc, raan, argp, ta = coe
    ca, sa = np.cos(ta), np.sin(ta)
    p = sma * (1.0 - ecc**2)
    rr = p / (1.0 + ecc * ca)
    vv = np.sqrt(mu / p)
    rp, rq, vp, vq = rr * ca, rr * sa, -vv * sa, vv * (ecc + ca)
    co, so, ci, si, cw, sw = np.cos(raan), np.sin(raan), np.cos(inc), np.sin(inc), np.cos(argp), np.sin(argp)
    # columns of rot3(-raan) rot1(-inc) rot3(-argp) with the reference's (active) rot matrices
    P = np.stack([co * cw - so * ci * sw, -so * cw - co * ci * sw, si * sw])
    Q = np.stack([co * sw + so * ci * cw, -so * sw + co * ci * cw, -si * cw])
    return np.concatenate([P * rp + Q * rq, P * vp + Q * vq], axis=0)


def lla_to_eci(lla: np.ndarray, jd: float = 0.0) -> np.ndarray:
    """(3,N) -> (6,N): ecef2eci(lla2ecef(.), jd) (transforms.py:34-90)."""
    lat, lon, alt = lla
    rd = (RE + alt) * np.cos(lat)
    x, y, z = rd * np.cos(lon), rd * np.sin(lon), (RE + alt) * np.sin(lat)
    c, s = np.cos(OMEGA_EARTH * jd), np.sin(OMEGA_EARTH * jd)
    rx, ry = c * x - s * y, s * x + c * y
    return np.stack([rx, ry, z, -OMEGA_EARTH * ry, OMEGA_EARTH * rx, np.zeros_like(z)])


@dataclass
class Catalog:
    sensors: np.ndarray       # (6, M) ECI at t = 0
    sensor_kinds: np.ndarray  # (M,) uint8, 0 = satellite, 1 = terrestrial
    targets: np.ndarray       # (6, N) truth ECI at t = 0
    est_x: np.ndarray         # (6, N) initial estimates
    est_p: np.ndarray         # (N, 6, 6)
    Q: np.ndarray
    R: np.ndarray
    regime: np.ndarray        # (N,) 0 LEO / 1 MEO / 2 GEO

    @property
    def M(self):
        return self.sensors.shape[1]

    @property
    def N(self):
        return self.targets.shape[1]


def synth_catalog(num_targets: int, num_sensors: int, seed: int, mix: str = "mixed",
                  ecc_max: float = 0.01, space_sensor_frac: float = 0.0,
                  sensor_seed: int | None = None) -> Catalog:
    """Targets come from default_rng((seed, 0)), sensors from default_rng((sensor_seed, 1)) so that
    shards with different target seeds can share one replicated sensor network."""
    rng = np.random.default_rng((seed, 0))
    srng = np.random.default_rng((seed if sensor_seed is None else sensor_seed, 1))
    N, M = int(num_targets), int(num_sensors)
    names = [r for r, _ in MIXES[mix]]
    probs = np.array([p for _, p in MIXES[mix]])
    reg = rng.choice(len(names), size=N, p=probs)
    lo = np.array([REGIMES[n][0] for n in names])[reg]
    hi = np.array([REGIMES[n][1] for n in names])[reg]
    coe = np.stack([rng.uniform(lo, hi), rng.uniform(0.0, ecc_max, N), rng.uniform(0.0, np.pi, N),
                    rng.uniform(0.0, 2 * np.pi, N), rng.uniform(0.0, 2 * np.pi, N),
                    rng.uniform(0.0, 2 * np.pi, N)])
    targets = coe_to_eci(coe)
    regime = np.array([{"LEO": 0, "MEO": 1, "GEO": 2}[names[k]] for k in reg], dtype=np.int8)
    n_space = int(round(space_sensor_frac * M))
    n_ground = M - n_space
    lla = np.stack([srng.uniform(-np.pi / 2, np.pi / 2, n_ground), srng.uniform(-np.pi, np.pi, n_ground),
                    np.full(n_ground, 1e-6)])
    ground = lla_to_eci(lla, 0.0)
    scoe = np.stack([srng.uniform(RE + 400, RE + 1000, n_space), np.zeros(n_space),
                     srng.uniform(0.0, np.pi, n_space), srng.uniform(0.0, 2 * np.pi, n_space),
                     srng.uniform(0.0, 2 * np.pi, n_space), srng.uniform(0.0, 2 * np.pi, n_space)])
    sensors = np.concatenate([ground, coe_to_eci(scoe)], axis=1)
    kinds = np.concatenate([np.ones(n_ground, np.uint8), np.zeros(n_space, np.uint8)])
    D = np.diag([1.0, 1.0, 1.0, 1e-2, 1e-2, 1e-2])
    Q, R, P0 = 0.1 * D, 1.0 * D, 10.0 * D
    est_x = targets + rng.normal(size=(6, N)) * np.sqrt(np.diag(P0))[:, None]
    est_p = np.broadcast_to(P0, (N, 6, 6)).copy()
    return Catalog(np.ascontiguousarray(sensors), kinds, np.ascontiguousarray(targets),
                   np.ascontiguousarray(est_x), est_p, Q, R, regime)
