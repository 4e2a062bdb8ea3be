"""Generate golden vectors from the LIVE reference (run in the build container only).

    python tests/golden/make_golden.py      # writes tests/golden/*.npz

The reference package cannot be imported as shipped (punchclock/__init__.py:5
imports gymnasium, which is not installed), so a bare ``punchclock`` module
object is pre-registered with ``__path__`` pointing at /root/reference; the
dynamics / transforms / estimation modules then run unmodified against this
container's numpy + scipy.  /root/reference does not
exist on the GPU box: tests only ever read the committed .npz files.
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference/punchclock"
pkg = types.ModuleType("punchclock")
pkg.__path__ = [REF]
sys.modules["punchclock"] = pkg

from punchclock.common.transforms import coe2eci, ecef2eci, eci2ecef, lla2ecef  # noqa: E402
from punchclock.dynamics.dynamics_classes import SatDynamicsModel, StaticTerrestrial  # noqa: E402
from punchclock.estimation.ez_ukf import ezUKF  # noqa: E402
from punchclock.estimation.ukf_v2 import UnscentedKalmanFilter  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
RE = 6378.1363
sat = SatDynamicsModel()
ter = StaticTerrestrial()


def catalog(rng, n, regime):
    lo, hi = {"LEO": (RE + 400, RE + 1000), "MEO": (20000, 30000), "GEO": (42000, 43000),
              "HEO": (15000, 30000)}[regime]
    out = []
    for _ in range(n):
        a = rng.uniform(lo, hi)
        e = rng.uniform(0, 0.01) if regime != "HEO" else rng.uniform(0.3, 0.7)
        if regime == "HEO":
            e = min(e, 1 - (RE + 300) / a)
        out.append(coe2eci(a, e, rng.uniform(0, np.pi), *rng.uniform(0, 2 * np.pi, 3)))
    return np.array(out).T


def mp_truth(x, dt, mu=398600.0):
    """Closed-form two-body solution (Lagrange f and g functions, Kepler's equation in the
    eccentric-anomaly difference) evaluated with 40-digit mpmath arithmetic."""
    import mpmath as mp
    mp.mp.dps = 40
    r0 = [mp.mpf(float(v)) for v in x[:3]]
    v0 = [mp.mpf(float(v)) for v in x[3:]]
    mu, dt = mp.mpf(mu), mp.mpf(dt)
    r = mp.sqrt(sum(c * c for c in r0))
    v2 = sum(c * c for c in v0)
    a = 1 / (2 / r - v2 / mu)
    n = mp.sqrt(mu / a**3)
    ec = 1 - r / a
    es = sum(p * q for p, q in zip(r0, v0)) / mp.sqrt(mu * a)
    dE = mp.findroot(lambda E: E - ec * mp.sin(E) + es * (1 - mp.cos(E)) - n * dt, n * dt)
    r1 = a * (1 - ec * mp.cos(dE) + es * mp.sin(dE))
    f = 1 - a * (1 - mp.cos(dE)) / r
    g = dt - (dE - mp.sin(dE)) / n
    fd = -mp.sqrt(mu * a) * mp.sin(dE) / (r * r1)
    gd = 1 - a * (1 - mp.cos(dE)) / r1
    return np.array([float(f * p + g * q) for p, q in zip(r0, v0)] + [float(fd * p + gd * q) for p, q in zip(r0, v0)])


def gold_propagate():
    rng = np.random.default_rng(20240625)
    d = {}
    # tests/dynamics/test_dynamics_classes.py:139-149 scenarios
    d["single_x0"] = np.array([8000.0, 0, 0, 0, 6, 0])
    d["single_out"] = sat.propagate(d["single_x0"], 0, 5)
    ts = np.linspace(0, 10000, 49)
    x = np.array([7000.0, 0, 0, 0, 7.5, 0])
    chain = [x]
    for t0, t1 in zip(ts[:-1], ts[1:]):
        chain.append(sat.propagate(chain[-1], t0, t1))
    d["chain_t"] = ts
    d["chain_x"] = np.array(chain).T
    d["chain_truth_end"] = mp_truth(x, 10000.0)
    # seeded catalogs, (6,K) batch API
    x0 = np.concatenate([catalog(rng, 12, r) for r in ("LEO", "MEO", "GEO", "HEO")], axis=1)
    d["cat_x0"] = x0
    for dt in (60.0, 100.0, 600.0):
        d[f"cat_dt{int(dt)}"] = sat.propagate(x0, 0.0, dt)
    d["cat_back"] = sat.propagate(x0, 100.0, 40.0)          # negative time step
    # 30-digit Taylor-series truth for the longest step: at dt = 600 s the reference's own error
    # (rtol = atol = 1e-10 on |x| ~ 1e4 km) reaches 1e-7 km, so parity there is judged against
    # the reference's error band and, separately, against this truth.
    d["cat_dt600_truth"] = np.array([mp_truth(x0[:, i], 600.0) for i in range(x0.shape[1])]).T
    d["cat_dt100_truth"] = np.array([mp_truth(x0[:, i], 100.0) for i in range(x0.shape[1])]).T
    d["col61_out"] = sat.propagate(x0[:, :1], 0.0, 60.0)     # (6,1) -> (6,)
    np.savez_compressed(os.path.join(HERE, "propagate.npz"), **d)


def gold_transforms():
    rng = np.random.default_rng(20240626)
    d = {}
    lla = np.stack([rng.uniform(-np.pi / 2, np.pi / 2, 10), rng.uniform(-np.pi, np.pi, 10),
                    rng.uniform(0, 2, 10)])
    d["lla"] = lla
    d["lla_ecef"] = np.array([lla2ecef(lla[:, i]) for i in range(10)]).T
    x = rng.normal(size=(6, 10)) * np.array([[7000, 7000, 7000, 5, 5, 5]]).T
    d["x"] = x
    for jd in (0.0, 1234.5, 86400.0 * 3):
        d[f"ecef2eci_{int(jd)}"] = ecef2eci(x, jd)
        d[f"eci2ecef_{int(jd)}"] = eci2ecef(x, jd)
    sens0 = np.array([ecef2eci(lla2ecef(lla[:, i]), 0) for i in range(10)]).T
    d["ter_x0"] = sens0
    d["ter_dt100"] = ter.propagate(sens0, 0.0, 100.0)
    d["ter_chain"] = ter.propagate(ter.propagate(sens0, 0.0, 100.0), 100.0, 4000.0)
    d["ter_single"] = ter.propagate(sens0[:, 0], 50.0, 110.0)
    coe = np.stack([rng.uniform(RE + 300, 45000, 16), rng.uniform(0, 0.7, 16), rng.uniform(0, np.pi, 16),
                    rng.uniform(0, 2 * np.pi, 16), rng.uniform(0, 2 * np.pi, 16), rng.uniform(0, 2 * np.pi, 16)])
    d["coe"] = coe
    d["coe_eci"] = np.array([coe2eci(*coe[:, i]) for i in range(16)]).T
    np.savez_compressed(os.path.join(HERE, "transforms.npz"), **d)


def run_filter(f, times, zs, d, key):
    recs = {k: [] for k in ("est_x", "est_p", "pred_x", "pred_p", "nis")}
    for t, z in zip(times, zs):
        f.predictAndUpdate(t, None if np.isnan(z[0]) else z)
        recs["est_x"].append(f.est_x.copy())
        recs["est_p"].append(f.est_p.copy())
        recs["pred_x"].append(f.pred_x.copy())
        recs["pred_p"].append(f.pred_p.copy())
        recs["nis"].append(float(f.nis) if np.size(f.nis) and not np.isnan(z[0]) else np.nan)
    for k, v in recs.items():
        d[f"{key}_{k}"] = np.array(v)


def gold_ukf():
    rng = np.random.default_rng(20240627)
    d = {}
    # tests/estimation/test_ez_ukf.py:25-40
    x = np.array([8000.0, 0, 0, 0, 8, 0])
    f = ezUKF({"x_init": x, "p_init": 0.1, "dynamics_type": "satellite", "Q": 0.1, "R": 0.1})
    d["w_gamma"] = np.array(f.gamma)
    d["w_mean"] = f.mean_weight
    d["w_cvr_diag"] = np.diag(f.cvr_weight)
    f.predictAndUpdate(5, x)
    d["ez_est_x"], d["ez_est_p"], d["ez_pred_p"], d["ez_nis"] = f.est_x, f.est_p, f.pred_p, np.array(f.nis)
    # tests/estimation/test_ukf_v2.py:254-313,408-478: orbit UKF, measurement every 5th step
    D3 = np.diag([1, 1, 1, 1e-3, 1e-3, 1e-3])
    Q, R, P0 = 2 * D3, 1.0 * D3, 20 * D3
    xt = np.array([7000.0, 0, 0, 0, 7.5, 0.5])
    x_est0 = xt + rng.normal(size=6) * np.sqrt(np.diag(P0))
    times = 100.0 * np.arange(1, 26)
    truth, zs = [], []
    for i, t in enumerate(times):
        xt = sat.propagate(xt, t - 100.0, t)
        truth.append(xt)
        if (i + 1) % 5 == 0:
            zs.append(xt + rng.normal(size=6) * np.sqrt(np.diag(R)))
        else:
            zs.append(np.full(6, np.nan))
    f = UnscentedKalmanFilter(0, x_est0, P0, sat, lambda s: s, Q, R)
    d.update(orbit_Q=Q, orbit_R=R, orbit_P0=P0, orbit_x0=x_est0, orbit_times=times,
             orbit_z=np.array(zs), orbit_truth=np.array(truth))
    run_filter(f, times, zs, d, "orbit")
    # reference default filter config (env_parameters.py:572-577), mixed catalog, z every step / none
    D = np.diag([1, 1, 1, 1e-2, 1e-2, 1e-2])
    Q, R, P0 = 0.1 * D, 1.0 * D, 10.0 * D
    x0 = np.concatenate([catalog(rng, 4, r) for r in ("LEO", "MEO", "GEO", "HEO")], axis=1)
    d.update(cat_Q=Q, cat_R=R, cat_P0=P0, cat_truth0=x0)
    est0 = x0 + rng.normal(size=x0.shape) * np.sqrt(np.diag(P0))[:, None]
    d["cat_est0"] = est0
    for mode in ("tasked", "untasked", "mixed"):
        for dt in (60.0, 100.0):
            out = {k: [] for k in ("est_x", "est_p", "pred_p", "nis", "z", "truth")}
            for i in range(x0.shape[1]):
                f = ezUKF({"x_init": est0[:, i], "p_init": P0, "dynamics_type": "satellite", "Q": Q, "R": R})
                xt = x0[:, i]
                ex, ep, pp, ni, zz, tt = [], [], [], [], [], []
                for s in range(4):
                    xt = sat.propagate(xt, s * dt, (s + 1) * dt)
                    has = mode == "tasked" or (mode == "mixed" and (s + i) % 2 == 0)
                    z = xt + rng.normal(size=6) * np.sqrt(np.diag(R)) if has else np.full(6, np.nan)
                    f.predictAndUpdate((s + 1) * dt, z if has else None)
                    ex.append(f.est_x.copy()); ep.append(f.est_p.copy()); pp.append(f.pred_p.copy())
                    ni.append(float(f.nis) if has else np.nan); zz.append(z); tt.append(xt)
                for k, v in zip(out, (ex, ep, pp, ni, zz, tt)):
                    out[k].append(v)
            for k, v in out.items():
                d[f"cat_{mode}_dt{int(dt)}_{k}"] = np.array(v)        # (targets, steps, ...)
    # non-PD covariance -> nearestPD fallback (ukf_v2.py:170-176)
    bad = P0.copy()
    bad[0, 1] = bad[1, 0] = 10.5
    f = ezUKF({"x_init": x0[:, 0], "p_init": P0, "dynamics_type": "satellite", "Q": Q, "R": R})
    d["npd_cov"] = bad
    d["npd_sigma"] = f.generateSigmaPoints(x0[:, 0], bad)
    np.savez_compressed(os.path.join(HERE, "ukf.npz"), **d)


if __name__ == "__main__":
    gold_propagate()
    gold_transforms()
    gold_ukf()
    print("golden vectors written to", HERE)
