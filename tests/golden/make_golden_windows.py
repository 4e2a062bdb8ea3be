"""Golden access-window forecasts from the LIVE reference AccessWindowCalculator
(schedule_tree/access_windows.py), satvis.isVis -> oracle restatement.  Build container only:

    python tests/golden/make_golden_windows.py      # writes tests/golden/windows.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_stubs  # noqa: E402,F401
import satvis.visibility_func as svf  # noqa: E402
import punchclock.schedule_tree.access_windows as aw  # noqa: E402
from punchclock.common.transforms import coe2eci, ecef2eci, lla2ecef  # noqa: E402

aw.isVis = svf.isVis                       # the module did `from satvis.visibility_func import isVis`
RE = 6378.1363
rng = np.random.default_rng(20240702)
out = {}
for tag, (M_g, M_s, N, horizon, dt, fixed, merge, t0) in {
        "a": (3, 1, 7, 6, 300.0, False, True, 0.0), "b": (2, 0, 5, 8, 600.0, True, False, 1200.0)}.items():
    sens = [ecef2eci(lla2ecef(np.array([rng.uniform(-1.2, 1.2), rng.uniform(-np.pi, np.pi), 1e-6])), t0)
            for _ in range(M_g)]
    sens += [coe2eci(rng.uniform(RE + 500, RE + 900), 0.001, rng.uniform(0, np.pi), *rng.uniform(0, 2 * np.pi, 3))
             for _ in range(M_s)]
    targ = [coe2eci(rng.uniform(RE + 400, 30000), rng.uniform(0, 0.05), rng.uniform(0, np.pi),
                    *rng.uniform(0, 2 * np.pi, 3)) for _ in range(N)]
    xs, xt = np.array(sens).T, np.array(targ).T
    dyn_s = ["terrestrial"] * M_g + ["satellite"] * M_s
    calc = aw.AccessWindowCalculator(num_sensors=M_g + M_s, num_targets=N, dynamics_sensors=dyn_s,
                                     dynamics_targets="satellite", horizon=horizon, dt=dt, fixed_horizon=fixed,
                                     merge_windows=merge)
    nw, vh, vht, th, ttw = calc.calcNumWindows(xs, xt, t0, return_vis_hist=True)
    out.update({f"{tag}_xs": xs, f"{tag}_xt": xt, f"{tag}_kinds": np.array([1] * M_g + [0] * M_s, np.uint8),
                f"{tag}_cfg": np.array([horizon, dt, float(fixed), float(merge), t0]), f"{tag}_num_windows": nw,
                f"{tag}_vis_hist": vh, f"{tag}_vis_hist_targets": vht, f"{tag}_time": th, f"{tag}_ttw": ttw})
    print(tag, vh.shape, nw)
np.savez_compressed(os.path.join(HERE, "windows.npz"), **out)
