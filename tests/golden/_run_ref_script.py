"""Run one of the REFERENCE's own test scripts (they are top-level scripts that print) on one of two stacks:

    python tests/golden/_run_ref_script.py ref  <reference tree> tests/estimation/test_ez_ukf.py
    python tests/golden/_run_ref_script.py seam <reference tree> tests/estimation/test_ez_ukf.py

`ref`: the all-reference stack under the stand-ins of _ref_stubs.py (what make_golden_ref_scripts.py records);
`seam`: the same script, unmodified, with `clockpunch_b200.install_as_punchclock()` active and the TEST-ONLY CPU
doubles of tests/_oracle_engine.py / tests/_oracle_lib.py in place of the CUDA library (tests/test_reference_scripts.py).
Every random source is seeded the same way in both (the scripts themselves seed nothing).
"""
import os
import random
import runpy
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (HERE, os.path.join(ROOT, "tests"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)
warnings.filterwarnings("ignore")
mode, ref_root, script = sys.argv[1], sys.argv[2], sys.argv[3]
os.environ["PUNCHCLOCK_REFERENCE"] = ref_root

import numpy as np  # noqa: E402

_default_rng = np.random.default_rng
_count = [0]


def _seeded_default_rng(seed=None):
    """`default_rng()` without a seed -> the k-th unseeded generator of the run gets seed 1000 + k."""
    if seed is None:
        _count[0] += 1
        seed = 1000 + _count[0]
    return _default_rng(seed)


np.random.default_rng = _seeded_default_rng
import numpy.random  # noqa: E402

numpy.random.default_rng = _seeded_default_rng
np.random.seed(0)
random.seed(0)
np.set_printoptions(precision=6, suppress=False, linewidth=200)

import _ref_stubs  # noqa: E402,F401

if mode == "seam":
    import clockpunch_b200 as cpb
    cpb.install_as_punchclock(ref_root)
    sys.modules["punchclock.ray.build_ray_policy"] = _ref_stubs._StubModule("punchclock.ray.build_ray_policy")
    import _oracle_engine
    _oracle_engine.install_cpu_doubles()
elif ref_root != "/root/reference":
    sys.modules["punchclock"].__path__ = [os.path.join(ref_root, "punchclock")]
np.random.seed(0)
runpy.run_path(os.path.join(ref_root, script), run_name="__main__")
