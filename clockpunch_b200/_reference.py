"""Access to the reference's OWN implementation of configurations that are not on the hot path.

The CUDA path of this package is the reference's hot-path configuration: 6-D state, identity
measurement, `SatDynamicsModel` / `StaticTerrestrial` (what `ezUKF` and `SSASchedulerParams` build).
`UnscentedKalmanFilter` (ukf_v2.py:58-128) and `simplePropagate` (propagator.py:12-81) are generic --
any state dimension, any Python measurement function or right-hand side -- and the reference's own
tests use them that way (tests/estimation/test_ukf_v2.py:54-170, tests/dynamics/test_propagator.py:10-68).
Those configurations are host-side Python by nature (a Python callback per sigma point); when the
reference tree is available (`find_reference`: install_as_punchclock(reference_path), $PUNCHCLOCK_REFERENCE,
or an importable `punchclock`) they are served by the reference's own classes, loaded from its
files and bound to the mirrored base classes; otherwise they raise `NotImplementedError`.  This is NOT
a fallback of the hot path: a hot-path configuration never comes here, and it still raises when the
CUDA library or a device is missing.
"""
from __future__ import annotations

import importlib.util
import os
import sys

_loaded: dict[str, object] = {}


def load(rel: str):
    """The reference module `punchclock/<rel>.py` ('estimation.ukf_v2'), executed from its own file under a
    private name, with its `punchclock.*` imports resolved through install_as_punchclock()."""
    if rel in _loaded:
        return _loaded[rel]
    import clockpunch_b200 as cpb
    ref = cpb.find_reference()
    if ref is None:
        raise NotImplementedError(
            "this configuration is not on the CUDA hot path (6-D state, identity measurement, built-in two-body / "
            "Earth-fixed dynamics) and no reference tree was found to serve it (set PUNCHCLOCK_REFERENCE or call "
            "clockpunch_b200.install_as_punchclock(reference_path)); clockpunch_b200 has no CPU implementation of its own")
    root = sys.modules.get("punchclock")
    if root is None or not getattr(root, "__clockpunch_b200__", False):
        cpb.install_as_punchclock(os.path.dirname(ref))
    path = os.path.join(ref, *rel.split(".")) + ".py"
    name = "clockpunch_b200._reference_modules." + rel.replace(".", "_")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    _loaded[rel] = mod
    return mod
