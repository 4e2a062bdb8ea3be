"""clockpunch_b200 -- B200-native hot path of punchclock (dylan906/clockpunch).

Batched orbit propagation, the (sensor, target) visibility map and the per-target
UKF predict/update run as hand-written FP64 CUDA kernels (sm_100a) in
``libpunch_b200.so``; this package is the thin ctypes layer over its C-ABI
(``include/punch_b200.h``) plus a mirror of the reference's Python module API
(``punchclock.dynamics``, ``punchclock.common``, ``punchclock.estimation``,
``punchclock.simulation``) so reference callers can switch with an import swap:

    import clockpunch_b200 as cpb
    cpb.install_as_punchclock()        # `import punchclock.dynamics...` now resolves here

There is no CPU fallback: every physics call raises if the CUDA library or a
CUDA device is missing.
"""
from __future__ import annotations

import importlib
import sys

__version__ = "0.1.0"

_MIRRORED = [
    "common", "common.constants", "common.transforms", "common.visibility", "common.agents", "common.metrics", "common.utilities",
    "dynamics", "dynamics.dynamics", "dynamics.propagator", "dynamics.dynamics_classes",
    "estimation", "estimation.filter_base_class", "estimation.ukf_v2", "estimation.ez_ukf",
    "simulation", "simulation.sim_utils",
    "environment", "environment.env", "environment.env_utils", "environment.env_parameters",
    "schedule_tree", "schedule_tree.access_windows",
]


def install_as_punchclock() -> None:
    """Alias this package's mirrored modules under the reference's import paths."""
    me = sys.modules[__name__]
    sys.modules.setdefault("punchclock", me)
    for name in _MIRRORED:
        try:
            mod = importlib.import_module(f"{__name__}.{name}")
        except ModuleNotFoundError:
            continue
        sys.modules.setdefault(f"punchclock.{name}", mod)


def _register_gym_env() -> None:
    """`gym.make("SSAScheduler-v0")` as the reference registers it (punchclock/__init__.py:7-10),
    pointing at the CUDA-backed env.  gymnasium is optional; without it this is a no-op."""
    try:
        from gymnasium.envs.registration import register, registry
    except ModuleNotFoundError:
        return
    if "SSAScheduler-v0" not in registry:
        register(id="SSAScheduler-v0", entry_point="clockpunch_b200.environment.env:SSAScheduler")


_register_gym_env()
