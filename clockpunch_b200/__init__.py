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

# Modules of this package that stand in for the reference module of the same path.  Each defines every
# public name of the module it replaces (checked by tests/test_seam.py against the reference tree).
_MIRRORED = [
    "common.constants", "common.transforms", "common.visibility", "common.agents", "common.metrics",
    "dynamics.dynamics", "dynamics.propagator", "dynamics.dynamics_classes",
    "estimation.filter_base_class", "estimation.ukf_v2", "estimation.ez_ukf",
    "simulation.sim_utils",
    "environment.env", "environment.env_utils", "environment.env_parameters",
    "schedule_tree.access_windows",
]
# Mirrored only as far as the hot path needs it (actionSpace2Array): aliased when there is NO reference
# tree to take the full module from.
_PARTIAL = ["common.utilities"]
_PACKAGES = ["common", "dynamics", "estimation", "simulation", "environment", "schedule_tree"]


def find_reference(reference_path: str | None = None) -> str | None:
    """Directory of the reference's `punchclock` package: `reference_path` (the package directory or
    its parent), else $PUNCHCLOCK_REFERENCE, else a `punchclock` that is importable or already
    imported (its __init__ is not executed: it needs gymnasium).  None when there is none."""
    import importlib.util
    import os
    cands = [reference_path, os.environ.get("PUNCHCLOCK_REFERENCE")]
    for c in cands:
        if not c:
            continue
        for d in (c, os.path.join(c, "punchclock")):
            if os.path.isfile(os.path.join(d, "environment", "env.py")):
                return os.path.abspath(d)
        if c is reference_path:
            raise FileNotFoundError(f"no punchclock package under {reference_path!r}")
    mod = sys.modules.get("punchclock")
    paths = list(getattr(mod, "__path__", [])) if mod is not None and mod is not sys.modules[__name__] else []
    if not paths:
        try:
            spec = importlib.util.find_spec("punchclock")
        except (ImportError, ValueError):
            spec = None
        paths = list(spec.submodule_search_locations or []) if spec is not None else []
    for d in paths:
        if os.path.isfile(os.path.join(d, "environment", "env.py")):
            return os.path.abspath(d)
    return None


def install_as_punchclock(reference_path: str | None = None) -> str | None:
    """Make `import punchclock....` resolve to this package for the hot-path modules and to the
    reference for everything else.

    `punchclock` and its sub-packages become namespace-like package objects whose `__path__` is the
    reference's directory (when one is found: see find_reference), so that modules this package does
    not mirror -- `punchclock.ray.*`, the Gymnasium wrappers, `punchclock.policies`, `punchclock.nets`,
    `punchclock.common.{math,utilities,orbits,...}`, `punchclock.simulation.{sim_runner,mc}` ... -- are
    imported from the reference, unmodified, by the ordinary import machinery; the mirrored modules
    are pre-registered in `sys.modules` (and set as attributes of their package), so both those
    reference modules and user code get the CUDA-backed classes under the reference's names.  Any
    `punchclock*` module imported before the call is dropped from `sys.modules` first, so that later
    imports bind to the mirrors (objects somebody already holds are not rebound: call this first).
    The reference's package `__init__` (which only registers the Gymnasium env id) is not executed;
    `SSAScheduler-v0` is registered for the mirrored env instead.  Returns the reference directory in
    use, or None (then only the mirrored modules are importable)."""
    import os
    import types
    ref = find_reference(reference_path)
    for name in [n for n in sys.modules if n == "punchclock" or n.startswith("punchclock.")]:
        del sys.modules[name]

    def package(name: str, rel: str) -> types.ModuleType:
        m = types.ModuleType(name)
        m.__package__ = name
        d = os.path.join(ref, rel) if ref else None
        m.__path__ = [d] if d and os.path.isdir(d) else []
        m.__file__ = os.path.join(d, "__init__.py") if m.__path__ else None
        m.__clockpunch_b200__ = True
        sys.modules[name] = m
        return m

    root = package("punchclock", "")
    root.__version__ = "0.8.5"
    pkgs = {}
    for sub in _PACKAGES:
        pkgs[sub] = package(f"punchclock.{sub}", sub)
        setattr(root, sub, pkgs[sub])
    for name in _MIRRORED + ([] if ref else _PARTIAL):
        mod = importlib.import_module(f"{__name__}.{name}")
        sys.modules[f"punchclock.{name}"] = mod
        sub, leaf = name.split(".")
        setattr(pkgs[sub], leaf, mod)
    _register_gym_env()
    return ref


def _register_gym_env() -> None:
    """`gym.make("SSAScheduler-v0")` as the reference registers it (punchclock/__init__.py:7-10),
    pointing at the CUDA-backed env.  gymnasium is optional; without it this is a no-op."""
    try:
        from gymnasium.envs.registration import register, registry
    except ModuleNotFoundError:
        return
    try:
        if "SSAScheduler-v0" not in registry:
            register(id="SSAScheduler-v0", entry_point="clockpunch_b200.environment.env:SSAScheduler")
    except TypeError:      # a stand-in gymnasium without a real registry (test stubs)
        pass


_register_gym_env()
