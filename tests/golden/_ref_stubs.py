"""TEST-ONLY stand-ins that let the reference package import in the build container
(gymnasium / ray / satvis / intervaltree / matplotlib are not installed): see make_golden_env.py.
Importing this module registers them; none of the stand-ins does arithmetic except the satvis
visibility functions, which are the oracle's restatement (parity unpinned)."""
import importlib.abc
import importlib.machinery
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
np.Inf = np.inf


class _AnyMeta(type):
    """Class-level attribute access on a fabricated class (`plt.subplots` where `plt` is one) works too."""
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


class _Anything(metaclass=_AnyMeta):
    """Permissive class: any attribute, call, subscript or two-way unpacking works."""
    def __init__(self, *a, **k):
        pass

    def __getitem__(self, item):
        return _Anything()

    def __iter__(self):
        return iter((_Anything(), _Anything()))

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __class_getitem__(cls, item):
        return cls

    def __mro_entries__(self, bases):
        return (_Anything,)


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (_Anything,), {})
        setattr(self, name, cls)
        return cls


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in ("ray", "satvis", "intervaltree", "matplotlib", "tensorflow", "pygame"):
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


# gymnasium: a functional stand-in (spaces, Env, Wrapper, FilterObservation ...), because the reference's
# wrappers really use them (tests/golden/_mini_gym.py); everything else missing is fabricated
sys.path.insert(0, HERE)
import _mini_gym  # noqa: E402

_mini_gym.install()
sys.meta_path.insert(0, _Finder())


class RandomEnv(sys.modules["gymnasium"].Env):
    """Stand-in for ray.rllib.examples.env.random_env.RandomEnv (the reference's wrapper tests wrap it): spaces
    from the config, observations / rewards sampled from them, never-ending episodes."""

    def __init__(self, config=None):
        gym = sys.modules["gymnasium"]
        config = config or {}
        self.action_space = config.get("action_space", gym.spaces.Discrete(2))
        self.observation_space = config.get("observation_space", gym.spaces.Discrete(2))
        self.reward_space = config.get("reward_space", gym.spaces.Box(low=-1.0, high=1.0, shape=(), dtype=np.float32))
        self.steps = 0

    def reset(self, *, seed=None, options=None):
        self.steps = 0
        return self.observation_space.sample(), {}

    def step(self, action):
        self.steps += 1
        return self.observation_space.sample(), float(self.reward_space.sample()), False, False, {}


_rand_mod = _StubModule("ray.rllib.examples.env.random_env")
_rand_mod.__path__ = []
_rand_mod.RandomEnv = RandomEnv
sys.modules["ray.rllib.examples.env.random_env"] = _rand_mod
pkg = types.ModuleType("punchclock")
pkg.__path__ = ["/root/reference/punchclock"]
sys.modules["punchclock"] = pkg
# satvis arithmetic -> oracle restatement
import satvis.visibility_func as svf  # noqa: E402
from oracle.visibility import calc_vis_and_der_vis, is_vis, visibility_func  # noqa: E402

svf.isVis = lambda r1, r2, RE, hg=0: is_vis(r1, r2, RE, hg)
svf.visibilityFunc = lambda r1, r2, RE, hg=0, tol=1e-13: visibility_func(r1, r2, RE, hg, tol)
svf.calcVisAndDerVis = lambda r1, r1dot, r1mag_dot, r2, r2dot, r2mag_dot, RE, hg=0: calc_vis_and_der_vis(
    r1, r1dot, r1mag_dot, r2, r2dot, r2mag_dot, RE, hg)
# the Ray-checkpoint policy loader drags RLlib + the torch nets in (sim_utils.py:15 -> ray/build_ray_policy.py
# -> nets/*: `torch, nn = try_import_torch()`): stub that one module; the custom-policy builder and the
# policies themselves are the reference's own
sys.modules["punchclock.ray.build_ray_policy"] = _StubModule("punchclock.ray.build_ray_policy")
