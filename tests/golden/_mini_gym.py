"""TEST-ONLY functional stand-in for the slice of gymnasium 0.26 that the reference's environment and
wrappers use (gymnasium is not installed in this image and cannot be: no network).

tests/golden/_ref_stubs.py fabricates *permissive* classes for missing packages, which is enough to
import the reference and step its bare SSAScheduler, but not to run its Gymnasium WRAPPERS
(punchclock/environment/{obs,info,misc,reward}_wrappers.py): those need spaces that really hold shapes
and bounds, a Wrapper that really forwards attributes, flatten()/flatten_space() ...  This module
provides exactly that, with gymnasium's semantics (attribute forwarding that refuses private names,
Dict spaces that sort plain-dict keys, one-hot flattening of MultiDiscrete, FilterObservation), so that
the same wrapper stack can be run (a) on the live reference env to produce golden rollouts and (b) on
the mirrored env through `install_as_punchclock()` -- tests/test_seam.py.  Nothing here does physics.

    import _mini_gym; _mini_gym.install()        # registers gymnasium, gymnasium.spaces, .wrappers, ...
"""
from __future__ import annotations

import sys
import types
from collections import OrderedDict
from copy import deepcopy

import numpy as np


# ------------------------------------------------------------------------------------------ spaces
class Space:
    def __init__(self, shape=None, dtype=None, seed=None):
        self._shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)
        self._np_random = np.random.default_rng(seed)

    @property
    def shape(self):
        return self._shape

    @property
    def np_random(self):
        return self._np_random

    def seed(self, seed=None):
        self._np_random = np.random.default_rng(seed)
        return [seed]

    def contains(self, x):
        raise NotImplementedError

    def __contains__(self, x):
        return self.contains(x)

    def sample(self, mask=None):
        raise NotImplementedError


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        dtype = np.dtype(dtype)
        if shape is None:
            shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
        shape = tuple(int(s) for s in shape)
        super().__init__(shape, dtype, seed)
        with np.errstate(invalid="ignore", over="ignore"):
            self.low = np.broadcast_to(np.asarray(low, dtype=np.float64), shape).astype(dtype).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=np.float64), shape).astype(dtype).copy()
        self.bounded_below = -np.inf < np.broadcast_to(np.asarray(low, dtype=np.float64), shape)
        self.bounded_above = np.inf > np.broadcast_to(np.asarray(high, dtype=np.float64), shape)

    def is_bounded(self, manner="both"):
        below, above = bool(np.all(self.bounded_below)), bool(np.all(self.bounded_above))
        return {"both": below and above, "below": below, "above": above}[manner]

    def contains(self, x):
        if not isinstance(x, np.ndarray):
            try:
                x = np.asarray(x, dtype=self.dtype)
            except (ValueError, TypeError):
                return False
        return bool(np.can_cast(x.dtype, self.dtype) and x.shape == self.shape
                    and np.all(x >= self.low) and np.all(x <= self.high))

    def sample(self, mask=None):
        lo = np.where(self.bounded_below, self.low, -1e3).astype(np.float64)
        hi = np.where(self.bounded_above, self.high, 1e3).astype(np.float64)
        return np.asarray(self.np_random.uniform(lo, hi)).astype(self.dtype)

    def __eq__(self, other):
        return (isinstance(other, Box) and self.shape == other.shape and self.dtype == other.dtype
                and np.array_equal(self.low, other.low) and np.array_equal(self.high, other.high))

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"


class Discrete(Space):
    def __init__(self, n, seed=None, start=0):
        self.n, self.start = int(n), int(start)
        super().__init__((), np.int64, seed)

    def contains(self, x):
        if isinstance(x, (int, np.integer)) or (isinstance(x, np.ndarray) and x.shape == () and np.issubdtype(x.dtype, np.integer)):
            return self.start <= int(x) < self.start + self.n
        return False

    def sample(self, mask=None):
        return int(self.start + self.np_random.integers(self.n))

    def __eq__(self, other):
        return isinstance(other, Discrete) and (self.n, self.start) == (other.n, other.start)


class MultiBinary(Space):
    def __init__(self, n, seed=None):
        if isinstance(n, (list, tuple, np.ndarray)):
            self.n = shape = tuple(int(i) for i in n)
        else:
            self.n = int(n)
            shape = (self.n,)
        super().__init__(shape, np.int8, seed)

    def contains(self, x):
        if isinstance(x, (list, tuple)):
            x = np.array(x)
        return bool(isinstance(x, np.ndarray) and x.shape == self.shape and np.all((x == 0) | (x == 1)))

    def sample(self, mask=None):
        return self.np_random.integers(0, 2, size=self.shape, dtype=self.dtype)

    def __eq__(self, other):
        return isinstance(other, MultiBinary) and self.n == other.n


class MultiDiscrete(Space):
    def __init__(self, nvec, dtype=np.int64, seed=None):
        self.nvec = np.array(nvec, dtype=dtype, copy=True)
        assert (self.nvec > 0).all(), "nvec (counts) have to be positive"
        super().__init__(self.nvec.shape, dtype, seed)

    def contains(self, x):
        if isinstance(x, (list, tuple)):
            x = np.array(x)
        return bool(isinstance(x, np.ndarray) and x.shape == self.shape and x.dtype != object
                    and np.all(0 <= x) and np.all(x < self.nvec))

    def sample(self, mask=None):
        return (self.np_random.random(self.nvec.shape) * self.nvec).astype(self.dtype)

    def __len__(self):
        return len(self.nvec)

    def __eq__(self, other):
        return isinstance(other, MultiDiscrete) and np.array_equal(self.nvec, other.nvec)


class Dict(Space):
    def __init__(self, spaces=None, seed=None, **spaces_kwargs):
        if isinstance(spaces, Dict):
            spaces = spaces.spaces
        if isinstance(spaces, dict) and not isinstance(spaces, OrderedDict):
            try:                                       # gymnasium sorts the keys of a plain dict
                spaces = OrderedDict(sorted(spaces.items()))
            except TypeError:
                spaces = OrderedDict(spaces.items())
        elif spaces is None:
            spaces = OrderedDict()
        else:
            spaces = OrderedDict(spaces)
        spaces.update(spaces_kwargs)
        for k, s in spaces.items():
            assert isinstance(s, Space), f"Dict space element is not an instance of Space: key={k!r}"
        self.spaces = spaces
        super().__init__(None, None, seed)

    def contains(self, x):
        if isinstance(x, dict) and x.keys() == self.spaces.keys():
            return all(x[k] in self.spaces[k] for k in self.spaces)
        return False

    def sample(self, mask=None):
        return OrderedDict((k, s.sample()) for k, s in self.spaces.items())

    def __getitem__(self, key):
        return self.spaces[key]

    def __setitem__(self, key, value):
        assert isinstance(value, Space), f"Trying to set {key} to Dict space with value that is not a gymnasium space"
        self.spaces[key] = value

    def __iter__(self):
        yield from self.spaces

    def __len__(self):
        return len(self.spaces)

    def keys(self):
        return self.spaces.keys()

    def values(self):
        return self.spaces.values()

    def items(self):
        return self.spaces.items()

    def get(self, key, default=None):
        return self.spaces.get(key, default)

    def __eq__(self, other):
        return isinstance(other, Dict) and self.spaces == other.spaces

    def __repr__(self):
        return "Dict(" + ", ".join(f"{k!r}: {s}" for k, s in self.spaces.items()) + ")"


class Tuple(Space):
    def __init__(self, spaces, seed=None):
        self.spaces = tuple(spaces)
        super().__init__(None, None, seed)

    def contains(self, x):
        return isinstance(x, (tuple, list)) and len(x) == len(self.spaces) and all(a in s for a, s in zip(x, self.spaces))

    def sample(self, mask=None):
        return tuple(s.sample() for s in self.spaces)


# ------------------------------------------------------------------------------------------ spaces.utils
def flatdim(space):
    if isinstance(space, (Box, MultiBinary)):
        return int(np.prod(space.shape))
    if isinstance(space, Discrete):
        return int(space.n)
    if isinstance(space, MultiDiscrete):
        return int(np.sum(space.nvec))
    if isinstance(space, Dict):
        return sum(flatdim(s) for s in space.spaces.values())
    if isinstance(space, Tuple):
        return sum(flatdim(s) for s in space.spaces)
    raise NotImplementedError(f"Unknown space: `{space}`")


def flatten(space, x):
    if isinstance(space, (Box, MultiBinary)):
        return np.asarray(x, dtype=space.dtype).flatten()
    if isinstance(space, Discrete):
        onehot = np.zeros(space.n, dtype=space.dtype)
        onehot[x - space.start] = 1
        return onehot
    if isinstance(space, MultiDiscrete):
        offsets = np.zeros((space.nvec.size + 1,), dtype=space.dtype)
        offsets[1:] = np.cumsum(space.nvec.flatten())
        onehot = np.zeros((offsets[-1],), dtype=space.dtype)
        onehot[offsets[:-1] + np.asarray(x).flatten()] = 1
        return onehot
    if isinstance(space, Dict):
        return np.concatenate([np.array(flatten(s, x[k])) for k, s in space.spaces.items()])
    if isinstance(space, Tuple):
        return np.concatenate([np.array(flatten(s, xi)) for xi, s in zip(x, space.spaces)])
    raise NotImplementedError(f"Unknown space: `{space}`")


def unflatten(space, x):
    if isinstance(space, (Box, MultiBinary)):
        return np.asarray(x, dtype=space.dtype).reshape(space.shape)
    if isinstance(space, Discrete):
        return int(space.start + np.nonzero(x)[0][0])
    if isinstance(space, MultiDiscrete):
        offsets = np.zeros((space.nvec.size + 1,), dtype=space.dtype)
        offsets[1:] = np.cumsum(space.nvec.flatten())
        (indices,) = np.nonzero(x)
        return np.asarray(indices - offsets[:-1], dtype=space.dtype).reshape(space.shape)
    if isinstance(space, Dict):
        dims = np.asarray([flatdim(s) for s in space.spaces.values()], dtype=np.int_)
        parts = np.split(x, np.cumsum(dims[:-1]))
        return OrderedDict((k, unflatten(s, p)) for p, (k, s) in zip(parts, space.spaces.items()))
    raise NotImplementedError(f"Unknown space: `{space}`")


def flatten_space(space):
    if isinstance(space, Box):
        return Box(space.low.flatten(), space.high.flatten(), dtype=space.dtype)
    if isinstance(space, (Discrete, MultiBinary, MultiDiscrete)):
        return Box(low=0, high=1, shape=(flatdim(space),), dtype=space.dtype)
    if isinstance(space, (Dict, Tuple)):
        subs = space.spaces.values() if isinstance(space, Dict) else space.spaces
        flat = [flatten_space(s) for s in subs]
        return Box(low=np.concatenate([s.low for s in flat]), high=np.concatenate([s.high for s in flat]),
                   dtype=np.result_type(*[s.dtype for s in flat]))
    raise NotImplementedError(f"Unknown space: `{space}`")


# ------------------------------------------------------------------------------------------ core
class Env:
    metadata = {"render_modes": []}
    render_mode = None
    reward_range = (-float("inf"), float("inf"))
    spec = None
    action_space = None
    observation_space = None
    _np_random = None

    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random = np.random.default_rng()
        return self._np_random

    def step(self, action):
        raise NotImplementedError

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self._np_random = np.random.default_rng(seed)

    def render(self):
        raise NotImplementedError

    def close(self):
        pass

    @property
    def unwrapped(self):
        return self

    def __enter__(self):
        return self

    def __exit__(self, *args):
        self.close()
        return False


class Wrapper(Env):
    def __init__(self, env):
        self.env = env
        self._action_space = None
        self._observation_space = None
        self._reward_range = None
        self._metadata = None

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(f"accessing private attribute '{name}' is prohibited")
        return getattr(self.env, name)

    @property
    def spec(self):
        return self.env.spec

    @classmethod
    def class_name(cls):
        return cls.__name__

    @property
    def action_space(self):
        return self.env.action_space if self._action_space is None else self._action_space

    @action_space.setter
    def action_space(self, space):
        self._action_space = space

    @property
    def observation_space(self):
        return self.env.observation_space if self._observation_space is None else self._observation_space

    @observation_space.setter
    def observation_space(self, space):
        self._observation_space = space

    @property
    def reward_range(self):
        return self.env.reward_range if self._reward_range is None else self._reward_range

    @reward_range.setter
    def reward_range(self, value):
        self._reward_range = value

    @property
    def metadata(self):
        return self.env.metadata if self._metadata is None else self._metadata

    @metadata.setter
    def metadata(self, value):
        self._metadata = value

    @property
    def render_mode(self):
        return self.env.render_mode

    @property
    def np_random(self):
        return self.env.np_random

    def step(self, action):
        return self.env.step(action)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def render(self, *args, **kwargs):
        return self.env.render(*args, **kwargs)

    def close(self):
        return self.env.close()

    def __str__(self):
        return f"<{type(self).__name__}{self.env}>"

    __repr__ = __str__

    @property
    def unwrapped(self):
        return self.env.unwrapped


class ObservationWrapper(Wrapper):
    def reset(self, **kwargs):
        obs, info = self.env.reset(**kwargs)
        return self.observation(obs), info

    def step(self, action):
        observation, reward, terminated, truncated, info = self.env.step(action)
        return self.observation(observation), reward, terminated, truncated, info

    def observation(self, observation):
        raise NotImplementedError


class RewardWrapper(Wrapper):
    def step(self, action):
        observation, reward, terminated, truncated, info = self.env.step(action)
        return observation, self.reward(reward), terminated, truncated, info

    def reward(self, reward):
        raise NotImplementedError


class ActionWrapper(Wrapper):
    def step(self, action):
        return self.env.step(self.action(action))

    def action(self, action):
        raise NotImplementedError


# ------------------------------------------------------------------------------------------ wrappers
class FilterObservation(ObservationWrapper):
    def __init__(self, env, filter_keys=None):
        super().__init__(env)
        wrapped = env.observation_space
        assert isinstance(wrapped, Dict), f"FilterObservationWrapper is only usable with dict observations, environment observation space is {type(wrapped)}"
        observation_keys = wrapped.spaces.keys()
        if filter_keys is None:
            filter_keys = tuple(observation_keys)
        missing = {k for k in filter_keys if k not in observation_keys}
        if missing:
            raise ValueError(f"All the filter_keys must be included in the original observation space.\n"
                             f"Filter keys: {filter_keys}\nObservation keys: {observation_keys}\nMissing keys: {missing}")
        self.observation_space = type(wrapped)([(name, deepcopy(space)) for name, space in wrapped.spaces.items()
                                                if name in filter_keys])
        self._env = env
        self._filter_keys = tuple(filter_keys)

    def observation(self, observation):
        return type(observation)([(name, value) for name, value in observation.items() if name in self._filter_keys])


class FlattenObservation(ObservationWrapper):
    def __init__(self, env):
        super().__init__(env)
        self.observation_space = flatten_space(env.observation_space)

    def observation(self, observation):
        return flatten(self.env.observation_space, observation)


class TransformReward(RewardWrapper):
    def __init__(self, env, f):
        super().__init__(env)
        assert callable(f)
        self.f = f

    def reward(self, reward):
        return self.f(reward)


class TransformObservation(ObservationWrapper):
    def __init__(self, env, f):
        super().__init__(env)
        assert callable(f)
        self.f = f

    def observation(self, observation):
        return self.f(observation)


# ------------------------------------------------------------------------------------------ registry
registry: dict = {}


def register(id, entry_point=None, **kwargs):  # noqa: A002
    registry[id] = {"entry_point": entry_point, **kwargs}


def make(id, **kwargs):  # noqa: A002
    import importlib
    ep = registry[id]["entry_point"]
    if isinstance(ep, str):
        mod, attr = ep.split(":")
        ep = getattr(importlib.import_module(mod), attr)
    return ep(**kwargs)


def install():
    """Register the stand-in under gymnasium's module names (idempotent)."""
    if isinstance(sys.modules.get("gymnasium"), types.ModuleType) and getattr(sys.modules["gymnasium"], "__mini_gym__", False):
        return sys.modules["gymnasium"]

    def module(name, **attrs):
        m = types.ModuleType(name)
        m.__path__ = []
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    space_names = dict(Space=Space, Box=Box, Discrete=Discrete, MultiBinary=MultiBinary, MultiDiscrete=MultiDiscrete,
                       Dict=Dict, Tuple=Tuple, flatdim=flatdim, flatten=flatten, unflatten=unflatten,
                       flatten_space=flatten_space)
    utils = module("gymnasium.spaces.utils", flatdim=flatdim, flatten=flatten, unflatten=unflatten,
                   flatten_space=flatten_space)
    spaces = module("gymnasium.spaces", utils=utils, **space_names)
    wrappers = module("gymnasium.wrappers", FilterObservation=FilterObservation, FlattenObservation=FlattenObservation,
                      TransformReward=TransformReward, TransformObservation=TransformObservation)
    wrappers.filter_observation = module("gymnasium.wrappers.filter_observation", FilterObservation=FilterObservation)
    registration = module("gymnasium.envs.registration", register=register, registry=registry, make=make)
    envs = module("gymnasium.envs", registration=registration, register=register, registry=registry)
    core = module("gymnasium.core", Env=Env, Wrapper=Wrapper, ObservationWrapper=ObservationWrapper,
                  RewardWrapper=RewardWrapper, ActionWrapper=ActionWrapper)
    # gymnasium.utils.env_checker.check_env: the reference's tests/environment/test_env.py calls it; here it only
    # resets and steps the env once with a sampled action
    def check_env(env, *a, **k):
        env.reset()
        env.step(env.action_space.sample())

    env_checker = module("gymnasium.utils.env_checker", check_env=check_env)
    gutils = module("gymnasium.utils", env_checker=env_checker)
    gym = module("gymnasium", utils=gutils, Env=Env, Wrapper=Wrapper, ObservationWrapper=ObservationWrapper,
                 RewardWrapper=RewardWrapper, ActionWrapper=ActionWrapper, Space=Space, spaces=spaces,
                 wrappers=wrappers, envs=envs, core=core, register=register, make=make, registry=registry,
                 __version__="0.26.3-mini", __mini_gym__=True)
    return gym
