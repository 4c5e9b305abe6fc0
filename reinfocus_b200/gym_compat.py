"""Gymnasium surface used by the env layer, with a built-in stand-in.

The reference depends on gymnasium ~=0.29 (``gymnasium.Env``,
``gymnasium.experimental.vector.VectorEnv``, ``gymnasium.spaces.{Box,Discrete,Space}``,
``gymnasium.vector.utils.batch_space``, ``gymnasium.envs.registration.register`` and
``gymnasium.make_vec``). gymnasium is not installable offline, and
1.x removed ``gymnasium.experimental``, so this module exports those names from the real
package when a compatible one is importable and from the minimal classes below otherwise.
Only the behaviour the env layer relies on is implemented (shapes, bounds, dtypes, seeding,
sampling, containment, batching, a registry)."""

from __future__ import annotations

import importlib
import sys
import types
from typing import Any

import numpy


# ----------------------------------------------------------------------------- stand-ins


class _Space:
    def __init__(self, shape=None, dtype=None, seed=None):
        self._shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else numpy.dtype(dtype)
        self._np_random = None
        if seed is not None:
            self.seed(seed)

    @property
    def shape(self):
        return self._shape

    @property
    def np_random(self):
        if self._np_random is None:
            self.seed()
        return self._np_random

    def seed(self, seed=None):
        self._np_random = numpy.random.Generator(numpy.random.PCG64(seed))
        return [seed]

    def sample(self):
        raise NotImplementedError

    def contains(self, x) -> bool:
        raise NotImplementedError

    def __contains__(self, x) -> bool:
        return self.contains(x)


class _Box(_Space):
    """Closed box in R^n. Scalar bounds without a shape give shape (1,), as gymnasium does."""

    def __init__(self, low, high, shape=None, dtype=numpy.float32, seed=None):
        if shape is not None:
            shape = tuple(int(s) for s in shape)
        elif isinstance(low, numpy.ndarray):
            shape = low.shape
        elif isinstance(high, numpy.ndarray):
            shape = high.shape
        else:
            shape = (1,)
        super().__init__(shape, dtype, seed)
        self.low = numpy.broadcast_to(numpy.asarray(low, dtype=self.dtype), shape).copy()
        self.high = numpy.broadcast_to(numpy.asarray(high, dtype=self.dtype), shape).copy()

    def sample(self):
        high = self.high if self.dtype.kind == "f" else self.high.astype(numpy.int64) + 1
        sample = self.np_random.uniform(self.low, high, size=self._shape)
        if self.dtype.kind in "iu":
            sample = numpy.floor(sample)
        return sample.astype(self.dtype)

    def contains(self, x) -> bool:
        x = numpy.asarray(x)
        return bool(x.shape == self._shape and numpy.all(x >= self.low) and numpy.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low}, {self.high}, {self._shape}, {self.dtype})"

    def __eq__(self, other):
        return (isinstance(other, _Box) and self._shape == other.shape
                and numpy.array_equal(self.low, other.low)
                and numpy.array_equal(self.high, other.high))


class _Discrete(_Space):
    def __init__(self, n: int, seed=None, start: int = 0):
        super().__init__((), numpy.int64, seed)
        self.n = int(n)
        self.start = int(start)

    def sample(self):
        return numpy.int64(self.start + self.np_random.integers(self.n))

    def contains(self, x) -> bool:
        try:
            value = int(x)
        except (TypeError, ValueError):
            return False
        return self.start <= value < self.start + self.n

    def __repr__(self):
        return f"Discrete({self.n})"

    def __eq__(self, other):
        return isinstance(other, _Discrete) and self.n == other.n and self.start == other.start


class _MultiDiscrete(_Space):
    def __init__(self, nvec, dtype=numpy.int64, seed=None, start=None):
        self.nvec = numpy.asarray(nvec, dtype=dtype)
        self.start = numpy.zeros_like(self.nvec) if start is None else numpy.asarray(start, dtype=dtype)
        super().__init__(self.nvec.shape, dtype, seed)

    def sample(self):
        return (self.start + self.np_random.random(self.nvec.shape) * self.nvec).astype(self.dtype)

    def contains(self, x) -> bool:
        x = numpy.asarray(x)
        return bool(x.shape == self._shape and numpy.all(x >= self.start)
                    and numpy.all(x < self.start + self.nvec))

    def __repr__(self):
        return f"MultiDiscrete({self.nvec})"


def _batch_space(space, n: int = 1):
    """gymnasium.vector.utils.batch_space for Box and Discrete."""

    if isinstance(space, _Box):
        repeats = (n,) + (1,) * len(space.shape)
        return _Box(numpy.tile(space.low, repeats), numpy.tile(space.high, repeats), dtype=space.dtype)
    if isinstance(space, _Discrete):
        return _MultiDiscrete(numpy.full((n,), space.n, dtype=numpy.int64),
                              start=numpy.full((n,), space.start, dtype=numpy.int64))
    raise TypeError(f"cannot batch {space!r}")


class _Env:
    metadata: dict[str, Any] = {"render_modes": []}
    render_mode = None
    spec = None
    observation_space = None
    action_space = None
    _np_random = None

    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random = numpy.random.Generator(numpy.random.PCG64())
        return self._np_random

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self._np_random = numpy.random.Generator(numpy.random.PCG64(seed))

    def step(self, action):
        raise NotImplementedError

    def render(self):
        return None

    def close(self):
        pass

    @property
    def unwrapped(self):
        return self


class _VectorEnv:
    """gymnasium.experimental.vector.VectorEnv as far as the reference uses it."""

    metadata: dict[str, Any] = {"render_modes": []}
    spec = None
    render_mode = None
    closed = False
    num_envs = 1
    observation_space = None
    action_space = None
    single_observation_space = None
    single_action_space = None
    _np_random = None

    def __init__(self):
        pass

    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random = numpy.random.Generator(numpy.random.PCG64())
        return self._np_random

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self._np_random = numpy.random.Generator(numpy.random.PCG64(seed))

    def step(self, actions):
        raise NotImplementedError

    def render(self):
        return None

    def close(self, **kwargs):
        self.closed = True

    @property
    def unwrapped(self):
        return self


class _EnvSpec:
    def __init__(self, id, entry_point, vector_entry_point=None, max_episode_steps=None, kwargs=None):
        # pylint: disable=redefined-builtin,too-many-arguments
        self.id = id
        self.entry_point = entry_point
        self.vector_entry_point = vector_entry_point
        self.max_episode_steps = max_episode_steps
        self.kwargs = dict(kwargs or {})


_registry: dict[str, _EnvSpec] = {}


def _load_entry_point(entry_point):
    if callable(entry_point):
        return entry_point
    module_name, attr = entry_point.split(":")
    return getattr(importlib.import_module(module_name), attr)


def _register(id, entry_point=None, vector_entry_point=None, max_episode_steps=None, **kwargs):
    # pylint: disable=redefined-builtin
    _registry[id] = _EnvSpec(id, entry_point, vector_entry_point, max_episode_steps,
                             kwargs.get("kwargs"))


class _TimeLimit:
    """Truncates single-env episodes after max_episode_steps (gymnasium.wrappers.TimeLimit)."""

    def __init__(self, env, max_episode_steps: int):
        self.env = env
        self._max_episode_steps = max_episode_steps
        self._elapsed = 0

    def __getattr__(self, name):
        return getattr(self.env, name)

    def reset(self, **kwargs):
        self._elapsed = 0
        return self.env.reset(**kwargs)

    def step(self, action):
        obs, reward, terminated, truncated, info = self.env.step(action)
        self._elapsed += 1
        if self._elapsed >= self._max_episode_steps:
            truncated = True
        return obs, reward, terminated, truncated, info

    @property
    def unwrapped(self):
        return self.env


def _make(id, **kwargs):
    # pylint: disable=redefined-builtin
    spec = _registry[id] if isinstance(id, str) else id
    env = _load_entry_point(spec.entry_point)(**{**spec.kwargs, **kwargs})
    env.spec = spec
    if spec.max_episode_steps is not None:
        env = _TimeLimit(env, spec.max_episode_steps)
    return env


def _make_vec(id, num_envs: int = 1, vectorization_mode: str = "custom", vector_kwargs=None, **kwargs):
    # pylint: disable=redefined-builtin
    spec = _registry[id] if isinstance(id, str) else id
    assert vectorization_mode == "custom" and spec.vector_entry_point is not None, (
        "the stand-in registry only builds custom vector envs (vector_entry_point)")
    call_kwargs = {**spec.kwargs, **kwargs, **(vector_kwargs or {}), "num_envs": num_envs}
    if spec.max_episode_steps is not None:
        call_kwargs.setdefault("max_episode_steps", spec.max_episode_steps)
    env = _load_entry_point(spec.vector_entry_point)(**call_kwargs)
    env.spec = spec
    return env


# ------------------------------------------------------------------------ name resolution


def _real_gymnasium():
    try:
        import gymnasium
        from gymnasium.experimental import vector as experimental_vector  # gymnasium 0.29 only
        from gymnasium.vector import utils
    except ImportError:
        return None
    return gymnasium, experimental_vector, utils


_real = _real_gymnasium()
USING_REAL_GYMNASIUM = _real is not None

if USING_REAL_GYMNASIUM:
    _gym, _exp_vector, _utils = _real
    spaces = _gym.spaces
    Env = _gym.Env
    VectorEnv = _exp_vector.VectorEnv
    batch_space = _utils.batch_space
    register = _gym.envs.registration.register
    make = _gym.make
    make_vec = _gym.make_vec
else:
    spaces = types.SimpleNamespace(Space=_Space, Box=_Box, Discrete=_Discrete,
                                   MultiDiscrete=_MultiDiscrete)
    Env = _Env
    VectorEnv = _VectorEnv
    batch_space = _batch_space
    register = _register
    make = _make
    make_vec = _make_vec


def install_as_gymnasium() -> None:
    """Registers the stand-ins under the ``gymnasium`` module names (harness use: lets code
    written against gymnasium 0.29 import in a container without it)."""

    if "gymnasium" in sys.modules:
        return

    def module(name, **attrs):
        mod = types.ModuleType(name)
        mod.__dict__.update(attrs)
        sys.modules[name] = mod
        return mod

    spaces_mod = module("gymnasium.spaces", Space=_Space, Box=_Box, Discrete=_Discrete,
                        MultiDiscrete=_MultiDiscrete)
    utils_mod = module("gymnasium.vector.utils", batch_space=_batch_space)
    vector_mod = module("gymnasium.vector", utils=utils_mod, VectorEnv=_VectorEnv)
    exp_vector_mod = module("gymnasium.experimental.vector", VectorEnv=_VectorEnv)
    exp_mod = module("gymnasium.experimental", vector=exp_vector_mod)
    registration_mod = module("gymnasium.envs.registration", register=_register, registry=_registry,
                              EnvSpec=_EnvSpec)
    envs_mod = module("gymnasium.envs", registration=registration_mod)
    module("gymnasium", spaces=spaces_mod, vector=vector_mod, experimental=exp_mod, envs=envs_mod,
           Env=_Env, make=_make, make_vec=_make_vec, register=_register)
