"""Adapter between this package's gymnasium-style vector envs and the stable-baselines3
``VecEnv`` protocol (reference reinfocus/environments/experimental/vector_shim.py:20-229).

gymnasium vector envs return ``(obs, reward, terminated, truncated, info_dict)`` from one
``step``; stable-baselines3 wants ``step_async`` / ``step_wait`` returning ``(obs, reward,
done, [info per env])`` with the last observation of a finished episode under
``"terminal_observation"``. stable-baselines3 is optional: with it installed ``SB3Wrapper``
is a real ``VecEnv`` (usable by its algorithms and by ``VecMonitor``); without it the same
class stands on a minimal base with the same attribute surface, so rollout code written
against the protocol (``examples/ppo.py``) runs unchanged.
"""

from collections.abc import Iterable
from typing import Any

import numpy

from reinfocus_b200 import gym_compat
from reinfocus_b200.environments import vector_environment

try:  # pragma: no cover - depends on the installation
    from stable_baselines3.common import monitor as _sb3_monitor
    from stable_baselines3.common import vec_env as _sb3_vec_env
    from stable_baselines3.common.vec_env import base_vec_env as _sb3_base
    from stable_baselines3.common.vec_env import vec_monitor as _sb3_vec_monitor

    HAVE_SB3 = True
    _VecEnvBase = _sb3_base.VecEnv
except ImportError:
    HAVE_SB3 = False

    class _VecEnvBase:  # type: ignore[no-redef]
        """What ``stable_baselines3.common.vec_env.VecEnv.__init__`` and ``step`` provide."""

        def __init__(self, num_envs: int, observation_space, action_space):
            self.num_envs = num_envs
            self.observation_space = observation_space
            self.action_space = action_space
            self.reset_infos: list[dict[str, Any]] = [{} for _ in range(num_envs)]
            self._seeds: list[int | None] = [None] * num_envs
            self._options: list[dict[str, Any]] = [{} for _ in range(num_envs)]

        def step(self, actions):
            self.step_async(actions)
            return self.step_wait()


class SB3Wrapper(_VecEnvBase):
    """Presents a :class:`VectorEnvironment` as a stable-baselines3 ``VecEnv``."""

    def __init__(self, env, render_mode: str | None):
        if not isinstance(env, vector_environment.VectorEnvironment):
            raise NotImplementedError

        self._env = env
        super().__init__(env.num_envs, env.single_observation_space, env.single_action_space)
        self.render_mode = render_mode
        self._actions = None

    @property
    def unwrapped_vector_env(self):
        """The wrapped gymnasium-style vector env."""

        return self._env

    def reset(self):
        return self._env.reset()[0]

    def step_async(self, actions: numpy.ndarray):
        self._actions = actions

    def step_wait(self):
        assert self._actions is not None

        obs, rewards, terminated, truncated, info = self._env.step(self._actions)
        dones = terminated | truncated
        per_env_keys = [key for key, value in info.items() if isinstance(value, numpy.ndarray)]
        infos: list[dict[str, Any]] = [
            {key: info[key][i] for key in per_env_keys} for i in range(self.num_envs)
        ]
        # the env auto-resets inside step, so obs[i] of a finished env is already the first
        # observation of its next episode -- the reference hands that same row out as the
        # terminal observation (vector_shim.py:86-87), and so does this
        for i in numpy.flatnonzero(dones):
            infos[i]["terminal_observation"] = obs[i]
        return obs, rewards, dones, infos

    def close(self):
        self._env.close()

    def get_attr(self, attr_name: str, indices=None) -> list:
        if hasattr(self._env, attr_name):
            return [getattr(self._env, attr_name)] * self._count(indices)
        raise NotImplementedError(f"{attr_name}, {indices}")

    def set_attr(self, attr_name: str, value: Any, indices=None):
        raise NotImplementedError(f"{attr_name}, {value}, {indices}")

    def env_method(self, method_name: str, *method_args, indices=None, **method_kwargs) -> list[Any]:
        raise NotImplementedError(f"{method_name}, {method_args}, {indices}, {method_kwargs}")

    def env_is_wrapped(self, wrapper_class, indices=None) -> list[bool]:
        return [False] * self._count(indices)

    def get_images(self):
        return [self._env.render()]

    def _count(self, indices) -> int:
        if isinstance(indices, int):
            return 1
        if isinstance(indices, Iterable):
            return len(list(indices))
        return self._env.num_envs


def rewrapper(naive_vec_env):
    """``vec_env_wrapper`` hook for rl_zoo3: swaps the ``DummyVecEnv`` of n single envs that
    stable-baselines3 builds for one custom vector env of the same spec (one render launch
    per step instead of n), keeping an outer ``Monitor`` as a ``VecMonitor``."""

    if not HAVE_SB3 or not isinstance(naive_vec_env, _sb3_vec_env.DummyVecEnv):
        return naive_vec_env

    first = naive_vec_env.envs[0]
    if first.spec is None:
        return naive_vec_env

    vector_kwargs: dict[str, Any] = {}
    if first.spec.max_episode_steps is not None:
        vector_kwargs["max_episode_steps"] = first.spec.max_episode_steps
    render_mode = "rgb_array" if first.render_mode == "human" else None
    vector_kwargs["render_mode"] = render_mode

    wrapped = SB3Wrapper(
        gym_compat.make_vec(first.spec, naive_vec_env.num_envs, vectorization_mode="custom",
                            vector_kwargs=vector_kwargs),
        render_mode,
    )
    if isinstance(first, _sb3_monitor.Monitor):
        return _sb3_vec_monitor.VecMonitor(wrapped, first.EXT, first.info_keywords)
    return wrapped
