"""Single-env Gymnasium environment composed of six strategy objects
(reference environments/environment.py). Internally it is a batch of one: the strategies
are the vector ones, and this class unwraps their first row."""

from typing import Any

import numpy

from reinfocus_b200 import gym_compat


def _first(*batched):
    return tuple(values[0] for values in batched)


class Environment(gym_compat.Env):
    # pylint: disable=too-many-instance-attributes
    """ender / initializer / observer / rewarder / transformer / visualizer, batch size 1
    (reference :19-140). Unlike the vector env it does not restart finished episodes: the
    caller resets."""

    metadata = {"render_modes": ["rgb_array"], "render_fps": 4}

    def __init__(self, ender, initializer, observer, rewarder, transformer, visualizer,
                 render_mode: str | None = None):
        # pylint: disable=too-many-arguments
        self._ender, self._initializer, self._observer = ender, initializer, observer
        self._rewarder, self._transformer, self._visualizer = rewarder, transformer, visualizer
        self.action_space = transformer.single_action_space
        self.observation_space = observer.single_observation_space
        assert render_mode in (None, *self.metadata["render_modes"])
        self.render_mode = render_mode
        self._state = None

    @property
    def _drawing(self) -> bool:
        return self.render_mode == "rgb_array"

    def reset(self, *, seed: int | None = None, options: dict[str, Any] | None = None):
        super().reset(seed=seed)
        state = self._state = self._initializer.initialize(1)
        self._ender.reset(state)
        observations = self._observer.reset(state)
        self._rewarder.reset(state, observations)
        if self._drawing:
            self._visualizer.reset(state, observations)
        return observations[0], {}

    def step(self, action):
        assert self._state is not None
        state = self._state = self._transformer.transform(self._state, numpy.array([action]))
        self._ender.step(state)
        observations = self._observer.observe(state)
        if self._drawing:
            self._visualizer.step(state, observations)
        rewards = self._rewarder.reward(state, observations)
        return (*_first(observations, rewards, self._ender.is_terminated(), self._ender.is_truncated()), {})

    def render(self):
        return self._visualizer.visualize() if self._drawing else None
