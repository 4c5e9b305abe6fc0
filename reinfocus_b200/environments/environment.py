"""Single-env Gymnasium environment composed of six strategy objects
(reference environments/environment.py)."""

from typing import Any

import numpy

from reinfocus_b200 import gym_compat


class Environment(gym_compat.Env):
    # pylint: disable=too-many-instance-attributes
    """ender / initializer / observer / rewarder / transformer / visualizer, batch size 1
    (reference :19-140)."""

    metadata = {"render_modes": ["rgb_array"], "render_fps": 4}

    def __init__(self, ender, initializer, observer, rewarder, transformer, visualizer,
                 render_mode: str | None = None):
        # pylint: disable=too-many-arguments
        self._ender = ender
        self._initializer = initializer
        self._observer = observer
        self._rewarder = rewarder
        self._transformer = transformer
        self._visualizer = visualizer
        self.observation_space = observer.single_observation_space
        self.action_space = transformer.single_action_space
        assert render_mode is None or render_mode in self.metadata["render_modes"]
        self.render_mode = render_mode
        self._state = None

    def reset(self, *, seed: int | None = None, options: dict[str, Any] | None = None):
        super().reset(seed=seed)
        self._state = self._initializer.initialize(1)
        self._ender.reset(self._state)
        observations = self._observer.reset(self._state)
        self._rewarder.reset(self._state, observations)
        if self.render_mode == "rgb_array":
            self._visualizer.reset(self._state, observations)
        return observations[0], {}

    def step(self, action):
        assert self._state is not None
        self._state = self._transformer.transform(self._state, numpy.array([action]))
        self._ender.step(self._state)
        observations = self._observer.observe(self._state)
        if self.render_mode == "rgb_array":
            self._visualizer.step(self._state, observations)
        return (observations[0], self._rewarder.reward(self._state, observations)[0],
                self._ender.is_terminated()[0], self._ender.is_truncated()[0], {})

    def render(self):
        if self.render_mode == "rgb_array":
            return self._visualizer.visualize()
        return None
