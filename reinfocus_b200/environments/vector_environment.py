"""Vector env: a batch of envs stepped together, with same-step auto-reset of finished
episodes (reference environments/vector_environment.py)."""

from typing import Any

import numpy

from reinfocus_b200 import gym_compat


class VectorEnvironment(gym_compat.VectorEnv):
    # pylint: disable=too-many-instance-attributes
    """(reference :19-176)"""

    metadata = {"render_modes": ["rgb_array"], "render_fps": 4}

    def __init__(self, ender, initializer, observer, rewarder, transformer, visualizer,
                 num_envs: int = 2, render_mode: str | None = None):
        # pylint: disable=too-many-arguments
        super().__init__()
        self._ender = ender
        self._initializer = initializer
        self._observer = observer
        self._rewarder = rewarder
        self._transformer = transformer
        self._visualizer = visualizer
        self.num_envs = num_envs
        self.action_space = transformer.action_space
        self.observation_space = observer.observation_space
        self.single_action_space = transformer.single_action_space
        self.single_observation_space = observer.single_observation_space
        assert render_mode is None or render_mode in self.metadata["render_modes"]
        self.render_mode = render_mode
        self._state = None

    def reset(self, *, seed: int | None = None, options: dict[str, Any] | None = None):
        super().reset(seed=seed)
        self._state = self._initializer.initialize(self.num_envs)
        self._ender.reset(self._state)
        observations = self._observer.reset(self._state, None)
        self._rewarder.reset(self._state, observations)
        if self.render_mode == "rgb_array":
            self._visualizer.reset(self._state, observations)
        return observations, {}

    def step(self, actions):
        assert self._state is not None
        self._state = self._transformer.transform(self._state, actions)
        self._ender.step(self._state)
        observations = self._observer.observe(self._state)
        # rewards come from the pre-reset observations (reference :128-130 precede :137-146)
        rewards = self._rewarder.reward(self._state, observations)
        terminated = self._ender.is_terminated()
        truncated = self._ender.is_truncated()
        done = terminated | truncated
        if any(done):
            new_state = self._initializer.initialize(done.sum())
            self._state[done] = new_state
            self._ender.reset(new_state, done)
            # a second, k-env render in the same step: batch positions 0..k-1
            new_observations = self._observer.reset(new_state, done)
            observations[done] = new_observations
            self._rewarder.reset(new_state, new_observations, done)
            if self.render_mode == "rgb_array":
                self._visualizer.reset(new_state, new_observations, done)
        if self.render_mode == "rgb_array":
            not_done = ~done
            self._visualizer.step(self._state[not_done], observations[not_done], not_done)
        return observations, rewards, terminated, truncated, {}

    def render(self):
        if self.render_mode == "rgb_array":
            return self._visualizer.visualize()
        return None
