"""Vector env: a batch of envs stepped together, with same-step auto-reset of finished
episodes (reference environments/vector_environment.py). The strategy objects (ender,
initializer, observer, rewarder, transformer, visualizer) do the work; this class only
sequences them - and the sequence is part of the parity contract:

    transform -> ender.step -> observe (render all n) -> reward (pre-reset observations)
    -> for the finished envs: new states, ender.reset, observe again (render k envs as batch
       positions 0..k-1), rewarder.reset -> visualizer bookkeeping
"""

from typing import Any

import numpy

from reinfocus_b200 import gym_compat

_STRATEGIES = ("ender", "initializer", "observer", "rewarder", "transformer", "visualizer")


class VectorEnvironment(gym_compat.VectorEnv):
    # pylint: disable=too-many-instance-attributes
    """(reference :19-176)"""

    metadata = {"render_modes": ["rgb_array"], "render_fps": 4}

    def __init__(self, ender, initializer, observer, rewarder, transformer, visualizer,
                 num_envs: int = 2, render_mode: str | None = None):
        # pylint: disable=too-many-arguments
        super().__init__()
        for name, strategy in zip(_STRATEGIES, (ender, initializer, observer, rewarder, transformer,
                                                visualizer)):
            setattr(self, f"_{name}", strategy)
        self.num_envs = num_envs
        self.single_action_space, self.action_space = (transformer.single_action_space,
                                                       transformer.action_space)
        self.single_observation_space, self.observation_space = (observer.single_observation_space,
                                                                 observer.observation_space)
        assert render_mode in (None, *self.metadata["render_modes"])
        self.render_mode = render_mode
        self._state = None

    @property
    def _drawing(self) -> bool:
        return self.render_mode == "rgb_array"

    def reset(self, *, seed: int | None = None, options: dict[str, Any] | None = None):
        super().reset(seed=seed)
        self._state = self._initializer.initialize(self.num_envs)
        observations = self._begin_episodes(self._state, None)
        return observations, {}

    def step(self, actions):
        assert self._state is not None
        self._state = self._transformer.transform(self._state, actions)
        self._ender.step(self._state)
        observations = self._observer.observe(self._state)
        # rewards come from the pre-reset observations (reference :128-130 precede :137-146)
        rewards = self._rewarder.reward(self._state, observations)
        terminated, truncated = self._ender.is_terminated(), self._ender.is_truncated()
        finished = terminated | truncated
        if finished.any():
            fresh = self._initializer.initialize(finished.sum())
            self._state[finished] = fresh
            # a second, k-env render in the same step: batch positions 0..k-1
            observations[finished] = self._begin_episodes(fresh, finished)
        if self._drawing:
            running = ~finished
            self._visualizer.step(self._state[running], observations[running], running)
        return observations, rewards, terminated, truncated, {}

    def render(self):
        return self._visualizer.visualize() if self._drawing else None

    def _begin_episodes(self, states, which):
        """First observations of the episodes starting in ``states`` (all envs if ``which`` is
        None, else the masked ones), with every strategy told about the restart."""

        if which is None:
            self._ender.reset(states)
            observations = self._observer.reset(states, None)
            self._rewarder.reset(states, observations)
            if self._drawing:
                self._visualizer.reset(states, observations)
        else:
            self._ender.reset(states, which)
            observations = self._observer.reset(states, which)
            self._rewarder.reset(states, observations, which)
            if self._drawing:
                self._visualizer.reset(states, observations, which)
        return observations
