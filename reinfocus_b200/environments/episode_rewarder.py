"""Rewarders (reference environments/episode_rewarder.py). They combine with ``+`` and
``*``; the arithmetic (dtype promotion included) follows the reference expression by
expression, because the reward sequence is part of the parity contract."""

from __future__ import annotations

from collections.abc import Callable
from typing import Protocol

import numpy
from numpy.typing import NDArray


class IEpisodeRewarder(Protocol):
    """The interface rewarders follow (reference episode_rewarder.py:15-60)."""

    def reset(self, states, observations, indices: NDArray[numpy.bool_] | None = None):
        ...

    def reward(self, states, observations) -> NDArray[numpy.float32]:
        ...


class BaseRewarder:
    def __add__(self, other) -> BaseRewarder:
        return OpRewarder(self, other, numpy.add)

    def __mul__(self, other) -> BaseRewarder:
        return OpRewarder(self, other, numpy.multiply)

    def reset(self, states, observations, indices: NDArray[numpy.bool_] | None = None):
        """Called with the first states/observations of restarted episodes."""

    def reward(self, states, observations) -> NDArray[numpy.float32]:
        raise NotImplementedError


class _RemembersElement(BaseRewarder):
    """Keeps the previous step's value of one state element. The reference stores a *view*
    of the state column (episode_rewarder.py:128,154) that the env later overwrites for
    reset envs (vector_environment.py:140); the net effect - previous value = the element at
    the end of the previous step, new episodes included - is what is kept here, on a copy."""

    def __init__(self, check_index: int):
        self._check_index = check_index
        self._old_states = None

    def reset(self, states, observations, indices=None):
        column = states[:, self._check_index]
        if self._old_states is not None and indices is not None:
            self._old_states[indices] = column
        else:
            self._old_states = column.copy()

    def _swap(self, states):
        assert self._old_states is not None
        previous = self._old_states
        self._old_states = states[:, self._check_index].copy()
        return previous


class DeltaRewarder(_RemembersElement):
    # pylint: disable=too-few-public-methods
    """``reward`` (default -1) per ``scale`` the element moved (reference :86-156)."""

    def __init__(self, check_index: int, scale: float, reward: float = -1.0):
        super().__init__(check_index)
        self._scale = scale
        self._reward = reward

    def reward(self, states, observations):
        previous = self._swap(states)
        return abs(states[:, self._check_index] - previous) * self._reward / self._scale


class DistanceRewarder(BaseRewarder):
    # pylint: disable=too-few-public-methods
    """Linear in the distance between two elements: ``high`` at 0, ``low`` at ``span``
    (reference :159-207)."""

    def __init__(self, check_indices: tuple[int, int], span: float, low: float = -1.0,
                 high: float = 0.0):
        self._check_indices = check_indices
        self._span = span
        self._low = low
        self._high = high

    def reward(self, states, observations):
        gap = abs(states[:, self._check_indices[0]] - states[:, self._check_indices[1]])
        return (1 - gap / self._span) * (self._high - self._low) + self._low


class ObservationRewarder(BaseRewarder):
    # pylint: disable=too-few-public-methods
    """One element of the observation is the reward (reference :210-238)."""

    def __init__(self, reward_observation_index: int):
        self._reward_observation_index = reward_observation_index

    def reward(self, states, observations):
        return observations[:, self._reward_observation_index]


class OnTargetRewarder(BaseRewarder):
    # pylint: disable=too-few-public-methods
    """``on`` while two elements are within ``span`` of each other, else ``off``
    (reference :241-292)."""

    def __init__(self, check_indices: tuple[int, int], span: float, off: float = 0.0,
                 on: float = 1.0):
        self._check_indices = check_indices
        self._span = span
        self._off = off
        self._delta = on - off

    def reward(self, states, observations):
        gap = abs(states[:, self._check_indices[0]] - states[:, self._check_indices[1]])
        return (gap < self._span) * self._delta + self._off


class OpRewarder(BaseRewarder):
    """Combines two rewarders with an elementwise operation (reference :295-358)."""

    def __init__(self, l_rewarder, r_rewarder,
                 op: Callable[[NDArray[numpy.float32], NDArray[numpy.float32]], NDArray[numpy.float32]]):
        self._l_rewarder = l_rewarder
        self._r_rewarder = r_rewarder
        self._op = op

    def reset(self, states, observations, indices=None):
        self._l_rewarder.reset(states, observations, indices)
        self._r_rewarder.reset(states, observations, indices)

    def reward(self, states, observations):
        return self._op(self._l_rewarder.reward(states, observations),
                        self._r_rewarder.reward(states, observations))


class StoppedRewarder(_RemembersElement):
    # pylint: disable=too-few-public-methods
    """``reward`` whenever the element moved less than ``threshold`` (reference :361-429)."""

    def __init__(self, check_index: int, threshold: float, reward: float = 1.0):
        super().__init__(check_index)
        self._threshold = abs(threshold)
        self._reward = reward

    def reward(self, states, observations):
        previous = self._swap(states)
        return (abs(states[:, self._check_index] - previous) < self._threshold) * self._reward
