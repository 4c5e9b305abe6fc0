"""Action -> state transformers (reference environments/state_transformer.py)."""

import abc
from collections.abc import Collection
from typing import Protocol

import numpy
from numpy.typing import NDArray

from reinfocus_b200 import gym_compat

spaces = gym_compat.spaces


class IStateTransformer(Protocol):
    # pylint: disable=too-few-public-methods
    """The interface transformers follow (reference state_transformer.py:18-38)."""

    action_space: object
    single_action_space: object

    def transform(self, states, actions):
        ...


class StateTransformer(abc.ABC):
    # pylint: disable=too-few-public-methods
    """Transformers for POMDPs sharing one action space (reference :57-85)."""

    def __init__(self, num_envs: int, single_action_space):
        self.single_action_space = single_action_space
        self.action_space = gym_compat.batch_space(single_action_space, num_envs)

    @abc.abstractmethod
    def transform(self, states: NDArray[numpy.float32], actions: NDArray) -> NDArray[numpy.float32]:
        """The states that result from taking ``actions`` in ``states``."""


class ContinuousJumpTransformer(StateTransformer):
    # pylint: disable=too-few-public-methods
    """Actions in [-1, 1] place one state element proportionally within ``limits``; jumps
    shorter than ``stop_threshold`` are ignored (reference :88-137)."""

    def __init__(self, num_envs: int, move_index: int, limits: tuple[float, float],
                 stop_threshold: float = 0.1):
        super().__init__(num_envs, spaces.Box(-1, 1, dtype=numpy.float32))
        self._limits = limits
        self._move_index = move_index
        self._stop_threshold = abs(stop_threshold)

    def transform(self, states, actions):
        result = states.copy()
        fraction = (actions.flatten() + 1) / 2.0
        destinations = fraction * (self._limits[1] - self._limits[0]) + self._limits[0]
        far_enough = abs(result[:, self._move_index] - destinations) > self._stop_threshold
        result[far_enough, self._move_index] = destinations[far_enough]
        return result


class ContinuousMoveTransformer(StateTransformer):
    # pylint: disable=too-few-public-methods
    """Actions in [-1, 1] move one state element by ``action * speed`` (reference :140-192)."""

    def __init__(self, num_envs: int, move_index: int, limits: tuple[float, float], speed: float,
                 stop_threshold: float = 0.1):
        # pylint: disable=too-many-arguments
        super().__init__(num_envs, spaces.Box(-1, 1, dtype=numpy.float32))
        self._limits = limits
        self._move_index = move_index
        self._speed = speed
        self._stop_threshold = abs(stop_threshold)

    def transform(self, states, actions):
        result = states.copy()
        moves = numpy.clip(actions.flatten(), -1, 1) * self._speed
        result[:, self._move_index] += (abs(moves) > self._stop_threshold) * moves
        return numpy.clip(result, *self._limits)


class DiscreteJumpTransformer(StateTransformer):
    # pylint: disable=too-few-public-methods
    """Each discrete action sets one state element to a fixed position (reference :195-219)."""

    def __init__(self, num_envs: int, move_index: int, limits: tuple[float, float],
                 action_set: Collection[float]):
        super().__init__(num_envs, spaces.Discrete(len(action_set)))
        self._limits = limits
        self._move_index = move_index
        self._action_set = numpy.asarray(action_set, dtype=numpy.float32)

    def transform(self, states, actions):
        result = states.copy()
        result[:, self._move_index] = self._action_set[numpy.asarray(actions).flatten()]
        return numpy.clip(result, *self._limits)


class DiscreteMoveTransformer(StateTransformer):
    # pylint: disable=too-few-public-methods
    """Each discrete action moves one state element by a fixed distance, clipped to
    ``limits`` (reference :222-266). The action set keeps its float64 dtype, as in the
    reference: the in-place add rounds state + move to float32 once."""

    def __init__(self, num_envs: int, move_index: int, limits: tuple[float, float],
                 action_set: Collection[float]):
        super().__init__(num_envs, spaces.Discrete(len(action_set)))
        self._limits = limits
        self._move_index = move_index
        self._action_set = numpy.asarray(action_set)

    def transform(self, states, actions):
        result = states.copy()
        result[:, self._move_index] += self._action_set[numpy.asarray(actions).flatten()]
        return numpy.clip(result, *self._limits)
