"""Episode enders: decide when episodes terminate or truncate
(reference environments/episode_ender.py). All of them report "never terminated"; the focus
problem has an unbounded horizon and episodes only ever truncate."""

from __future__ import annotations

import abc
from collections.abc import Callable
from typing import Any, Protocol

import numpy
from numpy.typing import NDArray

from reinfocus_b200 import histories


class IEpisodeEnder(Protocol):
    """The interface episode enders follow (reference episode_ender.py:18-83)."""

    def step(self, states: Any):
        ...

    def is_terminated(self) -> NDArray[numpy.bool_]:
        ...

    def is_truncated(self) -> NDArray[numpy.bool_]:
        ...

    def reset(self, states: Any, indices: NDArray[numpy.bool_] | None = None):
        ...

    def status(self, index: int) -> str:
        ...


def _all_envs(num_envs: int, indices):
    return numpy.full(num_envs, True) if indices is None else indices


class BaseEnder(abc.ABC):
    """Enders combine with ``&`` and ``|`` (reference :86-109)."""

    _num_envs: int

    def __and__(self, other) -> BaseEnder:
        return OpEnder(self, other, numpy.bitwise_and)

    def __or__(self, other) -> BaseEnder:
        return OpEnder(self, other, numpy.bitwise_or)

    def step(self, states: Any):
        """Called once per timestep with the new states."""

    def reset(self, states: Any, indices: NDArray[numpy.bool_] | None = None):
        """Called with the first states of the episodes that restarted (all if None)."""

    def is_terminated(self) -> NDArray[numpy.bool_]:
        return numpy.full(self._num_envs, False)

    @abc.abstractmethod
    def is_truncated(self) -> NDArray[numpy.bool_]:
        ...

    def status(self, index: int) -> str:
        # pylint: disable=unused-argument
        return ""


class DivergingEnder(BaseEnder):
    """Truncates once two state elements have moved apart (gap grown by more than
    ``threshold`` since the previous step) on ``early_end_steps`` steps, consecutive or not
    (reference :112-207)."""

    def __init__(self, num_envs: int, check_indices: tuple[int, int], threshold: float,
                 early_end_steps: int = 10):
        self._num_envs = num_envs
        self._check_indices = check_indices
        self._threshold = threshold
        self._early_end_steps = early_end_steps
        self._diverging_steps = numpy.zeros(num_envs, dtype=numpy.int32)
        self._last_diff = numpy.zeros(num_envs, dtype=numpy.float32)

    def _gap(self, states):
        return abs(states[:, self._check_indices[0]] - states[:, self._check_indices[1]])

    def step(self, states: NDArray[numpy.float32]):
        gap = self._gap(states)
        self._diverging_steps[gap > self._last_diff + self._threshold] += 1
        self._last_diff = gap

    def is_truncated(self):
        return self._diverging_steps >= self._early_end_steps

    def reset(self, states, indices=None):
        indices = _all_envs(self._num_envs, indices)
        self._diverging_steps[indices] = 0
        self._last_diff[indices] = self._gap(states)

    def status(self, index: int) -> str:
        steps = self._diverging_steps[index]
        return f"diverging {steps} / {self._early_end_steps}" if steps > 0 else ""


class EndlessEnder(BaseEnder):
    """Never ends an episode (reference :210-270)."""

    def __init__(self, num_envs: int):
        self._num_envs = num_envs

    def is_truncated(self):
        return numpy.full(self._num_envs, False)


class OnTargetEnder(BaseEnder):
    """Truncates after two state elements have stayed within ``early_end_radius`` of each
    other for ``early_end_steps`` consecutive steps (reference :273-369)."""

    def __init__(self, num_envs: int, check_indices: tuple[int, int], early_end_radius: float,
                 early_end_steps: int = 10):
        self._num_envs = num_envs
        self._check_indices = check_indices
        self._radius = early_end_radius
        self._early_end_steps = early_end_steps
        self._on_target_steps = numpy.zeros(num_envs, dtype=numpy.int32)

    def step(self, states: NDArray[numpy.float32]):
        close = abs(states[:, self._check_indices[0]] - states[:, self._check_indices[1]]) < self._radius
        self._on_target_steps[close] += 1
        self._on_target_steps[~close] = 0

    def is_truncated(self):
        return self._on_target_steps >= self._early_end_steps

    def reset(self, states, indices=None):
        self._on_target_steps[_all_envs(self._num_envs, indices)] = 0

    def status(self, index: int) -> str:
        steps = self._on_target_steps[index]
        return f"on target {steps} / {self._early_end_steps}" if steps > 0 else ""


class OpEnder(BaseEnder):
    """Combines two enders' answers with a logical operation (reference :372-463)."""

    def __init__(self, l_ender, r_ender,
                 op: Callable[[NDArray[numpy.bool_], NDArray[numpy.bool_]], NDArray[numpy.bool_]]):
        self._l_ender = l_ender
        self._r_ender = r_ender
        self._op = op

    def step(self, states):
        self._l_ender.step(states)
        self._r_ender.step(states)

    def is_terminated(self):
        return self._op(self._l_ender.is_terminated(), self._r_ender.is_terminated())

    def is_truncated(self):
        return self._op(self._l_ender.is_truncated(), self._r_ender.is_truncated())

    def reset(self, states, indices=None):
        self._l_ender.reset(states, indices)
        self._r_ender.reset(states, indices)

    def status(self, index: int) -> str:
        parts = [s for s in (self._l_ender.status(index), self._r_ender.status(index)) if s]
        return ", ".join(parts)


class StoppedEnder(BaseEnder):
    """Truncates when one state element has stayed within ``early_end_span`` over the last
    ``early_end_steps`` + 1 recorded positions (reference :466-587)."""

    def __init__(self, num_envs: int, check_index: int, early_end_span: float,
                 early_end_steps: int = 10):
        self._num_envs = num_envs
        self._check_index = check_index
        self._early_end_span = early_end_span
        self._early_end_steps = early_end_steps
        self._moves = histories.Histories(num_envs, early_end_steps + 1)

    def step(self, states: NDArray[numpy.float32]):
        self._moves.append_events(states[:, self._check_index])

    def is_truncated(self):
        data = self._moves.data
        full = ~numpy.any(numpy.isnan(data), 1)
        with numpy.errstate(invalid="ignore"), numpy.testing.suppress_warnings() as sup:
            sup.filter(RuntimeWarning)
            span = abs(numpy.nanmax(data, 1) - numpy.nanmin(data, 1))
        return (span < self._early_end_span) & full

    def reset(self, states, indices=None):
        indices = _all_envs(self._num_envs, indices)
        self._moves.reset(indices)
        self._moves.append_events(states[:, self._check_index], indices)

    def status(self, index: int) -> str:
        moves = self._moves.data[index]
        top = bottom = moves[-1]
        stopped = self._early_end_steps
        for i, move in enumerate(moves[self._early_end_steps - 1::-1]):
            if numpy.isnan(move):
                stopped = i
                break
            bottom, top = min(bottom, move), max(top, move)
            if top - bottom > self._early_end_span:
                stopped = i
                break
        return f"stopped {stopped} / {self._early_end_steps}" if stopped else ""


class TimeLimitEnder(BaseEnder):
    """Truncates after ``max_steps`` steps (reference :590-656)."""

    def __init__(self, num_envs: int, max_steps: int):
        self._num_envs = num_envs
        self._max_steps = max_steps
        self._steps = numpy.zeros(num_envs, dtype=numpy.int32)

    def step(self, states):
        self._steps += 1

    def is_truncated(self):
        return self._steps >= self._max_steps

    def reset(self, states, indices=None):
        self._steps[_all_envs(self._num_envs, indices)] = 0

    def status(self, index: int) -> str:
        return f"step {self._steps[index]} / {self._max_steps}"
