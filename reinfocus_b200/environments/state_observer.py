"""State observers (reference environments/state_observer.py). ``FocusObserver`` is the
per-step caller of the hot path: it hands the targets / focus planes to the renderer and
gets one focus value per env back; here that is a single C-ABI call that never moves a
frame off the GPU (FastRenderer.step_focus)."""

import functools
from collections.abc import Sequence
from typing import Protocol, SupportsFloat

import numpy
from numpy.typing import NDArray

from reinfocus_b200 import gym_compat
from reinfocus_b200.graphics import render

spaces = gym_compat.spaces


class IStateObserver(Protocol):
    """The interface observers follow (reference state_observer.py:21-54)."""

    observation_space: object
    single_observation_space: object

    def observe(self, states, indices: NDArray[numpy.bool_] | None = None):
        ...

    def reset(self, states, indices: NDArray[numpy.bool_] | None = None):
        ...


class BaseObserver:
    # pylint: disable=too-few-public-methods
    """Observations within [min_obs, max_obs] (reference :57-100)."""

    def __init__(self, num_envs: int, min_obs: SupportsFloat | NDArray[numpy.float32],
                 max_obs: SupportsFloat | NDArray[numpy.float32]):
        self.single_observation_space = spaces.Box(min_obs, max_obs, dtype=numpy.float32)
        self.observation_space = gym_compat.batch_space(self.single_observation_space, num_envs)

    def _indices(self, indices):
        if indices is None:
            return numpy.full(self.observation_space.shape[0], True)
        return indices

    def observe(self, states, indices: NDArray[numpy.bool_] | None = None):
        raise NotImplementedError

    def reset(self, states, indices: NDArray[numpy.bool_] | None = None):
        """Observations of the first states of restarted episodes (all envs if None)."""

        return self.observe(states, self._indices(indices))


class WrapperObserver(BaseObserver):
    """Stacks the observations of child observers side by side (reference :103-164)."""

    def __init__(self, observers: Sequence[BaseObserver], min_obs, max_obs):
        sizes = {observer.observation_space.shape[0] for observer in observers}
        assert len(sizes) == 1, "Appended observers must have the same number of environments"
        super().__init__(sizes.pop(), min_obs, max_obs)
        self._observers = observers

    def reset(self, states, indices=None):
        return numpy.hstack([observer.reset(states, indices) for observer in self._observers],
                            dtype=numpy.float32)

    def wrapped_observations(self, states, indices=None):
        return numpy.hstack([observer.observe(states, indices) for observer in self._observers],
                            dtype=numpy.float32)


def _child_bounds(observers):
    lows = numpy.hstack([o.single_observation_space.low for o in observers], dtype=numpy.float32)
    highs = numpy.hstack([o.single_observation_space.high for o in observers], dtype=numpy.float32)
    return lows, highs


class DeltaObserver(WrapperObserver):
    # pylint: disable=too-few-public-methods
    """Observes the change of the children's observations since the previous step,
    optionally preceded by the observations themselves; zeros on reset
    (reference :167-292)."""

    def __init__(self, observers: BaseObserver | Sequence[BaseObserver], include_original: bool = False,
                 max_change: SupportsFloat | NDArray[numpy.float32] | None = None):
        if not isinstance(observers, Sequence):
            observers = [observers]
        lows, highs = _child_bounds(observers)
        if max_change is None:
            change = highs - lows
        elif isinstance(max_change, numpy.ndarray):
            change = highs - lows
            given = numpy.isfinite(max_change)
            change[given] = max_change[given]
        else:
            change = numpy.full(len(lows), max_change, dtype=numpy.float32)
        if include_original:
            super().__init__(observers, numpy.append(lows, -change), numpy.append(highs, change))
        else:
            super().__init__(observers, -change, change)
        self._include_original = include_original
        self._old_wrapped_observations = numpy.full(
            (self.observation_space.shape[0], len(lows)), numpy.nan, dtype=numpy.float32)

    def _emit(self, current, change, indices):
        self._old_wrapped_observations[indices] = current
        if self._include_original:
            return numpy.hstack([current, change], dtype=numpy.float32)
        return change

    def observe(self, states, indices=None):
        indices = self._indices(indices)
        current = self.wrapped_observations(states, indices)
        return self._emit(current, current - self._old_wrapped_observations[indices], indices)

    def reset(self, states, indices=None):
        indices = self._indices(indices)
        current = super().reset(states, indices)
        return self._emit(current, numpy.zeros(current.shape, dtype=numpy.float32), indices)


def _new_renderer():
    return render.FastRenderer()


@functools.cache
def cached_focus_extrema(ends: tuple[float, float], frame_height: int):
    """Smallest / largest focus value over the range of target and focus-plane positions
    (reference :295-320): one 13-env render with a FRESH default renderer - envs 0-1 put the
    target and the focus plane at opposite ends (minimum), envs 2-12 put both at the same
    one of 11 evenly spaced positions (maximum). The two numbers fix the observation bounds
    and hence the normalisation of every observation and reward."""

    max_targets = numpy.linspace(*ends, 11)
    renderer = _new_renderer()
    focus_values = renderer.step_focus(numpy.append(ends, max_targets),
                                       numpy.append(ends[::-1], max_targets), frame_height)
    return min(focus_values[0:2]), max(focus_values[2:13])


class FocusObserver(BaseObserver):
    # pylint: disable=too-few-public-methods
    """Focus value of the rendered scene whose target and focus-plane positions are two
    elements of the state (reference :323-383)."""

    def __init__(self, num_envs: int, target_index: int, focus_plane_index: int,
                 ends: tuple[float, float], renderer: render.FastRenderer, frame_height: int = 300):
        # pylint: disable=too-many-arguments
        min_focus, max_focus = cached_focus_extrema(ends, frame_height)
        super().__init__(num_envs, min_focus, max_focus)
        self._target_index = target_index
        self._focus_plane_index = focus_plane_index
        self._renderer = renderer
        self._frame_height = frame_height

    def observe(self, states, indices=None):
        indices = self._indices(indices)
        # reference :377-383: update_targets, update_focus_planes, render, focus_values. A
        # partial reset passes only the k done envs, which render as batch positions 0..k-1.
        focus_values = self._renderer.step_focus(states[:, self._target_index],
                                                 states[:, self._focus_plane_index],
                                                 self._frame_height)
        return numpy.reshape(focus_values, (indices.sum(), self.observation_space.shape[1]))


class IndexedElementObserver(BaseObserver):
    # pylint: disable=too-few-public-methods
    """One element of the state, verbatim (reference :386-421)."""

    def __init__(self, num_envs: int, element_index: int, min_obs: float, max_obs: float):
        super().__init__(num_envs, min_obs, max_obs)
        self._element_index = element_index

    def observe(self, states, indices=None):
        indices = self._indices(indices)
        return states[:, self._element_index].reshape((indices.sum(), self.observation_space.shape[1]))


class NormalizedObserver(WrapperObserver):
    # pylint: disable=too-few-public-methods
    """Children's observations side by side, scaled to [-1, 1] by their bounds and clipped
    (reference :424-517)."""

    def __init__(self, observers: BaseObserver | Sequence[BaseObserver]):
        if not isinstance(observers, Sequence):
            observers = [observers]
        lows, highs = _child_bounds(observers)
        count = len(lows)
        super().__init__(observers, numpy.ones(count, dtype=numpy.float32) * -1,
                         numpy.ones(count, dtype=numpy.float32))
        spans = numpy.vstack([lows, highs], dtype=numpy.float32)
        self._mid = numpy.average(spans, axis=0)
        self._scale = numpy.diff(spans / 2, axis=0).reshape(count)

    def observe(self, states, indices=None):
        return self._normalize(self.wrapped_observations(states, indices))

    def reset(self, states, indices=None):
        return self._normalize(super().reset(states, indices))

    def _normalize(self, values):
        return numpy.clip((values - self._mid) / self._scale, -1, 1, dtype=numpy.float32)
