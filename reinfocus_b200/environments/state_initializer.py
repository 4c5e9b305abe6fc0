"""State initialisers (reference environments/state_initializer.py)."""

from collections.abc import Collection, Sequence
from typing import Protocol

import numpy
from numpy import random
from numpy.typing import NDArray


class IStateInitializer(Protocol):
    # pylint: disable=too-few-public-methods
    def initialize(self, num_envs: int) -> NDArray[numpy.float32]:
        """A batch of ``num_envs`` fresh states."""


class RangedInitializer:
    # pylint: disable=too-few-public-methods
    """Each state element is uniform within a range picked uniformly from that element's
    list of ranges, e.g. ``[[(-1, 1)], [(-1, -0.5), (0.5, 1)]]``
    (reference state_initializer.py:30-71).

    The reference's generator is an unseeded PCG64DXSM (state_initializer.py:50), so its
    episodes are not reproducible; ``seed`` / ``generator`` make them so (parity tests and
    benchmarks inject one). Draw order is the reference's: per env, per element, first the
    range choice, then the uniform."""

    def __init__(self, ranges: Collection[Sequence[tuple[float, float]]], seed: int | None = None,
                 generator: random.Generator | None = None):
        self._generator = generator or random.Generator(random.PCG64DXSM(seed))
        self._ranges = ranges

    def initialize(self, num_envs: int) -> NDArray[numpy.float32]:
        if all(len(options) == 1 for options in self._ranges):
            # one range per element (every example env): choosing among one option consumes
            # no randomness, so the per-env loop below draws exactly one double per element
            # in env-major order - which is what a single vectorised uniform() draws
            low = numpy.array([options[0][0] for options in self._ranges], dtype=numpy.float64)
            high = numpy.array([options[0][1] for options in self._ranges], dtype=numpy.float64)
            return self._generator.uniform(low, high, size=(num_envs, len(self._ranges))).astype(
                numpy.float32)
        states = numpy.empty((num_envs, len(self._ranges)), dtype=numpy.float32)
        for env in range(num_envs):
            for element, options in enumerate(self._ranges):
                low, high = self._generator.choice(options)
                states[env, element] = self._generator.uniform(low, high)
        return states


class FixedInitializer:
    # pylint: disable=too-few-public-methods
    """Replays a prescribed stream of initial states (test / benchmark helper; the reference
    has no counterpart because its initialiser cannot be seeded)."""

    def __init__(self, states: NDArray[numpy.float32]):
        self._states = numpy.asarray(states, dtype=numpy.float32)
        self._next = 0

    def initialize(self, num_envs: int) -> NDArray[numpy.float32]:
        assert self._next + num_envs <= len(self._states), "ran out of prescribed states"
        out = self._states[self._next:self._next + num_envs].copy()
        self._next += num_envs
        return out
