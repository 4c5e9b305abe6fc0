"""Type variables shared by the strategy protocols (reference environments/types.py)."""

from typing import TypeVar

import numpy

StateT = TypeVar("StateT")
ActionT = TypeVar("ActionT", bound=numpy.generic)
ObservationT = TypeVar("ObservationT", bound=numpy.generic)
