"""Types shared by the strategy protocols (reference environments/types.py)."""

from __future__ import annotations

from typing import Generic, Protocol, TypeVar

import numpy
from numpy.typing import NDArray

T = TypeVar("T")


class IState(Protocol, Generic[T]):
    # pylint: disable=too-few-public-methods
    """What a batched env state must support: boolean-mask reads and writes of sub-batches
    (a NumPy array does)."""

    def __getitem__(self, key: NDArray[numpy.bool_]) -> T:
        ...

    def __setitem__(self, key: NDArray[numpy.bool_], value: T):
        ...


StateT = TypeVar("StateT", bound=IState)
StateT_co = TypeVar("StateT_co", bound=IState, covariant=True)
StateT_contra = TypeVar("StateT_contra", bound=IState, contravariant=True)

ActionT = TypeVar("ActionT", bound=numpy.generic)
ActionT_contra = TypeVar("ActionT_contra", bound=numpy.generic, contravariant=True)

ObservationT = TypeVar("ObservationT", bound=numpy.generic)
ObservationT_co = TypeVar("ObservationT_co", bound=numpy.generic, covariant=True)
ObservationT_contra = TypeVar("ObservationT_contra", bound=numpy.generic, contravariant=True)
