"""Type vocabulary of the strategy protocols (reference environments/types.py): the batched
state protocol and the state / action / observation type variables in their invariant,
covariant and contravariant flavours."""

from __future__ import annotations

import typing

import numpy
from numpy.typing import NDArray

T = typing.TypeVar("T")
_Mask = NDArray[numpy.bool_]


class IState(typing.Protocol, typing.Generic[T]):
    # pylint: disable=too-few-public-methods
    """A batched env state: sub-batches are read and written through boolean masks over the
    envs (a NumPy array qualifies)."""

    def __getitem__(self, key: _Mask) -> T: ...

    def __setitem__(self, key: _Mask, value: T): ...


def _flavours(name: str, bound, invariant: bool = True):
    made = []
    if invariant:
        made.append(typing.TypeVar(name, bound=bound))
    made.append(typing.TypeVar(f"{name}_co", bound=bound, covariant=True))
    made.append(typing.TypeVar(f"{name}_contra", bound=bound, contravariant=True))
    return made


StateT, StateT_co, StateT_contra = _flavours("StateT", IState)
ActionT, _, ActionT_contra = _flavours("ActionT", numpy.generic)
ObservationT, ObservationT_co, ObservationT_contra = _flavours("ObservationT", numpy.generic)
