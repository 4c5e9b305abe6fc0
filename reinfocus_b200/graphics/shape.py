"""Shape tags and the host-side shape record (reference graphics/shape.py)."""

import dataclasses

import numpy
from numpy.typing import NDArray

SPHERE = 0
RECTANGLE = 1


@dataclasses.dataclass
class CpuShape:
    """A shape ready for upload: float32 parameters plus its type tag."""

    parameters: NDArray[numpy.float32]
    shape_type: int
