"""Per-env target parameters (reference graphics/world.py:85-123 FastWorlds)."""

import math

import numpy
from numpy.typing import NDArray

from reinfocus_b200.graphics import device_data


class FastWorlds(device_data.DeviceData):
    """One z-aligned checkerboard square per env: float32 [n, 2] = (half side, z)."""

    def __init__(self, r_size: float = 20):
        super().__init__()
        self._r_size = r_size
        # reference shape_factory.py:29-41 get_absolute_size: distance * tan(radians(r/2)).
        # `distance` is a numpy.float32 element and the tangent a Python float, which NumPy 2
        # treats as weak: the product is a float32 multiply by float32(tan).
        self._tan = numpy.float32(math.tan(math.radians(r_size / 2)))

    def _make_device_data(self, data: NDArray[numpy.float32]) -> NDArray[numpy.float32]:
        packed = numpy.empty((len(data), 2), dtype=numpy.float32)
        packed[:, 0] = data * self._tan
        packed[:, 1] = -data
        return packed
