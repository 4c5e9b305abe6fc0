"""Per-env scene parameters (reference graphics/world.py): FastWorlds for the env path, Worlds
for the general-scene tracer."""

import math
from collections.abc import Collection

import numpy
from numpy.typing import NDArray

from reinfocus_b200.graphics import device_data
from reinfocus_b200.graphics import shape


class Worlds:
    """Sets of shapes, one set per env, padded to the largest set (reference world.py:27-82):
    parameters float32 [n, S, P], types int32 [n, S], sizes int32 [n]."""

    def __init__(self, *env_shapes: Collection[shape.CpuShape]):
        self._num_envs = len(env_shapes)
        self._sizes = numpy.array([len(shapes) for shapes in env_shapes], dtype=numpy.int32)
        most = int(max(self._sizes))
        width = max(max(len(s.parameters) for s in shapes) for shapes in env_shapes)
        self._parameters = numpy.zeros((self._num_envs, most, width), dtype=numpy.float32)
        self._types = numpy.zeros((self._num_envs, most), dtype=numpy.int32)
        for env, shapes in enumerate(env_shapes):
            for index, item in enumerate(shapes):
                self._parameters[env, index, :len(item.parameters)] = item.parameters
                self._types[env, index] = item.shape_type

    def __len__(self) -> int:
        return self._num_envs

    def device_data(self):
        """(parameters, types, sizes) as host arrays; render() uploads them."""

        return self._parameters, self._types, self._sizes


class FastWorlds(device_data.DeviceData):
    """One z-aligned checkerboard square per env: float32 [n, 2] = (half side, z)."""

    def __init__(self, r_size: float = 20):
        super().__init__()
        self._r_size = r_size
        # reference shape_factory.py:29-41 get_absolute_size: distance * tan(radians(r/2)).
        # `distance` is a numpy.float32 element and the tangent a Python float, which NumPy 2
        # treats as weak: the product is a float32 multiply by float32(tan).
        self._tan = numpy.float32(math.tan(math.radians(r_size / 2)))

    @property
    def packing_constant(self) -> float:
        """float32(tan(radians(r_size / 2))), the factor of ``_make_device_data``."""

        return float(self._tan)

    def _make_device_data(self, data: NDArray[numpy.float32]) -> NDArray[numpy.float32]:
        packed = numpy.empty((len(data), 2), dtype=numpy.float32)
        packed[:, 0] = data * self._tan
        packed[:, 1] = -data
        return packed
