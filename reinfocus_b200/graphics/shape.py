"""Shape tags and the host-side shape record (reference graphics/shape.py). The tags are the
integers the general-scene kernel switches on (rf_render_generic: 0 sphere, 1 rectangle)."""

import numpy
from numpy.typing import NDArray

SPHERE, RECTANGLE = range(2)


class CpuShape:
    """One shape on its way to the GPU: ``parameters`` (float32, layout per type, see
    sphere.py / rectangle.py) and ``shape_type`` (SPHERE or RECTANGLE)."""

    __slots__ = ("parameters", "shape_type")

    def __init__(self, parameters: NDArray[numpy.float32], shape_type: int):
        self.parameters = parameters
        self.shape_type = shape_type

    def __repr__(self) -> str:
        return f"CpuShape(parameters={self.parameters!r}, shape_type={self.shape_type!r})"

    def __eq__(self, other) -> bool:
        if not isinstance(other, CpuShape):
            return NotImplemented
        return self.shape_type == other.shape_type and numpy.array_equal(self.parameters, other.parameters)

    __hash__ = None
