"""z-aligned rectangles for the general-scene tracer (reference graphics/rectangle.py host
part; hit / fast_hit / uv live in csrc/rf_generic.cuh and csrc/rf_tracer.cuh)."""

import numpy

from reinfocus_b200.graphics import shape
from reinfocus_b200.graphics import vector

# parameter layout of the general tracer's rectangles
X_MIN, X_MAX, Y_MIN, Y_MAX, Z_POS, FX, FY = range(7)

# parameter layout of the fast tracer's squares (FastWorlds)
FH_RADIUS, FH_ZPOS = 0, 1


def rectangle(x_span: vector.V2F, y_span: vector.V2F, z_pos: float,
              texture: vector.V2F = vector.v2f(16, 16)) -> shape.CpuShape:
    """[x_min, x_max, y_min, y_max, z, checker fx, checker fy] (reference rectangle.py:26-46)."""

    return shape.CpuShape(numpy.array([*x_span, *y_span, z_pos, *texture], dtype=numpy.float32),
                          shape.RECTANGLE)
