"""Spheres for the general-scene tracer (reference graphics/sphere.py host part; the device
functions hit / uv live in csrc/rf_generic.cuh)."""

import numpy

from reinfocus_b200.graphics import shape
from reinfocus_b200.graphics import vector

# parameter layout
X, Y, Z, R, FX, FY = range(6)


def sphere(centre: vector.V3F, radius: float, texture: vector.V2F = vector.v2f(16, 16)) -> shape.CpuShape:
    """[cx, cy, cz, radius, checker fx, checker fy] (reference sphere.py:23-37)."""

    return shape.CpuShape(numpy.array([*centre, radius, *texture], dtype=numpy.float32), shape.SPHERE)
