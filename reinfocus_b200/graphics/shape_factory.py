"""Scene builders (reference graphics/shape_factory.py)."""

import math
from collections.abc import Sequence
from typing import NamedTuple

from reinfocus_b200.graphics import rectangle
from reinfocus_b200.graphics import shape
from reinfocus_b200.graphics import sphere
from reinfocus_b200.graphics import vector


class ShapeParameters(NamedTuple):
    """distance from the origin; absolute size (0 = derive it from r_size, the degrees of
    field of view the shape should span); checkerboard frequency."""

    distance: float = 10.0
    size: float = 0.0
    r_size: float = 20.0
    texture_f: tuple[int, int] = (16, 16)


def get_absolute_size(parameters: ShapeParameters) -> float:
    """reference shape_factory.py:29-41"""

    if parameters.size != 0.0:
        return parameters.size
    return parameters.distance * math.tan(math.radians(parameters.r_size / 2))


_SIDE = math.tan(math.radians(15))  # lateral offset per unit distance of the side-by-side scenes


def _sphere_at(x: float, parameters: ShapeParameters) -> shape.CpuShape:
    return sphere.sphere(vector.v3f(x, 0, -parameters.distance), get_absolute_size(parameters),
                         vector.v2f(*parameters.texture_f))


def _rect_at(x: float, parameters: ShapeParameters) -> shape.CpuShape:
    size = get_absolute_size(parameters)
    return rectangle.rectangle(vector.v2f(x - size, x + size), vector.v2f(-size, size),
                               -parameters.distance, vector.v2f(*parameters.texture_f))


def one_sphere(parameters: ShapeParameters = ShapeParameters()) -> Sequence[shape.CpuShape]:
    """One sphere on the z axis (reference :44-62)."""

    return [_sphere_at(0, parameters)]


def two_sphere(left_parameters: ShapeParameters = ShapeParameters(20.0),
               right_parameters: ShapeParameters = ShapeParameters(5.0)) -> Sequence[shape.CpuShape]:
    """Spheres left and right at different distances (reference :65-100)."""

    return [_sphere_at(-left_parameters.distance * _SIDE, left_parameters),
            _sphere_at(right_parameters.distance * _SIDE, right_parameters)]


def one_rect(parameters: ShapeParameters = ShapeParameters()) -> Sequence[shape.CpuShape]:
    """One rectangle on the z axis (reference :103-122)."""

    size = get_absolute_size(parameters)
    return [rectangle.rectangle(vector.v2f(-size, size), vector.v2f(-size, size), -parameters.distance,
                                vector.v2f(*parameters.texture_f))]


def two_rect(left_parameters: ShapeParameters = ShapeParameters(20.0),
             right_parameters: ShapeParameters = ShapeParameters(5.0)) -> Sequence[shape.CpuShape]:
    """Rectangles left and right at different distances (reference :125-160). The left one's
    x span is written as (-offset - size, -offset + size), as in the reference."""

    left_offset = left_parameters.distance * _SIDE
    left_size = get_absolute_size(left_parameters)
    return [
        rectangle.rectangle(vector.v2f(-left_offset - left_size, -left_offset + left_size),
                            vector.v2f(-left_size, left_size), -left_parameters.distance,
                            vector.v2f(*left_parameters.texture_f)),
        _rect_at(right_parameters.distance * _SIDE, right_parameters),
    ]


def mixed(left_parameters: ShapeParameters = ShapeParameters(5.0),
          right_parameters: ShapeParameters = ShapeParameters()) -> Sequence[shape.CpuShape]:
    """A sphere on the left, a rectangle on the right (reference :163-196)."""

    return [_sphere_at(-left_parameters.distance * _SIDE, left_parameters),
            _rect_at(right_parameters.distance * _SIDE, right_parameters)]
