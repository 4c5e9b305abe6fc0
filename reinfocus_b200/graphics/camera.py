"""Thin-lens camera parameters (reference graphics/camera.py:94-179 FastCameras).

All cameras share position, orientation, aperture and field of view; only the focus
distance differs per env. The per-env part is float32 [n, 3, 3] = (lower left corner,
horizontal, vertical); origin, u, v and the float64 lens radius are passed by value."""

import ctypes
import math

import numpy
from numpy.typing import NDArray

from reinfocus_b200.graphics import device_data
from reinfocus_b200.graphics import vector


GpuCamera = tuple  # (lower_left, horizontal, vertical, origin, u, v, lens_radius)


def make_gpu_camera(aperture: float = 0.1, aspect_ratio: float = 1, focus_distance: float = 10,
                    look_at: vector.V3F = vector.v3f(0, 0, -10),
                    look_from: vector.V3F = vector.v3f(0, 0, 0), up: vector.V3F = vector.v3f(0, 1, 0),
                    vfov: float = 30) -> GpuCamera:
    # pylint: disable=too-many-arguments
    """A thin-lens camera for the general-scene tracer (reference camera.py:182-226); the same
    float32 arithmetic as FastCameras, with a Python-float focus distance (so the scalar
    products are float64 before they meet the float32 vectors)."""

    half_height = math.tan((vfov * math.pi / 180.0) / 2.0)
    half_width = aspect_ratio * half_height
    w = vector.norm_v3f(vector.sub_v3f(look_from, look_at))
    u = vector.norm_v3f(vector.cross_v3f(up, w))
    v = vector.cross_v3f(w, u)
    return (
        vector.sub_v3f(look_from, vector.add_v3f((vector.smul_v3f(u, half_width * focus_distance),
                                                  vector.smul_v3f(v, half_height * focus_distance),
                                                  vector.smul_v3f(w, focus_distance)))),
        vector.smul_v3f(u, 2.0 * half_width * focus_distance),
        vector.smul_v3f(v, 2.0 * half_height * focus_distance),
        look_from, u, v, numpy.divide(aperture, 2.0),
    )


class Cameras:
    # pylint: disable=too-few-public-methods
    """One camera per env as a float64 [n, 19] array: the six float32 vectors and the float64
    lens radius side by side, which numpy.hstack promotes to float64 (reference
    camera.py:59-91)."""

    def __init__(self, *cameras: GpuCamera):
        self._cameras = numpy.hstack(
            [[cam[i] for cam in cameras] for i in range(6)]
            + [numpy.reshape([cam[6] for cam in cameras], (len(cameras), 1))])

    def __len__(self) -> int:
        return len(self._cameras)

    def device_data(self) -> NDArray[numpy.float64]:
        return self._cameras


class FastCameras(device_data.DeviceData):
    def __init__(self, aspect_ratio: float = 1, look_from: vector.V3F = vector.v3f(0, 0, 0),
                 look_at: vector.V3F = vector.v3f(0, 0, -10),
                 up: vector.V3F = vector.v3f(0, 1, 0), aperture: float = 0.1,
                 vfov: float = 30):
        # pylint: disable=too-many-arguments
        super().__init__()
        self._look_from = look_from
        # reference camera.py:124: a numpy.float64, which is why the kernel's aperture
        # arithmetic is float64
        self._half_aperture = numpy.divide(aperture, 2.0)
        self._half_height = math.tan((vfov * math.pi / 180.0) / 2.0)
        self._half_width = aspect_ratio * self._half_height
        self._w = vector.norm_v3f(vector.sub_v3f(look_from, look_at))
        self._u = vector.norm_v3f(vector.cross_v3f(up, self._w))
        self._v = vector.cross_v3f(self._w, self._u)
        self._constants = None

    @property
    def statics(self):
        """(origin, u, v, lens radius) passed to the kernel by value."""

        return self._look_from, self._u, self._v, float(self._half_aperture)

    @property
    def packing_constants(self):
        """The float32 constants ``_make_device_data`` multiplies with, in the field order of
        rf_scene_packing after ``world_tan``: half width / height, full width / height,
        origin, u, v, w and the (float64) lens radius."""

        f32 = numpy.float32
        vec = lambda values: (ctypes.c_float * 3)(*[float(f32(x)) for x in values])
        return (float(f32(self._half_width)), float(f32(self._half_height)),
                float(f32(2.0 * self._half_width)), float(f32(2.0 * self._half_height)),
                vec(self._look_from), vec(self._u), vec(self._v), vec(self._w),
                float(self._half_aperture))

    def _make_device_data(self, data: NDArray[numpy.float32]) -> NDArray[numpy.float32]:
        # reference camera.py:132-179, one env at a time; vectorised here with the same
        # float32 operations in the same order:
        #   lower_left = look_from - ((u*(hw*f) + v*(hh*f)) + w*f)
        #   horizontal = u * (2*hw*f),  vertical = v * (2*hh*f)
        # Python floats are weak under NumPy 2, so hw, hh, 2*hw, 2*hh enter as float32.
        f32 = numpy.float32
        if self._constants is None:  # the per-camera constants, converted once
            self._constants = tuple(numpy.asarray(x, dtype=f32)[None, :]
                                    for x in (self._u, self._v, self._w, self._look_from)) + (
                f32(self._half_width), f32(self._half_height),
                f32(2.0 * self._half_width), f32(2.0 * self._half_height))
        u, v, w, origin, half_width, half_height, full_width, full_height = self._constants
        f = data[:, None]
        total = (u * (half_width * f) + v * (half_height * f)) + w * f
        packed = numpy.empty((len(data), 3, 3), dtype=f32)
        packed[:, 0, :] = origin - total
        packed[:, 1, :] = u * (full_width * f)
        packed[:, 2, :] = v * (full_height * f)
        return packed
