"""Stand-in for the reference's reinfocus/graphics/render.py (FastRenderer only): the same
class surface (reference graphics/render.py:122-188), with `render` / `_device_render` /
`_make_random_states` (render.py:165-257) replaced by calls into libreinfocus_b200.so.

The host-side packing stays the reference's own: `camera.FastCameras` and `world.FastWorlds`
(graphics/camera.py:94-179, world.py:85-123) are imported from the reference package; their
`cuda.to_device` upload has to hand back the host array instead (one-line
change in each `_make_device_data`; the runner patches `numba.cuda.to_device` to do that
without touching the reference sources)."""

import ctypes
import os
from typing import Collection

import numpy
import torch

from reinfocus.graphics import camera  # the reference's own modules
from reinfocus.graphics import world

_LIB_PATH = os.environ.get(
    "REINFOCUS_B200_LIB",
    os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))),
                 "reinfocus_b200", "libreinfocus_b200.so"))
_lib = ctypes.CDLL(_LIB_PATH)
_vp, _i, _d = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
_f3 = ctypes.c_float * 3
_lib.rf_create.argtypes = [ctypes.POINTER(_vp), _i]
_lib.rf_destroy.argtypes = [_vp]
_lib.rf_set_world.argtypes = [_vp, _i, _vp, _vp]
_lib.rf_set_cameras.argtypes = [_vp, _i, _vp, _f3, _f3, _f3, _d, _vp]
_lib.rf_render.argtypes = [_vp, _i, _i, _i, _i, _vp, _vp, _vp]
_lib.rf_focus.argtypes = [_vp, _i, _i, _i, _vp, _i, _vp, _vp]
_lib.rf_last_error.restype = ctypes.c_char_p
_lib.rf_last_error.argtypes = [_vp]
_lib.rf_last_global_error.restype = ctypes.c_char_p


def check(ctx, rc):
    """Error behaviour of the reference: AssertionError for call-order / argument errors
    (device_data.py:43), RuntimeError for CUDA failures."""

    if rc == 0:
        return
    message = (_lib.rf_last_error(ctx) if ctx else _lib.rf_last_global_error()).decode()
    if rc in (-1, -4):  # RF_ERR_INVALID / RF_ERR_NO_SCENE
        raise AssertionError(message)
    raise RuntimeError(message)


def new_context():
    ctx = _vp()
    check(None, _lib.rf_create(ctypes.byref(ctx), torch.cuda.current_device()))
    return ctx


class FastRenderer:
    """reference graphics/render.py:122-257 on the C-ABI."""

    def __init__(self, block_shape=(1, 16, 16), samples_per_pixel: int = 100, r_size: float = 20):
        self._block_shape = block_shape  # launch shapes are the library's business
        self._samples_per_pixel = samples_per_pixel
        self._cameras = camera.FastCameras()
        self._worlds = world.FastWorlds(r_size=r_size)
        self._ctx = new_context()

    def __del__(self):
        if getattr(self, "_ctx", None):
            _lib.rf_destroy(self._ctx)

    def update_targets(self, targets: Collection[float]):
        self._worlds.update(targets)

    def update_focus_planes(self, focus_planes: Collection[float]):
        self._cameras.update(focus_planes)

    def render(self, frame_height: int):
        n = len(self._worlds)
        world_data = numpy.ascontiguousarray(self._worlds.device_data(), numpy.float32)  # [n, 2]
        cam, origin, u, v, lens = self._cameras.device_data()  # [n, 3, 3] + static camera
        cam = numpy.ascontiguousarray(cam, numpy.float32)
        check(self._ctx, _lib.rf_set_world(self._ctx, n, world_data.ctypes.data, None))
        check(self._ctx, _lib.rf_set_cameras(self._ctx, n, cam.ctypes.data, _f3(*origin), _f3(*u), _f3(*v),
                                             float(lens), None))
        frames = torch.empty((n, frame_height, frame_height, 3), dtype=torch.uint8, device="cuda")
        # rf_render keeps the reference's RNG-cache policy (seed 0, re-created only on growth)
        check(self._ctx, _lib.rf_render(self._ctx, n, frame_height, frame_height, self._samples_per_pixel,
                                        frames.data_ptr(), None, None))
        return frames.cpu().numpy()
