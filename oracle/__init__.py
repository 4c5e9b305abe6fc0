"""TEST INFRASTRUCTURE ONLY - ctypes front end of the CPU oracle (oracle/rf_oracle.c).

The product package (reinfocus_b200/) never imports this module. Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do, as the
checker or the timed CPU baseline.

Host-side parameter packing (reference graphics/world.py:100-123 and
graphics/camera.py:122-179) is restated here in plain NumPy, independently of the
product's vectorised packing in reinfocus_b200/graphics, so that the two can be checked
against each other and against the golden vectors.
"""

import ctypes
import math
import os
import subprocess

import numpy

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "librf_oracle.so")

PROFILE_SIM = 0
PROFILE_GPU = 1

STATE_DTYPE = numpy.dtype([("s0", numpy.uint64), ("s1", numpy.uint64)], align=True)

_lib = None


def build(force: bool = False) -> str:
    """Compiles the C restatement with oracle/Makefile (gcc only)."""

    if force or not os.path.exists(_LIB_PATH):
        subprocess.run(["make", "-C", _HERE, "clean", "all"], check=True, capture_output=True)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        c_float_p = ctypes.POINTER(ctypes.c_float)
        _lib.rfo_rng_init.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint64]
        _lib.rfo_rng_init_doubling.argtypes = _lib.rfo_rng_init.argtypes
        _lib.rfo_uniform_float32.argtypes = [ctypes.c_void_p]
        _lib.rfo_uniform_float32.restype = ctypes.c_float
        _lib.rfo_next.argtypes = [ctypes.c_void_p]
        _lib.rfo_next.restype = ctypes.c_uint64
        _lib.rfo_render_fast.argtypes = [
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
            ctypes.c_void_p, ctypes.c_void_p, c_float_p, c_float_p, c_float_p,
            ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
        ]
        _lib.rfo_render_generic.argtypes = [
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_int,
        ]
        _lib.rfo_gray.argtypes = [ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
        _lib.rfo_focus_gray.argtypes = [
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
        ]
        _lib.rfo_focus_rgb.argtypes = [
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_int,
        ]
        _lib.rfo_step.argtypes = [
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
            ctypes.c_void_p, ctypes.c_void_p, c_float_p, c_float_p, c_float_p,
            ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
        ]
        _lib.rfo_max_threads.restype = ctypes.c_int
        _lib.rfo_set_threads.argtypes = [ctypes.c_int]
    return _lib


def max_threads() -> int:
    return int(lib().rfo_max_threads())


# --------------------------------------------------------------------------------------
# RNG (numba.cuda.random restated)
# --------------------------------------------------------------------------------------


def rng_states(n: int, seed: int = 0, doubling: bool = False) -> numpy.ndarray:
    """numba.cuda.random.create_xoroshiro128p_states(n, seed) as a host structured array."""

    states = numpy.zeros(n, dtype=STATE_DTYPE)
    fn = lib().rfo_rng_init_doubling if doubling else lib().rfo_rng_init
    fn(states.ctypes.data, n, ctypes.c_uint64(seed & 0xFFFFFFFFFFFFFFFF))
    return states


def uniform_float32(states: numpy.ndarray, index: int) -> numpy.float32:
    """numba.cuda.random.xoroshiro128p_uniform_float32(states, index)."""

    ptr = states.ctypes.data + index * STATE_DTYPE.itemsize
    return numpy.float32(lib().rfo_uniform_float32(ptr))


# --------------------------------------------------------------------------------------
# Host-side packing (reference world.py:100-123, camera.py:110-179) - scalar restatement
# --------------------------------------------------------------------------------------


def pack_world(targets, r_size: float = 20.0) -> numpy.ndarray:
    """FastWorlds._make_device_data: [[target * tan(radians(r_size / 2)), -target]].

    ``target`` is a numpy.float32 element and the tangent a Python float, so under NumPy 2
    (NEP 50) the product is a float32 multiply by float32(tan)."""

    targets = numpy.asarray(targets, dtype=numpy.float32)
    tan = math.tan(math.radians(r_size / 2))
    return numpy.array([[t * tan, -t] for t in targets], dtype=numpy.float32).reshape(-1, 2)


class CameraStatics:
    """FastCameras.__init__ (reference camera.py:99-130) for the static camera parts."""

    def __init__(self, aspect_ratio=1, look_from=(0, 0, 0), look_at=(0, 0, -10),
                 up=(0, 1, 0), aperture=0.1, vfov=30):
        f32 = numpy.float32
        look_from = tuple(f32(c) for c in look_from)
        look_at = tuple(f32(c) for c in look_at)
        up = tuple(f32(c) for c in up)

        def norm(vec):
            length = float(numpy.linalg.norm(numpy.asarray(vec)))
            r = numpy.multiply(vec, 1.0 / length)
            return (r[0], r[1], r[2])

        def cross(a, b):
            c = tuple(numpy.cross(numpy.asarray(a), numpy.asarray(b)))
            return (c[0], c[1], c[2])

        self.look_from = look_from
        self.half_aperture = numpy.divide(aperture, 2.0)
        self.half_height = math.tan((vfov * math.pi / 180.0) / 2.0)
        self.half_width = aspect_ratio * self.half_height
        sub = numpy.subtract(look_from, look_at)
        self.w = norm((sub[0], sub[1], sub[2]))
        self.u = norm(cross(up, self.w))
        self.v = cross(self.w, self.u)


def pack_cameras(focus_planes, statics: CameraStatics | None = None) -> numpy.ndarray:
    """FastCameras._make_device_data (reference camera.py:132-179), one env at a time."""

    st = statics or CameraStatics()
    focus_planes = numpy.asarray(focus_planes, dtype=numpy.float32)

    def smul(vec, s):
        r = numpy.multiply(vec, s)
        return (r[0], r[1], r[2])

    rows = []
    for f in focus_planes:
        total = numpy.sum(
            (smul(st.u, st.half_width * f), smul(st.v, st.half_height * f), smul(st.w, f)),
            axis=0,
        )
        lower_left = numpy.subtract(st.look_from, (total[0], total[1], total[2]))
        rows.append([
            (lower_left[0], lower_left[1], lower_left[2]),
            smul(st.u, 2.0 * st.half_width * f),
            smul(st.v, 2.0 * st.half_height * f),
        ])
    return numpy.array(rows, dtype=numpy.float32).reshape(-1, 3, 3)


# --------------------------------------------------------------------------------------
# Tracer + focus measure
# --------------------------------------------------------------------------------------


def _f3(values):
    return (ctypes.c_float * 3)(*[float(v) for v in values])


def render_fast(world, cam_dyn, frame_height, spp, states, profile=PROFILE_GPU,
                origin=(0, 0, 0), u=(1, 0, 0), v=(0, 1, 0), lens_radius=0.05, threads=0,
                frame_width=None):
    """FastRenderer._device_render over a batch; advances ``states`` in place and returns
    uint8 frames [n, H, W, 3]."""

    world = numpy.ascontiguousarray(world, dtype=numpy.float32).reshape(-1, 2)
    cam_dyn = numpy.ascontiguousarray(cam_dyn, dtype=numpy.float32).reshape(-1, 9)
    n = world.shape[0]
    assert cam_dyn.shape[0] >= n
    H = int(frame_height)
    W = int(frame_width) if frame_width is not None else H
    assert states.dtype == STATE_DTYPE and states.flags.c_contiguous
    assert len(states) >= n * H * W
    frames = numpy.empty((n, H, W, 3), dtype=numpy.uint8)
    lib().rfo_render_fast(profile, n, H, W, int(spp), world.ctypes.data, cam_dyn.ctypes.data,
                          _f3(origin), _f3(u), _f3(v), float(lens_radius),
                          states.ctypes.data, frames.ctypes.data, threads)
    return frames


def render_generic(shape_params, shape_types, env_sizes, cameras, frame_shape, spp, threads=0):
    """reference render.render (render.py:88-119): fresh seed-0 states, general scenes."""

    shape_params = numpy.ascontiguousarray(shape_params, dtype=numpy.float32)
    if shape_params.shape[2] < 7:
        shape_params = numpy.ascontiguousarray(
            numpy.pad(shape_params, ((0, 0), (0, 0), (0, 7 - shape_params.shape[2]))))
    shape_types = numpy.ascontiguousarray(shape_types, dtype=numpy.int32)
    env_sizes = numpy.ascontiguousarray(env_sizes, dtype=numpy.int32)
    cameras = numpy.ascontiguousarray(cameras, dtype=numpy.float64)
    n, max_shapes, _ = shape_params.shape
    H, W = int(frame_shape[0]), int(frame_shape[1])
    states = rng_states(n * H * W, 0, doubling=True)
    frames = numpy.empty((n, H, W, 3), dtype=numpy.uint8)
    lib().rfo_render_generic(n, H, W, int(spp), max_shapes, shape_params.ctypes.data,
                             shape_types.ctypes.data, env_sizes.ctypes.data, cameras.ctypes.data,
                             states.ctypes.data, frames.ctypes.data, threads)
    return frames


def gray(images):
    images = numpy.ascontiguousarray(images, dtype=numpy.uint8)
    assert images.shape[-1] == 3
    out = numpy.empty(images.shape[:-1], dtype=numpy.uint8)
    lib().rfo_gray(out.size, images.ctypes.data, out.ctypes.data)
    return out


def focus_values_gray(grays, threads=0, planes=False):
    grays = numpy.ascontiguousarray(grays, dtype=numpy.uint8)
    if grays.ndim == 2:
        grays = grays[None]
    n, H, W = grays.shape
    out = numpy.empty(n, dtype=numpy.float64)
    med = numpy.empty_like(grays) if planes else None
    lap = numpy.empty_like(grays) if planes else None
    lib().rfo_focus_gray(n, H, W, grays.ctypes.data, out.ctypes.data,
                         med.ctypes.data if planes else None,
                         lap.ctypes.data if planes else None, threads)
    return (out, med, lap) if planes else out


def focus_values(images, threads=0):
    """reference vision.py:28-39 focus_values over uint8 RGB images [n, H, W, 3]."""

    images = numpy.ascontiguousarray(images, dtype=numpy.uint8)
    if images.ndim == 3:
        images = images[None]
    n, H, W, _ = images.shape
    out = numpy.empty(n, dtype=numpy.float64)
    lib().rfo_focus_rgb(n, H, W, images.ctypes.data, out.ctypes.data, threads)
    return out


def step(world, cam_dyn, frame_height, spp, states, profile=PROFILE_GPU, origin=(0, 0, 0),
         u=(1, 0, 0), v=(0, 1, 0), lens_radius=0.05, threads=0):
    """One hot-path step on the CPU: render + focus values, states advanced in place."""

    world = numpy.ascontiguousarray(world, dtype=numpy.float32).reshape(-1, 2)
    cam_dyn = numpy.ascontiguousarray(cam_dyn, dtype=numpy.float32).reshape(-1, 9)
    n = world.shape[0]
    H = int(frame_height)
    out = numpy.empty(n, dtype=numpy.float64)
    lib().rfo_step(profile, n, H, H, int(spp), world.ctypes.data, cam_dyn.ctypes.data,
                   _f3(origin), _f3(u), _f3(v), float(lens_radius), states.ctypes.data,
                   out.ctypes.data, threads)
    return out


class OracleFastRenderer:
    """CPU statement of reference render.FastRenderer (render.py:122-257) including the
    RNG-state cache semantics (states persist; re-created from seed 0 only on growth)."""

    def __init__(self, samples_per_pixel=100, r_size=20, profile=PROFILE_GPU, threads=0):
        self.spp = samples_per_pixel
        self.r_size = r_size
        self.profile = profile
        self.threads = threads
        self.statics = CameraStatics()
        self.world = None
        self.cam = None
        self.states = None

    def update_targets(self, targets):
        self.world = pack_world(targets, self.r_size)

    def update_focus_planes(self, focus_planes):
        self.cam = pack_cameras(focus_planes, self.statics)

    def render(self, frame_height):
        assert self.world is not None and self.cam is not None
        total = len(self.world) * frame_height * frame_height
        if self.states is None or len(self.states) < total:
            self.states = rng_states(total, 0)
        return render_fast(self.world, self.cam, frame_height, self.spp, self.states,
                           self.profile, self.statics.look_from, self.statics.u,
                           self.statics.v, float(self.statics.half_aperture), self.threads)
