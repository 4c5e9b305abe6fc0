"""ctypes binding of libreinfocus_b200.so (include/reinfocus_b200.h).

This is the only place the package touches native code. There is no fallback: if the
library is missing or cannot be loaded, or no sm_100 GPU is visible, the product path
raises instead of computing anything on the CPU.
"""

import ctypes
import os
import threading

import numpy

from reinfocus_b200 import build as _build

RF_OK = 0
RF_ERR_INVALID = -1
RF_ERR_CUDA = -2
RF_ERR_NOMEM = -3
RF_ERR_NO_SCENE = -4

ABI_VERSION = 9

SELFTEST_CHECKER, SELFTEST_PIXEL_DIV, SELFTEST_INV_LENGTH, SELFTEST_CONST_DIV, SELFTEST_CHECKER_PAIR = 0, 1, 2, 3, 4
OPT_FORCE_GENERIC = 0
OPT_TRACE_CONTEXTS = 1
INFO_LAST_TRACE_KERNEL = 0
INFO_LAST_FOCUS_KERNEL = 1

ENV_DISCRETE_MOVE, ENV_CONTINUOUS_JUMP, ENV_CONTINUOUS_MOVE, ENV_DISCRETE_JUMP = 0, 1, 2, 3
(ENV_ENDER_TIME_LIMIT, ENV_ENDER_DIVERGING, ENV_ENDER_ON_TARGET, ENV_ENDER_STOPPED, ENV_ENDER_ENDLESS,
 ENV_ENDER_AND, ENV_ENDER_OR) = range(7)
(ENV_REWARD_DELTA, ENV_REWARD_DISTANCE, ENV_REWARD_OBSERVATION, ENV_REWARD_ON_TARGET, ENV_REWARD_STOPPED,
 ENV_REWARD_ADD, ENV_REWARD_MUL) = range(7)
ENV_OBS_ELEMENT, ENV_OBS_FOCUS, ENV_OBS_DELTA, ENV_OBS_NORMALIZED = range(4)
ENV_MAX_NODES = 24
ENV_MAX_OBS_NODES = 16
ENV_MAX_OBS_DIM = 16
ENV_MAX_OBS_VALUES = 32
ENV_MAX_WINDOW = 16
ENV_ACTIONS_INT32, ENV_ACTIONS_INT64, ENV_ACTIONS_FLOAT32 = 0, 1, 2

STATE_DTYPE = numpy.dtype([("s0", numpy.uint64), ("s1", numpy.uint64)], align=True)


class ScenePacking(ctypes.Structure):
    """rf_scene_packing."""

    _fields_ = [
        ("world_tan", ctypes.c_float),
        ("half_width", ctypes.c_float), ("half_height", ctypes.c_float),
        ("full_width", ctypes.c_float), ("full_height", ctypes.c_float),
        ("origin", ctypes.c_float * 3), ("u", ctypes.c_float * 3),
        ("v", ctypes.c_float * 3), ("w", ctypes.c_float * 3),
        ("lens_radius", ctypes.c_double),
    ]


class EnvEnder(ctypes.Structure):
    """rf_env_ender."""

    _fields_ = [("kind", ctypes.c_int), ("i0", ctypes.c_int), ("i1", ctypes.c_int),
                ("steps", ctypes.c_int), ("value", ctypes.c_float)]


class EnvReward(ctypes.Structure):
    """rf_env_reward."""

    _fields_ = [("kind", ctypes.c_int), ("i0", ctypes.c_int), ("i1", ctypes.c_int),
                ("f0", ctypes.c_float), ("f1", ctypes.c_float),
                ("d0", ctypes.c_double), ("d1", ctypes.c_double)]


class EnvObserver(ctypes.Structure):
    """rf_env_observer."""

    _fields_ = [("kind", ctypes.c_int), ("arg", ctypes.c_int), ("flag", ctypes.c_int),
                ("offset", ctypes.c_int)]


class EnvConfig(ctypes.Structure):
    """rf_env_config."""

    _fields_ = [
        ("num_envs", ctypes.c_int), ("frame_height", ctypes.c_int),
        ("samples_per_pixel", ctypes.c_int),
        ("transformer", ctypes.c_int),
        ("n_moves", ctypes.c_int),
        ("moves", ctypes.c_double * 32),
        ("limits", ctypes.c_float * 2),
        ("jump_span", ctypes.c_float),
        ("jump_threshold", ctypes.c_float),
        ("move_speed", ctypes.c_float),
        ("jumps", ctypes.c_float * 32),
        ("n_enders", ctypes.c_int), ("n_rewards", ctypes.c_int),
        ("enders", EnvEnder * ENV_MAX_NODES),
        ("rewards", EnvReward * ENV_MAX_NODES),
        ("n_observers", ctypes.c_int), ("observers", EnvObserver * ENV_MAX_OBS_NODES),
        ("obs_mid", ctypes.c_float * ENV_MAX_OBS_VALUES), ("obs_scale", ctypes.c_float * ENV_MAX_OBS_VALUES),
        ("init_options", ctypes.c_int * 2),
        ("init_low", (ctypes.c_double * 4) * 2), ("init_high", (ctypes.c_double * 4) * 2),
        ("packing", ScenePacking),
    ]

_c_float_p = ctypes.POINTER(ctypes.c_float)
_vp = ctypes.c_void_p

_SIGNATURES = {
    "rf_create": (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.c_int]),
    "rf_destroy": (ctypes.c_int, [_vp]),
    "rf_last_error": (ctypes.c_char_p, [_vp]),
    "rf_last_global_error": (ctypes.c_char_p, []),
    "rf_abi_version": (ctypes.c_int, []),
    "rf_sizeof": (ctypes.c_int, [ctypes.c_int]),
    "rf_device_info": (ctypes.c_int, [_vp] + [ctypes.POINTER(ctypes.c_int)] * 4),
    "rf_launch_count": (ctypes.c_int64, [_vp]),
    "rf_rng_ensure": (ctypes.c_int, [_vp, ctypes.c_int64, ctypes.c_uint64, _vp]),
    "rf_rng_reset": (ctypes.c_int, [_vp]),
    "rf_rng_count": (ctypes.c_int64, [_vp]),
    "rf_rng_export": (ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_int64]),
    "rf_rng_import": (ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_int64]),
    "rf_rng_init_device": (ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_uint64, _vp]),
    "rf_rng_uniform_device": (ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_int, _vp, _vp]),
    "rf_set_world": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp]),
    "rf_set_cameras": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _c_float_p, _c_float_p,
                                      _c_float_p, ctypes.c_double, _vp]),
    "rf_scene_envs": (ctypes.c_int, [_vp]),
    "rf_render": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                 _vp, _vp, _vp]),
    "rf_render_generic": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_int, _vp, _vp, _vp, _vp, ctypes.c_uint64, _vp, _vp]),
    "rf_focus": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp,
                                ctypes.c_int, _vp, _vp]),
    "rf_focus_planes": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp,
                                       ctypes.c_int, _vp, _vp, _vp, _vp]),
    "rf_step_positions_host": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, _vp,
                                              ctypes.POINTER(ScenePacking), _vp, _vp]),
    "rf_step_host": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, _vp,
                                    _vp, _vp]),
    "rf_step_device": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, _vp]),
    "rf_set_scene_device": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp, ctypes.c_int,
                                           ctypes.POINTER(ScenePacking), _vp]),
    "rf_env_create": (ctypes.c_int, [_vp, ctypes.POINTER(EnvConfig), ctypes.POINTER(_vp)]),
    "rf_env_destroy": (ctypes.c_int, [_vp]),
    "rf_env_set_generator": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_uint64),
                                            ctypes.POINTER(ctypes.c_uint64), ctypes.c_uint32, ctypes.c_uint32]),
    "rf_env_get_generator": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_uint64),
                                            ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint32),
                                            ctypes.POINTER(ctypes.c_uint32)]),
    "rf_env_reset": (ctypes.c_int, [_vp, _vp, _vp]),
    "rf_env_step": (ctypes.c_int, [_vp, _vp, ctypes.c_int, _vp, _vp, _vp,
                                   ctypes.POINTER(ctypes.c_int), _vp]),
    "rf_env_node_rows": (ctypes.c_int, [_vp]),
    "rf_env_obs_dim": (ctypes.c_int, [_vp]),
    "rf_env_delta_width": (ctypes.c_int, [_vp]),
    "rf_env_export": (ctypes.c_int, [_vp, _vp, _vp, _vp]),
    "rf_env_import": (ctypes.c_int, [_vp, _vp, _vp, _vp]),
    "rf_selftest": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int,
                                   ctypes.POINTER(ctypes.c_int64), _vp]),
    "rf_set_option": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int]),
    "rf_get_info": (ctypes.c_int, [_vp, ctypes.c_int]),
    "rf_measure_fp32_peak": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_double),
                                            ctypes.POINTER(ctypes.c_double)]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None
_lib_lock = threading.Lock()


class NativeLibraryError(RuntimeError):
    """libreinfocus_b200.so is missing, stale or unusable. There is no CPU fallback."""


def library_path() -> str:
    return _build.LIB_PATH


def load() -> ctypes.CDLL:
    """Loads the native library (no GPU needed for loading and symbol lookup)."""

    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        path = library_path()
        if not os.path.exists(path):
            raise NativeLibraryError(
                f"{path} not found: build it with `python -m reinfocus_b200.build` "
                "(needs nvcc). reinfocus_b200 has no CPU fallback.")
        try:
            lib = ctypes.CDLL(path)
        except OSError as error:
            raise NativeLibraryError(f"cannot load {path}: {error}") from error
        for name, (restype, argtypes) in _SIGNATURES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as error:
                raise NativeLibraryError(f"{path} does not export {name}; rebuild it") from error
            fn.restype = restype
            fn.argtypes = argtypes
        if lib.rf_abi_version() != ABI_VERSION:
            raise NativeLibraryError(
                f"{path} has ABI {lib.rf_abi_version()}, expected {ABI_VERSION}; rebuild it")
        for which, struct in enumerate((ScenePacking, EnvConfig)):
            if lib.rf_sizeof(which) != ctypes.sizeof(struct):
                raise NativeLibraryError(
                    f"{struct.__name__} is {ctypes.sizeof(struct)} bytes here but "
                    f"{lib.rf_sizeof(which)} in {path}; rebuild it")
        _lib = lib
        return lib


def _f3(values):
    return (ctypes.c_float * 3)(*[float(v) for v in values])


def _stream_ptr(stream=None, device=None):
    """The cudaStream_t to launch on: torch's current stream on the context's device unless
    one is given."""

    if stream is None:
        import torch

        return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
    if hasattr(stream, "cuda_stream"):
        return ctypes.c_void_p(stream.cuda_stream)
    return ctypes.c_void_p(int(stream))


class Context:
    """One rf_ctx: a GPU, one cache of RNG states and one scene (= one FastRenderer)."""

    def __init__(self, device: int | None = None):
        import torch

        if not torch.cuda.is_available():
            raise NativeLibraryError(
                "no CUDA device visible: reinfocus_b200 runs on B200 (sm_100a) only and has "
                "no CPU fallback")
        self._lib = load()
        self.device = torch.cuda.current_device() if device is None else int(device)
        handle = _vp()
        rc = self._lib.rf_create(ctypes.byref(handle), self.device)
        if rc != RF_OK:
            raise NativeLibraryError(self._lib.rf_last_global_error().decode())
        self._handle = handle

    def close(self):
        if getattr(self, "_handle", None):
            self._lib.rf_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # pylint: disable=broad-except
            pass

    # ------------------------------------------------------------------ error mapping
    def _check(self, rc: int):
        if rc == RF_OK:
            return
        message = self._lib.rf_last_error(self._handle).decode()
        if rc in (RF_ERR_INVALID, RF_ERR_NO_SCENE):
            # the reference raises AssertionError when rendering before an update
            # (graphics/device_data.py:43)
            raise AssertionError(message)
        if rc == RF_ERR_NOMEM:
            raise MemoryError(message)
        raise RuntimeError(message)

    # ----------------------------------------------------------------------- queries
    def device_info(self) -> dict:
        vals = [ctypes.c_int() for _ in range(4)]
        self._check(self._lib.rf_device_info(self._handle, *[ctypes.byref(v) for v in vals]))
        return {"sm_count": vals[0].value, "cc": (vals[1].value, vals[2].value),
                "clock_khz": vals[3].value}

    def launch_count(self) -> int:
        return int(self._lib.rf_launch_count(self._handle))

    # --------------------------------------------------------------------------- RNG
    def rng_ensure(self, n_states: int, seed: int = 0, stream=None):
        self._check(self._lib.rf_rng_ensure(self._handle, int(n_states),
                                            ctypes.c_uint64(seed & (2**64 - 1)),
                                            _stream_ptr(stream, self.device)))

    def rng_reset(self):
        self._check(self._lib.rf_rng_reset(self._handle))

    def rng_count(self) -> int:
        return int(self._lib.rf_rng_count(self._handle))

    def rng_export(self, first: int = 0, n: int | None = None) -> numpy.ndarray:
        n = self.rng_count() - first if n is None else n
        out = numpy.empty(n, dtype=STATE_DTYPE)
        self._check(self._lib.rf_rng_export(self._handle, out.ctypes.data, first, n))
        return out

    def rng_import(self, states: numpy.ndarray, first: int = 0):
        states = numpy.ascontiguousarray(states, dtype=STATE_DTYPE)
        self._check(self._lib.rf_rng_import(self._handle, states.ctypes.data, first, len(states)))

    def rng_init_device(self, d_states_ptr: int, n: int, seed: int, stream=None):
        self._check(self._lib.rf_rng_init_device(self._handle, _vp(d_states_ptr), int(n),
                                                 ctypes.c_uint64(seed & (2**64 - 1)),
                                                 _stream_ptr(stream, self.device)))

    def rng_uniform_device(self, d_states_ptr: int, n: int, draws: int, d_out_ptr: int,
                           stream=None):
        self._check(self._lib.rf_rng_uniform_device(self._handle, _vp(d_states_ptr), int(n),
                                                    int(draws), _vp(d_out_ptr),
                                                    _stream_ptr(stream, self.device)))

    # ------------------------------------------------------------------------- scene
    def set_world(self, world: numpy.ndarray, stream=None):
        world = numpy.ascontiguousarray(world, dtype=numpy.float32).reshape(-1, 2)
        self._check(self._lib.rf_set_world(self._handle, len(world), world.ctypes.data,
                                           _stream_ptr(stream, self.device)))
        # the copy is stream-ordered from pageable memory: CUDA stages it before returning

    def set_cameras(self, cam_dyn: numpy.ndarray, origin, u, v, lens_radius: float, stream=None):
        cam_dyn = numpy.ascontiguousarray(cam_dyn, dtype=numpy.float32).reshape(-1, 9)
        self._check(self._lib.rf_set_cameras(self._handle, len(cam_dyn), cam_dyn.ctypes.data,
                                             _f3(origin), _f3(u), _f3(v), float(lens_radius),
                                             _stream_ptr(stream, self.device)))

    def scene_envs(self) -> int:
        return int(self._lib.rf_scene_envs(self._handle))

    # ----------------------------------------------------------------------- kernels
    def render(self, n: int, height: int, width: int, spp: int, d_rgb: int | None,
               d_gray: int | None, stream=None):
        self._check(self._lib.rf_render(self._handle, n, height, width, spp, _vp(d_rgb),
                                        _vp(d_gray), _stream_ptr(stream, self.device)))

    def render_generic(self, shape_params: numpy.ndarray, shape_types: numpy.ndarray,
                       env_sizes: numpy.ndarray, cameras: numpy.ndarray, height: int, width: int,
                       spp: int, d_rgb: int, seed: int = 0, stream=None):
        shape_params = numpy.ascontiguousarray(shape_params, dtype=numpy.float32)
        shape_types = numpy.ascontiguousarray(shape_types, dtype=numpy.int32)
        env_sizes = numpy.ascontiguousarray(env_sizes, dtype=numpy.int32)
        cameras = numpy.ascontiguousarray(cameras, dtype=numpy.float64)
        n, max_shapes, width_params = shape_params.shape
        assert width_params == 7 and shape_types.shape == (n, max_shapes)
        assert env_sizes.shape == (n,) and cameras.shape == (n, 19)
        self._check(self._lib.rf_render_generic(
            self._handle, n, height, width, spp, max_shapes, shape_params.ctypes.data,
            shape_types.ctypes.data, env_sizes.ctypes.data, cameras.ctypes.data,
            ctypes.c_uint64(seed & (2**64 - 1)), _vp(d_rgb), _stream_ptr(stream, self.device)))

    def focus(self, n: int, height: int, width: int, d_img: int, channels: int, d_out: int,
              d_median: int | None = None, d_laplacian: int | None = None, stream=None):
        self._check(self._lib.rf_focus_planes(self._handle, n, height, width, _vp(d_img),
                                              channels, _vp(d_out), _vp(d_median),
                                              _vp(d_laplacian), _stream_ptr(stream, self.device)))

    def step_host(self, n: int, height: int, spp: int, h_world: int | None,
                  h_cam_dyn: int | None, h_focus: int, stream=None):
        self._check(self._lib.rf_step_host(self._handle, n, height, spp, _vp(h_world),
                                           _vp(h_cam_dyn), _vp(h_focus), _stream_ptr(stream, self.device)))

    def step_positions_host(self, n: int, height: int, spp: int, h_targets: int, h_planes: int,
                            packing: ScenePacking, h_focus: int, stream=None):
        self._check(self._lib.rf_step_positions_host(self._handle, n, height, spp, _vp(h_targets), _vp(h_planes),
                                                     ctypes.byref(packing), _vp(h_focus),
                                                     _stream_ptr(stream, self.device)))

    def step_device(self, n: int, height: int, spp: int, d_focus: int, stream=None):
        self._check(self._lib.rf_step_device(self._handle, n, height, spp, _vp(d_focus),
                                             _stream_ptr(stream, self.device)))

    def set_scene_device(self, n: int, d_targets: int, d_planes: int, stride: int,
                         packing: ScenePacking, stream=None):
        self._check(self._lib.rf_set_scene_device(self._handle, n, _vp(d_targets), _vp(d_planes),
                                                  stride, ctypes.byref(packing),
                                                  _stream_ptr(stream, self.device)))

    # ------------------------------------------------------------------- self-checks
    def selftest(self, which: int, arg: int = 0, stream=None) -> int:
        bad = ctypes.c_int64(-1)
        self._check(self._lib.rf_selftest(self._handle, which, arg, ctypes.byref(bad),
                                          _stream_ptr(stream, self.device)))
        return bad.value

    def selftest_checker(self, stream=None) -> int:
        return self.selftest(SELFTEST_CHECKER, 0, stream)

    def set_option(self, option: int, value: int):
        self._check(self._lib.rf_set_option(self._handle, option, value))

    def last_trace_kernel(self) -> int:
        return int(self._lib.rf_get_info(self._handle, INFO_LAST_TRACE_KERNEL))

    def last_focus_kernel(self) -> int:
        return int(self._lib.rf_get_info(self._handle, INFO_LAST_FOCUS_KERNEL))

    def measure_fp32_peak(self) -> tuple[float, float]:
        tflops, mhz = ctypes.c_double(), ctypes.c_double()
        self._check(self._lib.rf_measure_fp32_peak(self._handle, ctypes.byref(tflops),
                                                   ctypes.byref(mhz)))
        return tflops.value, mhz.value


class DeviceEnv:
    """One rf_env: the device-resident vector env bound to a context (= to one renderer)."""

    _MASK64 = (1 << 64) - 1

    def __init__(self, context: Context, config: EnvConfig):
        self._context = context
        self._lib = context._lib  # pylint: disable=protected-access
        self.num_envs = config.num_envs
        handle = _vp()
        context._check(  # pylint: disable=protected-access
            self._lib.rf_env_create(context._handle, ctypes.byref(config), ctypes.byref(handle)))
        self._handle = handle
        self.obs_dim = int(self._lib.rf_env_obs_dim(handle))

    def close(self):
        if getattr(self, "_handle", None):
            self._lib.rf_env_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # pylint: disable=broad-except
            pass

    def _check(self, rc: int):
        self._context._check(rc)  # pylint: disable=protected-access

    def set_generator(self, state: int, inc: int, has_uint32: int = 0, uinteger: int = 0):
        """numpy ``bit_generator.state`` of a PCG64DXSM: 128-bit state and increment, and the
        buffered 32-bit half."""

        pair = ctypes.c_uint64 * 2
        self._check(self._lib.rf_env_set_generator(
            self._handle, pair(state >> 64, state & self._MASK64), pair(inc >> 64, inc & self._MASK64),
            int(has_uint32), int(uinteger)))

    def get_generator(self) -> tuple[int, int, int, int]:
        state, inc = (ctypes.c_uint64 * 2)(), (ctypes.c_uint64 * 2)()
        has_uint32, uinteger = ctypes.c_uint32(), ctypes.c_uint32()
        self._check(self._lib.rf_env_get_generator(self._handle, state, inc, ctypes.byref(has_uint32),
                                                   ctypes.byref(uinteger)))
        return (state[0] << 64) | state[1], (inc[0] << 64) | inc[1], has_uint32.value, uinteger.value

    def reset(self, d_obs: int, stream=None):
        self._check(self._lib.rf_env_reset(self._handle, _vp(d_obs),
                                           _stream_ptr(stream, self._context.device)))

    def step(self, d_actions: int, action_kind: int, d_obs: int, d_rewards: int, d_truncated: int,
             stream=None) -> int:
        resets = ctypes.c_int(0)
        self._check(self._lib.rf_env_step(self._handle, _vp(d_actions), action_kind, _vp(d_obs),
                                          _vp(d_rewards), _vp(d_truncated), ctypes.byref(resets),
                                          _stream_ptr(stream, self._context.device)))
        return resets.value

    def _layout(self):
        rows = int(self._lib.rf_env_node_rows(self._handle))
        delta_width = int(self._lib.rf_env_delta_width(self._handle))
        return (("states", numpy.float32, (self.num_envs, 2)),
                ("old_obs", numpy.float32, (self.num_envs, delta_width)),
                ("node_state", numpy.uint32, (rows, self.num_envs)))

    def export(self) -> dict:
        """Host copies of the per-env episode state: states, the DeltaObserver's previous
        observations and the raw per-env rows of the ender / rewarder nodes."""

        arrays = {name: numpy.empty(shape, dtype=dtype) for name, dtype, shape in self._layout()}
        self._check(self._lib.rf_env_export(
            self._handle, *[arrays[name].ctypes.data for name, _, _ in self._layout()]))
        return arrays

    def load(self, arrays: dict):
        """Restores what ``export`` returned (stands in for a reset)."""

        ordered = []
        for name, dtype, shape in self._layout():
            array = numpy.ascontiguousarray(arrays[name], dtype=dtype)
            assert array.shape == shape, name
            ordered.append(array)
        self._check(self._lib.rf_env_import(self._handle, *[array.ctypes.data for array in ordered]))


_shared_contexts: dict[int, Context] = {}


def shared_context(device: int | None = None) -> Context:
    """A per-device context for stateless calls (vision.focus_values, make_random_states)."""

    import torch

    device = torch.cuda.current_device() if device is None else int(device)
    ctx = _shared_contexts.get(device)
    if ctx is None:
        ctx = _shared_contexts[device] = Context(device)
    return ctx
