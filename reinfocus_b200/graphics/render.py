"""Ray tracing front end (reference graphics/render.py).

``FastRenderer`` keeps the reference's three methods - ``update_targets``,
``update_focus_planes``, ``render(frame_height) -> uint8 (n, H, H, 3)`` on the host - and
adds device-resident fast paths so that the per-step caller (FocusObserver) never moves
images: ``render_device``, ``render_gray_device``, ``focus_values_device`` and
``step_focus``. All of them launch the same hand-written sm_100a tracer
(csrc/rf_tracer.cuh) through the C-ABI; there is no CPU path."""

from collections.abc import Collection

import numpy
from numpy.typing import NDArray

from reinfocus_b200 import _lib
from reinfocus_b200.graphics import camera
from reinfocus_b200.graphics import world


def make_render_target(frame_shape: tuple[int, ...] = (300, 600)):
    """A uint8 render target on the GPU (reference render.py:19-28), as a torch tensor."""

    import torch

    return torch.empty(tuple(frame_shape) + (3,), dtype=torch.uint8, device="cuda")


def render(world_data: world.Worlds, cameras: camera.Cameras, frame_shape: tuple[int, int] = (300, 600),
           block_shape: tuple[int, int, int] = (1, 16, 16),
           samples_per_pixel: int = 100) -> NDArray[numpy.uint8]:
    # pylint: disable=unused-argument
    """Ray traced images uint8 (n, H, W, 3) of general scenes (spheres, rectangles, up to 50
    bounces) seen by one camera per env (reference render.py:88-119). RNG states are created
    from seed 0 on every call, as in the reference. ``block_shape`` is accepted for signature
    compatibility."""

    import torch

    ctx = _lib.shared_context()
    n = len(world_data)
    assert len(cameras) >= n, "one camera per env is required"
    height, width = frame_shape
    frames = torch.empty((n, height, width, 3), dtype=torch.uint8, device=f"cuda:{ctx.device}")
    parameters, types, sizes = world_data.device_data()
    if parameters.shape[2] < 7:
        parameters = numpy.pad(parameters, ((0, 0), (0, 0), (0, 7 - parameters.shape[2])))
    ctx.render_generic(parameters, types, sizes, cameras.device_data()[:n], height, width,
                       samples_per_pixel, frames.data_ptr(), seed=0)
    return frames.cpu().numpy()


class FastRenderer:
    """Produces images of focus scenes (reference render.py:122-257).

    One instance owns one C-ABI context, i.e. its own cache of per-pixel RNG states: they
    persist from call to call and are re-created from seed 0 only when a call needs more
    states than exist (reference render.py:248-257)."""

    def __init__(
        self,
        block_shape: tuple[int, int, int] = (1, 16, 16),
        samples_per_pixel: int = 100,
        r_size: float = 20,
        device: int | None = None,
    ):
        # block_shape is accepted for signature compatibility; the CUDA kernel picks its
        # own launch shape (256-thread blocks, consecutive threads along x).
        self._block_shape = block_shape
        self._samples_per_pixel = samples_per_pixel

        self._cameras = camera.FastCameras()
        self._worlds = world.FastWorlds(r_size=r_size)

        self._ctx = _lib.Context(device)
        self._uploaded_world = -1
        self._uploaded_cameras = -1
        self._pinned = {}
        self._packing = None   # rf_scene_packing of this renderer's worlds / cameras
        self._pending = None   # (targets, focus planes) of a step_focus not yet in the DeviceData

    # ------------------------------------------------------------ reference methods
    def update_targets(self, targets: Collection[float]):
        """reference render.py:147-154"""

        self._flush_pending()
        self._worlds.update(targets)

    def update_focus_planes(self, focus_planes: Collection[float]):
        """reference render.py:156-163"""

        self._flush_pending()
        self._cameras.update(focus_planes)

    def _flush_pending(self):
        """``step_focus`` hands the positions straight to the library; the host-side
        ``DeviceData`` (what ``render`` / ``len`` / a later partial update see) catches up here."""

        if self._pending is not None:
            targets, focus_planes = self._pending
            self._pending = None
            self._worlds.update(targets)
            self._cameras.update(focus_planes)

    def render(self, frame_height: int) -> NDArray[numpy.uint8]:
        """reference render.py:165-188: uint8 RGB frames (n, H, H, 3) on the host."""

        return self.render_device(frame_height).cpu().numpy()

    # ------------------------------------------------------------------- fast paths
    @property
    def context(self) -> _lib.Context:
        return self._ctx

    @property
    def samples_per_pixel(self) -> int:
        return self._samples_per_pixel

    def __len__(self) -> int:
        self._flush_pending()
        return len(self._worlds)

    def _sync_scene(self) -> int:
        self._flush_pending()
        world_data = self._worlds.device_data()  # AssertionError before the first update
        cam_data = self._cameras.device_data()
        n = len(self._worlds)
        assert len(self._cameras) >= n, (
            f"{n} targets but only {len(self._cameras)} focus planes were set")
        if self._uploaded_world != self._worlds.version:
            self._ctx.set_world(world_data)
            self._uploaded_world = self._worlds.version
        if self._uploaded_cameras != self._cameras.version:
            self._ctx.set_cameras(cam_data, *self._cameras.statics)
            self._uploaded_cameras = self._cameras.version
        return n

    def render_device(self, frame_height: int):
        """The frames of ``render`` as a torch uint8 CUDA tensor (n, H, H, 3)."""

        import torch

        n = self._sync_scene()
        frames = torch.empty((n, frame_height, frame_height, 3), dtype=torch.uint8,
                             device=f"cuda:{self._ctx.device}")
        self._ctx.render(n, frame_height, frame_height, self._samples_per_pixel,
                         frames.data_ptr(), None)
        return frames

    def render_gray_device(self, frame_height: int):
        """cv2.cvtColor(frame, RGB2GRAY) of each frame, accumulated in registers: torch
        uint8 CUDA tensor (n, H, H). Consumes the RNG exactly like ``render``."""

        import torch

        n = self._sync_scene()
        gray = torch.empty((n, frame_height, frame_height), dtype=torch.uint8,
                           device=f"cuda:{self._ctx.device}")
        self._ctx.render(n, frame_height, frame_height, self._samples_per_pixel, None,
                         gray.data_ptr())
        return gray

    def focus_values_device(self, frame_height: int):
        """vision.focus_values(self.render(frame_height)) without leaving the GPU: torch
        float64 CUDA tensor (n,)."""

        import torch

        n = self._sync_scene()
        out = torch.empty((n,), dtype=torch.float64, device=f"cuda:{self._ctx.device}")
        self._ctx.step_device(n, frame_height, self._samples_per_pixel, out.data_ptr())
        return out

    def scene_packing(self) -> _lib.ScenePacking:
        """The constants of the host packing (FastWorlds / FastCameras ._make_device_data)
        for rf_set_scene_device, which does that packing on the GPU."""

        if self._packing is None:
            self._packing = _lib.ScenePacking(self._worlds.packing_constant, *self._cameras.packing_constants)
        return self._packing

    def scene_overwritten(self):
        """Tells the renderer that someone (rf_set_scene_device, a device env) replaced the
        scene inside its context: the next host-side render uploads again."""

        self._uploaded_world = -1
        self._uploaded_cameras = -1

    def step_focus_device(self, targets, focus_planes, frame_height: int = 300):
        """``step_focus`` for positions that already live on the GPU: float32 CUDA tensors
        (n,) in (any stride), float64 CUDA tensor (n,) out, no host copies. Scene packing
        (reference world.py:107-123, camera.py:132-179) runs on the device."""

        import torch

        assert targets.is_cuda and focus_planes.is_cuda, "device tensors expected"
        assert targets.dtype == torch.float32 and focus_planes.dtype == torch.float32
        assert targets.dim() == 1 and targets.shape == focus_planes.shape
        assert targets.stride(0) == focus_planes.stride(0)
        n = targets.shape[0]
        out = torch.empty((n,), dtype=torch.float64, device=targets.device)
        self._ctx.set_scene_device(n, targets.data_ptr(), focus_planes.data_ptr(),
                                   targets.stride(0), self.scene_packing())
        self.scene_overwritten()
        self._ctx.step_device(n, frame_height, self._samples_per_pixel, out.data_ptr())
        return out

    def step_focus(self, targets: Collection[float], focus_planes: Collection[float],
                   frame_height: int = 300) -> NDArray[numpy.float64]:
        """One FocusObserver.observe (reference state_observer.py:377-383) in a single
        C-ABI call: host targets / focus planes in, host float64 focus values out. Equivalent
        to update_targets + update_focus_planes + vision.focus_values(render(frame_height)):
        the positions go to the GPU as they are (8 bytes per env) and the packing of
        FastWorlds / FastCameras runs there (rf_step_positions_host), bit for bit what the host
        classes compute."""

        import torch

        targets = numpy.asarray(targets, dtype=numpy.float32)
        focus_planes = numpy.asarray(focus_planes, dtype=numpy.float32)
        assert targets.ndim == 1 and focus_planes.ndim == 1, "expected one value per environment"
        n = len(targets)
        assert len(focus_planes) >= n, f"{n} targets but only {len(focus_planes)} focus planes were set"
        if n == 0:
            self.update_targets(targets)
            self.update_focus_planes(focus_planes)
            return numpy.empty((0,), dtype=numpy.float64)
        # pinned staging buffers, grown on demand and reused for smaller batches (the vector
        # env alternates between all n envs and the k that restarted)
        if self._pinned.get("capacity", 0) < n:
            capacity = max(n, 2 * self._pinned.get("capacity", 0))
            self._pinned = {"capacity": capacity,
                            "targets": torch.empty((capacity,), dtype=torch.float32).pin_memory(),
                            "planes": torch.empty((capacity,), dtype=torch.float32).pin_memory(),
                            "focus": torch.empty((capacity,), dtype=torch.float64).pin_memory()}
            self._pinned["views"] = tuple(self._pinned[name].numpy() for name in ("targets", "planes", "focus"))
            self._pinned["pointers"] = tuple(self._pinned[name].data_ptr() for name in ("targets", "planes", "focus"))
        h_targets, h_planes, h_focus = self._pinned["views"]
        h_targets[:n] = targets
        h_planes[:n] = focus_planes[:n]
        self._ctx.step_positions_host(n, frame_height, self._samples_per_pixel, *self._pinned["pointers"][:2],
                                      self.scene_packing(), self._pinned["pointers"][2])
        # the context now holds this scene; the DeviceData objects learn about it lazily
        self._pending = (targets.copy(), focus_planes.copy())
        self.scene_overwritten()
        return h_focus[:n].copy()
