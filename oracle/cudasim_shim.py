"""TEST INFRASTRUCTURE ONLY - never imported by the product (reinfocus_b200/).

Harness that lets the *unmodified* reference (``/root/reference``) run in a container
without a GPU, under numba's CUDA simulator (``NUMBA_ENABLE_CUDASIM=1``). The reference
does not import under the simulator as shipped; three
harness-side patches fix that without touching the reference sources:

1. ``numba.cuda.cudadrv.devicearray.DeviceNDArray`` does not exist under CUDASIM but is
   imported by name at reference ``graphics/render.py:8``, ``camera.py:10``,
   ``world.py:8`` ... -> alias it to the simulator's ``FakeCUDAArray``.
2. The reference decorates every device function with a bare ``@cuda.jit`` (e.g.
   ``graphics/vector.py:29``); the simulator treats those as kernels and refuses to call
   them without a launch configuration -> when such a "kernel" is called from inside a
   running kernel, run it as a device function.
3. ``cutil.py:118`` does ``isinstance(index, (int, numba.int32))`` which is not legal in
   plain Python because ``numba.int32`` is a numba type instance -> alias to numpy.int32.

Also provides stand-ins for ``gymnasium`` (absent offline) so that the reference's
``reinfocus.environments`` can be imported for env-sequence golden vectors.

Usage (only in the build container, where /root/reference exists)::

    NUMBA_ENABLE_CUDASIM=1 python -c "import oracle.cudasim_shim as s; s.install(); ..."
"""

import os
import sys

REFERENCE_ROOT = os.environ.get("REINFOCUS_REFERENCE", "/root/reference")


def install(reference_root: str = REFERENCE_ROOT, with_gym_stub: bool = True) -> None:
    """Patches numba's simulator and puts the reference on sys.path."""

    assert os.environ.get("NUMBA_ENABLE_CUDASIM") == "1", "run with NUMBA_ENABLE_CUDASIM=1"
    assert os.path.isdir(reference_root), f"reference not found at {reference_root}"

    import numpy
    import numba
    import numba.cuda.cudadrv.devicearray as da
    from numba.cuda.simulator import kernel as K
    from numba.cuda.simulator.kernelapi import swapped_cuda_module

    if not hasattr(da, "DeviceNDArray") or da.DeviceNDArray is not da.FakeCUDAArray:
        da.DeviceNDArray = da.FakeCUDAArray

    if not getattr(K.FakeCUDAKernel, "_rf_patched", False):
        original_call = K.FakeCUDAKernel.__call__

        def call(self, *args):
            context = K._get_kernel_context()
            if context is not None and not self._device:
                with swapped_cuda_module(self.fn, context):
                    return self.fn(*args)
            return original_call(self, *args)

        K.FakeCUDAKernel.__call__ = call
        K.FakeCUDAKernel._rf_patched = True

    numba.int32 = numpy.int32

    if with_gym_stub:
        install_gym_stub()
        install_plot_stub()

    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)


def install_plot_stub() -> None:
    """matplotlib is imported at module import time by the reference's
    environments/episode_visualizer.py:6-10 (and therefore by anything that imports the env
    classes); it is not installed offline. An empty stand-in is enough to import - nothing
    in the golden-vector harness plots."""

    try:
        import matplotlib  # noqa: F401

        return
    except ImportError:
        pass
    import types

    root = types.ModuleType("matplotlib")
    root.colormaps = {}
    colors = types.ModuleType("matplotlib.colors")
    colors.Colormap = object
    pyplot = types.ModuleType("matplotlib.pyplot")
    root.colors, root.pyplot = colors, pyplot
    sys.modules.update({"matplotlib": root, "matplotlib.colors": colors,
                        "matplotlib.pyplot": pyplot})


def install_gym_stub() -> None:
    """Registers the in-repo gymnasium stand-in (reinfocus_b200.gym_compat) under the
    module names the reference imports (gymnasium is not installable offline)."""

    try:
        import gymnasium  # noqa: F401

        return
    except ImportError:
        pass

    repo_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if repo_root not in sys.path:
        sys.path.insert(0, repo_root)

    from reinfocus_b200 import gym_compat

    gym_compat.install_as_gymnasium()
