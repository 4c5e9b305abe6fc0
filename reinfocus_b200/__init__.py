"""reinfocus_b200: B200-native (sm_100a) implementation of the reinfocus per-step hot path.

Drop-in module layout (reference module -> this package):
    reinfocus.graphics.render  -> reinfocus_b200.graphics.render  (FastRenderer, render)
    reinfocus.graphics.random  -> reinfocus_b200.graphics.random  (make_random_states)
    reinfocus.vision           -> reinfocus_b200.vision           (focus_value, focus_values)
    reinfocus.environments.*   -> reinfocus_b200.environments.*   (Environment, VectorEnvironment
                                                                   and the strategy objects)
The compute lives in libreinfocus_b200.so (hand-written CUDA behind a C-ABI, see
include/reinfocus_b200.h); there is no CPU fallback.
"""

__version__ = "0.1.0"


def _check_numpy():
    """The host packing (graphics/world.py, camera.py) and the device packing kernel reproduce
    the reference's float32 scene parameters under NumPy 2 scalar promotion (NEP 50: Python
    floats are weak, so ``float32 * python_float`` stays float32) - the rules of the
    environment the parity goldens were recorded in (NumPy 2.3, numba 0.65). Under NumPy 1.x
    the same reference expressions take a float64 product, which can differ by one float32
    ulp: refuse to run there rather than be silently off."""

    import numpy

    if int(numpy.__version__.split(".")[0]) < 2:
        raise ImportError(
            f"reinfocus_b200 needs NumPy >= 2 (found {numpy.__version__}): its scene packing follows "
            "NumPy 2 scalar promotion, which is what the parity vectors were recorded under")


_check_numpy()
