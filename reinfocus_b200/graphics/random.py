"""RNG states on the GPU (reference graphics/random.py).

``make_random_states(n, seed)`` returns the same xoroshiro128+ states as
``numba.cuda.random.create_xoroshiro128p_states(n, seed)`` - state i is state 0 jumped
i * 2**64 steps - but builds them on the GPU by GF(2) matrix doubling instead of numba's
sequential CPU chain (csrc/rf_rng.cuh)."""

import numpy

from reinfocus_b200 import _lib


class RandomStates:
    """A device array of xoroshiro128+ states (stand-in for numba's DeviceNDArray)."""

    def __init__(self, tensor):
        self.tensor = tensor  # torch int64 [n, 2] on the GPU: (s0, s1) bit patterns

    def __len__(self) -> int:
        return int(self.tensor.shape[0])

    def copy_to_host(self) -> numpy.ndarray:
        return self.tensor.cpu().numpy().view(numpy.uint64).reshape(-1, 2).copy().view(
            _lib.STATE_DTYPE).reshape(-1)


def make_random_states(n: int, seed: int) -> RandomStates:
    """reference graphics/random.py:8-18"""

    import torch

    ctx = _lib.shared_context()
    tensor = torch.empty((n, 2), dtype=torch.int64, device=f"cuda:{ctx.device}")
    ctx.rng_init_device(tensor.data_ptr(), n, seed)
    return RandomStates(tensor)


def uniform_floats(states: RandomStates, draws: int):
    """``draws`` successive xoroshiro128p_uniform_float32 samples of every state (advances
    them); the batched host-callable form of reference graphics/random.py:21-33."""

    import torch

    ctx = _lib.shared_context()
    out = torch.empty((len(states), draws), dtype=torch.float32, device=states.tensor.device)
    ctx.rng_uniform_device(states.tensor.data_ptr(), len(states), draws, out.data_ptr())
    return out
