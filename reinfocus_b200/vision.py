"""Image focus measure (reference vision.py): variance of the 8-bit Laplacian of the
3x3-median-filtered gray image, computed by one CUDA launch for a whole batch
(csrc/rf_focus.cuh) instead of a Python loop of three OpenCV calls per image."""

from collections.abc import Sequence

import numpy

from reinfocus_b200 import _lib


def _as_device_uint8(images):
    import torch

    if isinstance(images, torch.Tensor):
        assert images.dtype == torch.uint8, "images must be uint8"
        if not images.is_cuda:
            images = images.cuda(non_blocking=True)
        return images.contiguous()
    images = numpy.ascontiguousarray(images)
    assert images.dtype == numpy.uint8, "images must be uint8"
    return torch.from_numpy(images).cuda()


def focus_values_device(images):
    """Focus values of a batch of RGB (n, H, W, 3) or gray (n, H, W) uint8 images as a torch
    float64 CUDA tensor (n,). Accepts NumPy arrays or torch tensors."""

    import torch

    dev = _as_device_uint8(images)
    if dev.ndim == 4:
        assert dev.shape[-1] == 3, "expected RGB images (n, H, W, 3)"
        channels = 3
    else:
        assert dev.ndim == 3, "expected gray images (n, H, W)"
        channels = 1
    n, height, width = int(dev.shape[0]), int(dev.shape[1]), int(dev.shape[2])
    ctx = _lib.shared_context(dev.device.index)
    out = torch.empty((n,), dtype=torch.float64, device=dev.device)
    if n:
        with torch.cuda.device(dev.device):
            ctx.focus(n, height, width, dev.data_ptr(), channels, out.data_ptr())
    return out


def focus_value(image) -> float:
    """reference vision.py:11-25: how 'in focus' an RGB image (H, W, 3) is."""

    return float(focus_values_device(image[None])[0])


def focus_values(images) -> Sequence[float]:
    """reference vision.py:28-39: focus values of a number of RGB images."""

    return focus_values_device(images).cpu().tolist()
