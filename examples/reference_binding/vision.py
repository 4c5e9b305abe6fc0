"""Stand-in for the reference's reinfocus/vision.py (vision.py:11-39): the cv2 loop
(cvtColor -> medianBlur -> Laplacian -> var per image) replaced by one rf_focus call."""

import numpy
import torch

from examples.reference_binding import render as _binding

_ctx = None


def focus_values(images):
    """uint8 (n, H, W, 3) images -> list of n focus values (float64)."""

    global _ctx
    if _ctx is None:
        _ctx = _binding.new_context()
    dev = torch.as_tensor(numpy.ascontiguousarray(images)).cuda().contiguous()
    out = torch.empty(len(dev), dtype=torch.float64, device="cuda")
    _binding.check(_ctx, _binding._lib.rf_focus(_ctx, len(dev), dev.shape[1], dev.shape[2], dev.data_ptr(), 3,
                                               out.data_ptr(), None))
    return out.cpu().tolist()


def focus_value(image):
    return focus_values(numpy.asarray(image)[None])[0]
