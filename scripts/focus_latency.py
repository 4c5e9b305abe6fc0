"""Focus-stencil time over batch sizes (gray input on the device), one B200."""

import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    import torch

    from reinfocus_b200 import _lib

    ctx = _lib.shared_context()
    only = None
    if "--only" in sys.argv:  # --only <height> <envs>: one case (profiler captures)
        at = sys.argv.index("--only")
        only = (int(sys.argv[at + 1]), int(sys.argv[at + 2]))
    for height in (300, 600):
        for n in (1, 8, 64, 256, 1024, 4096):
            if only and (height, n) != only:
                continue
            if n * height * height > 2**31:
                continue
            gray = torch.randint(0, 256, (n, height, height), dtype=torch.uint8, device="cuda")
            out = torch.empty(n, dtype=torch.float64, device="cuda")
            for _ in range(3):
                ctx.focus(n, height, height, gray.data_ptr(), 1, out.data_ptr())
            start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20
            start.record()
            for _ in range(reps):
                ctx.focus(n, height, height, gray.data_ptr(), 1, out.data_ptr())
            stop.record()
            torch.cuda.synchronize()
            ms = start.elapsed_time(stop) / reps
            print(json.dumps({"H": height, "envs": n, "focus_us": ms * 1e3,
                              "GB/s": n * height * height / (ms * 1e-3) / 1e9}), flush=True)


if __name__ == "__main__":
    main()
