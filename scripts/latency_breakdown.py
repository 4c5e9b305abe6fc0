"""Where a small-batch step goes: FastRenderer.step_focus (Python staging + C call) vs the C
call alone (rf_step_positions_host on pinned buffers) vs the two kernels (CUDA events)."""

import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    import numpy
    import torch

    from reinfocus_b200.graphics import render

    rng = numpy.random.Generator(numpy.random.PCG64DXSM(1234))
    for n in (1, 2, 4, 8, 16):
        renderer = render.FastRenderer()
        ctx = renderer.context
        targets = rng.uniform(5, 10, (40, n)).astype(numpy.float32)
        planes = rng.uniform(5, 10, (40, n)).astype(numpy.float32)
        for i in range(8):
            renderer.step_focus(targets[i], planes[i], 300)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(8, 40):
            renderer.step_focus(targets[i], planes[i], 300)
        full = (time.perf_counter() - t0) / 32 * 1e3
        h_targets, h_planes, h_focus = renderer._pinned["pointers"]
        packing = renderer.scene_packing()
        t0 = time.perf_counter()
        for i in range(32):
            ctx.step_positions_host(n, 300, 100, h_targets, h_planes, packing, h_focus)
        c_call = (time.perf_counter() - t0) / 32 * 1e3
        gray = torch.empty((n, 300, 300), dtype=torch.uint8, device="cuda")
        focus = torch.empty((n,), dtype=torch.float64, device="cuda")
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        trace, stencil = [], []
        for i in range(8):
            e[0].record()
            ctx.render(n, 300, 300, 100, None, gray.data_ptr())
            e[1].record()
            ctx.focus(n, 300, 300, gray.data_ptr(), 1, focus.data_ptr())
            e[2].record()
            torch.cuda.synchronize()
            trace.append(e[0].elapsed_time(e[1]))
            stencil.append(e[1].elapsed_time(e[2]))
        print(json.dumps({"envs": n, "step_focus_ms": full, "c_call_ms": c_call,
                          "trace_kernel_ms": float(numpy.median(trace)),
                          "focus_kernel_ms": float(numpy.median(stencil))}), flush=True)


if __name__ == "__main__":
    main()
