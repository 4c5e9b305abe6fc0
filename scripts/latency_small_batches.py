"""Step latency (host API, rf_step_host) for small vector envs, single- vs multi-context
tracer: picks the batch size from which several pixels per thread pay off."""

import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    import numpy
    import torch

    from reinfocus_b200 import _lib
    from reinfocus_b200.graphics import render

    rng = numpy.random.Generator(numpy.random.PCG64DXSM(1234))
    for n in (1, 2, 4, 8, 13, 16, 32, 64):
        row = {"envs": n}
        for contexts in (0, 4, 8):
            renderer = render.FastRenderer()
            renderer.context.set_option(_lib.OPT_TRACE_CONTEXTS, contexts)
            targets = rng.uniform(5, 10, (12, n)).astype(numpy.float32)
            planes = rng.uniform(5, 10, (12, n)).astype(numpy.float32)
            for i in range(4):
                renderer.step_focus(targets[i], planes[i], 300)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(4, 12):
                renderer.step_focus(targets[i], planes[i], 300)
            row[f"ms_ctx{contexts}"] = (time.perf_counter() - t0) / 8 * 1e3
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
