"""BASELINE config 3: render / focus_value sweep on one B200 - resolution x1/x2/x4 and
samples per pixel 1..64 at 64 envs - for roofline characterisation.

    python scripts/sweep.py > gpurun_out/sweep.jsonl   (also prints a markdown table to stderr)
"""

import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    import numpy
    import torch

    from reinfocus_b200.graphics import render

    n = 64
    rng = numpy.random.Generator(numpy.random.PCG64DXSM(1234))
    targets = rng.uniform(5, 10, n).astype(numpy.float32)
    planes = rng.uniform(5, 10, n).astype(numpy.float32)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    rows = []
    for height in (300, 600, 1200):
        for spp in (1, 2, 4, 8, 16, 32, 64, 100):
            renderer = render.FastRenderer(samples_per_pixel=spp)
            renderer.update_targets(targets)
            renderer.update_focus_planes(planes)
            ctx = renderer.context
            renderer._sync_scene()
            ctx.rng_ensure(n * height * height, 0)
            gray = torch.empty((n, height, height), dtype=torch.uint8, device="cuda")
            out = torch.empty((n,), dtype=torch.float64, device="cuda")
            trace_ms, focus_ms = [], []
            for rep in range(5):
                e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                e0.record()
                ctx.render(n, height, height, spp, None, gray.data_ptr())
                e1.record()
                ctx.focus(n, height, height, gray.data_ptr(), 1, out.data_ptr())
                e2.record()
                torch.cuda.synchronize()
                if rep >= 2:
                    trace_ms.append(e0.elapsed_time(e1))
                    focus_ms.append(e1.elapsed_time(e2))
            t, f = float(numpy.mean(trace_ms)), float(numpy.mean(focus_ms))
            rays = n * height * height * spp
            row = {"envs": n, "height": height, "spp": spp, "trace_ms": t, "focus_ms": f,
                   "rays_per_s": rays / (t * 1e-3),
                   "trace_state_gbs": n * height * height * 33 / (t * 1e-3) / 1e9,
                   "focus_gbs": n * height * height / (f * 1e-3) / 1e9,
                   "focus_frac_of_hbm": n * height * height / (f * 1e-3) / 1e9 / hbm}
            rows.append(row)
            print(json.dumps(row), flush=True)
            del renderer
            torch.cuda.empty_cache()
    lines = ["| H | spp | trace ms | Grays/s | state+gray GB/s | focus ms | focus GB/s | of HBM peak |",
             "|---|---|---|---|---|---|---|---|"]
    for r in rows:
        lines.append(f"| {r['height']} | {r['spp']} | {r['trace_ms']:.3f} | {r['rays_per_s'] / 1e9:.1f} | "
                     f"{r['trace_state_gbs']:.0f} | {r['focus_ms']:.3f} | {r['focus_gbs']:.0f} | "
                     f"{r['focus_frac_of_hbm']:.1%} |")
    sys.stderr.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
