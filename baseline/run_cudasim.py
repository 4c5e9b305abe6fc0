"""Times the UNMODIFIED reference's hot path under numba's CUDA simulator
(NUMBA_ENABLE_CUDASIM=1, the reference's only CPU path) on the
box's host cores: FastRenderer.render + vision.focus_values at a reduced size (the simulator
runs one Python thread per CUDA thread: ~300 rays/s), rate extrapolated to the full frame.

    NUMBA_ENABLE_CUDASIM=1 python baseline/run_cudasim.py [--height 32] [--spp 8] [--envs 1]

Needs a copy of the reference (baseline/_ref, or /root/reference in the build container) and
the harness shim oracle/cudasim_shim.py. Never imported by the product."""

import argparse
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--envs", type=int, default=1)
    parser.add_argument("--height", type=int, default=32)
    parser.add_argument("--spp", type=int, default=8)
    parser.add_argument("--full-height", type=int, default=300)
    parser.add_argument("--full-spp", type=int, default=100)
    args = parser.parse_args()
    assert os.environ.get("NUMBA_ENABLE_CUDASIM") == "1", "run with NUMBA_ENABLE_CUDASIM=1"

    reference = os.path.join(REPO, "baseline", "_ref")
    if not os.path.isdir(os.path.join(reference, "reinfocus")):
        reference = "/root/reference"
    from oracle import cudasim_shim

    cudasim_shim.install(reference_root=reference)

    import numpy

    from reinfocus import vision
    from reinfocus.graphics import render

    rng = numpy.random.Generator(numpy.random.PCG64DXSM(1234))
    renderer = render.FastRenderer(samples_per_pixel=args.spp)
    targets = rng.uniform(5, 10, args.envs).astype(numpy.float32)
    planes = rng.uniform(5, 10, args.envs).astype(numpy.float32)
    t0 = time.perf_counter()
    renderer.update_targets(targets)
    renderer.update_focus_planes(planes)
    frames = renderer.render(args.height)
    t1 = time.perf_counter()
    focus = vision.focus_values(frames)
    t2 = time.perf_counter()
    rays = args.envs * args.height * args.height * args.spp
    rays_per_s = rays / (t1 - t0)
    full_rays = args.full_height * args.full_height * args.full_spp
    print(json.dumps({
        "impl": "reference-cudasim", "envs": args.envs, "height": args.height, "spp": args.spp,
        "rays": rays, "render_s": t1 - t0, "focus_s": t2 - t1, "rays_per_s": rays_per_s,
        "env_steps_per_s_extrapolated": rays_per_s / full_rays,
        "extrapolated_to": f"{args.full_height}x{args.full_height} x {args.full_spp} spp per env-step "
                           f"({full_rays} rays), render only",
        "cores": 1, "nproc": os.cpu_count(), "focus_sample": [float(f) for f in focus[:2]],
    }))


if __name__ == "__main__":
    main()
