"""Times the UNMODIFIED reference's numba-CUDA hot path on the GPU box (reported baseline,
BASELINE.md section 4 item 1): FastRenderer + vision.focus_values driven exactly as
FocusObserver.observe does (reference state_observer.py:377-383).

    python baseline/run_numba_cuda.py --envs 64 --steps 3 [--out gpurun_out/numba_64.json]

Needs baseline/_ref (see oracle/gen_golden_gpu.py). Never imported by the product."""

import argparse
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "baseline", "_ref"))


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--envs", type=int, default=64)
    parser.add_argument("--steps", type=int, default=3)
    parser.add_argument("--warmup", type=int, default=1)
    parser.add_argument("--height", type=int, default=300)
    parser.add_argument("--spp", type=int, default=100)
    parser.add_argument("--out", default=None)
    parser.add_argument("--generic-calls", type=int, default=0,
                        help="also time render.render (general scenes, reference render.py:88-119) on its "
                             "default 300 x 600 x 100 spp frame of shape_factory.mixed(), this many calls")
    args = parser.parse_args()

    import numpy
    from numba import cuda

    from reinfocus import vision
    from reinfocus.graphics import render

    rng = numpy.random.Generator(numpy.random.PCG64DXSM(1234))
    renderer = render.FastRenderer(samples_per_pixel=args.spp)
    timings = {"pack": [], "render": [], "focus": [], "step": []}
    t_init = None
    for step in range(args.warmup + args.steps):
        targets = rng.uniform(5, 10, args.envs).astype(numpy.float32)
        planes = rng.uniform(5, 10, args.envs).astype(numpy.float32)
        t0 = time.perf_counter()
        renderer.update_targets(targets)
        renderer.update_focus_planes(planes)
        t1 = time.perf_counter()
        frames = renderer.render(args.height)  # includes RNG init + JIT on the first call
        t2 = time.perf_counter()
        focus = vision.focus_values(frames)
        t3 = time.perf_counter()
        if step == 0:
            t_init = t2 - t1
        if step >= args.warmup:
            timings["pack"].append(t1 - t0)
            timings["render"].append(t2 - t1)
            timings["focus"].append(t3 - t2)
            timings["step"].append(t3 - t0)
    mean = {k: float(numpy.mean(v)) for k, v in timings.items()}
    rays = args.envs * args.height * args.height * args.spp
    result = {
        "impl": "reference-numba-cuda",
        "device": cuda.get_current_device().name.decode()
        if isinstance(cuda.get_current_device().name, bytes) else str(cuda.get_current_device().name),
        "envs": args.envs, "height": args.height, "spp": args.spp, "steps": args.steps,
        "first_call_s_incl_jit_and_rng_init": t_init,
        "mean_s": mean,
        "env_steps_per_s": args.envs / mean["step"],
        "rays_per_s_render_call": rays / mean["render"],
        "focus_sample": [float(f) for f in focus[:4]],
    }
    if args.generic_calls:
        from reinfocus.graphics import camera, shape_factory, world

        worlds = world.Worlds(shape_factory.mixed())
        cameras = camera.Cameras(camera.make_gpu_camera())
        render.render(worlds, cameras)  # JIT
        times = []
        for _ in range(args.generic_calls):
            t0 = time.perf_counter()
            frames = render.render(worlds, cameras)
            times.append(time.perf_counter() - t0)
        result["generic"] = {"scene": "shape_factory.mixed(), one env, 300 x 600, 100 spp (render.render defaults)",
                             "calls": args.generic_calls, "mean_s": float(numpy.mean(times)),
                             "min_s": float(numpy.min(times)),
                             "rays_per_s": 300 * 600 * 100 / float(numpy.mean(times)),
                             "mean_colour": [float(c) for c in frames.reshape(-1, 3).mean(axis=0)]}
    line = json.dumps(result)
    print(line)
    if args.out:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out, "w") as f:
            f.write(line + "\n")


if __name__ == "__main__":
    main()
