"""TEST INFRASTRUCTURE ONLY - golden frames for the generic render path (reference
graphics/render.py:31-119: device_render / render with Worlds + Cameras, spheres and
rectangles, up to 50 bounces), produced ON THE GPU BOX by the unmodified reference under
numba-CUDA (needs baseline/_ref, see oracle/gen_golden_gpu.py).

    gpurun -- python oracle/gen_golden_generic_gpu.py   # -> gpurun_out/golden_gpu/gpu_generic_*.npz
"""

import hashlib
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(REPO, "gpurun_out", "golden_gpu")


def scenes(sf, camera, graphics=None):
    """name -> (list of per-env shape lists, list of per-env camera kwargs, frame_shape, spp).
    `graphics`: the package's sphere / rectangle / vector modules (for the hand-built scene)."""

    P = sf.ShapeParameters
    extra = {}
    if graphics is not None:
        sphere, rectangle, vector = graphics
        # three shapes per env, odd frame size, varied cameras: nothing the factories build
        extra["three_shapes"] = (
            [[sphere.sphere(vector.v3f(-1.5, 0.5, -7.0), 1.25, vector.v2f(6, 10)),
              rectangle.rectangle(vector.v2f(-0.5, 2.5), vector.v2f(-2.0, 0.25), -9.0, vector.v2f(5, 3)),
              sphere.sphere(vector.v3f(1.0, -0.75, -4.0), 0.5)],
             sf.two_rect(P(9.0, texture_f=(7, 7)), P(4.5))],
            [dict(aperture=0.4, focus_distance=6.0, vfov=45),
             dict(look_from=(-0.2, 0.1, 0.3), aspect_ratio=1.25)],
            (37, 53), 9)
    return {
        **extra,
        "one_rect": ([sf.one_rect(P(r_size=30))], [dict()], (40, 60), 8),
        "two_rect": ([sf.two_rect()], [dict(focus_distance=5.0)], (30, 44), 6),
        "one_sphere": ([sf.one_sphere()], [dict()], (40, 60), 8),
        "two_sphere": ([sf.two_sphere()], [dict(focus_distance=20.0, aperture=0.3)], (36, 50), 5),
        "mixed_batch": ([sf.mixed(), sf.one_rect(), sf.two_sphere(P(8.0, texture_f=(4, 8)), P(12.0)),
                         sf.one_sphere(P(distance=6.0, size=1.5))],
                        [dict(), dict(focus_distance=7.0), dict(aspect_ratio=1.5, vfov=40),
                         dict(look_from=(0.5, 0.25, 1.0), look_at=(0.0, 0.0, -6.0), aperture=0.05)],
                        (33, 47), 7),
        # reference tests/graphics/render_test.py:57-80 (test_average_colour)
        "ref_test_sphere": ([sf.one_sphere()], [dict()], (100, 200), 10),
        "default_size": ([sf.mixed()], [dict()], (300, 600), 100),
    }


def main():
    sys.path.insert(0, os.path.join(REPO, "baseline", "_ref"))
    import numpy
    from numba import cuda

    from reinfocus.graphics import camera, rectangle, render, shape_factory, sphere, world
    from reinfocus.graphics import vector

    only = set(sys.argv[1:])
    os.makedirs(OUT, exist_ok=True)
    for name, (env_shapes, cam_kwargs, frame_shape, spp) in scenes(
            shape_factory, camera, (sphere, rectangle, vector)).items():
        if only and name not in only:
            continue
        cams = []
        for kw in cam_kwargs:
            kw = dict(kw)
            for key in ("look_from", "look_at", "up"):
                if key in kw:
                    kw[key] = vector.v3f(*kw[key])
            cams.append(camera.make_gpu_camera(**kw))
        worlds = world.Worlds(*env_shapes)
        cameras = camera.Cameras(*cams)
        frames = render.render(worlds, cameras, frame_shape=frame_shape, samples_per_pixel=spp)
        params, types, sizes = (a.copy_to_host() for a in worlds.device_data())
        out = {"frames_sha256": numpy.array(hashlib.sha256(frames.tobytes()).hexdigest()),
               "frame_shape": numpy.array(frame_shape), "spp": numpy.int64(spp),
               "shape_params": params, "shape_types": types, "env_sizes": sizes,
               "cameras": cameras.device_data().copy_to_host(),
               "channel_means": frames.reshape(len(frames), -1, 3).mean(axis=1)}
        if frames.size <= 600000:
            out["frames"] = frames
        else:
            out["frames_head"] = frames[:, :16]
        numpy.savez_compressed(os.path.join(OUT, f"gpu_generic_{name}.npz"), **out)
        print("done", name, frames.shape, out["channel_means"][0], flush=True)

    if only:
        return
    kernel = render.device_render
    for _, ptx in kernel.inspect_asm().items():
        open(os.path.join(OUT, "numba_generic_render.ptx"), "w").write(ptx)
    try:
        for _, sass in kernel.inspect_sass().items():
            open(os.path.join(OUT, "numba_generic_render.sass"), "w").write(sass)
    except Exception as error:  # pylint: disable=broad-except
        print("inspect_sass failed:", error)


if __name__ == "__main__":
    main()
