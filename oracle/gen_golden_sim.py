"""TEST INFRASTRUCTURE ONLY - generates tests/golden/sim_*.npz and rng/focus fixtures.

Runs the UNMODIFIED reference (/root/reference) in the build container:
  * graphics (FastRenderer, FastWorlds, FastCameras) under numba's CUDA simulator via
    oracle/cudasim_shim.py -> "SIM profile" golden frames, device data and focus values;
  * numba.cuda.random's own CPU initialiser / uniform sampler (third-party, the RNG the
    reference calls at graphics/random.py:18,33) -> RNG known answers;
  * reinfocus.vision (cv2 + numpy, runs natively) -> focus-measure known answers.

The reference cannot travel to the GPU box, so the outputs are committed as small
fixtures. Re-run with:  python oracle/gen_golden_sim.py [case ...]
(about ten minutes on 8 cores; cases run in parallel processes).
"""

import multiprocessing
import os
import sys

os.environ["NUMBA_ENABLE_CUDASIM"] = "1"

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, REPO)

# name -> (samples_per_pixel, r_size, [ (targets, focus_planes, frame_height), ... ])
# Each case is ONE FastRenderer driven through a sequence of update/render calls, so that
# RNG-state persistence, growth re-seeding (render.py:256-257) and sub-batch renders
# (state_observer.py:377-383 on partial resets) are all pinned.
RENDER_CASES = {
    "a_persist": (2, 20, [([7.5], [7.5], 16), ([7.5], [7.5], 16)]),
    "b_two_envs": (8, 20, [([7.5, 7.5], [7.5, 5.0], 32), ([7.5, 7.5], [7.5, 5.0], 32)]),
    "c_grow": (3, 20, [([5.5, 9.0], [9.5, 6.0], 8), ([5.5, 9.0], [9.5, 6.0], 12),
                        ([5.5, 9.0], [9.5, 6.0], 8)]),
    "d_ends": (4, 20, [([5.0, 10.0, 6.3], [10.0, 5.0, 6.3], 24)]),
    "e_partial": (4, 20, [([6.0, 7.0, 8.0], [6.5, 7.0, 9.5], 10), ([9.25], [5.75], 10),
                           ([6.0, 7.0, 8.0], [6.25, 7.0, 9.5], 10)]),
    "f_spp100": (100, 20, [([8.125], [7.9], 8)]),
    "g_rsize": (5, 35, [([5.0, 9.99], [5.0, 9.99], 12)]),
}


def run_render_case(name):
    import numpy

    import oracle.cudasim_shim as shim

    shim.install(with_gym_stub=False)

    from reinfocus import vision
    from reinfocus.graphics import render

    spp, r_size, calls = RENDER_CASES[name]
    renderer = render.FastRenderer(samples_per_pixel=spp, r_size=r_size)
    out = {"spp": numpy.int64(spp), "r_size": numpy.float64(r_size),
           "n_calls": numpy.int64(len(calls))}
    for i, (targets, planes, height) in enumerate(calls):
        renderer.update_targets(targets)
        renderer.update_focus_planes(planes)
        frames = renderer.render(height)
        cams = renderer._cameras.device_data()
        out[f"targets_{i}"] = numpy.asarray(targets, dtype=numpy.float64)
        out[f"planes_{i}"] = numpy.asarray(planes, dtype=numpy.float64)
        out[f"height_{i}"] = numpy.int64(height)
        out[f"frames_{i}"] = frames
        out[f"world_{i}"] = renderer._worlds.device_data().copy_to_host()
        out[f"cam_dyn_{i}"] = cams[0].copy_to_host()
        out[f"cam_static_{i}"] = numpy.array([*cams[1], *cams[2], *cams[3]], dtype=numpy.float32)
        out[f"lens_{i}"] = numpy.float64(cams[4])
        out[f"focus_{i}"] = numpy.asarray(vision.focus_values(frames), dtype=numpy.float64)
        out[f"n_states_{i}"] = numpy.int64(len(renderer._random_states))
    numpy.savez_compressed(os.path.join(GOLDEN, f"sim_render_{name}.npz"), **out)
    return name


def run_rng():
    """numba's own CPU-side initialiser and sampler, compiled (not simulated)."""

    import subprocess

    code = r"""
import numpy, sys
from numba.cuda.random import (init_xoroshiro128p_states_cpu, xoroshiro128p_dtype,
                               xoroshiro128p_uniform_float32, xoroshiro128p_next)
out = {}
for seed in (0, 1, 12345, 2**63 + 5):
    st = numpy.empty(3000, dtype=xoroshiro128p_dtype)
    init_xoroshiro128p_states_cpu(st, numpy.uint64(seed), numpy.uint64(0))
    out[f"states_seed{seed}"] = numpy.stack([st["s0"], st["s1"]], axis=1)
    draws = numpy.empty((8, 16), dtype=numpy.float32)
    for i in range(8):
        for k in range(16):
            draws[i, k] = xoroshiro128p_uniform_float32(st, i)
    out[f"uniform_seed{seed}"] = draws
    out[f"after_seed{seed}"] = numpy.stack([st["s0"][:8], st["s1"][:8]], axis=1)
numpy.savez_compressed(sys.argv[1], **out)
"""
    env = dict(os.environ)
    env.pop("NUMBA_ENABLE_CUDASIM", None)
    subprocess.run([sys.executable, "-c", code, os.path.join(GOLDEN, "rng_numba.npz")],
                   check=True, env=env)
    return "rng"


def run_focus():
    """reinfocus.vision (cv2 + numpy) on assorted images; small inputs are stored, large
    ones are regenerated from a recorded PCG64 seed."""

    import numpy

    sys.path.insert(0, "/root/reference")
    from reinfocus import vision

    out = {}
    # the reference's own vision tests (tests/vision_test.py:14-34)
    checker = numpy.zeros((10, 10, 3), dtype=numpy.uint8)
    checker[::2, ::2] = 255
    checker[1::2, 1::2] = 255
    small = {
        "zeros": numpy.zeros((5, 5, 3), dtype=numpy.uint8),
        "ones": numpy.ones((5, 5, 3), dtype=numpy.uint8),
        "checker10": checker,
    }
    rng = numpy.random.Generator(numpy.random.PCG64(2024))
    for shape in ((1, 1), (1, 7), (7, 1), (2, 2), (3, 3), (4, 9), (16, 16), (31, 17), (64, 48)):
        small[f"rand_{shape[0]}x{shape[1]}"] = rng.integers(
            0, 256, size=shape + (3,), dtype=numpy.uint8)
    # smooth-ish images so that the Laplacian is not saturated everywhere
    yy, xx = numpy.mgrid[0:40, 0:56]
    wave = (127.5 + 100 * numpy.sin(xx / 3.0) * numpy.cos(yy / 5.0))
    small["wave_40x56"] = numpy.stack([wave, wave[::-1], wave.T[:40, :40].repeat(2, 1)[:, :56]],
                                      axis=-1).astype(numpy.uint8)
    names = sorted(small)
    out["small_names"] = numpy.array(names)
    for name in names:
        out[f"img_{name}"] = small[name]
        out[f"fv_{name}"] = numpy.float64(vision.focus_value(small[name]))
    # large random images: seed + shape recorded, pixels regenerated by the tests
    large = [(300, 300, 11), (300, 300, 12), (600, 600, 13), (257, 301, 14)]
    out["large_specs"] = numpy.array(large, dtype=numpy.int64)
    fvs = []
    for h, w, seed in large:
        g = numpy.random.Generator(numpy.random.PCG64(seed))
        # low-contrast noise around a gradient keeps the Laplacian unsaturated
        base = numpy.linspace(0, 200, w)[None, :, None] + g.integers(0, 40, size=(h, w, 3))
        img = base.astype(numpy.uint8)
        fvs.append(vision.focus_value(img))
    out["large_fv"] = numpy.array(fvs, dtype=numpy.float64)
    batch = rng.integers(0, 256, size=(3, 20, 20, 3), dtype=numpy.uint8)
    out["batch_imgs"] = batch
    out["batch_fv"] = numpy.asarray(vision.focus_values(batch), dtype=numpy.float64)
    numpy.savez_compressed(os.path.join(GOLDEN, "focus_cv2.npz"), **out)
    return "focus"


def _dispatch(job):
    if job == "rng":
        return run_rng()
    if job == "focus":
        return run_focus()
    return run_render_case(job)


if __name__ == "__main__":
    os.makedirs(GOLDEN, exist_ok=True)
    jobs = sys.argv[1:] or (["rng", "focus"] + list(RENDER_CASES))
    with multiprocessing.get_context("spawn").Pool(min(8, len(jobs))) as pool:
        for done in pool.imap_unordered(_dispatch, jobs):
            print("done", done, flush=True)
