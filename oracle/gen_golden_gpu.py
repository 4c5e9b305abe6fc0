"""TEST INFRASTRUCTURE ONLY - generates tests/golden/gpu_*.npz ON THE GPU BOX.

Runs the UNMODIFIED reference (a plain copy of /root/reference/reinfocus placed in the
git-ignored baseline/_ref/, the offline stand-in for `pip install --target baseline/_ref`,
which fails here because the reference's build backend `hatchling` is not installed)
through its real numba-CUDA path on a B200 and records frames, focus values and RNG
states. These are the "GPU profile" known answers that pin oracle/rf_oracle.c
(RFO_PROFILE_GPU) and the CUDA kernels.

    gpurun -- python oracle/gen_golden_gpu.py      # writes gpurun_out/golden_gpu/*.npz
    cp gpurun_out/golden_gpu/*.npz tests/golden/   # then commit

Also dumps the PTX/SASS numba generated for the kernel (profiles/ evidence for the FMA
contraction study).
"""

import hashlib
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(REPO, "baseline", "_ref")
OUT = os.path.join(REPO, "gpurun_out", "golden_gpu")

# name -> (samples_per_pixel, r_size, [(targets, focus_planes, frame_height), ...], store)
# store: "frames" keeps every frame, "hash" keeps sha256 + focus values + a few frames
CASES = {
    "a_persist": (2, 20, [([7.5], [7.5], 16), ([7.5], [7.5], 16)], "frames"),
    "b_two_envs": (8, 20, [([7.5, 7.5], [7.5, 5.0], 32), ([7.5, 7.5], [7.5, 5.0], 32)], "frames"),
    "c_grow": (3, 20, [([5.5, 9.0], [9.5, 6.0], 8), ([5.5, 9.0], [9.5, 6.0], 12),
                        ([5.5, 9.0], [9.5, 6.0], 8)], "frames"),
    "d_ends": (4, 20, [([5.0, 10.0, 6.3], [10.0, 5.0, 6.3], 24)], "frames"),
    "e_partial": (4, 20, [([6.0, 7.0, 8.0], [6.5, 7.0, 9.5], 10), ([9.25], [5.75], 10),
                           ([6.0, 7.0, 8.0], [6.25, 7.0, 9.5], 10)], "frames"),
    "f_spp100": (100, 20, [([8.125], [7.9], 8)], "frames"),
    "g_rsize": (5, 35, [([5.0, 9.99], [5.0, 9.99], 12)], "frames"),
    "h_full_frame": (100, 20, [([7.5], [7.0], 300)], "frames"),
    "i_odd_size": (7, 20, [([6.0, 9.5, 5.25], [9.0, 9.5, 5.0], 75)], "frames"),
    "j_full_two_calls": (100, 20, [([5.5, 9.0], [5.75, 6.0], 300), ([5.5, 9.0], [5.5, 9.0], 300)],
                         "hash"),
}


def extrema_case():
    """cached_focus_extrema((5, 10), 300) inputs (reference state_observer.py:295-320)."""

    import numpy

    ends = (5.0, 10.0)
    max_targets = numpy.linspace(*ends, 11)
    return (100, 20, [(list(numpy.append(ends, max_targets)),
                       list(numpy.append(ends[::-1], max_targets)), 300)], "hash")


def main():
    sys.path.insert(0, REF)
    import numpy
    from numba import cuda

    from reinfocus import vision
    from reinfocus.graphics import render

    os.makedirs(OUT, exist_ok=True)
    print("numba device:", cuda.get_current_device().name,
          cuda.get_current_device().compute_capability, flush=True)

    cases = dict(CASES)
    cases["k_extrema"] = extrema_case()
    for name, (spp, r_size, calls, store) in cases.items():
        t0 = time.time()
        renderer = render.FastRenderer(samples_per_pixel=spp, r_size=r_size)
        out = {"spp": numpy.int64(spp), "r_size": numpy.float64(r_size),
               "n_calls": numpy.int64(len(calls)), "store": numpy.array(store)}
        for i, (targets, planes, height) in enumerate(calls):
            renderer.update_targets(targets)
            renderer.update_focus_planes(planes)
            frames = renderer.render(height)
            out[f"targets_{i}"] = numpy.asarray(targets, dtype=numpy.float64)
            out[f"planes_{i}"] = numpy.asarray(planes, dtype=numpy.float64)
            out[f"height_{i}"] = numpy.int64(height)
            out[f"focus_{i}"] = numpy.asarray(vision.focus_values(frames), dtype=numpy.float64)
            out[f"sha256_{i}"] = numpy.array(hashlib.sha256(frames.tobytes()).hexdigest())
            out[f"n_states_{i}"] = numpy.int64(len(renderer._random_states))
            if store == "frames":
                out[f"frames_{i}"] = frames
            else:
                out[f"frames_{i}_first"] = frames[:1]
                out[f"channel_sums_{i}"] = frames.reshape(len(frames), -1, 3).sum(axis=1)
            states = renderer._random_states.copy_to_host()
            out[f"states_head_{i}"] = numpy.stack([states["s0"][:64], states["s1"][:64]], axis=1)
        numpy.savez_compressed(os.path.join(OUT, f"gpu_render_{name}.npz"), **out)
        print(f"done {name} in {time.time() - t0:.1f}s", flush=True)

    # what numba compiled: PTX and SASS of the kernel actually launched
    kernel = render.FastRenderer._device_render
    for sig, ptx in kernel.inspect_asm().items():
        with open(os.path.join(OUT, "numba_device_render.ptx"), "w") as f:
            f.write(ptx)
    try:
        for sig, sass in kernel.inspect_sass().items():
            with open(os.path.join(OUT, "numba_device_render.sass"), "w") as f:
                f.write(sass)
    except Exception as error:  # nvdisasm missing etc.
        print("inspect_sass failed:", error)


if __name__ == "__main__":
    main()
