"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C-ABI
exactly as a user of the reference-shaped Python API would, against the CPU oracle on the
same seeded inputs, against the committed golden fixtures, and - at full size - through
size-independent properties. Integer/byte work is compared bit-exactly."""

import glob
import hashlib
import os

import numpy
import pytest

import oracle

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(REPO, "tests", "golden")


@pytest.fixture(scope="module")
def torch():
    import torch

    assert torch.cuda.is_available(), "these tests need a GPU"
    return torch


@pytest.fixture(scope="module")
def ctx(torch):
    from reinfocus_b200 import _lib

    return _lib.shared_context()


def _renderer(**kwargs):
    from reinfocus_b200.graphics import render

    return render.FastRenderer(**kwargs)


# --------------------------------------------------------------------------- RNG (a3, a6)


@pytest.mark.parametrize("n,seed", [(1, 0), (2, 0), (3, 5), (255, 0), (256, 1), (257, 2),
                                    (1000, 12345), (90000, 0), (300007, 2**63 + 5)])
def test_rng_init_matches_oracle(torch, n, seed):
    from reinfocus_b200.graphics import random

    states = random.make_random_states(n, seed)
    assert len(states) == n
    numpy.testing.assert_array_equal(states.copy_to_host(), oracle.rng_states(n, seed, doubling=True))


def test_rng_init_and_draws_match_numba_golden(torch):
    from reinfocus_b200.graphics import random

    gold = numpy.load(os.path.join(GOLDEN, "rng_numba.npz"))
    for seed in (0, 1, 12345, 2**63 + 5):
        states = random.make_random_states(3000, seed)
        host = states.copy_to_host()
        got = numpy.stack([host["s0"], host["s1"]], axis=1)
        numpy.testing.assert_array_equal(got[8:], gold[f"states_seed{seed}"][8:])
        draws = random.uniform_floats(states, 16).cpu().numpy()
        numpy.testing.assert_array_equal(draws[:8], gold[f"uniform_seed{seed}"])
        after = states.copy_to_host()
        numpy.testing.assert_array_equal(
            numpy.stack([after["s0"][:8], after["s1"][:8]], axis=1), gold[f"after_seed{seed}"])


def test_uniform_draws_match_oracle_including_small_values(torch):
    """The integer->float32 conversion must round exactly like float32(float64(k) * 2**-53),
    including results that need the low bits (tiny k) and the rounding-up to 1.0."""

    from reinfocus_b200 import _lib
    from reinfocus_b200.graphics import random

    # craft states whose next output r = s0 + s1 covers edge patterns
    outputs = [0, 1 << 11, (1 << 11) - 1, (1 << 64) - 1, (1 << 64) - (1 << 11), 0xFFFFFFFFFFFFF800,
               0x0000000000FFF800, 0x00000100000007FF, 0x8000008000000000, 0x8000007FFFFFFFFF,
               0x0000000080000000, 0x00000000FFFFFFFF]
    host = numpy.zeros(len(outputs), dtype=_lib.STATE_DTYPE)
    host["s0"] = numpy.array(outputs, dtype=numpy.uint64)
    states = random.RandomStates(
        torch.from_numpy(host.view(numpy.uint64).reshape(-1, 2).view(numpy.int64).copy()).cuda())
    got = random.uniform_floats(states, 3).cpu().numpy()
    want_states = host.copy()
    want = numpy.array([[oracle.uniform_float32(want_states, i) for _ in range(3)]
                        for i in range(len(outputs))], dtype=numpy.float32)
    numpy.testing.assert_array_equal(got, want)
    numpy.testing.assert_array_equal(states.copy_to_host(), want_states)


# ----------------------------------------------------------------- checker table (a10)


def test_checker_table_matches_float64_sin_for_every_float32(ctx):
    assert ctx.selftest_checker() == 0


@pytest.mark.parametrize("size", [300, 7, 75, 600])
def test_hoisted_reciprocal_division_equals_ddiv_for_every_input(ctx, size):
    """float32((x + U) / size) for every pixel coordinate x and every float32 U in [0, 1]."""

    from reinfocus_b200 import _lib

    assert ctx.selftest(_lib.SELFTEST_PIXEL_DIV, size) == 0


def test_fma_pipe_checker_parity_equals_the_table_for_every_float32(ctx):
    """The multi-pixel tracer reads the checker cell parity off RZ / RU fused multiply-adds
    (2^23 + floor(32 u) in the mantissa): same parity as the table-based cell, and "exact"
    claimed iff 32 u is not an integer, for every float32 u in [0, 1]."""

    from reinfocus_b200 import _lib

    assert ctx.selftest(_lib.SELFTEST_CHECKER_PAIR) == 0


def test_branch_free_inverse_length_equals_intrinsics_for_every_input(ctx):
    from reinfocus_b200 import _lib

    assert ctx.selftest(_lib.SELFTEST_INV_LENGTH) == 0


def test_hoisted_reciprocal_float32_division_equals_fdiv(ctx):
    """(r + P) / (r + r) with the reciprocal refinement hoisted: every numerator in [0, c] for
    96 divisors c spread over [0.5, 8)."""

    from reinfocus_b200 import _lib

    assert ctx.selftest(_lib.SELFTEST_CONST_DIV, 96) == 0


@pytest.mark.parametrize("targets,planes,height,spp", [
    ([7.5, 5.0, 10.0, 6.3], [7.5, 10.0, 5.0, 6.3], 40, 12),
    ([9.0, 5.5], [5.25, 9.75], 97, 5),
])
def test_specialised_kernel_equals_literal_kernel(torch, targets, planes, height, spp):
    """A/B on the device: the default-camera kernel (float32 shortcuts) against the
    any-camera kernel (the literal float64-typed statement), same scene, same states."""

    from reinfocus_b200 import _lib

    fast, literal = _renderer(samples_per_pixel=spp), _renderer(samples_per_pixel=spp)
    literal.context.set_option(_lib.OPT_FORCE_GENERIC, 1)
    for renderer in (fast, literal):
        renderer.update_targets(targets)
        renderer.update_focus_planes(planes)
    for _ in range(2):
        a, b = fast.render(height), literal.render(height)
        numpy.testing.assert_array_equal(a, b)
    assert fast.context.last_trace_kernel() >= 1 and literal.context.last_trace_kernel() == 0
    numpy.testing.assert_array_equal(fast.context.rng_export(), literal.context.rng_export())


@pytest.mark.parametrize("contexts", [0, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("height,n", [(36, 3), (50, 2), (75, 1)])
def test_multi_context_kernel_equals_literal_kernel(torch, contexts, height, n):
    """Every pixels-per-thread setting of the default-camera kernel against the literal
    kernel: RGB frames, gray frames and the RNG states left behind (ragged sizes: 36*36,
    50*50 and 75*75 are not multiples of the block's pixel count)."""

    from reinfocus_b200 import _lib

    spp = 6
    targets, planes = [7.5, 5.0, 9.5][:n], [7.0, 10.0, 9.5][:n]
    tested, literal = _renderer(samples_per_pixel=spp), _renderer(samples_per_pixel=spp)
    tested.context.set_option(_lib.OPT_TRACE_CONTEXTS, contexts)
    literal.context.set_option(_lib.OPT_FORCE_GENERIC, 1)
    for renderer in (tested, literal):
        renderer.update_targets(targets)
        renderer.update_focus_planes(planes)
    numpy.testing.assert_array_equal(tested.render(height), literal.render(height))
    assert tested.context.last_trace_kernel() == (contexts or 1)
    gray_a = tested.render_gray_device(height).cpu().numpy()
    gray_b = literal.render_gray_device(height).cpu().numpy()
    numpy.testing.assert_array_equal(gray_a, gray_b)
    numpy.testing.assert_array_equal(tested.context.rng_export(), literal.context.rng_export())


@pytest.mark.parametrize("contexts", [4, 7, 8])
@pytest.mark.parametrize("height,spp,n", [(1200, 2, 2), (601, 3, 1), (300, 100, 1)])
def test_multi_context_kernel_at_sweep_sizes_matches_oracle(torch, height, spp, n, contexts):
    """The multi-pixel tracer (4 and 8 pixels per thread are what the library picks by batch
    size, 7 was round 1's) at the sweep's frame sizes (and an odd one), gray path, against the
    oracle: pixels and the RNG states left behind."""

    from reinfocus_b200 import _lib

    targets, planes = [6.25, 9.5][:n], [7.0, 9.0][:n]
    gpu = _renderer(samples_per_pixel=spp)
    gpu.context.set_option(_lib.OPT_TRACE_CONTEXTS, contexts)
    cpu = oracle.OracleFastRenderer(samples_per_pixel=spp, profile=oracle.PROFILE_GPU)
    for renderer in (gpu, cpu):
        renderer.update_targets(targets)
        renderer.update_focus_planes(planes)
    gray = gpu.render_gray_device(height).cpu().numpy()
    assert gpu.context.last_trace_kernel() == contexts
    numpy.testing.assert_array_equal(gray, oracle.gray(cpu.render(height)))
    numpy.testing.assert_array_equal(gpu.context.rng_export(), cpu.states)


def test_frames_beyond_the_multi_pixel_kernels_coordinate_range(torch):
    """The multi-pixel tracer keeps pixel coordinates as half floats (exact below 2048); a
    2052-pixel frame must take the one-pixel kernel, whatever the option says, and match the
    oracle."""

    from reinfocus_b200 import _lib

    height, spp = 2052, 1
    gpu = _renderer(samples_per_pixel=spp)
    gpu.context.set_option(_lib.OPT_TRACE_CONTEXTS, 8)
    cpu = oracle.OracleFastRenderer(samples_per_pixel=spp, profile=oracle.PROFILE_GPU)
    for renderer in (gpu, cpu):
        renderer.update_targets([8.0])
        renderer.update_focus_planes([7.0])
    gray = gpu.render_gray_device(height).cpu().numpy()
    assert gpu.context.last_trace_kernel() == 1
    numpy.testing.assert_array_equal(gray, oracle.gray(cpu.render(height)))
    # 2048 itself is inside the range: the largest frame of the multi-pixel kernel
    edge = _renderer(samples_per_pixel=spp)
    edge.context.set_option(_lib.OPT_TRACE_CONTEXTS, 8)
    edge_cpu = oracle.OracleFastRenderer(samples_per_pixel=spp, profile=oracle.PROFILE_GPU)
    for renderer in (edge, edge_cpu):
        renderer.update_targets([8.0])
        renderer.update_focus_planes([7.0])
    gray = edge.render_gray_device(2048).cpu().numpy()
    assert edge.context.last_trace_kernel() == 8
    numpy.testing.assert_array_equal(gray, oracle.gray(edge_cpu.render(2048)))
    with pytest.raises(AssertionError):
        edge.context.set_option(_lib.OPT_TRACE_CONTEXTS, 9)


def test_pixels_per_thread_follow_the_batch_size(torch):
    """Default option: one pixel per thread for one or two envs, four from three on (in small
    blocks that spread over the SMs), eight once the batch fills the GPU; all leave the same
    frames as the oracle (below), here only the choice is checked."""

    for envs, kernel in ((1, 1), (2, 1), (3, 4), (5, 4), (8, 4), (13, 4), (24, 4), (32, 8), (64, 8)):
        renderer = _renderer(samples_per_pixel=1)
        renderer.update_targets([7.0] * envs), renderer.update_focus_planes([6.0] * envs)
        renderer.render_gray_device(300)
        assert renderer.context.last_trace_kernel() == kernel, envs


@pytest.mark.parametrize("envs", [3, 8, 24, 64])
def test_default_kernel_choice_matches_oracle_at_every_batch_size_class(torch, envs):
    """The kernels the library picks by itself for small and medium batches (4 x 64, 4 x 128
    and 8 x 256 pixels x threads) against the oracle: gray frames and RNG states."""

    height, spp = 300, 2
    rng = numpy.random.default_rng(envs)
    targets = rng.uniform(5, 10, envs).astype(numpy.float32)
    planes = rng.uniform(5, 10, envs).astype(numpy.float32)
    gpu = _renderer(samples_per_pixel=spp)
    cpu = oracle.OracleFastRenderer(samples_per_pixel=spp, profile=oracle.PROFILE_GPU)
    for renderer in (gpu, cpu):
        renderer.update_targets(targets)
        renderer.update_focus_planes(planes)
    for _ in range(2):  # twice: the states of the first call feed the second
        gray = gpu.render_gray_device(height).cpu().numpy()
        assert gpu.context.last_trace_kernel() in (4, 8)
        numpy.testing.assert_array_equal(gray, oracle.gray(cpu.render(height)))
    numpy.testing.assert_array_equal(gpu.context.rng_export(), cpu.states)


def test_literal_kernel_handles_a_non_default_camera(torch):
    """Any origin / basis / lens radius goes through the literal kernel and still matches
    the oracle bit for bit."""

    from reinfocus_b200 import _lib

    n, height, spp = 2, 33, 6
    world = oracle.pack_world([6.0, 8.0])
    statics = oracle.CameraStatics(look_from=(0.3, -0.2, 0.5), look_at=(0.1, 0.4, -9.0),
                                   up=(0.1, 1.0, 0.05), aperture=0.23, vfov=28)
    cam = oracle.pack_cameras([6.5, 7.5], statics)
    ctx = _lib.Context()
    ctx.set_world(world)
    ctx.set_cameras(cam, statics.look_from, statics.u, statics.v, float(statics.half_aperture))
    frames = torch.empty((n, height, height, 3), dtype=torch.uint8, device="cuda")
    ctx.render(n, height, height, spp, frames.data_ptr(), None)
    assert ctx.last_trace_kernel() == 0
    states = oracle.rng_states(n * height * height, 0)
    want = oracle.render_fast(world, cam, height, spp, states, oracle.PROFILE_GPU,
                              statics.look_from, statics.u, statics.v,
                              float(statics.half_aperture))
    numpy.testing.assert_array_equal(frames.cpu().numpy(), want)
    numpy.testing.assert_array_equal(ctx.rng_export(), states)


# ------------------------------------------------------------------------ tracer (a4-a11)


CASES = [
    # targets, focus planes, height, spp
    ([7.5], [7.5], 16, 2),
    ([7.5, 7.5], [7.5, 5.0], 32, 8),
    ([5.0, 10.0, 6.3], [10.0, 5.0, 6.3], 25, 4),  # 1875 pixels: exercises ragged tails
    ([8.125], [7.9], 8, 100),
    ([6.0, 9.5, 5.25, 7.0, 8.0], [9.0, 9.5, 5.0, 7.25, 5.5], 75, 7),
]


@pytest.mark.parametrize("targets,planes,height,spp", CASES)
def test_frames_match_oracle_bit_exactly(torch, targets, planes, height, spp):
    gpu = _renderer(samples_per_pixel=spp)
    cpu = oracle.OracleFastRenderer(samples_per_pixel=spp, profile=oracle.PROFILE_GPU)
    for renderer in (gpu, cpu):
        renderer.update_targets(targets)
        renderer.update_focus_planes(planes)
    for _ in range(2):  # second call: persisted, advanced RNG states
        got = gpu.render(height)
        want = cpu.render(height)
        assert got.dtype == numpy.uint8 and got.shape == (len(targets), height, height, 3)
        numpy.testing.assert_array_equal(got, want)
    numpy.testing.assert_array_equal(gpu.context.rng_export(), cpu.states)


def test_full_size_frame_matches_oracle(torch):
    """BASELINE config 1: one env, 300 x 300, 100 samples per pixel."""

    gpu = _renderer()
    cpu = oracle.OracleFastRenderer(profile=oracle.PROFILE_GPU)
    for renderer in (gpu, cpu):
        renderer.update_targets([7.5])
        renderer.update_focus_planes([7.0])
    got, want = gpu.render(300), cpu.render(300)
    mismatch = int((got != want).sum())
    assert mismatch == 0, f"{mismatch} of {got.size} bytes differ"
    numpy.testing.assert_array_equal(gpu.context.rng_export(), cpu.states)


@pytest.mark.parametrize("height,spp,n", [(1, 9, 2), (2, 5, 3), (3, 4, 1), (5, 3, 4), (1200, 2, 1), (601, 1, 2)])
def test_extreme_frame_sizes_match_oracle(torch, height, spp, n):
    """Tiny frames (1 .. 5 pixels a side: every block is ragged) and the largest sweep size."""

    targets = [5.0, 10.0, 7.25, 8.5][:n]
    planes = [10.0, 5.0, 7.25, 6.0][:n]
    gpu = _renderer(samples_per_pixel=spp)
    cpu = oracle.OracleFastRenderer(samples_per_pixel=spp, profile=oracle.PROFILE_GPU)
    for renderer in (gpu, cpu):
        renderer.update_targets(targets)
        renderer.update_focus_planes(planes)
    got, want = gpu.render(height), cpu.render(height)
    assert int((got != want).sum()) == 0
    numpy.testing.assert_array_equal(gpu.step_focus(targets, planes, height),
                                     oracle.focus_values(cpu.render(height)))


def test_empty_batches(torch):
    from reinfocus_b200 import vision

    assert vision.focus_values(numpy.zeros((0, 8, 8, 3), dtype=numpy.uint8)) == []
    renderer = _renderer()
    renderer.update_targets([])
    renderer.update_focus_planes([])
    with pytest.raises(AssertionError):
        renderer.render(8)


def test_rng_state_cache_semantics(torch):
    """reference render.py:248-257: states persist, and are re-created from seed 0 only when
    a call needs more than exist; a smaller later call reuses the larger cache."""

    gpu = _renderer(samples_per_pixel=3)
    cpu = oracle.OracleFastRenderer(samples_per_pixel=3, profile=oracle.PROFILE_GPU)
    for renderer in (gpu, cpu):
        renderer.update_targets([5.5, 9.0])
        renderer.update_focus_planes([9.5, 6.0])
    for height in (8, 12, 8):
        numpy.testing.assert_array_equal(gpu.render(height), cpu.render(height))
        assert gpu.context.rng_count() == len(cpu.states)
    # partial-reset style sub-batch: fewer envs use the states of batch positions 0..k-1
    for renderer in (gpu, cpu):
        renderer.update_targets([9.25])
        renderer.update_focus_planes([5.75])
    numpy.testing.assert_array_equal(gpu.render(12), cpu.render(12))


def test_render_before_update_raises_assertion(torch):
    renderer = _renderer()
    with pytest.raises(AssertionError):  # reference device_data.py:43
        renderer.render(8)
    renderer.update_targets([7.0])
    with pytest.raises(AssertionError):
        renderer.render(8)


def test_gray_output_is_cv2_gray_of_rgb(torch):
    a = _renderer(samples_per_pixel=6)
    b = _renderer(samples_per_pixel=6)
    for renderer in (a, b):
        renderer.update_targets([6.0, 8.5, 9.9])
        renderer.update_focus_planes([6.5, 5.0, 9.9])
    rgb = a.render(41)  # 3 * 41 * 41 = 5043 pixels: not a multiple of 4
    gray = b.render_gray_device(41).cpu().numpy()
    numpy.testing.assert_array_equal(gray, oracle.gray(rgb))


@pytest.mark.parametrize("height", [8, 44, 48, 100])
def test_gray_only_render_equals_oracle(torch, height):
    """Gray-only renders (the step path's output) at several sizes: the pixels, and the states
    they leave behind, are the oracle's."""

    spp = 7
    gpu = _renderer(samples_per_pixel=spp)
    cpu = oracle.OracleFastRenderer(samples_per_pixel=spp, profile=oracle.PROFILE_GPU)
    for renderer in (gpu, cpu):
        renderer.update_targets([6.5, 9.0, 5.0])
        renderer.update_focus_planes([6.0, 9.0, 10.0])
    for _ in range(2):
        gray = gpu.render_gray_device(height).cpu().numpy()
        numpy.testing.assert_array_equal(gray, oracle.gray(cpu.render(height)))
    numpy.testing.assert_array_equal(gpu.context.rng_export(), cpu.states)


def test_batch_split_property_at_64_envs(torch):
    """Size-independent property: rendering a batch equals rendering each env alone with
    that env's slice of the RNG states (env e owns states [e*H*W, (e+1)*H*W))."""

    height, spp, n = 48, 5, 64
    rng = numpy.random.Generator(numpy.random.PCG64DXSM(1234))
    targets = rng.uniform(5, 10, n).astype(numpy.float32)
    planes = rng.uniform(5, 10, n).astype(numpy.float32)
    whole = _renderer(samples_per_pixel=spp)
    whole.update_targets(targets)
    whole.update_focus_planes(planes)
    frames = whole.render(height)
    all_states = oracle.rng_states(n * height * height, 0, doubling=True)
    single = _renderer(samples_per_pixel=spp)
    for e in (0, 1, 17, 63):
        single.update_targets(targets[e:e + 1])
        single.update_focus_planes(planes[e:e + 1])
        single.context.rng_ensure(height * height)
        single.context.rng_import(all_states[e * height * height:(e + 1) * height * height])
        numpy.testing.assert_array_equal(single.render(height)[0], frames[e])


def test_full_benchmark_batch_matches_oracle_on_sampled_envs(torch):
    """BASELINE config at full size - 4096 envs x 300x300 x 100 spp, one step - checked where
    the oracle can follow in seconds: for sampled envs (first, last, a restart-sized prefix
    boundary, random ones) the focus value and the RNG states left behind equal the oracle's
    step of that env alone, started from that env's slice of the initial states."""

    n, height, spp = 4096, 300, 100
    pixels = height * height
    rng = numpy.random.Generator(numpy.random.PCG64DXSM(1234))
    targets = rng.uniform(5, 10, n).astype(numpy.float32)
    planes = rng.uniform(5, 10, n).astype(numpy.float32)
    renderer = _renderer(samples_per_pixel=spp)
    ctx = renderer.context
    ctx.rng_ensure(n * pixels)
    sampled = [0, 1, 512, 513, 2047, 3333, 4095]
    before = {e: ctx.rng_export(e * pixels, pixels) for e in sampled}
    focus = renderer.step_focus(targets, planes, height)
    assert ctx.last_trace_kernel() == 8 and ctx.last_focus_kernel() == 1
    assert focus.shape == (n,) and numpy.isfinite(focus).all()
    world, cam = oracle.pack_world(targets), oracle.pack_cameras(planes)
    for e in sampled:
        states = before[e].copy()
        want = oracle.step(world[e:e + 1], cam[e:e + 1], height, spp, states)
        assert focus[e] == want[0], (e, focus[e], want[0])
        numpy.testing.assert_array_equal(ctx.rng_export(e * pixels, pixels), states)
    del renderer


def test_pixel_index_beyond_2_31(torch):
    """24 000 envs x 300x300 = 2.16e9 pixels (34.6 GB of RNG states): the pixel index is
    64-bit as in the reference (render.py:217). Envs on both sides of the 2^31 boundary, and
    the last one, against the oracle from their own slices of the initial states."""

    n, height, spp = 24000, 300, 1
    pixels = height * height
    assert n * pixels > 2**31
    rng = numpy.random.Generator(numpy.random.PCG64DXSM(99))
    targets = rng.uniform(5, 10, n).astype(numpy.float32)
    planes = rng.uniform(5, 10, n).astype(numpy.float32)
    renderer = _renderer(samples_per_pixel=spp)
    ctx = renderer.context
    ctx.rng_ensure(n * pixels)
    boundary = 2**31 // pixels  # the env whose pixels straddle index 2^31
    sampled = [0, boundary - 1, boundary, boundary + 1, n - 1]
    before = {e: ctx.rng_export(e * pixels, pixels) for e in sampled}
    # the state at index 2^31 + 5 is what the jump chain says it is
    reference_states = oracle.rng_states(4096, 0)
    numpy.testing.assert_array_equal(ctx.rng_export(0, 4096), reference_states)
    focus = renderer.step_focus(targets, planes, height)
    world, cam = oracle.pack_world(targets), oracle.pack_cameras(planes)
    for e in sampled:
        states = before[e].copy()
        want = oracle.step(world[e:e + 1], cam[e:e + 1], height, spp, states)
        assert focus[e] == want[0], (e, focus[e], want[0])
        numpy.testing.assert_array_equal(ctx.rng_export(e * pixels, pixels), states)
    del renderer
    torch.cuda.empty_cache()


# ------------------------------------------------------------------- focus measure (a12)


@pytest.mark.parametrize("shape", [(1, 1), (1, 7), (7, 1), (2, 2), (3, 3), (4, 9), (5, 5), (16, 16),
                                   (31, 17), (64, 48), (75, 75), (300, 300), (257, 301), (600, 600)])
def test_focus_matches_oracle_and_cv2(torch, shape):
    from reinfocus_b200 import vision

    rng = numpy.random.default_rng(shape[0] * 1000 + shape[1])
    n = 3
    base = numpy.linspace(0, 200, shape[1])[None, None, :, None]
    imgs = (base + rng.integers(0, 56, size=(n,) + shape + (3,))).astype(numpy.uint8)
    got = numpy.array(vision.focus_values(imgs))
    want = oracle.focus_values(imgs)
    numpy.testing.assert_array_equal(got, want)  # same exact integer sums, same division
    gray = oracle.gray(imgs)
    got_gray = vision.focus_values_device(torch.from_numpy(gray).cuda()).cpu().numpy()
    numpy.testing.assert_array_equal(got_gray, want)
    try:
        import cv2
    except ImportError:
        return
    ref = [cv2.Laplacian(cv2.medianBlur(cv2.cvtColor(i, cv2.COLOR_RGB2GRAY), 3), cv2.CV_8U).var()
           for i in imgs]
    numpy.testing.assert_allclose(got, ref, rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("shape", [(2, 8), (5, 8), (33, 12), (7, 56), (9, 60), (10, 64), (40, 116), (40, 120),
                                   (40, 124), (40, 128), (12, 180), (37, 240), (64, 244), (31, 364),
                                   (300, 300), (75, 1200)])
@pytest.mark.parametrize("channels", [1, 3])
def test_packed_focus_kernel_equals_staged_kernel_and_oracle(torch, shape, channels):
    """The warp-marching u16x2 kernel (W % 4 == 0) against the staged general kernel (A/B on
    the device) and the oracle, gray and RGB input, widths around the 120-column tile and its
    60-column half (a last segment that narrow is run for two envs per warp; three envs leave
    the last upper half-warp without one)."""

    from reinfocus_b200 import _lib

    rng = numpy.random.default_rng(shape[0] * 7919 + shape[1] + channels)
    n = 3
    base = numpy.linspace(0, 180, shape[1])[None, None, :, None] * numpy.linspace(0.3, 1, shape[0])[None, :, None, None]
    imgs = (base + rng.integers(0, 76, size=(n,) + shape + (3,))).astype(numpy.uint8)
    imgs[1, ::3] = 255  # saturating rows: exercises both clamps of the Laplacian
    imgs[2, :, ::5] = 0
    data = imgs if channels == 3 else oracle.gray(imgs)
    want = oracle.focus_values(imgs) if channels == 3 else oracle.focus_values_gray(data)
    dev = torch.from_numpy(data).cuda()
    out = torch.empty(n, dtype=torch.float64, device="cuda")
    packed, staged = _lib.Context(), _lib.Context()
    staged.set_option(_lib.OPT_FORCE_GENERIC, 1)
    for context, kernel in ((packed, 1), (staged, 0)):
        for _ in range(2):  # twice: the accumulators must be re-armed by the last tile
            out.zero_()
            context.focus(n, shape[0], shape[1], dev.data_ptr(), channels, out.data_ptr())
            numpy.testing.assert_array_equal(out.cpu().numpy(), want)
        assert context.last_focus_kernel() == kernel


def test_packed_focus_kernel_random_shapes(torch):
    """60 random (envs, height, width, channels) cases around every tiling rule of the packed
    kernel: 1 to 4 column segments, last segment wider or narrower than 60 columns (shared by
    env pairs or not), odd and even batches, bands of 1 to H rows."""

    from reinfocus_b200 import _lib

    rng = numpy.random.default_rng(2024)
    ctx = _lib.Context()
    for case in range(60):
        envs = int(rng.integers(1, 10))
        height = int(rng.integers(2, 90))
        width = 4 * int(rng.integers(2, 110))
        channels = (1, 3)[case % 2]
        imgs = rng.integers(0, 256, size=(envs, height, width, 3), dtype=numpy.uint8)
        if case % 3 == 0:  # piecewise-constant content: medians keep the edges, Laplacian saturates
            imgs = numpy.repeat(numpy.repeat(imgs[:, ::3, ::4], 3, axis=1), 4, axis=2)[:, :height, :width]
            imgs = numpy.ascontiguousarray(imgs)
        data = imgs if channels == 3 else oracle.gray(imgs)
        want = oracle.focus_values(imgs) if channels == 3 else oracle.focus_values_gray(data)
        dev = torch.from_numpy(data).cuda()
        out = torch.zeros(envs, dtype=torch.float64, device="cuda")
        ctx.focus(envs, height, width, dev.data_ptr(), channels, out.data_ptr())
        assert ctx.last_focus_kernel() == 1
        numpy.testing.assert_array_equal(out.cpu().numpy(), want, err_msg=f"{envs} x {height} x {width} x {channels}")


def test_packed_focus_kernel_on_a_large_batch(torch):
    """4096-env-like launch shape at reduced height: many envs, grid.y > tiles."""

    from reinfocus_b200 import vision

    rng = numpy.random.default_rng(11)
    gray = rng.integers(0, 256, size=(701, 36, 300), dtype=numpy.uint8)
    for envs in (700, 701):  # even / odd: the 60-column segment pairs envs per warp
        got = vision.focus_values_device(torch.from_numpy(gray[:envs]).cuda()).cpu().numpy()
        numpy.testing.assert_array_equal(got, oracle.focus_values_gray(gray[:envs]))


def test_packed_focus_kernel_tall_saturated_frames_in_a_large_batch(torch):
    """Tall, high-contrast frames in a batch large enough that the launcher would give one
    warp the whole column band: the tile sums of the squared Laplacian (120 x 255^2 per
    row) pass 2^32 beyond 550 rows, so the band height is capped (kPackedMaxBand)."""

    from reinfocus_b200 import vision

    rng = numpy.random.default_rng(5)
    height, width, distinct, copies = 2400, 120, 3, 1600
    blocks = rng.integers(0, 2, size=(distinct, height // 3, width // 3), dtype=numpy.uint8) * 255
    gray = numpy.repeat(numpy.repeat(blocks, 3, axis=1), 3, axis=2)  # 3 x 3 blocks survive the median
    want = oracle.focus_values_gray(gray)
    assert (want > 8000).all()  # mostly saturated Laplacian
    batch = torch.from_numpy(gray).cuda().repeat(copies, 1, 1)
    got = vision.focus_values_device(batch).cpu().numpy()
    numpy.testing.assert_array_equal(got, numpy.tile(want, copies))


def test_focus_planes_match_oracle(ctx, torch):
    rng = numpy.random.default_rng(3)
    gray = rng.integers(0, 256, size=(4, 37, 53), dtype=numpy.uint8)
    d_gray = torch.from_numpy(gray).cuda()
    out = torch.empty(4, dtype=torch.float64, device="cuda")
    med = torch.empty_like(d_gray)
    lap = torch.empty_like(d_gray)
    ctx.focus(4, 37, 53, d_gray.data_ptr(), 1, out.data_ptr(), med.data_ptr(), lap.data_ptr())
    want, wmed, wlap = oracle.focus_values_gray(gray, planes=True)
    numpy.testing.assert_array_equal(med.cpu().numpy(), wmed)
    numpy.testing.assert_array_equal(lap.cpu().numpy(), wlap)
    numpy.testing.assert_array_equal(out.cpu().numpy(), want)


def test_focus_reference_vision_tests(torch):
    """reference tests/vision_test.py:14-34."""

    from reinfocus_b200 import vision

    assert vision.focus_value(numpy.zeros((5, 5, 3), dtype=numpy.uint8)) == 0.0
    assert vision.focus_value(numpy.ones((5, 5, 3), dtype=numpy.uint8)) == 0.0
    checker = numpy.zeros((10, 10, 3), dtype=numpy.uint8)
    checker[::2, ::2] = 255
    checker[1::2, 1::2] = 255
    assert vision.focus_value(checker) == pytest.approx(16230.240000000005, rel=1e-14)


def test_focus_golden_cv2(torch):
    from reinfocus_b200 import vision

    gold = numpy.load(os.path.join(GOLDEN, "focus_cv2.npz"))
    for name in gold["small_names"]:
        got = vision.focus_value(gold[f"img_{name}"])
        assert got == pytest.approx(float(gold[f"fv_{name}"]), rel=1e-13, abs=1e-13), name
    numpy.testing.assert_allclose(vision.focus_values(gold["batch_imgs"]), gold["batch_fv"], rtol=1e-13)


# ----------------------------------------------------------------------- fused step (a13)


def test_step_focus_equals_render_then_focus(torch):
    targets = [5.5, 7.5, 9.0, 6.25]
    planes = [5.75, 7.5, 6.0, 9.5]
    fused = _renderer(samples_per_pixel=9)
    cpu = oracle.OracleFastRenderer(samples_per_pixel=9, profile=oracle.PROFILE_GPU)
    cpu.update_targets(targets)
    cpu.update_focus_planes(planes)
    for _ in range(2):
        got = fused.step_focus(targets, planes, 60)
        want = oracle.focus_values(cpu.render(60))
        numpy.testing.assert_array_equal(got, want)
    # device-resident variant, same numbers for the next call
    got = fused.focus_values_device(60).cpu().numpy()
    numpy.testing.assert_array_equal(got, oracle.focus_values(cpu.render(60)))


def test_focus_is_monotonic_towards_the_target(torch):
    """reference tests/vision_test.py:40-56."""

    from reinfocus_b200 import vision

    renderer = _renderer()
    renderer.update_targets([10] * 5)
    renderer.update_focus_planes([40, 20, 10, 5, 1])
    fv = vision.focus_values(renderer.render(100))
    assert fv[0] < fv[1] < fv[2] > fv[3] > fv[4]


# --------------------------------------------- golden vectors from the reference on a B200


def _gpu_golden():
    return sorted(glob.glob(os.path.join(GOLDEN, "gpu_render_*.npz")))


@pytest.mark.parametrize("path", _gpu_golden(), ids=lambda p: os.path.basename(p)[11:-4])
def test_frames_match_reference_numba_cuda_golden(torch, path):
    from reinfocus_b200 import vision

    gold = numpy.load(path)
    renderer = _renderer(samples_per_pixel=int(gold["spp"]), r_size=float(gold["r_size"]))
    for i in range(int(gold["n_calls"])):
        renderer.update_targets(gold[f"targets_{i}"])
        renderer.update_focus_planes(gold[f"planes_{i}"])
        frames = renderer.render(int(gold[f"height_{i}"]))
        assert renderer.context.rng_count() == int(gold[f"n_states_{i}"])
        if f"frames_{i}" in gold:
            want = gold[f"frames_{i}"]
            mismatch = int((frames != want).sum())
            assert mismatch == 0, f"call {i}: {mismatch} of {want.size} bytes differ"
        else:
            assert hashlib.sha256(frames.tobytes()).hexdigest() == str(gold[f"sha256_{i}"])
        numpy.testing.assert_allclose(vision.focus_values(frames), gold[f"focus_{i}"], rtol=1e-13)
        head = renderer.context.rng_export(0, 64)
        numpy.testing.assert_array_equal(
            numpy.stack([head["s0"], head["s1"]], axis=1), gold[f"states_head_{i}"])


# ----------------------------------------- env sequences vs the reference on a B200 (a13, b)


def _env_gold(name):
    path = os.path.join(GOLDEN, f"gpu_env_{name}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated yet (oracle/gen_golden_env.py gpu)")
    return numpy.load(path)


def _compare_rollout(got, gold):
    """Terminations, truncations, observations and rewards identical to the reference's run.
    (The focus value under every observation is the exactly rounded variance here and numpy's
    pairwise float64 `var()` in the reference: the two sit within 2 ulp of each other in
    float64 and the observers cast to float32, so a difference can survive only when the
    float64 value lies within ~1e-16 relative of a float32 rounding boundary - about one
    observation in 1e8. The goldens hold a few thousand: every one must match.)"""

    for key in ("term", "trunc", "obs0", "obs", "rew"):
        assert got[key].dtype == gold[key].dtype and got[key].shape == gold[key].shape, key
        mismatches = int((got[key] != gold[key]).sum())
        assert mismatches == 0, f"{key}: {mismatches} of {gold[key].size} values differ from the reference"


def test_vector_discrete_steps_sequences_match_reference(torch):
    """DiscreteSteps-v0 vector env, 8 envs, 50 steps with auto-resets: observations, rewards,
    terminations and truncations against the reference's numba-CUDA run on a B200."""

    from examples import custom_environments
    from oracle import gen_golden_env

    gold = _env_gold("vector_discrete_steps")
    env = custom_environments.VectorDiscreteSteps(max_episode_steps=20, num_envs=8)
    focus = env._observer._observers[0]._observers[1].single_observation_space
    numpy.testing.assert_allclose(focus.low, gold["obs_low"], rtol=1e-6)
    numpy.testing.assert_allclose(focus.high, gold["obs_high"], rtol=1e-6)
    gen_golden_env._seed_initializer(env, 77)
    got = gen_golden_env._rollout(env, gold["actions"], True)
    assert gold["trunc"].any()
    _compare_rollout(got, gold)


def test_discrete_steps_sequences_match_reference(torch):
    from examples import custom_environments
    from oracle import gen_golden_env

    gold = _env_gold("discrete_steps")
    env = custom_environments.DiscreteSteps()
    gen_golden_env._seed_initializer(env, 78)
    got = gen_golden_env._rollout(env, gold["actions"], False)
    _compare_rollout(got, gold)


@pytest.mark.parametrize("kind", ["vector", "single"])
def test_sequences_with_interleaved_visualizer_renders_match_reference(torch, kind):
    """env.render() between steps goes through HistoryVisualizer.visualize ->
    renderer.render(600) on the renderer the observer shares (reference
    episode_visualizer.py:197): it re-creates the RNG states at 600 x 600 (render.py:256-257)
    and advances them, so every later observation depends on it. Sequences and the rendered
    scenes against the reference's numba-CUDA run on a B200 (one env per env object: the
    reference's visualizer cannot take a restart of only some envs, see the generator)."""

    from examples import custom_environments
    from oracle import gen_golden_env

    gold = _env_gold(f"{kind}_with_renders")
    if kind == "vector":
        env = custom_environments.VectorDiscreteSteps(max_episode_steps=7, num_envs=1, render_mode="rgb_array")
        gen_golden_env._seed_initializer(env, 79)
        got = gen_golden_env.rollout_with_renders(env, gold["actions"], True)
        assert gold["trunc"].any()
    else:
        env = custom_environments.DiscreteSteps(render_mode="rgb_array")
        gen_golden_env._seed_initializer(env, 80)
        got = gen_golden_env.rollout_with_renders(env, gold["actions"], False, render_every=3, render_phase=0)
    assert len(gold["render_at"]) >= 5
    numpy.testing.assert_array_equal(got["render_at"], gold["render_at"])
    numpy.testing.assert_array_equal(got["render_shape"], gold["render_shape"])
    assert got["render_sha256"].tolist() == gold["render_sha256"].tolist()
    _compare_rollout(got, gold)


def test_reference_env_layer_on_the_c_abi_binding(torch):
    """INTEGRATION.md section B executed: the reference's own FocusObserver / VectorEnvironment /
    VectorDiscreteSteps (baseline/_ref, unmodified) with only graphics.render and vision bound
    to the C-ABI (examples/reference_binding), replaying the reference's B200 sequence."""

    import json
    import subprocess
    import sys

    if not os.path.isdir(os.path.join(REPO, "baseline", "_ref", "reinfocus")):
        pytest.skip("baseline/_ref (copy of the reference) is not present")
    _env_gold("vector_discrete_steps")
    done = subprocess.run([sys.executable, os.path.join(REPO, "scripts", "run_reference_binding.py")],
                          capture_output=True, text=True, timeout=600, cwd=REPO)
    assert done.returncode == 0, done.stderr[-2000:]
    result = json.loads([line for line in done.stdout.splitlines() if line.startswith("{")][-1])
    assert result["env_class"].endswith("VectorDiscreteSteps")
    assert result["observer_module"].startswith(os.path.join("baseline", "_ref"))
    for key in ("obs0", "obs", "rew", "term", "trunc"):
        assert result[f"{key}_mismatches"] == 0, result


def test_registry_makes_the_example_envs(torch):
    import examples  # noqa: F401
    from reinfocus_b200 import gym_compat

    env = gym_compat.make_vec("DiscreteSteps-v0", num_envs=3, vectorization_mode="custom")
    obs, _ = env.reset()
    assert obs.shape == (3, 4) and obs.dtype == numpy.float32
    obs, rew, term, trunc, _ = env.step(numpy.array([6, 0, 12]))
    assert obs.shape == (3, 4) and rew.shape == (3,) and not term.any()
    assert numpy.all(numpy.abs(obs) <= 1)


# ------------------------------------------------- general-scene tracer (a14): render.render


def _generic_scene(name):
    from oracle import gen_golden_generic_gpu
    from reinfocus_b200.graphics import camera, rectangle, shape_factory, sphere, vector, world

    env_shapes, cam_kwargs, frame_shape, spp = gen_golden_generic_gpu.scenes(
        shape_factory, camera, (sphere, rectangle, vector))[name]
    cams = []
    for kw in cam_kwargs:
        kw = dict(kw)
        for key in ("look_from", "look_at", "up"):
            if key in kw:
                kw[key] = vector.v3f(*kw[key])
        cams.append(camera.make_gpu_camera(**kw))
    return world.Worlds(*env_shapes), camera.Cameras(*cams), frame_shape, spp


GENERIC_SCENES = ["one_rect", "two_rect", "one_sphere", "two_sphere", "mixed_batch", "ref_test_sphere",
                  "default_size", "three_shapes"]


@pytest.mark.parametrize("name", GENERIC_SCENES)
def test_generic_render_matches_reference_numba_cuda_golden(torch, name):
    """render.render (spheres + rectangles, 50 bounces) against frames the reference's
    device_render produced under numba-CUDA on a B200 (oracle/gen_golden_generic_gpu.py)."""

    from reinfocus_b200.graphics import render

    gold = numpy.load(os.path.join(GOLDEN, f"gpu_generic_{name}.npz"))
    worlds, cameras, frame_shape, spp = _generic_scene(name)
    numpy.testing.assert_array_equal(worlds.device_data()[0], gold["shape_params"])
    numpy.testing.assert_array_equal(cameras.device_data(), gold["cameras"])
    frames = render.render(worlds, cameras, frame_shape=frame_shape, samples_per_pixel=spp)
    assert frames.shape == (len(worlds),) + tuple(frame_shape) + (3,)
    if "frames" in gold:
        mismatch = int((frames != gold["frames"]).sum())
        assert mismatch == 0, f"{mismatch} of {frames.size} bytes differ"
    else:
        numpy.testing.assert_array_equal(frames[:, :16], gold["frames_head"])
        assert hashlib.sha256(frames.tobytes()).hexdigest() == str(gold["frames_sha256"])


def test_generic_render_reference_average_colour_tests(torch):
    """reference tests/graphics/render_test.py:30-80."""

    from reinfocus_b200.graphics import camera, render, shape_factory, world

    # test_device_average_colour: an r_size=30 rectangle fills the 30 degree frame
    frames = render.render(world.Worlds(shape_factory.one_rect(shape_factory.ShapeParameters(r_size=30))),
                           camera.Cameras(camera.make_gpu_camera()), frame_shape=(100, 100),
                           samples_per_pixel=10)
    means = frames.reshape(-1, 3).mean(axis=0)
    assert 0.25 * 255 <= means[0] <= 0.5 * 255 and 0.25 * 255 <= means[1] <= 0.5 * 255
    assert means[2] == 0
    # test_average_colour: an r_size=30 sphere in front of the sky, the reference's own scene,
    # frame shape and sample count (render_test.py:57-80)
    frames = render.render(world.Worlds(shape_factory.one_sphere(shape_factory.ShapeParameters(r_size=30))),
                           camera.Cameras(camera.make_gpu_camera()), frame_shape=(300, 300))
    means = frames.reshape(-1, 3).mean(axis=0) / 255
    assert 0.4 <= means[0] <= 0.6 and 0.4 <= means[1] <= 0.6 and 0.1 <= means[2] <= 0.2, means


def test_generic_render_matches_oracle_on_a_random_scene(torch):
    """CUDA general-scene tracer against the CPU oracle on a scene no golden vector covers:
    three shapes per env, odd frame size, varied cameras."""

    from reinfocus_b200.graphics import camera, render, rectangle, shape_factory, sphere, vector, world

    P = shape_factory.ShapeParameters
    env_shapes = [
        [sphere.sphere(vector.v3f(-1.5, 0.5, -7.0), 1.25, vector.v2f(6, 10)),
         rectangle.rectangle(vector.v2f(-0.5, 2.5), vector.v2f(-2.0, 0.25), -9.0, vector.v2f(5, 3)),
         sphere.sphere(vector.v3f(1.0, -0.75, -4.0), 0.5)],
        shape_factory.two_rect(P(9.0, texture_f=(7, 7)), P(4.5)),
    ]
    cams = [camera.make_gpu_camera(aperture=0.4, focus_distance=6.0, vfov=45),
            camera.make_gpu_camera(look_from=vector.v3f(-0.2, 0.1, 0.3), aspect_ratio=1.25)]
    worlds, cameras = world.Worlds(*env_shapes), camera.Cameras(*cams)
    got = render.render(worlds, cameras, frame_shape=(37, 53), samples_per_pixel=9)
    params, types, sizes = worlds.device_data()
    want = oracle.render_generic(params, types, sizes, cameras.device_data(), (37, 53), 9)
    mismatch = int((got != want).sum())
    assert mismatch == 0, f"{mismatch} of {got.size} bytes differ"
    # the same scene against the reference's own frames (numba-CUDA on a B200)
    gold = numpy.load(os.path.join(GOLDEN, "gpu_generic_three_shapes.npz"))
    numpy.testing.assert_array_equal(got, gold["frames"])


# ------------------------------------------------------ PPO rollout collection (config 5)


def test_ppo_rollout_and_update_run_on_the_vector_env(torch):
    """examples/ppo.py: one 32-step rollout on 6 envs with frame stack + normalisation, then a
    few PPO minibatches; the buffer has the ppo_tuned.yml shapes and the losses are finite."""

    from examples import ppo

    cfg = ppo.PPOConfig.from_yml(os.path.join(REPO, "examples", "ppo_tuned.yml"), "DiscreteSteps-v0")
    assert (cfg.n_steps, cfg.frame_stack, cfg.net_arch if isinstance(cfg.net_arch, tuple) else None) == (
        32, 5, (256, 256))
    history = ppo.train(num_envs=6, rollouts=1, config=cfg, max_minibatches=3, log=lambda e: None)
    entry = history[0]
    assert entry["env_steps"] == 6 * 32
    assert all(numpy.isfinite(entry[k]) for k in ("loss", "pg", "vf", "entropy", "mean_reward"))
    assert 0 < entry["entropy"] <= numpy.log(13) + 1e-6
