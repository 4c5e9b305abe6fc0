"""Benchmark of the reinfocus hot path on B200 (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--envs 4096] [--impl ours|reference]

One "step" = one FocusObserver.observe over the whole vector env: set targets / focus
planes, ray trace every env's 300 x 300 frame at 100 samples per pixel, reduce each frame
to its focus value. Metric: env-steps/s (BASELINE.json), whole job over all N GPUs, 4096
envs in total (strong scaling: each rank owns 4096/N envs).

`value`   : device-resident loop (scene already in HBM), CUDA-event timed, max over ranks.
`e2e`     : the public API (FastRenderer.step_focus) with host buffers: host->device copy of
            the step's targets / focus planes (pinned, 8 B per env) and device->host read of
            its focus values (8 B per env) every step.
`roofline`: the tracer kernel (dominant) against the FP32 FFMA peak measured live.
`cpu_baseline` / `--impl reference`: the CPU oracle (C restatement of the reference, all
            host cores) on a bounded sample of the same workload.
`numba_cuda_baseline`: the UNMODIFIED reference (baseline/_ref) with its numba-CUDA kernel and
            cv2 focus loop on this same GPU, in this same run, after the timed regions
            (N = 1 only); `cudasim_baseline`: the reference under NUMBA_ENABLE_CUDASIM=1 on
            one host core at a reduced frame. Reported baselines, not targets.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

METRIC = "env-steps/sec (render+focus value) at 4096 envs"
UNIT = "env-steps/s"
HEIGHT = 300
SPP = 100
FLOP_PER_RAY = 120.0
THEORETICAL_FP32_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12


def synthetic_inputs(num_envs, steps, seed=1234):
    """targets / focus planes ~ U[5, 10] float32 from PCG64DXSM(1234)."""

    import numpy

    rng = numpy.random.Generator(numpy.random.PCG64DXSM(seed))
    targets = rng.uniform(5, 10, (steps, num_envs)).astype(numpy.float32)
    planes = rng.uniform(5, 10, (steps, num_envs)).astype(numpy.float32)
    return targets, planes


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                     "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                fields = [f.strip() for f in out.strip().split(",")]
                if len(fields) >= 7:
                    self.samples.append(fields)
            except Exception:  # pylint: disable=broad-except
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=5)

    def summary(self):
        import statistics

        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        smax = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i] == "Active" for s in self.samples)]
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None, "reasons": reasons,
                "samples": len(self.samples)}


def cpu_baseline(steps, warmup, cores_hint=None):
    """The CPU oracle (oracle/, C restatement of the reference's numba kernel + cv2 focus
    measure) on a bounded sample of the workload, all host threads."""

    import numpy

    import oracle

    # torchrun exports OMP_NUM_THREADS=1: size the thread pool from the CPUs this process may
    # use, not from the OpenMP default
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    sample_envs = int(min(512, max(16, 8 * cores)))
    targets, planes = synthetic_inputs(sample_envs, steps + warmup)
    oracle.lib().rfo_set_threads(cores)
    states = oracle.rng_states(sample_envs * HEIGHT * HEIGHT, 0, doubling=True)
    times = []
    for i in range(steps + warmup):
        t0 = time.perf_counter()
        world = oracle.pack_world(targets[i])
        cam = oracle.pack_cameras(planes[i])
        oracle.step(world, cam, HEIGHT, SPP, states, profile=oracle.PROFILE_GPU, threads=cores)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    mean = float(numpy.mean(times))
    return {
        "value": sample_envs / mean, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": f"{sample_envs} envs x {HEIGHT}x{HEIGHT} x {SPP} spp per step "
                  f"({sample_envs * HEIGHT * HEIGHT * SPP:.3g} rays), {steps} timed steps",
        "ms_per_step": mean * 1e3,
    }


def _run_json(cmd, env=None, timeout=600):
    """Runs a helper script and returns the JSON object on its last stdout line."""

    try:
        done = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=REPO)
    except subprocess.TimeoutExpired:
        return {"unavailable": f"timed out after {timeout} s"}
    lines = [line for line in done.stdout.splitlines() if line.startswith("{")]
    if done.returncode != 0 or not lines:
        tail = (done.stderr or done.stdout).strip().splitlines()[-1:] or ["no output"]
        return {"unavailable": f"exit {done.returncode}: {tail[0][:200]}"}
    return json.loads(lines[-1])


def numba_cuda_baseline(gpu_index, envs, steps):
    """The unmodified reference's numba-CUDA path on this GPU, in this run: FastRenderer +
    vision.focus_values driven as FocusObserver.observe does (baseline/run_numba_cuda.py).
    Steady-state steps only; the first call (JIT + sequential RNG-state init) is listed
    separately. `envs` is a bounded sample: the path's cost is linear in the env count."""

    if not os.path.isdir(os.path.join(REPO, "baseline", "_ref", "reinfocus")):
        return {"unavailable": "baseline/_ref (copy of the reference) is not present"}
    env = dict(os.environ)
    env["CUDA_VISIBLE_DEVICES"] = str(gpu_index) if "CUDA_VISIBLE_DEVICES" not in env else env["CUDA_VISIBLE_DEVICES"]
    env.pop("NUMBA_ENABLE_CUDASIM", None)
    with ClockSampler(gpu_index) as clocks:
        result = _run_json([sys.executable, os.path.join("baseline", "run_numba_cuda.py"), "--envs", str(envs),
                            "--steps", str(steps), "--warmup", "1", "--height", str(HEIGHT), "--spp", str(SPP),
                            "--generic-calls", "3"],
                           env=env, timeout=900)
    if "unavailable" in result:
        return result
    return {
        "value": result["env_steps_per_s"], "unit": UNIT, "measured_in_run": True,
        "what": "unmodified reference: numba-CUDA FastRenderer.render + cv2 vision.focus_values, same GPU",
        "sample": f"{envs} envs x {HEIGHT}x{HEIGHT} x {SPP} spp per step, {steps} timed steps after 1 warm-up",
        "ms_per_step": result["mean_s"]["step"] * 1e3,
        "breakdown_ms": {k: v * 1e3 for k, v in result["mean_s"].items()},
        "rays_per_s_render_call": result["rays_per_s_render_call"],
        "first_call_s_incl_jit_and_rng_init": result["first_call_s_incl_jit_and_rng_init"],
        "generic": result.get("generic"),
        "clocks": clocks.summary(),
    }


def generic_leg(calls=5):
    """render.render on its defaults (general scenes: spheres + rectangles, 50 bounces; reference
    graphics/render.py:88-119): shape_factory.mixed(), one env, 300 x 600, 100 spp. `ms_per_call` is the
    public API (host scene in, host frames out), `kernel_ms` the tracer launch alone."""

    import numpy
    import torch

    from reinfocus_b200 import _lib
    from reinfocus_b200.graphics import camera, render, shape_factory, world

    worlds = world.Worlds(shape_factory.mixed())
    cameras = camera.Cameras(camera.make_gpu_camera())
    height, width, spp = 300, 600, 100
    render.render(worlds, cameras)
    torch.cuda.synchronize()
    times = []
    for _ in range(calls):
        t0 = time.perf_counter()
        frames = render.render(worlds, cameras)
        times.append(time.perf_counter() - t0)
    ctx = _lib.shared_context()
    parameters, types, sizes = worlds.device_data()
    if parameters.shape[2] < 7:
        parameters = numpy.pad(parameters, ((0, 0), (0, 0), (0, 7 - parameters.shape[2])))
    out = torch.empty((1, height, width, 3), dtype=torch.uint8, device="cuda")
    kernel = []
    for _ in range(calls):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ctx.render_generic(parameters, types, sizes, cameras.device_data()[:1], height, width, spp,
                           out.data_ptr(), seed=0)
        e1.record()
        torch.cuda.synchronize()
        kernel.append(e0.elapsed_time(e1))
    rays = height * width * spp
    return {"what": "render.render defaults: shape_factory.mixed(), 1 env, 300 x 600, 100 spp, host frames out",
            "ms_per_call": float(numpy.mean(times)) * 1e3, "device_ms": float(numpy.mean(kernel)),
            "rays_per_s": rays / float(numpy.mean(times)),
            "rays_per_s_device": rays / (float(numpy.mean(kernel)) * 1e-3),
            "mean_colour": [float(c) for c in frames.reshape(-1, 3).mean(axis=0)]}


def cudasim_baseline():
    """The reference under NUMBA_ENABLE_CUDASIM=1 (its CPU path) on one host core, reduced
    frame, rate extrapolated to the full frame (baseline/run_cudasim.py)."""

    env = dict(os.environ)
    env["NUMBA_ENABLE_CUDASIM"] = "1"
    result = _run_json([sys.executable, os.path.join("baseline", "run_cudasim.py"), "--envs", "1",
                        "--height", "32", "--spp", "4"], env=env, timeout=600)
    if "unavailable" in result:
        return result
    return {
        "value": result["env_steps_per_s_extrapolated"], "unit": UNIT, "measured_in_run": True,
        "what": "unmodified reference under NUMBA_ENABLE_CUDASIM=1 (one Python thread per CUDA thread)",
        "sample": f"1 env x 32x32 x 4 spp ({result['rays']} rays), extrapolated to {result['extrapolated_to']}",
        "rays_per_s": result["rays_per_s"], "cores": 1, "nproc": result["nproc"],
    }


def run_reference(args):
    """--impl reference: the reference's algorithm on the host cores (the reference itself is
    numba-CUDA only; its CPU statement is the oracle port, kind "port")."""

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = cpu_baseline(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": base["ms_per_step"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32+f64+u64", "data": "synthetic",
        "config": {"workload": f"DiscreteSteps-v0 hot path, {args.envs} envs, {HEIGHT}x{HEIGHT}, "
                               f"{SPP} spp (CPU arm: bounded sample)"},
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def run_ours(args):
    # NCCL_DEBUG=VERSION makes NCCL print its banner on stdout, which must carry one JSON line
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    import numpy
    import torch
    import torch.distributed as dist

    from reinfocus_b200 import parallel
    from reinfocus_b200.graphics import render

    rank, world_size, local_rank = parallel.init_from_env()
    assert world_size == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world_size}"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)

    first, last = parallel.shard_bounds(args.envs, world_size, rank)
    n_local = last - first
    total_steps = args.steps + args.warmup
    targets, planes = synthetic_inputs(args.envs, 2 * total_steps)
    targets, planes = targets[:, first:last], planes[:, first:last]

    renderer = render.FastRenderer(samples_per_pixel=SPP, device=local_rank)
    ctx = renderer.context
    if args.contexts is not None:
        from reinfocus_b200 import _lib as native

        ctx.set_option(native.OPT_TRACE_CONTEXTS, args.contexts)
    info = ctx.device_info()

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(value):
        if world_size == 1:
            return value
        t = torch.tensor([value], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # one-time setup outside any timed region: RNG states for every pixel of the shard
    t_init0 = time.perf_counter()
    ctx.rng_ensure(n_local * HEIGHT * HEIGHT, 0)
    torch.cuda.synchronize()
    rng_init_s = time.perf_counter() - t_init0

    focus_dev = torch.empty((n_local,), dtype=torch.float64, device=device)
    gray_dev = torch.empty((n_local, HEIGHT, HEIGHT), dtype=torch.uint8, device=device)

    # ------------------------------------------------------------- value: device-resident
    renderer.update_targets(targets[0])
    renderer.update_focus_planes(planes[0])
    renderer._sync_scene()
    for _ in range(args.warmup):
        ctx.step_device(n_local, HEIGHT, SPP, focus_dev.data_ptr())
        parallel.gather_observations(focus_dev, args.envs)
    launches_before = ctx.launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with ClockSampler(local_rank) as clocks:
        start.record()
        for _ in range(args.steps):
            ctx.step_device(n_local, HEIGHT, SPP, focus_dev.data_ptr())
            gathered = parallel.gather_observations(focus_dev, args.envs)
        stop.record()
        barrier()
    device_ms = max_over_ranks(start.elapsed_time(stop)) / args.steps
    launches = ctx.launch_count() - launches_before
    checksum = float(gathered.sum().item())

    # ------------------------------------------- per-kernel timing (tracer, focus stencil)
    # (and, for N > 1, the obs gather and the host-side gap between steps: what the scaling
    # run needs to name its limiter - max over ranks of each part)
    trace_ms, focus_ms, gather_ms, step_wall_ms = [], [], [], []
    barrier()
    for _ in range(min(args.steps, 3)):
        e0, e1, e2, e3 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
        t0 = time.perf_counter()
        e0.record()
        ctx.render(n_local, HEIGHT, HEIGHT, SPP, None, gray_dev.data_ptr())
        e1.record()
        ctx.focus(n_local, HEIGHT, HEIGHT, gray_dev.data_ptr(), 1, focus_dev.data_ptr())
        e2.record()
        parallel.gather_observations(focus_dev, args.envs)
        e3.record()
        torch.cuda.synchronize()
        step_wall_ms.append((time.perf_counter() - t0) * 1e3)
        trace_ms.append(e0.elapsed_time(e1))
        focus_ms.append(e1.elapsed_time(e2))
        gather_ms.append(e2.elapsed_time(e3))
    trace_ms = float(numpy.mean(trace_ms))
    focus_ms = float(numpy.mean(focus_ms))
    breakdown = {"trace_ms": max_over_ranks(trace_ms), "focus_ms": max_over_ranks(focus_ms),
                 "gather_ms": max_over_ranks(float(numpy.mean(gather_ms))),
                 "step_wall_ms": max_over_ranks(float(numpy.mean(step_wall_ms))),
                 "what": "max over ranks, one step at a time with a device sync after each: tracer launch, focus "
                         "stencil, obs all-gather (includes waiting for the slowest rank), host wall clock"}

    # --------------------------------------------------- e2e: public API with host buffers
    for i in range(args.warmup):
        renderer.step_focus(targets[total_steps + i], planes[total_steps + i], HEIGHT)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        fv = renderer.step_focus(targets[total_steps + args.warmup + i],
                                 planes[total_steps + args.warmup + i], HEIGHT)
        if world_size > 1:
            parallel.gather_observations(torch.from_numpy(fv).to(device), args.envs)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps

    # ------------- the real vector env (DiscreteSteps-v0): transform, observe (render + focus),
    # reward, end, and the same-step re-render of the envs that were reset
    from examples import custom_environments
    from reinfocus_b200.environments import state_initializer

    env = custom_environments.VectorDiscreteSteps(
        max_episode_steps=20, num_envs=n_local,
        initializer=state_initializer.RangedInitializer([[custom_environments.ENDS]] * 2, seed=1234 + rank))
    action_rng = numpy.random.Generator(numpy.random.PCG64DXSM(4321 + rank))
    env.reset()
    resets = 0
    for _ in range(args.warmup):
        env.step(action_rng.integers(0, 13, n_local))
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, _, terminated, truncated, _ = env.step(action_rng.integers(0, 13, n_local))
        resets += int((terminated | truncated).sum())
    barrier()
    env_loop_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
    del env

    # the same env with the whole step on the device (rf_env_step): actions are device
    # tensors (a policy on the GPU), observations / rewards stay there; one sync at the end
    env = custom_environments.DeviceVectorDiscreteSteps(
        max_episode_steps=20, num_envs=n_local,
        initializer=state_initializer.RangedInitializer([[custom_environments.ENDS]] * 2, seed=1234 + rank))
    action_rng = numpy.random.Generator(numpy.random.PCG64DXSM(4321 + rank))
    device_actions = torch.from_numpy(
        action_rng.integers(0, 13, (args.warmup + args.steps, n_local))).to(device)
    env.reset()
    device_resets = 0
    for i in range(args.warmup):
        env.step(device_actions[i])
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        env.step(device_actions[args.warmup + i])
        device_resets += env.last_resets
    torch.cuda.synchronize()
    barrier()
    device_env_loop_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
    env.close()
    del env

    if rank != 0:
        return

    rays_local = n_local * HEIGHT * HEIGHT * SPP
    fp32_peak, implied_mhz = ctx.measure_fp32_peak()
    achieved_tflops = rays_local * FLOP_PER_RAY / (trace_ms * 1e-3) / 1e12
    peaks = {}
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    traffic = {}
    try:
        with open(os.path.join(REPO, "profiles", "ncu_traffic.json")) as f:
            traffic = json.load(f)
    except OSError:
        pass

    def scaled_traffic(key):
        per_env = traffic.get(key)
        return per_env * n_local if per_env else None

    focus_bytes = n_local * HEIGHT * HEIGHT + 8 * n_local
    base = cpu_baseline(max(1, min(args.steps, 2)), 1) if not args.no_cpu_baseline else None

    line = {
        "metric": METRIC if args.envs == 4096 else METRIC.replace("4096", str(args.envs)), "value": args.envs / (device_ms * 1e-3), "unit": UNIT,
        "n_gpus": world_size, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": device_ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32+f64+u64", "data": "synthetic",
        "scene": "static (`value` re-renders one scene from HBM-resident parameters; `e2e` takes new host "
                 "targets / focus planes every step)",
        "config": {
            "workload": f"DiscreteSteps-v0 hot path (FocusObserver.observe): {args.envs} envs "
                        f"sharded over {world_size} GPU(s), {HEIGHT}x{HEIGHT}, {SPP} spp, "
                        f"render + focus value",
            "envs_per_gpu": n_local, "rays_per_step": args.envs * HEIGHT * HEIGHT * SPP,
            "l2": "inputs larger than L2 (RNG states: 16 B/pixel = "
                  f"{n_local * HEIGHT * HEIGHT * 16 / 1e9:.2f} GB per GPU, read+written every step)",
            "parallelism": f"env-sharded x{world_size}, obs all-gather",
        },
        "rays_per_s": args.envs * HEIGHT * HEIGHT * SPP / (device_ms * 1e-3),
        "e2e": {"value": args.envs / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": args.envs * 2 * 4, "d2h_bytes_per_step": args.envs * 8,
                "bytes_note": "whole job; each rank copies 1/n_gpus of it"},
        "gpu_launches": launches,
        "clocks": clocks.summary(),
        "roofline": {
            "kernel": "rf::trace_mp_kernel" if ctx.last_trace_kernel() > 1 else "rf::trace_kernel",
            "pixels_per_thread": max(ctx.last_trace_kernel(), 1), "bound": "fp32",
            "achieved": achieved_tflops, "peak": fp32_peak, "unit": "TFLOP/s",
            "frac": achieved_tflops / fp32_peak if fp32_peak else None,
            "traffic": scaled_traffic("trace_kernel_bytes_per_env"),
            "traffic_source": traffic.get("note", "static ncu capture (profiles/ncu_traffic.json)") +
                              "; scaled by envs per launch",
            "algorithmic_bytes_per_launch": n_local * HEIGHT * HEIGHT * 33,
            "peak_source": f"FFMA loop measured live on this GPU (implies {implied_mhz:.0f} MHz); "
                           f"theoretical 148 SM x 128 x 2 x 1.965 GHz = {THEORETICAL_FP32_TFLOPS:.1f}",
            "algorithmic_flop_per_ray": FLOP_PER_RAY, "rays_per_launch": rays_local,
            "launch_ms": trace_ms, "rays_per_s": rays_local / (trace_ms * 1e-3),
            # what actually binds this kernel: warp instructions issued per second
            # against 4 schedulers x 1 instruction / clock / SM; instruction count from ncu
            "issue": issue_view(traffic, n_local, trace_ms, info["sm_count"], clocks.summary()),
            # ... and the pipe that binds it: LOP3 / SHF / IADD3 of xoroshiro128+ on the ALU pipe
            # (16 lanes per scheduler: one warp instruction per 2 clocks)
            "alu": alu_view(traffic, n_local, trace_ms, info["sm_count"], clocks.summary()),
            "hbm": {"achieved": n_local * HEIGHT * HEIGHT * 33 / (trace_ms * 1e-3) / 1e9,
                    "peak": hbm_peak, "unit": "GB/s"},
        },
        "roofline_focus": {
            "kernel": "rf::focus_packed_kernel", "bound": "hbm", "achieved": focus_bytes / (focus_ms * 1e-3) / 1e9,
            "peak": hbm_peak, "unit": "GB/s",
            "frac": focus_bytes / (focus_ms * 1e-3) / 1e9 / hbm_peak,
            "traffic": scaled_traffic("focus_kernel_bytes_per_env"),
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650",
            "launch_ms": focus_ms, "algorithmic_bytes_per_launch": focus_bytes,
            # what binds it: VIMNMX3.U16x2 / PRMT / IADD3 of the packed median and Laplacian on
            # the ALU pipe, not memory
            "alu": alu_view(traffic, n_local, focus_ms, info["sm_count"], clocks.summary(),
                            key="focus_kernel_alu_warp_inst_per_env",
                            source="static ncu capture at 4096 envs: pipe_alu active share x active cycles "
                                   "(profiles/r02/focus_4096envs_ncu_summary.txt); peak = SMs x 4 schedulers x "
                                   "sampled SM clock / 2"),
        },
        "env_loop": {
            "what": "VectorDiscreteSteps.step with uniform random actions: transform + render + "
                    "focus + rewards + enders + same-step re-render of reset envs, host API",
            "value": args.envs / (env_loop_ms * 1e-3), "unit": UNIT, "ms_per_step": env_loop_ms,
            "resets_per_step_rank0": resets / args.steps,
        },
        "env_loop_device": {
            "what": "DeviceVectorDiscreteSteps.step (rf_env_step): the same env with transformer, "
                    "enders, observers, rewarders and auto-reset on the GPU, device actions in, "
                    "device observations / rewards out",
            "value": args.envs / (device_env_loop_ms * 1e-3), "unit": UNIT,
            "ms_per_step": device_env_loop_ms,
            "resets_per_step_rank0": device_resets / args.steps,
        },
        "step_breakdown": dict(breakdown, timed_loop_ms_per_step=device_ms,
                               blocks_per_gpu=n_local * ((HEIGHT * HEIGHT + 2047) // 2048),
                               resident_blocks_per_gpu=info["sm_count"] * 3,
                               note="the tracer runs 3 blocks of 2048 pixels x 100 samples per SM; a block lasts "
                                    "trace_ms / waves, and the launch ends with up to one block time of partly "
                                    "idle SMs - a fixed cost that weighs more the fewer envs a GPU owns"),
        "rng_init_s": rng_init_s,
        "device": {"sm_count": info["sm_count"], "cc": list(info["cc"])},
        "checksum": checksum,
    }
    if base is not None:
        line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
    if world_size == 1:
        line["generic"] = generic_leg()
    if world_size == 1 and not args.no_reference_baselines:
        # the reference itself, same run, same GPU (after every timed region of this arm)
        ctx.rng_reset()  # hand the 5.9 GB of RNG states back before the reference allocates its own
        torch.cuda.empty_cache()
        line["numba_cuda_baseline"] = numba_cuda_baseline(local_rank, args.numba_envs, 3)
        if "value" in line["numba_cuda_baseline"]:
            line["vs_numba_cuda"] = line["value"] / line["numba_cuda_baseline"]["value"]
            line["e2e"]["vs_numba_cuda"] = line["e2e"]["value"] / line["numba_cuda_baseline"]["value"]
            numba_generic = line["numba_cuda_baseline"].get("generic")
            if numba_generic:
                line["generic"]["vs_numba_cuda"] = numba_generic["mean_s"] * 1e3 / line["generic"]["ms_per_call"]
        line["cudasim_baseline"] = cudasim_baseline()
    emit(line)
    if world_size > 1:
        dist.destroy_process_group()


def issue_view(traffic, n_local, trace_ms, sm_count, clock_summary):
    per_env = traffic.get("trace_kernel_warp_inst_per_env")
    mhz = clock_summary.get("sm_mhz") or clock_summary.get("sm_max_mhz")
    if not per_env or not mhz:
        return None
    achieved = per_env * n_local / (trace_ms * 1e-3) / 1e9
    peak = sm_count * 4 * mhz * 1e6 / 1e9
    return {"warp_inst_per_launch": per_env * n_local, "achieved": achieved, "peak": peak,
            "unit": "G warp-inst/s", "frac": achieved / peak,
            "source": "static ncu capture: smsp__inst_executed.sum per env at 4096 envs "
                      "(profiles/ncu_traffic.json); peak = SMs x 4 schedulers x sampled SM clock",
            "ncu_pct": traffic.get("trace_kernel_ncu_pct")}


def alu_view(traffic, n_local, trace_ms, sm_count, clock_summary, key="trace_kernel_alu_warp_inst_per_env",
             source=None):
    per_env = traffic.get(key)
    mhz = clock_summary.get("sm_mhz") or clock_summary.get("sm_max_mhz")
    if not per_env or not mhz:
        return None
    achieved = per_env * n_local / (trace_ms * 1e-3) / 1e9
    peak = sm_count * 4 * mhz * 1e6 / 2 / 1e9
    return {"warp_inst_per_launch": per_env * n_local, "achieved": achieved, "peak": peak,
            "unit": "G ALU-pipe warp-inst/s", "frac": achieved / peak,
            "source": source or
                      "static ncu capture: executed LOP3 / SHF / IADD3 / ISETP / FSETP ... per env at 4096 envs "
                      "(profiles/r02/final_4096envs_opcode_mix.txt); peak = SMs x 4 schedulers x sampled SM clock / 2 "
                      "(measured: tools/pipe_microbench.cu, 2.0 clocks per warp instruction)"}


_RESULT_STREAM = None


def claim_stdout():
    """stdout carries exactly one JSON line: everything else that writes to file descriptor 1
    (NCCL's version banner, library chatter) is sent to stderr for the life of the process."""

    global _RESULT_STREAM
    sys.stdout.flush()
    _RESULT_STREAM = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    stream = _RESULT_STREAM or sys.stdout
    stream.write(json.dumps(line) + "\n")
    stream.flush()


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--gpus", type=int, default=1)
    parser.add_argument("--steps", type=int, default=5)
    parser.add_argument("--warmup", type=int, default=3)
    parser.add_argument("--envs", type=int, default=4096)
    parser.add_argument("--impl", choices=["ours", "reference"], default="ours")
    parser.add_argument("--no-cpu-baseline", action="store_true")
    parser.add_argument("--no-reference-baselines", action="store_true",
                        help="skip the numba-CUDA and CUDASIM legs (the unmodified reference, N = 1 only)")
    parser.add_argument("--numba-envs", type=int, default=512,
                        help="env count of the numba-CUDA leg (its sequential RNG init costs 30 us per env-pixel row)")
    parser.add_argument("--contexts", type=int, default=None,
                        help="pixels per thread of the tracer (0, 2..8); default: library default")
    args = parser.parse_args()
    assert args.warmup >= 1
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
