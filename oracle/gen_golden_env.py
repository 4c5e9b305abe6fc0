"""TEST INFRASTRUCTURE ONLY - golden sequences for the env layer (strategy objects, auto
reset, observation normalisation), from the UNMODIFIED reference classes.

Two modes:
  python oracle/gen_golden_env.py sim   (build container, no GPU)
      The reference's env / strategy classes with an ANALYTIC stand-in for FocusObserver
      (the real one needs hours per frame under CUDASIM): pins everything around the hot
      path - transformer, enders, rewarders, Delta/Normalized observers, same-step reset.
      -> tests/golden/env_sim_*.npz
  python oracle/gen_golden_env.py gpu   (GPU box, needs baseline/_ref)
      The reference's own DiscreteSteps / VectorDiscreteSteps (examples/custom_environments)
      with the real numba-CUDA renderer at full size (300 x 300, 100 spp).
      -> gpurun_out/golden_gpu/gpu_env_*.npz  (copy to tests/golden/ and commit)

The reference's RangedInitializer is unseeded (state_initializer.py:50); its private
generator is replaced by a seeded PCG64DXSM so the runs can be reproduced.
"""

import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def _seed_initializer(env, seed):
    import numpy

    env._initializer._generator = numpy.random.Generator(numpy.random.PCG64DXSM(seed))


def _rollout(env, actions, vector):
    import numpy

    obs, _ = env.reset()
    out = {"obs0": numpy.asarray(obs), "actions": actions}
    rows = {"obs": [], "rew": [], "term": [], "trunc": []}
    for action in actions:
        o, r, te, tr, _ = env.step(action if vector else action)
        rows["obs"].append(numpy.array(o, copy=True))
        rows["rew"].append(numpy.array(r, copy=True))
        rows["term"].append(numpy.array(te, copy=True))
        rows["trunc"].append(numpy.array(tr, copy=True))
        if not vector and (te or tr):
            o, _ = env.reset()
            rows["obs"][-1] = numpy.array(o, copy=True)  # observation after the manual reset
    out.update({k: numpy.stack(v) for k, v in rows.items()})
    return out


# ------------------------------------------------------------------------------ sim mode


def analytic_focus(states, target_index=0, plane_index=1):
    """A smooth, peaked stand-in for the focus value (float64, like vision.focus_values)."""

    import numpy

    gap = states[:, target_index].astype(numpy.float64) - states[:, plane_index].astype(numpy.float64)
    return 50.0 + 400.0 / (1.0 + gap * gap)


def sim_cases():
    """name -> builder(modules) returning (env, actions, vector). Shared with the tests,
    which build the same compositions from reinfocus_b200.environments."""

    import numpy

    ends = (5.0, 10.0)
    moves = 5.0 / 2.0 ** numpy.arange(6)
    move_set = numpy.concatenate([-moves, [0], moves[::-1]])

    def observer(m, n, focus_cls):
        return m.state_observer.NormalizedObserver(
            m.state_observer.DeltaObserver(
                [m.state_observer.IndexedElementObserver(n, 1, *ends), focus_cls(n)],
                True, numpy.array([5.0, numpy.nan])))

    def discrete_vector(m, focus_cls, n=16, steps=90, seed=11):
        ender = m.episode_ender.TimeLimitEnder(n, 20) | m.episode_ender.DivergingEnder(
            n, (0, 1), 0.125, early_end_steps=3)
        env = m.vector_environment.VectorEnvironment(
            ender=ender,
            initializer=m.state_initializer.RangedInitializer([[ends]] * 2),
            observer=observer(m, n, focus_cls),
            rewarder=m.episode_rewarder.DeltaRewarder(1, 0.5)
            + m.episode_rewarder.ObservationRewarder(1)
            + m.episode_rewarder.OnTargetRewarder((0, 1), 0.25),
            transformer=m.state_transformer.DiscreteMoveTransformer(n, 1, ends, move_set),
            visualizer=None, num_envs=n)
        actions = numpy.random.Generator(numpy.random.PCG64(seed)).integers(0, 13, (steps, n))
        return env, actions, True

    def discrete_single(m, focus_cls, steps=70, seed=12):
        env = m.environment.Environment(
            ender=m.episode_ender.DivergingEnder(1, (0, 1), 0.125, early_end_steps=3),
            initializer=m.state_initializer.RangedInitializer([[ends]] * 2),
            observer=observer(m, 1, focus_cls),
            rewarder=m.episode_rewarder.DeltaRewarder(1, 0.5)
            + m.episode_rewarder.ObservationRewarder(1)
            + m.episode_rewarder.OnTargetRewarder((0, 1), 0.25),
            transformer=m.state_transformer.DiscreteMoveTransformer(1, 1, ends, move_set),
            visualizer=None)
        actions = numpy.random.Generator(numpy.random.PCG64(seed)).integers(0, 13, steps)
        return env, actions, False

    def continuous_jumps(m, focus_cls, steps=70, seed=13):
        env = m.environment.Environment(
            ender=m.episode_ender.DivergingEnder(1, (0, 1), 0.125, early_end_steps=3),
            initializer=m.state_initializer.RangedInitializer([[ends]] * 2),
            observer=observer(m, 1, focus_cls),
            rewarder=m.episode_rewarder.ObservationRewarder(1)
            + m.episode_rewarder.StoppedRewarder(1, 0.125)
            * m.episode_rewarder.OnTargetRewarder((0, 1), 0.25),
            transformer=m.state_transformer.ContinuousJumpTransformer(1, 1, ends, 0.125),
            visualizer=None)
        raw = numpy.random.Generator(numpy.random.PCG64(seed)).uniform(-1, 1, (steps, 1))
        raw[::5] *= 0.01  # some tiny jumps, below the stop threshold
        return env, raw.astype(numpy.float32), False

    def other_strategies(m, focus_cls, n=6, steps=60, seed=14):
        ender = (m.episode_ender.OnTargetEnder(n, (0, 1), 0.4, early_end_steps=2)
                 | m.episode_ender.StoppedEnder(n, 1, 0.05, early_end_steps=3)) & (
            m.episode_ender.EndlessEnder(n) | m.episode_ender.TimeLimitEnder(n, 3))
        env = m.vector_environment.VectorEnvironment(
            ender=ender,
            initializer=m.state_initializer.RangedInitializer([[(5.0, 6.0), (9.0, 10.0)], [ends]]),
            observer=m.state_observer.DeltaObserver(
                [m.state_observer.IndexedElementObserver(n, 0, *ends), focus_cls(n)]),
            rewarder=m.episode_rewarder.DistanceRewarder((0, 1), 5.0, -2.0, 1.0)
            * m.episode_rewarder.OnTargetRewarder((0, 1), 1.0, 0.5, 2.0),
            transformer=m.state_transformer.ContinuousMoveTransformer(n, 1, ends, 2.0, 0.3),
            visualizer=None, num_envs=n)
        raw = numpy.random.Generator(numpy.random.PCG64(seed)).uniform(-1.2, 1.2, (steps, n))
        return env, raw.astype(numpy.float32), True

    def discrete_jump(m, focus_cls, n=4, steps=40, seed=15):
        env = m.vector_environment.VectorEnvironment(
            ender=m.episode_ender.TimeLimitEnder(n, 7),
            initializer=m.state_initializer.RangedInitializer([[ends]] * 2),
            observer=m.state_observer.NormalizedObserver(focus_cls(n)),
            rewarder=m.episode_rewarder.ObservationRewarder(0),
            transformer=m.state_transformer.DiscreteJumpTransformer(
                n, 1, ends, [4.0, 5.5, 7.0, 8.5, 11.0]),
            visualizer=None, num_envs=n)
        actions = numpy.random.Generator(numpy.random.PCG64(seed)).integers(0, 5, (steps, n))
        return env, actions, True

    return {"discrete_vector": discrete_vector, "discrete_single": discrete_single,
            "continuous_jumps": continuous_jumps, "other_strategies": other_strategies,
            "discrete_jump": discrete_jump}


def make_analytic_observer(state_observer):
    class AnalyticFocusObserver(state_observer.BaseObserver):
        def __init__(self, num_envs):
            super().__init__(num_envs, 50.0, 450.0)

        def observe(self, states, indices=None):
            import numpy

            if indices is None:
                indices = numpy.full(self.observation_space.shape[0], True)
            return numpy.reshape(analytic_focus(states), (indices.sum(), 1))

    return AnalyticFocusObserver


class _Modules:
    def __init__(self, package):
        import importlib

        for name in ("environment", "vector_environment", "episode_ender", "episode_rewarder",
                     "state_initializer", "state_observer", "state_transformer"):
            setattr(self, name, importlib.import_module(f"{package}.environments.{name}"))


def run_sim():
    os.environ["NUMBA_ENABLE_CUDASIM"] = "1"
    import numpy

    import oracle.cudasim_shim as shim

    shim.install()
    modules = _Modules("reinfocus")
    focus_cls = make_analytic_observer(modules.state_observer)
    golden = os.path.join(REPO, "tests", "golden")
    for name, builder in sim_cases().items():
        env, actions, vector = builder(modules, focus_cls)
        _seed_initializer(env, 2024)
        numpy.savez_compressed(os.path.join(golden, f"env_sim_{name}.npz"),
                               **_rollout(env, actions, vector))
        print("done", name)


# ------------------------------------------------------------------------------ gpu mode


def run_gpu():
    import numpy

    sys.path.insert(0, os.path.join(REPO, "baseline", "_ref"))
    import oracle.cudasim_shim as shim

    shim.install_gym_stub()
    shim.install_plot_stub()
    from examples import custom_environments as reference_envs  # baseline/_ref/examples

    out_dir = os.path.join(REPO, "gpurun_out", "golden_gpu")
    os.makedirs(out_dir, exist_ok=True)

    env = reference_envs.VectorDiscreteSteps(max_episode_steps=20, num_envs=8)
    _seed_initializer(env, 77)
    actions = numpy.random.Generator(numpy.random.PCG64(5)).integers(0, 13, (50, 8))
    data = _rollout(env, actions, True)
    data["obs_low"] = env._observer._observers[0]._observers[1].single_observation_space.low
    data["obs_high"] = env._observer._observers[0]._observers[1].single_observation_space.high
    numpy.savez_compressed(os.path.join(out_dir, "gpu_env_vector_discrete_steps.npz"), **data)
    print("done vector", data["obs_low"], data["obs_high"])

    env = reference_envs.DiscreteSteps()
    _seed_initializer(env, 78)
    actions = numpy.random.Generator(numpy.random.PCG64(6)).integers(0, 13, 40)
    numpy.savez_compressed(os.path.join(out_dir, "gpu_env_discrete_steps.npz"),
                           **_rollout(env, actions, False))
    print("done single")


def rollout_with_renders(env, actions, vector=True, render_every=4, render_phase=1):
    """Vector-env rollout with env.render() interleaved (HistoryVisualizer.visualize ->
    renderer.render(600) through the SHARED renderer, reference episode_visualizer.py:197):
    the 600-px render re-creates the RNG states (render.py:256-257) and advances them, which
    changes every later observation. Records the sequences plus the sha256 of the rendered
    scene part (left 600 columns) of every frame."""

    import hashlib

    import numpy

    obs, _ = env.reset()
    out = {"obs0": numpy.asarray(obs), "actions": actions}
    rows = {"obs": [], "rew": [], "term": [], "trunc": []}
    shas, shapes, at = [], [], []
    for step, action in enumerate(actions):
        o, r, te, tr, _ = env.step(action)
        rows["obs"].append(numpy.array(o, copy=True))
        rows["rew"].append(numpy.array(r, copy=True))
        rows["term"].append(numpy.array(te, copy=True))
        rows["trunc"].append(numpy.array(tr, copy=True))
        if not vector and (te or tr):
            o, _ = env.reset()
            rows["obs"][-1] = numpy.array(o, copy=True)  # observation after the manual reset
        if step % render_every == render_phase:
            frame = env.render()
            scene = numpy.ascontiguousarray(frame[:, :600])
            shas.append(hashlib.sha256(scene.tobytes()).hexdigest())
            shapes.append(scene.shape)
            at.append(step)
    out.update({k: numpy.stack(v) for k, v in rows.items()})
    out.update({"render_sha256": numpy.array(shas), "render_shape": numpy.array(shapes),
                "render_at": numpy.array(at)})
    return out


def run_gpu_visualizer():
    """The reference's VectorDiscreteSteps(render_mode="rgb_array") with renders between the
    steps, numba-CUDA on the GPU box. matplotlib is not installable offline, so the graph
    panel of HistoryVisualizer is replaced by a blank image; its renderer.render(600) call -
    the part with side effects on the env - is the reference's own."""

    import numpy

    sys.path.insert(0, os.path.join(REPO, "baseline", "_ref"))
    import oracle.cudasim_shim as shim

    shim.install_gym_stub()
    shim.install_plot_stub()
    from examples import custom_environments as reference_envs  # baseline/_ref/examples
    from reinfocus.environments import episode_visualizer

    def blank_panel(self, env_index, frame_height=600):
        return numpy.full((frame_height, frame_height * 4 // 3, 3), 255, dtype=numpy.uint8)

    episode_visualizer.HistoryVisualizer._visualize_single_history = blank_panel

    out_dir = os.path.join(REPO, "gpurun_out", "golden_gpu")
    os.makedirs(out_dir, exist_ok=True)
    # One env per vector env: with render_mode="rgb_array" the reference's
    # HistoryVisualizer.reset appends the focus history without its `indices`
    # (episode_visualizer.py:186), so a restart of only some envs raises in numpy.hstack;
    # restarts of all envs at once (always the case with one env) work.
    env = reference_envs.VectorDiscreteSteps(max_episode_steps=7, num_envs=1, render_mode="rgb_array")
    _seed_initializer(env, 79)
    actions = numpy.random.Generator(numpy.random.PCG64(7)).integers(0, 13, (26, 1))
    data = rollout_with_renders(env, actions, True)
    numpy.savez_compressed(os.path.join(out_dir, "gpu_env_vector_with_renders.npz"), **data)
    print("done vector + visualizer", data["render_shape"].tolist(), int(data["trunc"].sum()), "truncations")

    env = reference_envs.DiscreteSteps(render_mode="rgb_array")
    _seed_initializer(env, 80)
    actions = numpy.random.Generator(numpy.random.PCG64(8)).integers(0, 13, 30)
    data = rollout_with_renders(env, actions, False, render_every=3, render_phase=0)
    numpy.savez_compressed(os.path.join(out_dir, "gpu_env_single_with_renders.npz"), **data)
    print("done single + visualizer", data["render_shape"].tolist(), int(data["term"].sum()), "terminations")


if __name__ == "__main__":
    {"sim": run_sim, "gpu": run_gpu, "gpu_vis": run_gpu_visualizer}[sys.argv[1]]()
