"""GPU tests of the device-resident vector env and of the on-device
scene packing (a1, a2): both must be bit-identical to the host classes, which the CPU suite
pins to the reference's own classes and the other GPU tests pin to the reference's numba-CUDA
run."""

import os

import numpy
import pytest

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(REPO, "tests", "golden")
ENDS = (5.0, 10.0)


@pytest.fixture(scope="module")
def torch():
    import torch

    assert torch.cuda.is_available(), "these tests need a GPU"
    return torch


def _generator(seed):
    return numpy.random.Generator(numpy.random.PCG64DXSM(seed))


def _strategies(num_envs, renderer, seed, frame_height, kind="steps", max_steps=20, observer_kind="example"):
    from reinfocus_b200.environments import (episode_ender, episode_rewarder, state_initializer,
                                             state_observer, state_transformer)

    focus = state_observer.FocusObserver(num_envs, 0, 1, ENDS, renderer, frame_height)
    plane = state_observer.IndexedElementObserver(num_envs, 1, *ENDS)
    target = state_observer.IndexedElementObserver(num_envs, 0, *ENDS)
    if observer_kind == "example":
        observer = state_observer.NormalizedObserver(state_observer.DeltaObserver(
            [plane, focus], True, numpy.array([5.0, numpy.nan])))
    elif observer_kind == "changes":  # three raw changes, neither the values nor a normalisation
        observer = state_observer.DeltaObserver([target, focus, plane])
    elif observer_kind == "levels":  # normalised values without a DeltaObserver
        observer = state_observer.NormalizedObserver([focus, plane])
    elif observer_kind == "nested":
        # wrappers nested three deep and side by side: normalised [plane, change of the
        # normalised focus value, [target, its change]]
        observer = state_observer.NormalizedObserver([
            plane,
            state_observer.DeltaObserver(state_observer.NormalizedObserver(focus)),
            state_observer.DeltaObserver([target], True)])
    else:  # "stacked": a change of changes over a mix of wrappers and base observers
        observer = state_observer.DeltaObserver([
            state_observer.DeltaObserver([plane, focus], True, numpy.array([5.0, numpy.nan])),
            state_observer.NormalizedObserver([target, plane]),
            state_observer.IndexedElementObserver(num_envs, 0, *ENDS)], True)
    ender = episode_ender.DivergingEnder(num_envs, (0, 1), 0.125, early_end_steps=3)
    if max_steps:
        ender = episode_ender.TimeLimitEnder(num_envs, max_steps) | ender
    if kind == "steps":
        moves = 5.0 / 2.0 ** numpy.arange(6)
        transformer = state_transformer.DiscreteMoveTransformer(
            num_envs, 1, ENDS, numpy.concatenate([-moves, [0], moves[::-1]]))
        rewarder = (episode_rewarder.DeltaRewarder(1, 0.5) + episode_rewarder.ObservationRewarder(1)
                    + episode_rewarder.OnTargetRewarder((0, 1), 0.25))
    elif kind == "other":
        # every ender kind and the remaining rewarders in one tree; the rewards stay float32
        ender = (episode_ender.OnTargetEnder(num_envs, (0, 1), 0.4, early_end_steps=2)
                 | episode_ender.StoppedEnder(num_envs, 1, 0.05, early_end_steps=3)) & (
            episode_ender.EndlessEnder(num_envs) | episode_ender.TimeLimitEnder(num_envs, 3))
        transformer = state_transformer.ContinuousMoveTransformer(num_envs, 1, ENDS, 2.0, 0.3)
        rewarder = (episode_rewarder.DistanceRewarder((0, 1), 5.0, -2.0, 1.0)
                    * episode_rewarder.ObservationRewarder(3) + episode_rewarder.DeltaRewarder(0, 0.25, 2.0))
    elif kind == "mixed":
        # float64 only through the product: Stopped * Distance, plus OnTarget with an offset
        ender = episode_ender.StoppedEnder(num_envs, 1, 0.2, early_end_steps=2) | (
            episode_ender.TimeLimitEnder(num_envs, 9)
            & episode_ender.DivergingEnder(num_envs, (1, 0), 0.05, early_end_steps=1))
        moves = 5.0 / 2.0 ** numpy.arange(6)
        transformer = state_transformer.DiscreteMoveTransformer(
            num_envs, 1, ENDS, numpy.concatenate([-moves, [0], moves[::-1]]))
        rewarder = (episode_rewarder.StoppedRewarder(1, 0.2, 3.0)
                    * episode_rewarder.DistanceRewarder((1, 0), 4.0, -1.0, 0.5)
                    + episode_rewarder.OnTargetRewarder((0, 1), 1.0, 0.5, 2.0))
    elif kind in ("moves", "positions"):
        if kind == "moves":
            transformer = state_transformer.ContinuousMoveTransformer(num_envs, 1, ENDS, 2.5, 0.125)
        else:
            transformer = state_transformer.DiscreteJumpTransformer(
                num_envs, 1, (5.5, 9.5), numpy.linspace(5.0, 10.0, 9))
        rewarder = (episode_rewarder.DeltaRewarder(1, 0.5) + episode_rewarder.ObservationRewarder(1)
                    + episode_rewarder.OnTargetRewarder((0, 1), 0.25))
    else:
        transformer = state_transformer.ContinuousJumpTransformer(num_envs, 1, ENDS, 0.125)
        rewarder = (episode_rewarder.ObservationRewarder(1)
                    + episode_rewarder.StoppedRewarder(1, 0.125)
                    * episode_rewarder.OnTargetRewarder((0, 1), 0.25))
    ranges = [[ENDS]] * 2
    if kind == "mixed":  # several ranges per element: Generator.choice before every uniform
        ranges = [[(5.0, 6.0), (9.0, 10.0), (7.0, 7.5)], [(5.0, 7.0), (8.0, 10.0)]]
    return {"ender": ender,
            "initializer": state_initializer.RangedInitializer(ranges, generator=_generator(seed)),
            "observer": observer, "rewarder": rewarder, "transformer": transformer}


def _pair(num_envs, seed, frame_height=300, spp=100, **kwargs):
    """The host env and the device env over the same strategies, each with its own fresh
    renderer (= its own seed-0 RNG state cache) and its own copy of the generator."""

    from reinfocus_b200.environments import device_vector_environment, vector_environment
    from reinfocus_b200.graphics import render

    host = vector_environment.VectorEnvironment(
        **_strategies(num_envs, render.FastRenderer(samples_per_pixel=spp), seed, frame_height, **kwargs),
        visualizer=None, num_envs=num_envs)
    device = device_vector_environment.DeviceVectorEnvironment(
        **_strategies(num_envs, render.FastRenderer(samples_per_pixel=spp), seed, frame_height, **kwargs),
        num_envs=num_envs)
    return host, device


def _host_renderer(observer):
    """The renderer of the FocusObserver somewhere below ``observer``."""

    if hasattr(observer, "_renderer"):
        return observer._renderer
    for child in getattr(observer, "_observers", []):
        found = _host_renderer(child)
        if found is not None:
            return found
    return None


def _assert_same_rollout(host, device, actions):
    want_obs, _ = host.reset()
    got_obs, _ = device.reset()
    assert got_obs.dtype.is_floating_point and got_obs.is_cuda
    numpy.testing.assert_array_equal(got_obs.cpu().numpy(), want_obs)
    resets = 0
    for step, step_actions in enumerate(actions):
        want = host.step(step_actions)
        got = device.step(step_actions)
        for name, w, g in zip(("obs", "rewards", "terminated", "truncated"), want, got):
            g = g.cpu().numpy()
            assert g.dtype == w.dtype, (name, g.dtype, w.dtype)
            numpy.testing.assert_array_equal(g, w, err_msg=f"{name} at step {step}")
        assert device.last_resets == int(want[3].sum())
        resets += device.last_resets
    exported = device.export_state()
    numpy.testing.assert_array_equal(exported["states"], host._state)
    host_generator = host._initializer._generator.bit_generator.state
    assert device.generator_state() == (host_generator["state"]["state"], host_generator["state"]["inc"],
                                        host_generator["has_uint32"], host_generator["uinteger"])
    numpy.testing.assert_array_equal(device._renderer.context.rng_export(0, 4096),
                                     _host_renderer(host._observer).context.rng_export(0, 4096))
    return resets


def test_discrete_device_env_equals_host_env(torch):
    """8 envs x 60 steps of the DiscreteSteps composition at the real frame size: every
    observation, reward and truncation, the states, the initializer's generator and the
    renderer's RNG states end up identical."""

    host, device = _pair(8, seed=31)
    actions = numpy.random.Generator(numpy.random.PCG64(3)).integers(0, 13, (60, 8))
    assert _assert_same_rollout(host, device, actions) > 8


def test_device_env_matches_reference_numba_cuda_golden(torch):
    """The example device env against the sequence the unmodified reference produced on a
    B200 (oracle/gen_golden_env.py gpu): same tolerance as the host env's test, because the
    reference's numpy.var() may sit an ulp off the exactly rounded variance."""

    from examples import custom_environments
    from reinfocus_b200.environments import state_initializer

    path = os.path.join(GOLDEN, "gpu_env_vector_discrete_steps.npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated yet")
    gold = numpy.load(path)
    env = custom_environments.DeviceVectorDiscreteSteps(
        max_episode_steps=20, num_envs=8,
        initializer=state_initializer.RangedInitializer([[ENDS]] * 2, generator=_generator(77)))
    obs0, _ = env.reset()
    numpy.testing.assert_allclose(obs0.cpu().numpy(), gold["obs0"], rtol=0, atol=2.5e-7)
    for step, step_actions in enumerate(gold["actions"]):
        obs, rewards, terminated, truncated, _ = env.step(torch.as_tensor(step_actions, device="cuda"))
        numpy.testing.assert_array_equal(truncated.cpu().numpy(), gold["trunc"][step])
        numpy.testing.assert_array_equal(terminated.cpu().numpy(), gold["term"][step])
        numpy.testing.assert_allclose(obs.cpu().numpy(), gold["obs"][step], rtol=0, atol=2.5e-7)
        numpy.testing.assert_allclose(rewards.cpu().numpy(), gold["rew"][step], rtol=0, atol=2.5e-7)
    assert gold["trunc"].any()


def test_jump_device_env_equals_host_env(torch):
    """ContinuousJumps composition (float32 actions, Stopped * OnTarget rewards), no time
    limit, including jumps below the stop threshold."""

    host, device = _pair(6, seed=32, frame_height=64, spp=10, kind="jumps", max_steps=0)
    raw = numpy.random.Generator(numpy.random.PCG64(4)).uniform(-1, 1, (50, 6, 1))
    raw[::4] *= 0.01
    actions = raw.astype(numpy.float32)
    assert _assert_same_rollout(host, device, actions) > 0


def test_continuous_move_device_env_equals_host_env(torch):
    """ContinuousMoveTransformer: float32 actions beyond [-1, 1] (clipped), moves below the
    stop threshold (ignored) and moves into the limits (clipped)."""

    host, device = _pair(6, seed=38, frame_height=48, spp=8, kind="moves")
    raw = numpy.random.Generator(numpy.random.PCG64(9)).uniform(-1.6, 1.6, (50, 6, 1))
    raw[::3] *= 0.03
    assert _assert_same_rollout(host, device, raw.astype(numpy.float32)) > 0


def test_discrete_jump_device_env_equals_host_env(torch):
    """DiscreteJumpTransformer: positions from a float32 set, some outside the clip limits."""

    host, device = _pair(6, seed=39, frame_height=48, spp=8, kind="positions")
    actions = numpy.random.Generator(numpy.random.PCG64(10)).integers(0, 9, (50, 6))
    assert _assert_same_rollout(host, device, actions) > 0


@pytest.mark.parametrize("observer_kind,reward_column,columns",
                         [("changes", 1, 3), ("levels", 0, 2), ("nested", 1, 4), ("stacked", 1, 14)])
def test_device_env_with_other_observer_layouts(torch, observer_kind, reward_column, columns):
    """Observer layouts beyond the example envs': a bare DeltaObserver over three base
    observers (changes only, not normalised), a NormalizedObserver without a DeltaObserver,
    and wrappers nested inside one another (reference state_observer.py:103-164 allows any
    nesting); the ObservationRewarder reads a focus column of each."""

    from reinfocus_b200.environments import device_vector_environment, episode_rewarder, vector_environment
    from reinfocus_b200.graphics import render

    def build(env_class, **extra):
        parts = _strategies(6, render.FastRenderer(samples_per_pixel=6), 41, 40, observer_kind=observer_kind)
        parts["rewarder"] = (episode_rewarder.ObservationRewarder(reward_column)
                             + episode_rewarder.OnTargetRewarder((0, 1), 0.25))
        return env_class(**parts, num_envs=6, **extra)

    host = build(vector_environment.VectorEnvironment, visualizer=None)
    device = build(device_vector_environment.DeviceVectorEnvironment)
    assert device.reset()[0].shape == (6, columns) == host.reset()[0].shape
    host, device = (build(vector_environment.VectorEnvironment, visualizer=None),
                    build(device_vector_environment.DeviceVectorEnvironment))
    actions = numpy.random.Generator(numpy.random.PCG64(12)).integers(0, 13, (50, 6))
    assert _assert_same_rollout(host, device, actions) > 6


@pytest.mark.parametrize("kind", ["other", "mixed"])
def test_device_env_with_every_ender_and_rewarder_kind(torch, kind):
    """Ender / rewarder trees beyond the example envs: OnTarget, Stopped and Endless enders
    under & and |, Distance / Stopped rewarders under * and +, float32 and float64 results."""

    host, device = _pair(7, seed=40, frame_height=40, spp=6, kind=kind)
    rng = numpy.random.Generator(numpy.random.PCG64(11))
    if kind == "other":
        actions = rng.uniform(-1.2, 1.2, (70, 7, 1)).astype(numpy.float32)
        actions[::3] *= 0.05
    else:
        actions = rng.integers(0, 13, (70, 7))
        actions[::4] = 6  # the zero move: stopped episodes
    assert _assert_same_rollout(host, device, actions) > 7


def test_device_env_with_large_ender_and_rewarder_trees(torch):
    """Strategy trees of 17 ender nodes and 21 rewarder nodes (the programs hold 24)."""

    from reinfocus_b200.environments import (device_vector_environment, episode_ender, episode_rewarder,
                                             vector_environment)
    from reinfocus_b200.graphics import render

    n = 5

    def build(env_class, **extra):
        parts = _strategies(n, render.FastRenderer(samples_per_pixel=5), 43, 36)
        ender = episode_ender.TimeLimitEnder(n, 11)
        for k in range(2):
            ender = (ender | episode_ender.DivergingEnder(n, (0, 1), 0.05 * (k + 1), early_end_steps=2 + k)) & (
                episode_ender.EndlessEnder(n) | episode_ender.OnTargetEnder(n, (0, 1), 0.3 + 0.1 * k, early_end_steps=2)
                | episode_ender.StoppedEnder(n, 1, 0.04 * (k + 1), early_end_steps=2))
        rewarder = episode_rewarder.ObservationRewarder(1)
        for k in range(3):
            rewarder = rewarder + episode_rewarder.DeltaRewarder(1, 0.5 + k) * episode_rewarder.DistanceRewarder(
                (0, 1), 5.0, -1.0, 1.0 + k) + episode_rewarder.OnTargetRewarder((0, 1), 0.25 * (k + 1))
        rewarder = rewarder + episode_rewarder.StoppedRewarder(1, 0.1, 2.0)
        parts["ender"], parts["rewarder"] = ender, rewarder
        return env_class(**parts, num_envs=n, **extra)

    host = build(vector_environment.VectorEnvironment, visualizer=None)
    device = build(device_vector_environment.DeviceVectorEnvironment)
    actions = numpy.random.Generator(numpy.random.PCG64(21)).integers(0, 13, (60, n))
    actions[::5] = 6
    assert _assert_same_rollout(host, device, actions) > n


def test_device_env_with_more_envs_than_one_scan_chunk(torch):
    """2500 envs (three chunks of the ordered restart scan) at a small frame: restarts keep
    drawing their first states in env order from the one generator."""

    host, device = _pair(2500, seed=33, frame_height=12, spp=3)
    actions = numpy.random.Generator(numpy.random.PCG64(5)).integers(0, 13, (25, 2500))
    assert _assert_same_rollout(host, device, actions) > 2500


def test_device_env_with_70000_envs(torch):
    """More envs than one grid row of the old stencil launch (65535) and restart ranks beyond
    16 bits: 70 000 envs at 8x8 pixels, every env restarting at least once."""

    host, device = _pair(70000, seed=36, frame_height=8, spp=2, max_steps=4)
    actions = numpy.random.Generator(numpy.random.PCG64(7)).integers(0, 13, (6, 70000))
    assert _assert_same_rollout(host, device, actions) > 70000


def test_device_env_takes_device_actions_and_rejects_bad_ones(torch):
    # negative indices wrap around like NumPy's, int32 equals int64
    host, twin = _pair(4, seed=37, frame_height=32, spp=4)
    host.reset(), twin.reset()
    for actions in ([-1, -13, 0, 5], [12, -7, -6, 3]):
        want = host.step(numpy.array(actions))
        got = twin.step(torch.tensor(actions, dtype=torch.int32, device="cuda"))
        for w, g in zip(want[:4], got[:4]):
            numpy.testing.assert_array_equal(g.cpu().numpy(), w)

    _, device = _pair(4, seed=34, frame_height=32, spp=4)
    with pytest.raises(AssertionError):
        device.step(numpy.zeros(4, dtype=numpy.int64))  # before reset
    device.reset()
    obs, rewards, terminated, truncated, info = device.step(
        torch.tensor([0, 6, 12, -1], dtype=torch.int32, device="cuda"))
    assert obs.shape == (4, 4) and obs.dtype == torch.float32 and bool((obs.abs() <= 1).all())
    assert rewards.dtype == torch.float64 and truncated.dtype == torch.bool
    assert not bool(terminated.any()) and info == {}
    with pytest.raises(AssertionError, match="action"):
        device.step(numpy.array([0, 1, 13, 2]))
    with pytest.raises(AssertionError):
        device.step(numpy.array([0, 1, 2, 3]))  # must be reset after a failed step
    device.reset()
    device.step(numpy.array([0, 1, 2, 3]))
    with pytest.raises(AssertionError):
        device.step(numpy.array([0, 1, 2]))


def test_scene_packing_on_device_equals_host_packing(torch):
    """rf_set_scene_device against FastWorlds / FastCameras._make_device_data + upload: the
    default camera through FastRenderer, a tilted one through the context."""

    from reinfocus_b200 import _lib
    from reinfocus_b200.graphics import camera, render, vector, world

    rng = numpy.random.Generator(numpy.random.PCG64(6))
    targets = rng.uniform(5, 10, 7).astype(numpy.float32)
    planes = rng.uniform(5, 10, 7).astype(numpy.float32)
    host, device = render.FastRenderer(samples_per_pixel=5), render.FastRenderer(samples_per_pixel=5)
    states = torch.as_tensor(numpy.stack([targets, planes], axis=1), device="cuda")
    for _ in range(2):
        want = host.step_focus(targets, planes, 48)
        got = device.step_focus_device(states[:, 0], states[:, 1], 48)
        numpy.testing.assert_array_equal(got.cpu().numpy(), want)
    # host-side use of the same renderer afterwards uploads its own scene again
    numpy.testing.assert_array_equal(device.step_focus(targets[:3], planes[:3], 48),
                                     host.step_focus(targets[:3], planes[:3], 48))

    cameras = camera.FastCameras(aspect_ratio=1, look_from=vector.v3f(0.3, -0.2, 0.5),
                                 look_at=vector.v3f(0.1, 0.4, -9.0), up=vector.v3f(0.1, 1.0, 0.05),
                                 aperture=0.23, vfov=28)
    worlds = world.FastWorlds(r_size=17)
    cameras.update(planes)
    worlds.update(targets)
    a, b = _lib.Context(), _lib.Context()
    a.set_world(worlds.device_data())
    a.set_cameras(cameras.device_data(), *cameras.statics)
    b.set_scene_device(7, states[:, 0].data_ptr(), states[:, 1].data_ptr(), 2,
                       _lib.ScenePacking(worlds.packing_constant, *cameras.packing_constants))
    frames = [torch.empty((7, 40, 40, 3), dtype=torch.uint8, device="cuda") for _ in range(2)]
    a.render(7, 40, 40, 6, frames[0].data_ptr(), None)
    b.render(7, 40, 40, 6, frames[1].data_ptr(), None)
    assert a.last_trace_kernel() == 0 and b.last_trace_kernel() == 0
    assert int(frames[0].max()) > 0
    assert torch.equal(frames[0], frames[1])


def test_ppo_rollout_on_the_device_env_equals_the_host_rollout(torch):
    """examples/ppo.py with device_env=True: the rollout buffer (observations after frame
    stacking and running normalisation, actions, rewards, dones, advantages) equals the one
    collected through the host env with the same seeds, up to float64 summation order in the
    running moments."""

    from examples import custom_environments, ppo
    from reinfocus_b200.environments import state_initializer

    cfg = ppo.PPOConfig(n_steps=24)
    device = torch.device("cuda", torch.cuda.current_device())
    data = {}
    for name, env_cls, collector_cls in (
            ("host", custom_environments.VectorDiscreteSteps, ppo.RolloutCollector),
            ("device", custom_environments.DeviceVectorDiscreteSteps, ppo.DeviceRolloutCollector)):
        env = env_cls(max_episode_steps=20, num_envs=6,
                      initializer=state_initializer.RangedInitializer([[ENDS]] * 2, seed=9))
        torch.manual_seed(0)
        policy = ppo.ActorCritic(4 * cfg.frame_stack, 13, cfg.net_arch).to(device)
        collector = collector_cls(env, policy, cfg, device)
        torch.manual_seed(1)
        data[name] = {k: torch.as_tensor(v).cpu().numpy() for k, v in collector.collect().items()}
    assert data["host"]["done"].any()
    numpy.testing.assert_array_equal(data["device"]["act"], data["host"]["act"])
    numpy.testing.assert_array_equal(data["device"]["done"], data["host"]["done"])
    for key in ("obs", "rew", "val", "logp", "adv", "ret"):
        assert data["device"][key].shape == data["host"][key].shape
        numpy.testing.assert_allclose(data["device"][key], data["host"][key], rtol=1e-4, atol=1e-4,
                                      err_msg=key)
    history = ppo.train(num_envs=6, rollouts=1, config=cfg, max_minibatches=2, log=lambda e: None,
                        device_env=True)
    assert history[0]["device_env"] and numpy.isfinite(history[0]["loss"])


def test_device_env_checkpoint_and_resume(torch):
    """state_dict() after 12 steps, loaded into a brand-new env (new context, new renderer):
    the next 15 steps are those of the uninterrupted run, bit for bit."""

    _, env = _pair(5, seed=35, frame_height=40, spp=6)
    actions = numpy.random.Generator(numpy.random.PCG64(8)).integers(0, 13, (27, 5))
    env.reset()
    for step_actions in actions[:12]:
        env.step(step_actions)
    checkpoint = env.state_dict()
    want = [[t.cpu().numpy() for t in env.step(step_actions)[:4]] for step_actions in actions[12:]]
    assert any(w[3].any() for w in want), "the continuation should include restarts"

    _, resumed = _pair(5, seed=999, frame_height=40, spp=6)  # different generator on purpose
    resumed.load_state_dict(checkpoint)
    for step_actions, w in zip(actions[12:], want):
        got = [t.cpu().numpy() for t in resumed.step(step_actions)[:4]]
        for g, ww in zip(got, w):
            numpy.testing.assert_array_equal(g, ww)
    assert resumed.generator_state() == env.generator_state()
    for key, value in env.export_state().items():
        numpy.testing.assert_array_equal(resumed.export_state()[key], value, err_msg=key)
