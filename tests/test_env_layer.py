"""CPU tests of the env layer (reinfocus_b200.environments) against golden sequences that
the unmodified reference classes produced (oracle/gen_golden_env.py sim): same compositions
of strategy objects, same seeded initial states, same scripted actions. The focus value
comes from an analytic stand-in here (the real FocusObserver needs the GPU; its sequences
are covered by the -m gpu tests), so this pins everything around the hot path: transformer,
enders, rewarders, Delta/Normalized observers and the same-step auto-reset. All comparisons
are exact."""

import os

import functools

import numpy
import pytest

from oracle import gen_golden_env
from reinfocus_b200 import gym_compat, histories

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def modules():
    return gen_golden_env._Modules("reinfocus_b200")


@pytest.mark.parametrize("name", sorted(gen_golden_env.sim_cases()))
def test_env_sequences_equal_the_reference(modules, name):
    gold = numpy.load(os.path.join(GOLDEN, f"env_sim_{name}.npz"))
    focus_cls = gen_golden_env.make_analytic_observer(modules.state_observer)
    env, actions, vector = gen_golden_env.sim_cases()[name](modules, focus_cls)
    gen_golden_env._seed_initializer(env, 2024)
    got = gen_golden_env._rollout(env, actions, vector)
    for key in ("obs0", "obs", "rew", "term", "trunc"):
        assert got[key].dtype == gold[key].dtype, (key, got[key].dtype, gold[key].dtype)
        numpy.testing.assert_array_equal(got[key], gold[key], err_msg=key)
    assert gold["trunc"].any(), "the scenario should exercise episode ends"


def test_spaces_follow_gymnasium_conventions(modules):
    spaces = gym_compat.spaces
    box = spaces.Box(numpy.float64(1.5), numpy.float64(2.5), dtype=numpy.float32)
    assert box.shape == (1,) and box.dtype == numpy.float32  # scalar bounds -> shape (1,)
    batched = gym_compat.batch_space(box, 4)
    assert batched.shape == (4, 1) and batched.low.dtype == numpy.float32
    assert numpy.all(batched.low == 1.5) and numpy.all(batched.high == 2.5)
    discrete = spaces.Discrete(13)
    assert gym_compat.batch_space(discrete, 3).shape == (3,)
    assert discrete.contains(12) and not discrete.contains(13)
    observer = modules.state_observer.IndexedElementObserver(5, 1, 5.0, 10.0)
    assert observer.observation_space.shape == (5, 1)
    assert observer.single_observation_space.shape == (1,)


def test_histories_ring_buffer():
    h = histories.Histories(3, 4)
    assert numpy.isnan(h.data).all()
    h.append_events([1.0, 2.0, 3.0])
    h.append_events([4.0, 5.0], numpy.array([True, False, True]))
    numpy.testing.assert_array_equal(h.most_recent_events(), [4.0, 2.0, 5.0])
    numpy.testing.assert_array_equal(h.get_history(0), [1.0, 4.0])
    numpy.testing.assert_array_equal(h.get_history(1), [2.0])
    h.reset([False, True, False])
    assert len(h.get_history(1)) == 0 and len(h.get_history(2)) == 2
    for value in range(10):
        h.append_events([value, value, value])
    numpy.testing.assert_array_equal(h.get_history(0), [6.0, 7.0, 8.0, 9.0])


def test_registry_builds_example_envs_lazily():
    """examples/__init__.py registers the env ids; building them needs the GPU, so only the
    registration and entry-point resolution are checked here."""

    import examples  # noqa: F401

    if gym_compat.USING_REAL_GYMNASIUM:
        pytest.skip("real gymnasium registry in use")
    spec = gym_compat._registry["DiscreteSteps-v0"]
    assert spec.max_episode_steps == 20
    assert gym_compat._load_entry_point(spec.vector_entry_point).__name__ == "VectorDiscreteSteps"
    assert gym_compat._load_entry_point(spec.entry_point).__name__ == "DiscreteSteps"


def test_vectorised_initializer_draws_the_same_states_as_the_reference_loop(modules):
    """RangedInitializer's fast path (one range per element) against the reference's draw
    order: per env, per element, choice() then uniform() on one PCG64DXSM stream."""

    ranges = [[(5.0, 10.0)], [(5.0, 10.0)], [(-1.0, 2.5)]]
    fast = modules.state_initializer.RangedInitializer(ranges, seed=99)
    reference_order = numpy.random.Generator(numpy.random.PCG64DXSM(99))
    for count in (1, 7, 300):
        want = numpy.array([[reference_order.uniform(*reference_order.choice(r)) for r in ranges]
                            for _ in range(count)], dtype=numpy.float32)
        numpy.testing.assert_array_equal(fast.initialize(count), want)


def test_sb3_adapter_follows_the_vec_env_protocol(modules):
    """SB3Wrapper (reference vector_shim.py:20-186) over the discrete vector composition:
    step_async/step_wait return what VectorEnvironment.step returns with done = terminated
    | truncated, per-env info dicts and the terminal observation row on finished envs."""

    from reinfocus_b200.environments.experimental import vector_shim

    focus_cls = gen_golden_env.make_analytic_observer(modules.state_observer)
    cases = gen_golden_env.sim_cases()
    env, actions, _ = cases["discrete_vector"](modules, focus_cls)
    twin, _, _ = cases["discrete_vector"](modules, focus_cls)
    gen_golden_env._seed_initializer(env, 5)
    gen_golden_env._seed_initializer(twin, 5)
    wrapped = vector_shim.SB3Wrapper(env, None)
    assert wrapped.num_envs == 16 and wrapped.render_mode is None
    assert wrapped.observation_space.shape == (4,) and wrapped.action_space.n == 13
    numpy.testing.assert_array_equal(wrapped.reset(), twin.reset()[0])
    saw_done = False
    for step_actions in actions[:40]:
        obs, rewards, dones, infos = wrapped.step(step_actions)
        want_obs, want_rew, term, trunc, _ = twin.step(step_actions)
        numpy.testing.assert_array_equal(obs, want_obs)
        numpy.testing.assert_array_equal(rewards, want_rew)
        numpy.testing.assert_array_equal(dones, term | trunc)
        assert len(infos) == 16
        for i, info in enumerate(infos):
            assert ("terminal_observation" in info) == bool(dones[i])
            if dones[i]:
                numpy.testing.assert_array_equal(info["terminal_observation"], obs[i])
        saw_done |= bool(dones.any())
    assert saw_done
    assert wrapped.get_attr("num_envs") == [16] * 16
    assert wrapped.get_attr("num_envs", 3) == [16] and wrapped.get_attr("num_envs", [1, 2]) == [16, 16]
    assert wrapped.env_is_wrapped(object) == [False] * 16
    with pytest.raises(NotImplementedError):
        wrapped.get_attr("no_such_attribute")
    with pytest.raises(NotImplementedError):
        wrapped.set_attr("x", 1)
    with pytest.raises(NotImplementedError):
        wrapped.env_method("anything")
    with pytest.raises(NotImplementedError):
        vector_shim.SB3Wrapper(object(), None)
    # without stable-baselines3 there is no DummyVecEnv to swap: the hook is the identity
    if not vector_shim.HAVE_SB3:
        marker = object()
        assert vector_shim.rewrapper(marker) is marker


def test_device_env_reads_strategy_trees_and_refuses_the_rest(modules):
    """DeviceVectorEnvironment flattens the ender / rewarder trees into postfix programs and
    reads the other strategies' parameters; what it cannot express must raise instead of
    being approximated."""

    from reinfocus_b200 import _lib
    from reinfocus_b200.environments import device_vector_environment as dve

    m = modules
    config = _lib.EnvConfig()
    enders = m.episode_ender
    time_limit = enders.TimeLimitEnder(2, 20)
    diverging = enders.DivergingEnder(2, (0, 1), 0.125, early_end_steps=3)
    dve._read_ender(time_limit | diverging, config)
    kinds = [config.enders[i].kind for i in range(config.n_enders)]
    assert kinds == [_lib.ENV_ENDER_TIME_LIMIT, _lib.ENV_ENDER_DIVERGING, _lib.ENV_ENDER_OR]
    assert (config.enders[0].steps, config.enders[1].steps, config.enders[1].value) == (20, 3, 0.125)
    tree = (enders.OnTargetEnder(2, (0, 1), 0.4, early_end_steps=2)
            | enders.StoppedEnder(2, 1, 0.05, early_end_steps=3)) & (enders.EndlessEnder(2) | time_limit)
    dve._read_ender(tree, config)
    kinds = [config.enders[i].kind for i in range(config.n_enders)]
    assert kinds == [_lib.ENV_ENDER_ON_TARGET, _lib.ENV_ENDER_STOPPED, _lib.ENV_ENDER_OR,
                     _lib.ENV_ENDER_ENDLESS, _lib.ENV_ENDER_TIME_LIMIT, _lib.ENV_ENDER_OR, _lib.ENV_ENDER_AND]
    for ender in (enders.DivergingEnder(2, (0, 2), 0.1), enders.StoppedEnder(2, 1, 0.05, early_end_steps=16),
                  enders.OpEnder(time_limit, diverging, numpy.bitwise_xor),
                  functools.reduce(lambda a, b: a | b, [time_limit] * 13)):  # 25 nodes, the program holds 24
        with pytest.raises(NotImplementedError):
            dve._read_ender(ender, config)

    rewarders = m.episode_rewarder
    steps = (rewarders.DeltaRewarder(1, 0.5) + rewarders.ObservationRewarder(1)
             + rewarders.OnTargetRewarder((0, 1), 0.25))
    assert dve._read_rewarder(steps, config) is True  # bool * Python float makes it float64
    kinds = [config.rewards[i].kind for i in range(config.n_rewards)]
    assert kinds == [_lib.ENV_REWARD_DELTA, _lib.ENV_REWARD_OBSERVATION, _lib.ENV_REWARD_ADD,
                     _lib.ENV_REWARD_ON_TARGET, _lib.ENV_REWARD_ADD]
    assert (config.rewards[0].f0, config.rewards[0].f1) == (-1.0, 0.5)
    assert (config.rewards[3].f0, config.rewards[3].d0, config.rewards[3].d1) == (0.25, 0.0, 1.0)
    narrow = rewarders.DistanceRewarder((0, 1), 5.0, -2.0, 1.0) * rewarders.ObservationRewarder(3)
    assert dve._read_rewarder(narrow, config) is False  # float32 all the way
    assert (config.rewards[0].f0, config.rewards[0].f1, config.rewards[0].d0) == (5.0, 3.0, -2.0)
    for rewarder in (rewarders.DeltaRewarder(2, 0.5), rewarders.ObservationRewarder(4),
                     rewarders.OpRewarder(steps, steps, numpy.subtract)):
        with pytest.raises(NotImplementedError):
            dve._read_rewarder(rewarder, config)

    transformers = m.state_transformer
    dve._read_transformer(transformers.DiscreteMoveTransformer(2, 1, (5.0, 10.0), [-1.0, 0.0, 1.0]), config)
    assert (config.transformer, config.n_moves, list(config.moves[:3])) == (0, 3, [-1.0, 0.0, 1.0])
    dve._read_transformer(transformers.ContinuousJumpTransformer(2, 1, (5.0, 10.0), 0.125), config)
    assert (config.transformer, config.jump_span, config.jump_threshold) == (1, 5.0, 0.125)
    dve._read_transformer(transformers.ContinuousMoveTransformer(2, 1, (5.0, 10.0), 1.5, 0.25), config)
    assert (config.transformer, config.move_speed, config.jump_threshold) == (2, 1.5, 0.25)
    dve._read_transformer(transformers.DiscreteJumpTransformer(2, 1, (5.0, 10.0), [5.0, 7.5, 10.0]), config)
    assert (config.transformer, config.n_moves, list(config.jumps[:3])) == (3, 3, [5.0, 7.5, 10.0])
    for transformer in (transformers.DiscreteMoveTransformer(2, 0, (5.0, 10.0), [0.0]),
                        transformers.ContinuousMoveTransformer(2, 0, (5.0, 10.0), 1.0),
                        transformers.DiscreteJumpTransformer(2, 1, (5.0, 10.0), numpy.zeros(33)),
                        transformers.DiscreteMoveTransformer(2, 1, (5.0, 10.0), numpy.zeros(33))):
        with pytest.raises(NotImplementedError):
            dve._read_transformer(transformer, config)

    initializers = m.state_initializer
    generator = dve._read_initializer(initializers.RangedInitializer([[(5.0, 10.0)]] * 2, seed=3), config)
    assert generator.bit_generator.state["bit_generator"] == "PCG64DXSM"
    assert [config.init_low[i][0] for i in range(2)] == [5.0, 5.0]
    assert [config.init_high[i][0] for i in range(2)] == [10.0, 10.0] and list(config.init_options) == [1, 1]
    dve._read_initializer(initializers.RangedInitializer([[(5.0, 6.0), (8.0, 9.0)], [(5.0, 10.0)]], seed=4), config)
    assert list(config.init_options) == [2, 1] and list(config.init_low[0][:2]) == [5.0, 8.0]
    assert list(config.init_high[0][:2]) == [6.0, 9.0]
    for initializer in (initializers.RangedInitializer([[(5.0, 10.0)] * 5, [(5.0, 10.0)]]),
                        initializers.RangedInitializer([[(5.0, 10.0)]] * 3),
                        initializers.FixedInitializer(numpy.zeros((4, 2))),
                        initializers.RangedInitializer(
                            [[(5.0, 10.0)]] * 2, generator=numpy.random.Generator(numpy.random.PCG64(1)))):
        with pytest.raises(NotImplementedError):
            dve._read_initializer(initializer, config)

    focus_cls = gen_golden_env.make_analytic_observer(modules.state_observer)
    env, _, _ = gen_golden_env.sim_cases()["discrete_vector"](modules, focus_cls)
    with pytest.raises(NotImplementedError):  # the focus value must come from FocusObserver
        dve._read_observer(env._observer, config)

    # nested observer wrappers flatten to a postfix program (children before parents)
    observers = m.state_observer
    observers.cached_focus_extrema.cache_clear()

    class NoRenderer:  # pylint: disable=too-few-public-methods
        """Stands in for a FastRenderer: reading the tree never renders."""

    original = observers.cached_focus_extrema
    observers.cached_focus_extrema = lambda ends, height: (50.0, 450.0)
    try:
        focus = observers.FocusObserver(2, 0, 1, (5.0, 10.0), NoRenderer(), 40)
    finally:
        observers.cached_focus_extrema = original
    plane = observers.IndexedElementObserver(2, 1, 5.0, 10.0)
    target = observers.IndexedElementObserver(2, 0, 5.0, 10.0)
    nested = observers.NormalizedObserver([plane, observers.DeltaObserver(observers.NormalizedObserver(focus)),
                                           observers.DeltaObserver([target], True)])
    renderer, columns = dve._read_observer(nested, config)
    assert isinstance(renderer, NoRenderer) and columns == 4 == nested.observation_space.shape[1]
    program = [(o.kind, o.arg, o.flag, o.offset) for o in config.observers[:config.n_observers]]
    assert program == [(_lib.ENV_OBS_ELEMENT, 1, 0, 0), (_lib.ENV_OBS_FOCUS, 0, 0, 0),
                       (_lib.ENV_OBS_NORMALIZED, 1, 0, 0), (_lib.ENV_OBS_DELTA, 1, 0, 0),
                       (_lib.ENV_OBS_ELEMENT, 0, 0, 0), (_lib.ENV_OBS_DELTA, 1, 1, 0),
                       (_lib.ENV_OBS_NORMALIZED, 3, 0, 1)]
    assert list(config.obs_mid[:1]) == [250.0] and list(config.obs_scale[:1]) == [200.0]
    for bad in (observers.NormalizedObserver([plane, target]),                       # no FocusObserver
                observers.DeltaObserver([focus, observers.NormalizedObserver(focus)]),  # two renders per step
                observers.DeltaObserver([observers.DeltaObserver([plane, focus, target, plane, target], True)] * 2,
                                        True)):                                       # 40 columns
        with pytest.raises(NotImplementedError):
            dve._read_observer(bad, config)
