"""Vector env whose whole step stays on the GPU.

``VectorEnvironment`` (reference environments/vector_environment.py:75-164) runs its strategy
objects in NumPy on the host and only the render + focus measure on the GPU.
``DeviceVectorEnvironment`` takes the SAME strategy objects, reads their parameters and runs
their arithmetic in two small CUDA kernels around the render (rf_env_step), so states,
observations and rewards live in device memory and a policy on the GPU can act on them
without a host round trip. Sequences are bit-identical to ``VectorEnvironment`` driven by the
same generator (tests/test_gpu_device_env.py).

What is understood (anything else raises ``NotImplementedError`` - use ``VectorEnvironment``):

* state ``[target, focus plane]``;
* any of the four transformers (discrete / continuous, move / jump) on the focus plane;
* any ``&`` / ``|`` tree of ``TimeLimitEnder``, ``DivergingEnder``, ``OnTargetEnder``,
  ``StoppedEnder``, ``EndlessEnder`` (up to 24 nodes);
* observer: any tree of ``DeltaObserver``s and ``NormalizedObserver``s, nested in any order,
  over ``IndexedElementObserver``s and one ``FocusObserver`` (up to 16 nodes and 16 columns);
* any ``+`` / ``*`` tree of ``DeltaRewarder``, ``DistanceRewarder``, ``ObservationRewarder``,
  ``OnTargetRewarder``, ``StoppedRewarder`` (up to 24 nodes), with NumPy's result types;
* initializer ``RangedInitializer`` (up to four ranges per element) on a PCG64DXSM generator.
"""

from typing import Any

import numpy

from reinfocus_b200 import _lib
from reinfocus_b200 import gym_compat
from reinfocus_b200.environments import episode_ender
from reinfocus_b200.environments import episode_rewarder
from reinfocus_b200.environments import state_initializer
from reinfocus_b200.environments import state_observer
from reinfocus_b200.environments import state_transformer

TARGET, FOCUS_PLANE = 0, 1


def _unsupported(what: str):
    raise NotImplementedError(
        f"DeviceVectorEnvironment does not understand this {what}; use VectorEnvironment")


def _f32(value) -> float:
    return float(numpy.float32(value))


def _read_transformer(transformer, config: _lib.EnvConfig):
    # pylint: disable=protected-access
    if transformer._move_index != FOCUS_PLANE:
        _unsupported("transformer (it must move the focus plane)")
    config.limits[0], config.limits[1] = (_f32(limit) for limit in transformer._limits)
    if isinstance(transformer, state_transformer.DiscreteMoveTransformer):
        moves = numpy.asarray(transformer._action_set, dtype=numpy.float64)
        if len(moves) > 32:
            _unsupported("transformer (more than 32 moves)")
        config.transformer = _lib.ENV_DISCRETE_MOVE
        config.n_moves = len(moves)
        for i, move in enumerate(moves):
            config.moves[i] = float(move)
    elif isinstance(transformer, state_transformer.ContinuousJumpTransformer):
        config.transformer = _lib.ENV_CONTINUOUS_JUMP
        config.jump_span = _f32(transformer._limits[1] - transformer._limits[0])
        config.jump_threshold = _f32(transformer._stop_threshold)
    elif isinstance(transformer, state_transformer.ContinuousMoveTransformer):
        config.transformer = _lib.ENV_CONTINUOUS_MOVE
        config.move_speed = _f32(transformer._speed)
        config.jump_threshold = _f32(transformer._stop_threshold)
    elif isinstance(transformer, state_transformer.DiscreteJumpTransformer):
        positions = numpy.asarray(transformer._action_set, dtype=numpy.float32)
        if len(positions) > 32:
            _unsupported("transformer (more than 32 positions)")
        config.transformer = _lib.ENV_DISCRETE_JUMP
        config.n_moves = len(positions)
        for i, position in enumerate(positions):
            config.jumps[i] = float(position)
    else:
        _unsupported("transformer")


def _state_index(index) -> int:
    index = int(index)
    if index not in (TARGET, FOCUS_PLANE):
        _unsupported("strategy (the state is [target, focus plane])")
    return index


def _read_ender(ender, config: _lib.EnvConfig):
    """Flattens the ender tree into the postfix program of rf_env_config."""

    # pylint: disable=protected-access
    program: list[_lib.EnvEnder] = []

    def emit(node):
        if isinstance(node, episode_ender.OpEnder):
            kinds = {numpy.bitwise_and: _lib.ENV_ENDER_AND, numpy.bitwise_or: _lib.ENV_ENDER_OR}
            if node._op not in kinds:
                _unsupported("ender (only & and | combine enders)")
            emit(node._l_ender)
            emit(node._r_ender)
            program.append(_lib.EnvEnder(kinds[node._op], 0, 0, 0, 0.0))
        elif isinstance(node, episode_ender.TimeLimitEnder):
            program.append(_lib.EnvEnder(_lib.ENV_ENDER_TIME_LIMIT, 0, 0, int(node._max_steps), 0.0))
        elif isinstance(node, episode_ender.DivergingEnder):
            a, b = (_state_index(i) for i in node._check_indices)
            program.append(_lib.EnvEnder(_lib.ENV_ENDER_DIVERGING, a, b, int(node._early_end_steps),
                                         _f32(node._threshold)))
        elif isinstance(node, episode_ender.OnTargetEnder):
            a, b = (_state_index(i) for i in node._check_indices)
            program.append(_lib.EnvEnder(_lib.ENV_ENDER_ON_TARGET, a, b, int(node._early_end_steps),
                                         _f32(node._radius)))
        elif isinstance(node, episode_ender.StoppedEnder):
            if node._early_end_steps + 1 > _lib.ENV_MAX_WINDOW:
                _unsupported(f"ender (StoppedEnder windows hold at most {_lib.ENV_MAX_WINDOW} positions)")
            program.append(_lib.EnvEnder(_lib.ENV_ENDER_STOPPED, _state_index(node._check_index), 0,
                                         int(node._early_end_steps), _f32(node._early_end_span)))
        elif isinstance(node, episode_ender.EndlessEnder):
            program.append(_lib.EnvEnder(_lib.ENV_ENDER_ENDLESS, 0, 0, 0, 0.0))
        else:
            _unsupported("ender")

    emit(ender)
    if len(program) > _lib.ENV_MAX_NODES:
        _unsupported(f"ender (more than {_lib.ENV_MAX_NODES} nodes)")
    config.n_enders = len(program)
    for i, node in enumerate(program):
        config.enders[i] = node


def _read_rewarder(rewarder, config: _lib.EnvConfig, columns: int = 4) -> bool:
    """Flattens the rewarder tree into the postfix program of rf_env_config. Returns whether
    NumPy would make the rewards float64 (a float32 tree stays float32)."""

    # pylint: disable=protected-access
    program: list[_lib.EnvReward] = []

    def emit(node) -> bool:
        if isinstance(node, episode_rewarder.OpRewarder):
            kinds = {numpy.add: _lib.ENV_REWARD_ADD, numpy.multiply: _lib.ENV_REWARD_MUL}
            if node._op not in kinds:
                _unsupported("rewarder (only + and * combine rewarders)")
            wide_l = emit(node._l_rewarder)
            wide_r = emit(node._r_rewarder)
            program.append(_lib.EnvReward(kinds[node._op], 0, 0, 0.0, 0.0, 0.0, 0.0))
            return wide_l or wide_r
        if isinstance(node, episode_rewarder.DeltaRewarder):
            program.append(_lib.EnvReward(_lib.ENV_REWARD_DELTA, _state_index(node._check_index), 0,
                                          _f32(node._reward), _f32(node._scale), 0.0, 0.0))
            return False
        if isinstance(node, episode_rewarder.DistanceRewarder):
            a, b = (_state_index(i) for i in node._check_indices)
            program.append(_lib.EnvReward(_lib.ENV_REWARD_DISTANCE, a, b, _f32(node._span),
                                          _f32(node._high - node._low), float(node._low), 0.0))
            return False
        if isinstance(node, episode_rewarder.ObservationRewarder):
            column = int(node._reward_observation_index)
            if not 0 <= column < columns:
                _unsupported("rewarder (observation column out of range)")
            program.append(_lib.EnvReward(_lib.ENV_REWARD_OBSERVATION, column, 0, 0.0, 0.0, 0.0, 0.0))
            return False
        if isinstance(node, episode_rewarder.OnTargetRewarder):
            a, b = (_state_index(i) for i in node._check_indices)
            program.append(_lib.EnvReward(_lib.ENV_REWARD_ON_TARGET, a, b, _f32(node._span), 0.0,
                                          float(node._off), float(node._delta)))
            return True
        if isinstance(node, episode_rewarder.StoppedRewarder):
            program.append(_lib.EnvReward(_lib.ENV_REWARD_STOPPED, _state_index(node._check_index), 0,
                                          _f32(node._threshold), 0.0, float(node._reward), 0.0))
            return True
        _unsupported("rewarder")
        return False

    wide = emit(rewarder)
    if len(program) > _lib.ENV_MAX_NODES:
        _unsupported(f"rewarder (more than {_lib.ENV_MAX_NODES} nodes)")
    config.n_rewards = len(program)
    for i, node in enumerate(program):
        config.rewards[i] = node
    return wide


def _read_observer(observer, config: _lib.EnvConfig):
    """Flattens the observer tree - DeltaObservers and NormalizedObservers, nested in any
    order, over IndexedElementObservers and exactly one FocusObserver - into the postfix
    program of rf_env_config. Returns the FocusObserver's renderer and the number of
    observation columns."""

    # pylint: disable=protected-access
    program: list[_lib.EnvObserver] = []
    focus_observers = []
    normalized_columns = 0
    alive_peak = 0

    def emit(node, alive: int) -> int:
        """Appends `node`'s subtree; `alive`: values to the left that are still on the stack.
        Returns the width of the vector the node leaves."""

        nonlocal normalized_columns, alive_peak
        if isinstance(node, state_observer.FocusObserver):
            if (node._target_index, node._focus_plane_index) != (TARGET, FOCUS_PLANE):
                _unsupported("observer (the FocusObserver must read [target, focus plane])")
            focus_observers.append(node)
            program.append(_lib.EnvObserver(_lib.ENV_OBS_FOCUS, 0, 0, 0))
            alive_peak = max(alive_peak, alive + 1)
            return 1
        if isinstance(node, state_observer.IndexedElementObserver):
            program.append(_lib.EnvObserver(_lib.ENV_OBS_ELEMENT, _state_index(node._element_index), 0, 0))
            alive_peak = max(alive_peak, alive + 1)
            return 1
        if not isinstance(node, (state_observer.DeltaObserver, state_observer.NormalizedObserver)):
            _unsupported("observer")
        children = list(node._observers)
        width = 0
        for child in children:
            width += emit(child, alive + width)
        if isinstance(node, state_observer.DeltaObserver):
            include = bool(node._include_original)
            program.append(_lib.EnvObserver(_lib.ENV_OBS_DELTA, len(children), int(include), 0))
            width *= 2 if include else 1
        else:
            assert node._mid.dtype == numpy.float32 and node._scale.dtype == numpy.float32
            assert len(node._mid) == width == len(node._scale)
            if normalized_columns + width > _lib.ENV_MAX_OBS_VALUES:
                _unsupported(f"observer (more than {_lib.ENV_MAX_OBS_VALUES} normalised columns)")
            for i in range(width):
                config.obs_mid[normalized_columns + i] = float(node._mid[i])
                config.obs_scale[normalized_columns + i] = float(node._scale[i])
            program.append(_lib.EnvObserver(_lib.ENV_OBS_NORMALIZED, len(children), 0, normalized_columns))
            normalized_columns += width
        alive_peak = max(alive_peak, alive + width)
        return width

    columns = emit(observer, 0)
    if len(focus_observers) != 1:
        _unsupported("observer (exactly one FocusObserver renders per step)")
    if len(program) > _lib.ENV_MAX_OBS_NODES:
        _unsupported(f"observer (more than {_lib.ENV_MAX_OBS_NODES} nodes)")
    if columns > _lib.ENV_MAX_OBS_DIM or alive_peak > _lib.ENV_MAX_OBS_VALUES:
        _unsupported(f"observer (more than {_lib.ENV_MAX_OBS_DIM} columns)")
    config.n_observers = len(program)
    for i, node in enumerate(program):
        config.observers[i] = node
    focus = focus_observers[0]
    config.frame_height = int(focus._frame_height)
    return focus._renderer, columns


def _read_initializer(initializer, config: _lib.EnvConfig):
    # pylint: disable=protected-access
    if not (isinstance(initializer, state_initializer.RangedInitializer) and len(initializer._ranges) == 2
            and all(1 <= len(options) <= 4 for options in initializer._ranges)):
        _unsupported("initializer")
    generator = initializer._generator
    if generator.bit_generator.state["bit_generator"] != "PCG64DXSM":
        _unsupported("initializer (its generator must be a PCG64DXSM)")
    for i, options in enumerate(initializer._ranges):
        config.init_options[i] = len(options)
        for k, (low, high) in enumerate(options):
            config.init_low[i][k], config.init_high[i][k] = float(low), float(high)
    return generator


class DeviceVectorEnvironment(gym_compat.VectorEnv):
    # pylint: disable=too-many-instance-attributes
    """``VectorEnvironment`` with the step on the GPU. ``reset`` / ``step`` return torch CUDA
    tensors: observations float32 (n, columns), rewards float64 or float32 (n,), terminated / truncated bool
    (n,). ``step`` takes actions as a torch CUDA tensor (int32 / int64 for discrete moves,
    float32 for jumps) or anything ``numpy.asarray`` understands (copied to the GPU).

    The initializer's NumPy generator is read once, at construction: from then on the env
    owns the stream on the device (``generator_state`` returns where it stands)."""

    metadata = {"render_modes": []}

    def __init__(self, ender, initializer, observer, rewarder, transformer, num_envs: int = 2):
        # pylint: disable=too-many-arguments
        import torch

        super().__init__()
        self.num_envs = num_envs
        self.action_space = transformer.action_space
        self.observation_space = observer.observation_space
        self.single_action_space = transformer.single_action_space
        self.single_observation_space = observer.single_observation_space
        self.render_mode = None

        config = _lib.EnvConfig()
        config.num_envs = num_envs
        _read_transformer(transformer, config)
        _read_ender(ender, config)
        renderer, self._columns = _read_observer(observer, config)
        self._wide_rewards = _read_rewarder(rewarder, config, self._columns)
        generator = _read_initializer(initializer, config)
        config.samples_per_pixel = renderer.samples_per_pixel
        config.packing = renderer.scene_packing()

        self._renderer = renderer
        self._discrete = config.transformer in (_lib.ENV_DISCRETE_MOVE, _lib.ENV_DISCRETE_JUMP)
        self._device = torch.device(f"cuda:{renderer.context.device}")
        self._env = _lib.DeviceEnv(renderer.context, config)
        snapshot = generator.bit_generator.state
        self._env.set_generator(int(snapshot["state"]["state"]), int(snapshot["state"]["inc"]),
                                int(snapshot["has_uint32"]), int(snapshot["uinteger"]))
        self._started = False
        self.last_resets = 0

    # ------------------------------------------------------------------------ gym surface
    def reset(self, *, seed: int | None = None, options: dict[str, Any] | None = None):
        import torch

        super().reset(seed=seed)
        observations = torch.empty((self.num_envs, self._columns), dtype=torch.float32, device=self._device)
        self._renderer.scene_overwritten()
        self._env.reset(observations.data_ptr())
        self._started = True
        return observations, {}

    def step(self, actions):
        import torch

        assert self._started, "reset the env first"
        actions, kind = self._device_actions(actions)
        observations = torch.empty((self.num_envs, self._columns), dtype=torch.float32, device=self._device)
        rewards = torch.empty((self.num_envs,), dtype=torch.float64, device=self._device)
        truncated = torch.empty((self.num_envs,), dtype=torch.bool, device=self._device)
        terminated = torch.zeros((self.num_envs,), dtype=torch.bool, device=self._device)
        self._renderer.scene_overwritten()
        self.last_resets = self._env.step(actions.data_ptr(), kind, observations.data_ptr(),
                                          rewards.data_ptr(), truncated.data_ptr())
        if not self._wide_rewards:
            rewards = rewards.to(torch.float32)  # exact: a float32 tree was evaluated in float32
        return observations, rewards, terminated, truncated, {}

    def close(self, **kwargs):
        self._env.close()

    # ------------------------------------------------------------------------- inspection
    def generator_state(self) -> tuple[int, int, int, int]:
        """(state, increment, has_uint32, uinteger) of the PCG64DXSM stream the restarts draw
        from, as in numpy's ``bit_generator.state``."""

        return self._env.get_generator()

    def export_state(self) -> dict:
        """Host copies of the states and ender counters (tests, checkpoints)."""

        return self._env.export()

    def state_dict(self) -> dict:
        """Everything a resumed run needs to continue bit-identically: the per-env episode
        state, the initializer's generator and the renderer's RNG states (the reference
        cannot checkpoint an env: its RNG states live in a numba device array)."""

        return {"env": self._env.export(), "generator": self._env.get_generator(),
                "rng_states": self._renderer.context.rng_export()}

    def load_state_dict(self, checkpoint: dict):
        self._env.load(checkpoint["env"])
        self._env.set_generator(*checkpoint["generator"])
        context = self._renderer.context
        context.rng_ensure(len(checkpoint["rng_states"]))
        context.rng_import(checkpoint["rng_states"])
        self._renderer.scene_overwritten()
        self._started = True

    # ---------------------------------------------------------------------------- helpers
    def _device_actions(self, actions):
        import torch

        if not isinstance(actions, torch.Tensor):
            actions = numpy.asarray(actions).reshape(-1)
            if self._discrete:
                actions = actions.astype(numpy.int64, copy=False)
            else:
                assert actions.dtype == numpy.float32, (
                    "continuous actions must be float32 (the jump arithmetic takes their dtype)")
            actions = torch.from_numpy(numpy.ascontiguousarray(actions)).to(self._device)
        actions = actions.reshape(-1).contiguous()
        assert actions.is_cuda and actions.shape[0] == self.num_envs, "one action per env"
        if self._discrete:
            kinds = {torch.int32: _lib.ENV_ACTIONS_INT32, torch.int64: _lib.ENV_ACTIONS_INT64}
            assert actions.dtype in kinds, "discrete actions must be int32 or int64"
            return actions, kinds[actions.dtype]
        assert actions.dtype == torch.float32, "continuous actions must be float32"
        return actions, _lib.ENV_ACTIONS_FLOAT32
