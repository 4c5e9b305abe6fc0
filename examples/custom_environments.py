"""The example environments of the reference (examples/custom_environments.py), built on the
B200 hot path: every step ray traces each env's scene and scores its sharpness on the GPU.

Common to all three:
* state = [target position, focus plane], initialised uniformly in [5, 10];
* observation = [focus plane, focus value, focus plane change, focus value change],
  normalised to [-1, 1];
* episodes truncate once target and focus plane have drifted apart by more than .125 on
  three steps (plus a time limit for the vector env).
"""

import numpy

from reinfocus_b200.environments import device_vector_environment
from reinfocus_b200.environments import environment
from reinfocus_b200.environments import episode_ender
from reinfocus_b200.environments import episode_rewarder
from reinfocus_b200.environments import episode_visualizer
from reinfocus_b200.environments import state_initializer
from reinfocus_b200.environments import state_observer
from reinfocus_b200.environments import state_transformer
from reinfocus_b200.environments import vector_environment
from reinfocus_b200.graphics import render

ENDS = (5.0, 10.0)
TARGET_RADIUS = 0.25
MAX_MOVE = 5.0

# state layout
TARGET, FOCUS_PLANE = 0, 1
# observation layout: focus plane, focus value, focus plane change, focus value change
FOCUS_VALUE_OBS = 1


def _observer(num_envs: int, renderer):
    return state_observer.NormalizedObserver(
        state_observer.DeltaObserver(
            [
                state_observer.IndexedElementObserver(num_envs, FOCUS_PLANE, *ENDS),
                state_observer.FocusObserver(num_envs, TARGET, FOCUS_PLANE, ENDS, renderer),
            ],
            True,
            numpy.array([MAX_MOVE, numpy.nan]),
        )
    )


def _visualizer(num_envs: int, renderer, ender):
    return episode_visualizer.HistoryVisualizer(
        num_envs, TARGET, FOCUS_PLANE, FOCUS_VALUE_OBS, renderer, ENDS, ender=ender,
        target_radius=TARGET_RADIUS)


def _step_rewarder():
    # -1 per .5 of focus-plane travel + the focus value + 1 when within .25 of the target
    return (episode_rewarder.DeltaRewarder(FOCUS_PLANE, TARGET_RADIUS * 2)
            + episode_rewarder.ObservationRewarder(FOCUS_VALUE_OBS)
            + episode_rewarder.OnTargetRewarder((TARGET, FOCUS_PLANE), TARGET_RADIUS))


def _step_moves():
    # 13 actions: -5, -2.5, ..., -5/32, 0, 5/32, ..., 2.5, 5
    moves = MAX_MOVE / 2.0 ** numpy.arange(6)
    return numpy.concatenate([-moves, [0], moves[::-1]])


def _initializer(initializer):
    return initializer or state_initializer.RangedInitializer([[ENDS]] * 2)


class DiscreteSteps(environment.Environment):
    # pylint: disable=too-few-public-methods
    """13 discrete focus-plane steps (reference examples/custom_environments.py:16-111)."""

    def __init__(self, render_mode: str | None = None, initializer=None):
        renderer = render.FastRenderer()
        ender = episode_ender.DivergingEnder(1, (TARGET, FOCUS_PLANE), TARGET_RADIUS / 2,
                                             early_end_steps=3)
        super().__init__(
            ender=ender,
            initializer=_initializer(initializer),
            observer=_observer(1, renderer),
            rewarder=_step_rewarder(),
            transformer=state_transformer.DiscreteMoveTransformer(1, FOCUS_PLANE, ENDS, _step_moves()),
            visualizer=_visualizer(1, renderer, ender),
            render_mode=render_mode,
        )


class VectorDiscreteSteps(vector_environment.VectorEnvironment):
    # pylint: disable=too-few-public-methods
    """DiscreteSteps vectorised: ``num_envs`` envs rendered by one launch per step, with a
    built-in time limit (reference examples/custom_environments.py:114-241)."""

    def __init__(self, max_episode_steps: int = 20, num_envs: int = 1,
                 render_mode: str | None = None, initializer=None):
        renderer = render.FastRenderer()
        ender = episode_ender.TimeLimitEnder(num_envs, max_episode_steps) | episode_ender.DivergingEnder(
            num_envs, (TARGET, FOCUS_PLANE), TARGET_RADIUS / 2, early_end_steps=3)
        super().__init__(
            ender=ender,
            initializer=_initializer(initializer),
            observer=_observer(num_envs, renderer),
            rewarder=_step_rewarder(),
            transformer=state_transformer.DiscreteMoveTransformer(num_envs, FOCUS_PLANE, ENDS,
                                                                  _step_moves()),
            visualizer=_visualizer(num_envs, renderer, ender),
            num_envs=num_envs,
            render_mode=render_mode,
        )


class DeviceVectorDiscreteSteps(device_vector_environment.DeviceVectorEnvironment):
    # pylint: disable=too-few-public-methods
    """VectorDiscreteSteps with the whole step on the GPU: same strategies, same sequences,
    but states / observations / rewards are torch CUDA tensors and never visit the host
    (no visualizer, hence no render mode)."""

    def __init__(self, max_episode_steps: int = 20, num_envs: int = 1, initializer=None):
        renderer = render.FastRenderer()
        super().__init__(
            ender=episode_ender.TimeLimitEnder(num_envs, max_episode_steps) | episode_ender.DivergingEnder(
                num_envs, (TARGET, FOCUS_PLANE), TARGET_RADIUS / 2, early_end_steps=3),
            initializer=_initializer(initializer),
            observer=_observer(num_envs, renderer),
            rewarder=_step_rewarder(),
            transformer=state_transformer.DiscreteMoveTransformer(num_envs, FOCUS_PLANE, ENDS,
                                                                  _step_moves()),
            num_envs=num_envs,
        )


class ContinuousJumps(environment.Environment):
    # pylint: disable=too-few-public-methods
    """Actions in [-1, 1] jump the focus plane anywhere in [5, 10]; rewarded by the focus value
    plus 1 for staying put on target (reference examples/custom_environments.py:244-340)."""

    def __init__(self, render_mode: str | None = None, initializer=None):
        min_move = TARGET_RADIUS / 2
        renderer = render.FastRenderer()
        ender = episode_ender.DivergingEnder(1, (TARGET, FOCUS_PLANE), min_move, early_end_steps=3)
        super().__init__(
            ender=ender,
            initializer=_initializer(initializer),
            observer=_observer(1, renderer),
            rewarder=episode_rewarder.ObservationRewarder(FOCUS_VALUE_OBS)
            + episode_rewarder.StoppedRewarder(FOCUS_PLANE, min_move)
            * episode_rewarder.OnTargetRewarder((TARGET, FOCUS_PLANE), TARGET_RADIUS),
            transformer=state_transformer.ContinuousJumpTransformer(1, FOCUS_PLANE, ENDS,
                                                                    TARGET_RADIUS / 2.0),
            visualizer=_visualizer(1, renderer, ender),
            render_mode=render_mode,
        )
