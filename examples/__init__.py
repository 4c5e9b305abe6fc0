"""Registers the example environments (reference examples/__init__.py:6-18) with whichever
registry is active (gymnasium if installed, the built-in stand-in otherwise)."""

from reinfocus_b200 import gym_compat

gym_compat.register(
    id="DiscreteSteps-v0",
    entry_point="examples.custom_environments:DiscreteSteps",
    vector_entry_point="examples.custom_environments:VectorDiscreteSteps",
    max_episode_steps=20,
)

gym_compat.register(
    id="ContinuousJumps-v0",
    entry_point="examples.custom_environments:ContinuousJumps",
    max_episode_steps=20,
)
