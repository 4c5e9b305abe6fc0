"""Registers the example environments (reference examples/__init__.py:6-18) with whichever
registry is active (gymnasium if installed, the built-in stand-in otherwise)."""

from reinfocus_b200 import gym_compat

_MODULE = "examples.custom_environments"
_EPISODE_STEPS = 20

# id -> (single-env class, vector-env class or None)
_ENVIRONMENTS = {
    "DiscreteSteps-v0": ("DiscreteSteps", "VectorDiscreteSteps"),
    "ContinuousJumps-v0": ("ContinuousJumps", None),
}

for _id, (_single, _vector) in _ENVIRONMENTS.items():
    _spec = {"id": _id, "entry_point": f"{_MODULE}:{_single}", "max_episode_steps": _EPISODE_STEPS}
    if _vector is not None:
        _spec["vector_entry_point"] = f"{_MODULE}:{_vector}"
    gym_compat.register(**_spec)
