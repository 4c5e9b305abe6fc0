"""Executes INTEGRATION.md section B: the reference's OWN env layer (baseline/_ref:
reinfocus/environments/*, graphics/camera.py, world.py, device_data.py and
examples/custom_environments.py, unmodified) with only `reinfocus.graphics.render` and
`reinfocus.vision` bound to libreinfocus_b200.so (examples/reference_binding/), replaying the
golden sequence the reference recorded with numba-CUDA on a B200
(tests/golden/gpu_env_vector_discrete_steps.npz). Prints one JSON line with mismatch counts.

Needs a GPU and baseline/_ref. Test infrastructure (tests/test_gpu_parity.py runs it)."""

import importlib.util
import json
import os
import sys
import types

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.path.join(REPO, "baseline", "_ref")


def main():
    import numpy

    sys.path.insert(0, REPO)
    sys.path.insert(1, REFERENCE)
    from oracle import cudasim_shim, gen_golden_env

    cudasim_shim.install_gym_stub()   # gymnasium 0.29 / matplotlib are not installable offline
    cudasim_shim.install_plot_stub()

    # FastWorlds / FastCameras._make_device_data end in cuda.to_device(host array): section B
    # has them return the host array (the library uploads it)
    from numba import cuda

    cuda.to_device = lambda array, *args, **kwargs: array

    import reinfocus  # the reference package, from baseline/_ref
    assert os.path.realpath(reinfocus.__file__).startswith(os.path.realpath(REFERENCE)), reinfocus.__file__
    import reinfocus.graphics  # noqa: F401

    from examples.reference_binding import render as bound_render
    from examples.reference_binding import vision as bound_vision

    sys.modules["reinfocus.graphics.render"] = bound_render
    reinfocus.graphics.render = bound_render
    sys.modules["reinfocus.vision"] = bound_vision
    reinfocus.vision = bound_vision

    spec = importlib.util.spec_from_file_location(
        "reference_custom_environments", os.path.join(REFERENCE, "examples", "custom_environments.py"))
    reference_envs = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(reference_envs)

    from reinfocus.environments import state_observer, vector_environment
    for module in (state_observer, vector_environment):
        assert os.path.realpath(module.__file__).startswith(os.path.realpath(REFERENCE))
    assert state_observer.render is bound_render and state_observer.vision is bound_vision

    gold = numpy.load(os.path.join(REPO, "tests", "golden", "gpu_env_vector_discrete_steps.npz"))
    env = reference_envs.VectorDiscreteSteps(max_episode_steps=20, num_envs=8)
    assert type(env).__mro__[1] is vector_environment.VectorEnvironment
    gen_golden_env._seed_initializer(env, 77)
    got = gen_golden_env._rollout(env, gold["actions"], True)
    result = {"env_class": f"{type(env).__module__}.{type(env).__name__}",
              "observer_module": state_observer.__file__.replace(REPO + os.sep, ""),
              "steps": int(len(gold["actions"]))}
    for key in ("obs0", "obs", "rew", "term", "trunc"):
        a, b = numpy.asarray(got[key]), gold[key]
        result[f"{key}_mismatches"] = int((a != b).sum()) if a.shape == b.shape else -1
        result[f"{key}_count"] = int(b.size)
    print(json.dumps(result))


if __name__ == "__main__":
    main()
