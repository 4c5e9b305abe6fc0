"""Runs the REFERENCE's own unit tests against this package (drop-in check).

`reinfocus` is aliased to `reinfocus_b200` module by module, gymnasium / matplotlib stand-ins
are installed where the real packages are missing, and the reference's test modules are
loaded from its tests directory:

    python scripts/run_reference_tests.py [--tests-root DIR] [--gpu]

--tests-root defaults to /root/reference (build container) or baseline/_ref (GPU box, where
oracle/gen_golden_gpu.py's copy of the reference lives; add its tests/ directory there).
Without --gpu only the host-side suites run (env layer, histories, device_data,
shape_factory); with --gpu also vision and FocusObserver (and the episode visualizer tests
when matplotlib is installed). The
reference's numba device-function tests (camera/physics/rectangle/... via ad-hoc @cuda.jit
kernels) have no counterpart here: those functions are inlined in the CUDA kernels and are
covered by the bit-exact frame comparisons instead."""

import argparse
import importlib
import os
import sys
import unittest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HOST_SUITES = [
    "tests.histories_test",
    "tests.graphics.device_data_test",
    "tests.graphics.shape_factory_test",
    "tests.environments.episode_ender_test",
    "tests.environments.episode_rewarder_test",
    "tests.environments.state_initializer_test",
    "tests.environments.state_transformer_test",
    "tests.environments.environment_test",
    "tests.environments.vector_environment_test",
]
# state_observer_test mixes host-only cases with FocusObserverTest (needs the GPU)
OBSERVER_SUITE = "tests.environments.state_observer_test"
GPU_SUITES = ["tests.vision_test"]
# needs the real matplotlib (colormaps); not installable offline
PLOT_SUITES = ["tests.environments.episode_visualizer_test"]


def install_aliases():
    sys.path.insert(0, REPO)
    from reinfocus_b200 import gym_compat

    if not gym_compat.USING_REAL_GYMNASIUM:
        gym_compat.install_as_gymnasium()
    import oracle.cudasim_shim as shim  # harness helper: matplotlib stand-in only

    shim.install_plot_stub()
    import reinfocus_b200

    alias = {"reinfocus": reinfocus_b200}
    for name in ("histories", "vision", "environments", "graphics"):
        alias[f"reinfocus.{name}"] = importlib.import_module(f"reinfocus_b200.{name}")
    for name in ("environment", "vector_environment", "episode_ender", "episode_rewarder",
                 "episode_visualizer", "state_initializer", "state_observer", "state_transformer", "types"):
        alias[f"reinfocus.environments.{name}"] = importlib.import_module(f"reinfocus_b200.environments.{name}")
    for name in ("render", "camera", "world", "device_data", "random", "vector", "shape", "sphere",
                 "rectangle", "shape_factory"):
        alias[f"reinfocus.graphics.{name}"] = importlib.import_module(f"reinfocus_b200.graphics.{name}")
    sys.modules.update(alias)


def run(tests_root: str, gpu: bool, verbosity: int = 1):
    install_aliases()
    sys.path.insert(0, tests_root)
    suite = unittest.TestSuite()
    loader = unittest.defaultTestLoader
    try:
        import matplotlib.colors

        have_plot = hasattr(matplotlib.colors, "LinearSegmentedColormap")
    except ImportError:
        have_plot = False
    for name in HOST_SUITES + (GPU_SUITES if gpu else []) + (PLOT_SUITES if gpu and have_plot else []):
        suite.addTests(loader.loadTestsFromName(name))
    observer = loader.loadTestsFromName(OBSERVER_SUITE)

    def keep(test):
        return gpu or "FocusObserverTest" not in test.id()

    def flatten(tests):
        for test in tests:
            if isinstance(test, unittest.TestSuite):
                yield from flatten(test)
            else:
                yield test

    suite.addTests(t for t in flatten(observer) if keep(t))
    return unittest.TextTestRunner(verbosity=verbosity, stream=sys.stderr).run(suite)


def default_tests_root():
    for root in ("/root/reference", os.path.join(REPO, "baseline", "_ref")):
        if os.path.isdir(os.path.join(root, "tests", "environments")):
            return root
    return None


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("--tests-root", default=default_tests_root())
    parser.add_argument("--gpu", action="store_true")
    args = parser.parse_args()
    if args.tests_root is None:
        sys.exit("reference tests not found")
    result = run(args.tests_root, args.gpu, verbosity=1)
    print(f"reference tests against reinfocus_b200: ran {result.testsRun}, "
          f"failures {len(result.failures)}, errors {len(result.errors)}, skipped {len(result.skipped)}")
    sys.exit(0 if result.wasSuccessful() else 1)
