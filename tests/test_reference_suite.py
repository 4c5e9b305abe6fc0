"""Drop-in check: the REFERENCE's own unit tests, run against this package with `reinfocus`
aliased to `reinfocus_b200` (scripts/run_reference_tests.py). The reference's test files are
not part of this repo: they are read from /root/reference in the build container or from
the git-ignored baseline/_ref copy on the GPU box, and the tests skip when neither exists."""

import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "scripts"))

import run_reference_tests  # noqa: E402


def _run(gpu: bool):
    root = run_reference_tests.default_tests_root()
    if root is None:
        pytest.skip("reference tests not available here")
    cmd = [sys.executable, os.path.join(REPO, "scripts", "run_reference_tests.py"), "--tests-root", root]
    if gpu:
        cmd.append("--gpu")
    result = subprocess.run(cmd, capture_output=True, text=True, cwd=REPO)
    assert result.returncode == 0, result.stdout[-2000:] + result.stderr[-4000:]
    return result.stdout


def test_reference_host_side_tests_pass_against_this_package():
    out = _run(gpu=False)
    ran = int(out.split("ran ")[1].split(",")[0])
    assert ran >= 100, out


@pytest.mark.gpu
def test_reference_gpu_tests_pass_against_this_package():
    """Adds the reference's vision tests and FocusObserverTest (real renders on the GPU)."""

    out = _run(gpu=True)
    ran = int(out.split("ran ")[1].split(",")[0])
    assert ran >= 110, out
