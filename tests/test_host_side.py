"""CPU tests of the host side: parameter packing against the reference's golden device
data, DeviceData memoisation semantics, and the C-ABI surface (load + symbols only; no
compute without a GPU)."""

import ctypes
import glob
import os
import re

import numpy
import pytest

import oracle
from reinfocus_b200 import _lib
from reinfocus_b200.graphics import camera
from reinfocus_b200.graphics import world

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = sorted(glob.glob(os.path.join(REPO, "tests", "golden", "sim_render_*.npz")))


@pytest.mark.parametrize("path", CASES, ids=lambda p: os.path.basename(p)[11:-4])
def test_vectorised_packing_equals_reference_device_data(path):
    gold = numpy.load(path)
    worlds = world.FastWorlds(r_size=float(gold["r_size"]))
    cameras = camera.FastCameras()
    for i in range(int(gold["n_calls"])):
        worlds.update(gold[f"targets_{i}"])
        cameras.update(gold[f"planes_{i}"])
        numpy.testing.assert_array_equal(worlds.device_data(), gold[f"world_{i}"])
        numpy.testing.assert_array_equal(cameras.device_data(), gold[f"cam_dyn_{i}"])
        origin, u, v, lens = cameras.statics
        numpy.testing.assert_array_equal(
            numpy.array([*origin, *u, *v], dtype=numpy.float32), gold[f"cam_static_{i}"])
        assert lens == float(gold[f"lens_{i}"])


def test_vectorised_packing_equals_scalar_restatement_on_random_inputs():
    rng = numpy.random.Generator(numpy.random.PCG64DXSM(1234))
    targets = rng.uniform(5, 10, 4096).astype(numpy.float32)
    planes = rng.uniform(5, 10, 4096).astype(numpy.float32)
    worlds = world.FastWorlds()
    cameras = camera.FastCameras()
    worlds.update(targets)
    cameras.update(planes)
    numpy.testing.assert_array_equal(worlds.device_data(), oracle.pack_world(targets))
    numpy.testing.assert_array_equal(cameras.device_data(), oracle.pack_cameras(planes))


def test_device_data_semantics():
    worlds = world.FastWorlds()
    assert len(worlds) == 0
    with pytest.raises(AssertionError):  # reference device_data.py:43
        worlds.device_data()
    worlds.update([5.0, 6.0, 7.0])
    assert len(worlds) == 3
    version = worlds.version
    first = worlds.device_data()
    worlds.update(numpy.array([5.0, 6.0, 7.0]))  # unchanged: no repack (device_data.py:57-62)
    assert worlds.version == version and worlds.device_data() is first
    worlds.update([5.0, 6.0])  # a shorter batch (partial reset) repacks
    assert len(worlds) == 2 and worlds.version == version + 1


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    assert lib.rf_abi_version() == _lib.ABI_VERSION
    header = open(os.path.join(REPO, "include", "reinfocus_b200.h")).read()
    declared = set(re.findall(r"^(?:int|int64_t|const char \*)\s*\*?\s*(rf_\w+)\s*\(", header, re.M))
    assert len(declared) >= 20
    raw = ctypes.CDLL(_lib.library_path())
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in include/reinfocus_b200.h but not exported"
    assert declared == set(_lib.EXPORTED_SYMBOLS)


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under reinfocus_b200/ may reference it."""

    offenders = []
    for root, _, files in os.walk(os.path.join(REPO, "reinfocus_b200")):
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, name)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", text, re.M) or "rf_oracle" in text:
                    offenders.append(name)
    assert not offenders, offenders


def test_missing_gpu_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    with pytest.raises(_lib.NativeLibraryError):
        _lib.Context()


def test_checker_boundary_mask():
    """The kernel's checkerboard table (csrc/rf_api.cu kCheckerBelowMask): bit k is set iff
    fl64(fl64(32 pi) * k/32) < k * pi, decided here with exact rationals and 80 digits of pi,
    and cross-checked against libm's float64 sin, which is what the oracle uses."""

    import math
    from fractions import Fraction

    pi = Fraction("3.14159265358979323846264338327950288419716939937510582097494459230781640628620899")
    c = 32.0 * math.pi
    mask = 0
    for k in range(1, 33):
        x = c * (k / 32.0)
        below = Fraction(x) < k * pi
        if below:
            mask |= 1 << k
        # sin(x) > 0 iff floor(x / pi) is even: x just below k*pi lies in cell k-1
        cell = k - 1 if below else k
        assert (math.sin(x) > 0) == (cell % 2 == 0), k
    source = open(os.path.join(REPO, "reinfocus_b200", "csrc", "rf_api.cu")).read()
    declared = int(re.search(r"kCheckerBelowMask = (0x[0-9a-f]+)ull", source).group(1), 16)
    assert declared == mask
