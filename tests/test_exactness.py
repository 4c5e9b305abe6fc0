"""CPU proof-by-exhaustion of the float32 forms the CUDA tracer uses in place of the
reference's float64-typed sub-expressions (csrc/rf_tracer.cuh add_sky and the aperture
offset). tests/exhaustive/exactness_sweep.c walks the float32 input domain; the unit test
runs it with a stride (every 97th float, a few seconds), REINFOCUS_FULL_SWEEP=1 runs every
float (about a minute). The GPU-only shortcuts (hoisted-reciprocal division, branch-free
1/length, checker table) have device-side sweeps in tests/test_gpu_parity.py."""

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def test_float32_forms_equal_the_float64_typed_reference_expressions(tmp_path):
    source = os.path.join(HERE, "exhaustive", "exactness_sweep.c")
    binary = str(tmp_path / "exactness_sweep")
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.run([gcc, "-O2", "-fopenmp", "-ffp-contract=off", "-o", binary, source, "-lm"],
                   check=True)
    stride = "1" if os.environ.get("REINFOCUS_FULL_SWEEP") == "1" else "97"
    result = subprocess.run([binary, stride], capture_output=True, text=True)
    lines = dict((line.split()[0], line.split()[1:]) for line in result.stdout.splitlines())
    assert result.returncode == 0, result.stdout
    assert int(lines["sky"][0]) > 1e7 and [int(v) for v in lines["sky"][1:]] == [0, 0, 0, 0]
    assert int(lines["lens"][0]) > 1e6 and int(lines["lens"][1]) == 0
    # the float32 split of float64(0.05) hard-coded in the kernel
    assert lines["lens"][2:] == ["0x1.99999ap-5", "-0x1.99999ap-31"]
