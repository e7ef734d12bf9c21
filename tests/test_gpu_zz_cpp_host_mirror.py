"""GPU: the C++ host mirror (include/mantapress.hpp) driven like the reference's test_0100_psolve.py / test_0110_mgsolve.py from C++:
solvePressure with every preconditioner, a smoke step and the liquid neighbours, each checked against the CPU oracle on the same
inputs (tests/cpp/host_mirror_test.cpp)."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

from test_cpp_host_mirror import ROOT, build_host_mirror_test  # noqa: E402


def test_cpp_host_mirror_parity(tmp_path):
    exe = build_host_mirror_test(str(tmp_path / "host_mirror_test"))
    oracle = os.path.join(ROOT, "oracle", "libmf_oracle_f32.so")
    assert os.path.exists(oracle), "oracle/libmf_oracle_f32.so is missing (make -C oracle oracle)"
    r = subprocess.run([exe, oracle], capture_output=True, text=True, timeout=600)
    print(r.stdout)
    assert r.returncode == 0 and "all checks passed" in r.stdout, r.stdout + r.stderr
