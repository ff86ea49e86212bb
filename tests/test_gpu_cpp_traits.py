"""-m gpu: runs the C++ trait-mirror test binary (tests/cpp/test_traits.cpp): the reference's own unit
tests restated in the compiled host language over include/fastlanes_b200.hpp -> C ABI -> CUDA kernels."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_cpp_trait_mirror():
    exe = os.path.join(ROOT, "build", "test_traits")
    assert os.path.exists(exe), "build/test_traits missing: run `make build/test_traits`"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ALL OK" in r.stdout
