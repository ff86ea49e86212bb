"""CPU: the bit arithmetic of the fused scan kernels (fastlanes_b200/csrc/fl_scan_bits.h — the very functions the
CUDA kernels call) run on an emulated warp and compared with a brute-force bitmap built from
index(row, lane) (src/macros.rs:20-24).  Needs g++ only."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_scan_bit_helpers_on_emulated_warp():
    subprocess.run(["make", "-s", "-C", ROOT, "build/test_scan_bits"], check=True)
    r = subprocess.run([os.path.join(ROOT, "build", "test_scan_bits")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "scan bits ok" in r.stdout
