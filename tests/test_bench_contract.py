"""not gpu: bench.py's reference arm prints ONE JSON line with the contract's keys (runs the CPU oracle on a tiny
sample); the roofline arithmetic helpers match SURVEY.md §8(d)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--cpu-log2-blocks", "10"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "Gint/s" and d["higher_is_better"] is True and d["dtype"] == "u32"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["value"] > 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_roofline_arithmetic():
    sys.path.insert(0, ROOT)
    import bench

    # u32 unpack moves 128*(W+32) bytes per block
    assert bench.algorithmic_bytes_per_block(1) == 4224
    assert bench.algorithmic_bytes_per_block(16) == 6144
    assert bench.algorithmic_bytes_per_block(32) == 8192
    peak, src = bench.load_peak()
    assert peak > 1000 and ("measured" in src or "fallback" in src)
