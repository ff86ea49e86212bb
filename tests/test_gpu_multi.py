"""-m gpu, needs >= 2 GPUs (skipped otherwise): the block shard / gather of SURVEY.md §8(e) over NCCL — scatter a packed
column from rank 0, fused scan + decode per rank, gather the per-block counts, all-reduce a checksum — against the same
work on one GPU (tools/mgpu_scan.py under torchrun)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


def test_nccl_scatter_scan_gather():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29733", os.path.join(ROOT, "tools", "mgpu_scan.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["counts_match"] and line["checksum_match"] and line["world"] == world
