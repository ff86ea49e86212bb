#!/usr/bin/env python
"""tools/mgpu_scan.py — block-sharded scan over the GPUs of one box (run under torchrun, one rank per GPU, NCCL).

rank 0 holds a packed u32 column on its GPU -> scatter_blocks (the one collective, outside the decode) -> every rank
runs the fused filter and the plain unpack on its shard -> gather_blocks of the per-block counts + all-reduce of a
checksum of the decoded values -> rank 0 recomputes both on the whole column with a single GPU and compares.
Prints one JSON line on rank 0; exit code 0 iff everything matches.  Used by tests/test_gpu_multi.py."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import fastlanes_b200 as fl  # noqa: E402
from fastlanes_b200.shard import block_shard, gather_blocks, scatter_blocks, sum_over_ranks  # noqa: E402


def decode(packed, n, width, ref, lo, hi):
    bitmap = torch.empty(n * 128, dtype=torch.uint8, device=packed.device)
    counts = torch.empty(n, dtype=torch.int32, device=packed.device)
    fl.Scan.filter_range(width, packed, ref, lo, hi, bitmap, counts)
    out = torch.empty(n * 1024, dtype=torch.int32, device=packed.device)
    fl.FoR.unfor_pack(width, packed, ref, out)
    chk = int((out.to(torch.int64) & 0xFFFFFFFF).sum().item())  # < 2^24 blocks * 2^10 * 2^32: exact in int64
    return counts, chk


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n_blocks, width, ref, lo, hi = (1 << 14) + 37, 13, 1000, 2000, 5000
    per = 32 * width
    src = None
    if rank == 0:
        g = torch.Generator(device=dev); g.manual_seed(7)
        src = torch.empty(n_blocks * per, dtype=torch.int32, device=dev).random_(-(1 << 31), (1 << 31) - 1, generator=g)
    mine = scatter_blocks(src, n_blocks, per, dist, device=dev)
    b0, b1 = block_shard(n_blocks, rank, world)
    counts, chk = decode(mine, b1 - b0, width, ref, lo, hi)
    all_counts = gather_blocks(counts, n_blocks, 1, dist)
    total_chk = sum_over_ranks(chk, dist, dev)
    ok = True
    if rank == 0:
        want_counts, want_chk = decode(src, n_blocks, width, ref, lo, hi)
        counts_ok = bool(torch.equal(all_counts, want_counts))
        ok = counts_ok and total_chk == want_chk
        print(json.dumps({"world": world, "n_blocks": n_blocks, "width": width, "selected": int(all_counts.sum().item()),
                          "counts_match": counts_ok, "checksum_match": total_chk == want_chk}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
