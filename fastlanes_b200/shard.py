"""Host-side sharding of a batch of independent 1024-value blocks over the GPUs of one box.

Every op of the hot path reads and writes exactly one block (plus its own base / reference); the
reference has no cross-block state.  So the multi-GPU plan is pure partitioning:
rank r of `world` owns the contiguous block range block_shard(n, r, world) of the packed, unpacked and
base arrays alike, and NO collective sits on the data path.  torch.distributed (NCCL on GPUs, gloo in
the CPU tests) is used only for the timing barrier, the max-over-ranks of the timings and an optional
checksum all-reduce.
"""
from __future__ import annotations


def block_shard(n_blocks: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced partition: ranks differ by at most one block; the union is [0, n_blocks)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    return n_blocks * rank // world, n_blocks * (rank + 1) // world


def element_range(n_blocks: int, rank: int, world: int, elems_per_block: int) -> tuple[int, int]:
    """Element offsets of a rank's shard in an array holding `elems_per_block` elements per block
    (1024 unpacked, 1024*W/T packed, LANES for bases)."""
    b0, b1 = block_shard(n_blocks, rank, world)
    return b0 * elems_per_block, b1 * elems_per_block


def waves(n_blocks: int, wave_blocks: int):
    """Split a shard into fixed-size waves (the last may be short) — used when a shard exceeds HBM."""
    b = 0
    while b < n_blocks:
        yield b, min(wave_blocks, n_blocks - b)
        b += wave_blocks


def max_over_ranks(value: float, dist=None, device=None) -> float:
    """Max of a per-rank scalar (timings are reported as the max over ranks)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch

    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: int, dist=None, device=None) -> int:
    """Sum of a per-rank integer (block counts, wrapping 63-bit checksums)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return int(value)
    import torch

    t = torch.tensor([int(value)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())
