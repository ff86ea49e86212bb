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


def scatter_blocks(src, n_blocks: int, elems_per_block: int, dist, device=None, root: int = 0):
    """The "trivial block shard" of north_star: rank `root` holds a contiguous array of `n_blocks` blocks
    (`elems_per_block` elements each: 1024 unpacked, 1024*W/T packed, LANES bases); every rank receives its own
    contiguous shard block_shard(n_blocks, rank, world).  One collective (NCCL scatter over NVLink on GPUs, gloo in
    the CPU tests), OUTSIDE the decode; shards are padded to the largest shard for the collective and trimmed after.
    `src` is only read on `root` (pass None elsewhere); returns this rank's shard as a new tensor."""
    import torch

    world, rank = dist.get_world_size(), dist.get_rank()
    b0, b1 = block_shard(n_blocks, rank, world)
    per = -(-n_blocks // world) * elems_per_block  # padded shard length
    meta = [None]
    if rank == root:
        meta = [(src.dtype, )]
    dist.broadcast_object_list(meta, src=root)
    dtype = meta[0][0]
    if device is None:
        device = src.device if rank == root else torch.device("cpu")
    out = torch.empty(per, dtype=dtype, device=device)
    chunks = None
    if rank == root:
        chunks = []
        for r in range(world):
            r0, r1 = block_shard(n_blocks, r, world)
            c = torch.zeros(per, dtype=dtype, device=device)
            c[: (r1 - r0) * elems_per_block] = src[r0 * elems_per_block: r1 * elems_per_block]
            chunks.append(c)
    dist.scatter(out, chunks, src=root)
    return out[: (b1 - b0) * elems_per_block].clone() if (b1 - b0) * elems_per_block != per else out


def gather_blocks(shard, n_blocks: int, elems_per_block: int, dist):
    """Inverse of scatter_blocks for results that are small enough to be worth moving (sampled blocks, bitmaps,
    counts, per-block statistics): all-gather of the padded shards, then the padding is dropped.  Every rank gets the
    whole array.  Decoded columns should stay sharded: NVLink (0.9 TB/s) is 7x slower than the decode itself."""
    import torch

    world = dist.get_world_size()
    per = -(-n_blocks // world) * elems_per_block
    padded = torch.zeros(per, dtype=shard.dtype, device=shard.device)
    padded[: shard.numel()] = shard
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    keep = []
    for r in range(world):
        r0, r1 = block_shard(n_blocks, r, world)
        keep.append(parts[r][: (r1 - r0) * elems_per_block])
    return torch.cat(keep)
