"""not gpu: the N>1 host logic (block sharding, max-over-ranks timing, checksum reduction) with
torch.distributed `gloo`, world size 2, on CPU.  Each rank decodes ITS shard with the CPU oracle (the
checker, standing in for the GPU kernel here); the concatenated shards must equal the single-rank decode."""
import os
import socket

import numpy as np
import pytest

from fastlanes_b200.shard import block_shard, element_range, waves


def test_block_shard_is_a_partition():
    for n in (0, 1, 7, 8, 1000, 1 << 26):
        for world in (1, 2, 3, 4, 8):
            edges = [block_shard(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for (a0, a1), (b0, b1) in zip(edges, edges[1:]):
                assert a1 == b0 and a0 <= a1
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
    assert element_range(10, 1, 2, 512) == (5 * 512, 10 * 512)
    assert list(waves(10, 4)) == [(0, 4), (4, 4), (8, 2)]
    with pytest.raises(ValueError):
        block_shard(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_blocks, width, tmpdir):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fastlanes_b200.shard import block_shard, max_over_ranks, sum_over_ranks
        from oracle import fl_oracle as oracle

        rng = np.random.default_rng(123)  # same seed on every rank: the "global" packed buffer
        packed = rng.integers(0, 1 << 32, size=n_blocks * 32 * width, dtype=np.uint32)
        b0, b1 = block_shard(n_blocks, rank, world)
        mine = oracle.unpack(packed[b0 * 32 * width: b1 * 32 * width], width, n_blocks=b1 - b0)
        np.save(os.path.join(tmpdir, f"shard{rank}.npy"), mine)
        # timing reduction: max over ranks
        t = max_over_ranks(10.0 + rank, dist)
        assert t == 10.0 + world - 1
        # verification reductions: block count and a wrapping checksum
        assert sum_over_ranks(b1 - b0, dist) == n_blocks
        chk = sum_over_ranks(int(mine.astype(np.uint64).sum() & np.uint64((1 << 62) - 1)), dist)
        np.save(os.path.join(tmpdir, f"chk{rank}.npy"), np.array([chk], dtype=np.int64))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_decode_matches_single_rank(tmp_path, oracle):
    import torch.multiprocessing as mp

    n_blocks, width, world = 101, 13, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_blocks, width, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(123)
    packed = rng.integers(0, 1 << 32, size=n_blocks * 32 * width, dtype=np.uint32)
    whole = oracle.unpack(packed, width)
    got = np.concatenate([np.load(tmp_path / f"shard{r}.npy") for r in range(world)])
    assert np.array_equal(got, whole)
    expect_chk = int(whole.astype(np.uint64).sum() & np.uint64((1 << 62) - 1))
    chks = [int(np.load(tmp_path / f"chk{r}.npy")[0]) for r in range(world)]
    # each rank's partial checksum is masked before the sum, so compare modulo 2^62 per shard sum
    parts = [int(np.load(tmp_path / f"shard{r}.npy").astype(np.uint64).sum() & np.uint64((1 << 62) - 1)) for r in range(world)]
    assert chks[0] == chks[1] == sum(parts)
    assert (sum(parts) - expect_chk) % (1 << 62) == 0


def _scatter_worker(rank, world, port, n_blocks, width, tmpdir):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fastlanes_b200.shard import block_shard, gather_blocks, scatter_blocks
        from oracle import fl_oracle as oracle

        per = 32 * width
        src = None
        if rank == 0:  # only the root holds the column
            rng = np.random.default_rng(321)
            src = torch.from_numpy(rng.integers(0, 1 << 31, size=n_blocks * per, dtype=np.int64).astype(np.int32))
        mine = scatter_blocks(src, n_blocks, per, dist)
        b0, b1 = block_shard(n_blocks, rank, world)
        assert mine.numel() == (b1 - b0) * per
        # decode the shard (the CPU oracle stands in for the GPU kernel), reduce it to per-block counts, gather those
        vals = oracle.unpack(mine.numpy().view(np.uint32), width, n_blocks=b1 - b0).reshape(b1 - b0, 1024)
        counts = torch.from_numpy((vals < (1 << (width - 1))).sum(1).astype(np.int32))
        allc = gather_blocks(counts, n_blocks, 1, dist)
        assert allc.numel() == n_blocks
        np.save(os.path.join(tmpdir, f"counts{rank}.npy"), allc.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_blocks", [101, 2, 1])
def test_scatter_decode_gather_two_ranks(tmp_path, oracle, n_blocks):
    """north_star's "NCCL only for the trivial block shard/gather": scatter a packed column from rank 0, decode per
    rank, gather a small per-block result — ragged and degenerate shard sizes (one rank may own nothing)."""
    import torch.multiprocessing as mp

    width, world = 11, 2
    mp.spawn(_scatter_worker, args=(world, _free_port(), n_blocks, width, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(321)
    packed = rng.integers(0, 1 << 31, size=n_blocks * 32 * width, dtype=np.int64).astype(np.int32).view(np.uint32)
    want = (oracle.unpack(packed, width, n_blocks=n_blocks).reshape(n_blocks, 1024) < (1 << (width - 1))).sum(1)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"counts{r}.npy"), want.astype(np.int32))
