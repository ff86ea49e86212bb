"""-m gpu: the HOST-buffer C-ABI family (fl_host_*), i.e. the drop-in for the reference's trait calls.

These read like the reference's own tests (src/bitpacking.rs:249-315, src/delta.rs:81-107,
src/ffor.rs:67-88, README.md:14-47), with the GPU library in place of the crate and the CPU oracle as checker.
"""
import numpy as np
import pytest

from gpu_util import DT, mask, rand_bytes

pytestmark = pytest.mark.gpu

ALL_TW = [(tb, w) for tb in (8, 16, 32, 64) for w in range(tb + 1)]


@pytest.fixture(scope="module")
def fl():
    import fastlanes_b200

    assert fastlanes_b200.device_count() >= 1
    return fastlanes_b200


def test_readme_example_config1(fl, oracle):
    # BASELINE config 1 / README.md:14-47: u16, W=3, values[i] = i % 8
    W = 3
    values = (np.arange(1024) % (1 << W)).astype(np.uint16)
    packed = np.zeros(128 * W // 2, dtype=np.uint16)
    fl.BitPacking.pack(W, values, packed)
    assert np.array_equal(packed, oracle.pack(values, W))  # byte-exact wire format
    unpacked = np.zeros(1024, dtype=np.uint16)
    fl.BitPacking.unpack(W, packed, unpacked)
    assert np.array_equal(values, unpacked)
    for i in range(0, 1024, 37):
        assert fl.BitPacking.unpack_single(W, packed, i) == values[i]
    gi = np.arange(1024, dtype=np.uint64)
    singles = np.zeros(1024, dtype=np.uint16)
    fl.BitPacking.unpack_gather(W, packed, gi, singles)
    assert np.array_equal(singles, values)


@pytest.mark.parametrize("tb", [8, 16, 32, 64])
def test_round_trip_every_width(fl, oracle, tb):
    # src/bitpacking.rs:273-315: values[i] = i % (1 << (W % T)); pack -> unpack -> every unpack_single
    gi = np.arange(1024, dtype=np.uint64)
    for w in range(tb + 1):
        values = (np.arange(1024, dtype=np.uint64) % np.uint64(1 << (w % tb))).astype(DT[tb])
        packed = np.zeros(1024 * w // tb, dtype=DT[tb])
        fl.BitPacking.pack(w, values, packed)
        assert np.array_equal(packed, oracle.pack(values, w)), (tb, w)
        unpacked = np.full(1024, 0xEE, dtype=DT[tb])
        fl.BitPacking.unpack(w, packed, unpacked)
        assert np.array_equal(unpacked, values), (tb, w)
        singles = np.full(1024, 0xEE, dtype=DT[tb])
        fl.BitPacking.unpack_gather(w, packed, gi, singles) if w else singles.fill(0)
        assert np.array_equal(singles, values), (tb, w)


def test_unchecked_pack_u32_iota_w10(fl):
    # src/bitpacking.rs:249-256
    values = np.arange(1024, dtype=np.uint32)
    packed = np.zeros(320, dtype=np.uint32)
    fl.BitPacking.unchecked_pack(10, values, packed)
    out = np.zeros(1024, dtype=np.uint32)
    fl.BitPacking.unchecked_unpack(10, packed, out)
    assert np.array_equal(values, out)
    assert [int(x) for x in packed[:4]] == [0x10020000, 0x50120401, 0x90220802, 0xD0320C03]


def test_unpack_single_u32_iota_w16(fl):
    # src/bitpacking.rs:259-271
    values = np.arange(1024, dtype=np.uint32)
    packed = np.zeros(512, dtype=np.uint32)
    fl.BitPacking.pack(16, values, packed)
    for i in (0, 1, 31, 32, 511, 777, 1023):
        assert fl.BitPacking.unpack_single(16, packed, i) == values[i]
        assert fl.BitPacking.unchecked_unpack_single(16, packed, i) == values[i]
    with pytest.raises(fl.FastLanesError) as e:
        fl.BitPacking.unpack_single(16, packed, 1024)  # assert!(index < 1024) bitpacking.rs:152
    assert e.value.status == 3


def test_delta_like_the_reference(fl, oracle):
    # src/delta.rs:81-107
    W = 15
    values = (np.arange(1024) // 8).astype(np.uint16)
    transposed = np.zeros(1024, dtype=np.uint16)
    fl.Transpose.transpose(values, transposed)
    assert np.array_equal(transposed, oracle.transpose(values))
    base = np.zeros(64, dtype=np.uint16)
    deltas = np.zeros(1024, dtype=np.uint16)
    fl.Delta.delta(transposed, base, deltas)
    packed = np.zeros(128 * W // 2, dtype=np.uint16)
    fl.BitPacking.pack(W, deltas, packed)
    unpacked = np.zeros(1024, dtype=np.uint16)
    fl.Delta.undelta_pack(W, packed, base, unpacked)  # fused
    assert np.array_equal(transposed, unpacked)
    fl.BitPacking.unpack(W, packed, unpacked)  # unfused
    undelta = np.zeros(1024, dtype=np.uint16)
    fl.Delta.undelta(unpacked, base, undelta)
    assert np.array_equal(transposed, undelta)
    back = np.zeros(1024, dtype=np.uint16)
    fl.Transpose.untranspose(undelta, back)
    assert np.array_equal(back, values)


def test_ffor_like_the_reference(fl):
    # src/ffor.rs:67-88
    W = 15
    values = (np.arange(1024) % (1 << W)).astype(np.uint16)
    packed = np.zeros(128 * W // 2, dtype=np.uint16)
    fl.FoR.for_pack(W, values, 10, packed)
    unpacked = np.zeros(1024, dtype=np.uint16)
    fl.BitPacking.unpack(W, packed, unpacked)
    assert np.array_equal(unpacked, ((values.astype(np.int64) - 10) & mask(W)).astype(np.uint16))
    fl.FoR.unfor_pack(W, packed, 10, unpacked)
    assert np.array_equal(unpacked, ((((values.astype(np.int64) - 10) & mask(W)) + 10) & 0xFFFF).astype(np.uint16))


def test_host_pipeline_many_chunks(fl, oracle):
    # more blocks than one pipelined chunk, ragged tail, small chunks to force slot reuse
    fl.host_configure(chunk_blocks=100, n_streams=3)
    try:
        rng = np.random.default_rng(11)
        n = 1037
        for tb, w in ((32, 13), (64, 37), (16, 9), (8, 5)):
            packed = rand_bytes(rng, n * 128 * w, tb)
            out = np.zeros(n * 1024, dtype=DT[tb])
            fl.BitPacking.unpack(w, packed, out)
            assert np.array_equal(out, oracle.unpack(packed, w, threads=8))
            base = rand_bytes(rng, n * 128, tb)
            fl.Delta.undelta_pack(w, packed, base, out)
            assert np.array_equal(out, oracle.undelta_pack(packed, base, w, threads=8))
            back = np.zeros_like(packed)
            fl.BitPacking.unpack(w, packed, out)
            fl.BitPacking.pack(w, out, back)
            assert np.array_equal(back, packed)
    finally:
        fl.host_configure(0, 0)


def test_pinned_buffers(fl, oracle):
    rng = np.random.default_rng(12)
    n, w = 300, 21
    packed = fl.pinned_empty(n * 32 * w, np.uint32)
    packed[:] = rand_bytes(rng, n * 128 * w, 32)
    out = fl.pinned_empty(n * 1024, np.uint32)
    fl.BitPacking.unpack(w, packed, out)
    assert np.array_equal(out, oracle.unpack(np.array(packed), w, threads=4))


@pytest.mark.parametrize("tb", [8, 16, 32, 64])
def test_direct_path_page_locked_buffers(fl, oracle, tb):
    """Mid-size calls whose buffers are ALL page-locked run as one launch on the caller's memory (no staging copies):
    every op family, widths incl. 0 and T, block counts above the 256 KiB low-latency limit, against the oracle; then the
    same call with one pageable buffer (the chunked pipeline) must give the same bytes."""
    rng = np.random.default_rng(1200 + tb)
    n = 2 * ((256 << 10) // (128 * tb)) + 37
    for w in (0, 1, tb // 2 + 1, tb - 1, tb):
        values = fl.pinned_empty(n * 1024, DT[tb]); values[:] = rand_bytes(rng, n * 128 * tb, tb)
        base = fl.pinned_empty(n * (1024 // tb), DT[tb]); base[:] = rand_bytes(rng, n * 128, tb)
        packed = fl.pinned_empty(max(1, n * 1024 * w // tb), DT[tb])[: n * 1024 * w // tb]
        out = fl.pinned_empty(n * 1024, DT[tb])
        fl.BitPacking.pack(w, values, packed)
        assert np.array_equal(packed, oracle.pack(np.array(values), w, threads=4)), (tb, w)
        out[:] = 0x5A
        fl.BitPacking.unpack(w, packed, out)
        assert np.array_equal(out, oracle.unpack(np.array(packed), w, n_blocks=n, threads=4)), (tb, w)
        ref = int(rng.integers(0, 1 << min(tb, 62)))
        fl.FoR.unfor_pack(w, packed, ref, out)
        assert np.array_equal(out, oracle.unfor_pack(np.array(packed), ref, w, n_blocks=n)), (tb, w)
        fl.Delta.undelta_pack(w, packed, base, out)
        want = oracle.undelta_pack(np.array(packed), np.array(base), w, n_blocks=n)
        assert np.array_equal(out, want), (tb, w)
        pageable_out = np.zeros(n * 1024, dtype=DT[tb])
        fl.Delta.undelta_pack(w, packed, base, pageable_out)
        assert np.array_equal(pageable_out, want), (tb, w)
        fl.Delta.transpose_delta_pack(w, values, base, packed)
        fl.Delta.undelta_pack_untranspose(w, packed, base, out)
        if w == tb:
            assert np.array_equal(out, values), (tb, w)
    fl.Transpose.transpose(values, out)
    assert np.array_equal(out, oracle.transpose(np.array(values))), tb


def test_small_calls_every_op_and_size_boundary(fl, oracle):
    """The low-latency path (zero-copy page-locked staging, one launch) takes calls whose in + base + out fit in 256 KiB;
    the block counts here straddle that limit for every type so both paths are compared with the oracle."""
    rng = np.random.default_rng(21)
    for tb in (8, 16, 32, 64):
        per_block = 128 * tb * 2 + 128  # roughly in + out + base bytes at W = T
        edge = (256 << 10) // per_block
        for n in (1, 2, edge - 1, edge, edge + 1, 2 * edge + 3):
            for w in (0, 1, tb // 2 + 1, tb):
                values = rand_bytes(rng, n * 128 * tb, tb)
                base = rand_bytes(rng, n * 128, tb)
                packed = np.zeros(n * 1024 * w // tb, dtype=DT[tb])
                fl.BitPacking.pack(w, values, packed)
                assert np.array_equal(packed, oracle.pack(values, w)), (tb, n, w)
                out = np.full(n * 1024, 0x5A, dtype=DT[tb])
                fl.Delta.undelta_pack(w, packed, base, out)
                assert np.array_equal(out, oracle.undelta_pack(packed, base, w, n_blocks=n)), (tb, n, w)
                fl.Delta.undelta_pack_untranspose(w, packed, base, out)
                assert np.array_equal(out, oracle.untranspose(oracle.undelta_pack(packed, base, w, n_blocks=n))), (tb, n, w)
            values = rand_bytes(rng, n * 128 * tb, tb)
            out = np.zeros_like(values)
            fl.Transpose.transpose(values, out)
            assert np.array_equal(out, oracle.transpose(values)), (tb, n)


def test_gather_moves_only_referenced_blocks(fl, oracle):
    """fl_host_unpack_gather: few indices into a large column (zero-copy path), many indices (compacted distinct blocks
    through device staging), more indices than blocks (whole-column copy), out-of-range index."""
    rng = np.random.default_rng(22)
    for tb, w, n_blocks in ((16, 9, 5000), (32, 17, 3000), (64, 33, 700), (8, 5, 4000)):
        packed = rand_bytes(rng, n_blocks * 128 * w, tb)
        for n in (1, 7, 130, 2500, n_blocks + 17):
            gi = rng.integers(0, n_blocks * 1024, size=n, dtype=np.uint64)
            if n >= 130:  # runs of consecutive blocks and repeated blocks
                gi[: n // 2] = np.sort(gi[: n // 2])
                gi[n // 2: n // 2 + 20] = gi[0]
            out = np.zeros(n, dtype=DT[tb])
            fl.BitPacking.unpack_gather(w, packed, gi, out)
            assert np.array_equal(out, oracle.unpack_gather(packed, w, gi)), (tb, w, n)
        gi = np.array([5, n_blocks * 1024], dtype=np.uint64)
        with pytest.raises(fl.FastLanesError) as e:
            fl.BitPacking.unpack_gather(w, packed, gi, np.zeros(2, dtype=DT[tb]))
        assert e.value.status == 3
    # the reference's own loop: 1024 unpack_single calls on one block (src/bitpacking.rs:259-271)
    values = np.arange(1024, dtype=np.uint32)
    packed = np.zeros(512, dtype=np.uint32)
    fl.BitPacking.pack(16, values, packed)
    assert [int(fl.BitPacking.unpack_single(16, packed, i)) for i in range(1024)] == list(range(1024))


def test_copy_probe_and_placement_queries(fl):
    from fastlanes_b200 import _lib

    n = 3000
    src = fl.pinned_empty(n * 32 * 8, np.uint32)
    dst = fl.pinned_empty(n * 1024, np.uint32)
    src.fill(1)
    assert _lib.lib().fl_host_copy_probe(128 * 8, 4096, n, src.ctypes.data, dst.ctypes.data) == 0
    assert _lib.lib().fl_host_copy_probe(0, 4096, n, None, dst.ctypes.data) == 0
    node = _lib.lib().fl_device_numa_node(0)
    got = fl.buffer_node(dst)
    assert got >= -1
    if node >= 0 and got >= 0 and got != node:  # MPOL_PREFERRED is a preference: a full node may spill
        import warnings

        warnings.warn(f"fl_host_alloc placed the buffer on node {got}, device 0 is on node {node}")
