"""-m gpu: the multi-device context family (fl_ctx_*, include/fastlanes_b200.h) and single-process multi-device use.

SURVEY.md §8(e): blocks are independent, so a batch shards by contiguous block ranges with no exchange between devices.
The context runs shard i on device i from its own worker thread.  On a 1-GPU box the context lists device 0 twice
(and three times) — each entry is an independent shard worker with its own pipeline — so the sharding logic, the
offsets of every array and the error paths are exercised everywhere; with >= 2 GPUs the real devices are used too.
Every result is compared bit-exactly with the CPU oracle on the whole batch."""
import numpy as np
import pytest

from gpu_util import DT, rand_bytes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fl():
    import fastlanes_b200

    assert fastlanes_b200.device_count() >= 1
    return fastlanes_b200


def device_lists(fl):
    n = fl.device_count()
    lists = [[0, 0], [0, 0, 0]]
    if n >= 2:
        lists += [list(range(n)), [1, 0]]
    return lists


def test_block_range_matches_python_shard(fl):
    from fastlanes_b200.shard import block_shard

    with fl.Context([0, 0, 0]) as ctx:
        assert ctx.devices == [0, 0, 0]
        for n in (0, 1, 2, 3, 7, 1000, (1 << 26) + 5):
            for i in range(3):
                assert ctx.block_range(n, i) == block_shard(n, i, 3)
        with pytest.raises(fl.FastLanesError) as e:
            ctx.block_range(10, 3)
        assert e.value.status == 3  # FL_ERR_INDEX


@pytest.mark.parametrize("tb", [8, 16, 32, 64])
def test_ctx_every_op_vs_oracle(fl, oracle, tb):
    rng = np.random.default_rng(900 + tb)
    dt = DT[tb]
    lanes = 1024 // tb
    for devs in device_lists(fl):
        with fl.Context(devs) as ctx:
            for n in (1, 2, 5, 67):  # fewer blocks than shards, ragged shards
                for w in sorted({0, 1, tb // 2 + 1, tb - 1, tb}):
                    values = rand_bytes(rng, n * 128 * tb, tb)
                    base = rand_bytes(rng, n * 128, tb)
                    ref = int(rand_bytes(rng, tb // 8, tb)[0])
                    packed = np.zeros(n * 1024 * w // tb, dtype=dt)
                    ctx.pack(w, values, packed)
                    assert np.array_equal(packed, oracle.pack(values, w)), (devs, n, w, "pack")
                    out = np.full(n * 1024, 0xEE, dtype=dt)
                    ctx.unpack(w, packed, out)
                    assert np.array_equal(out, oracle.unpack(packed, w, n_blocks=n)), (devs, n, w, "unpack")
                    fp = np.zeros_like(packed)
                    ctx.for_pack(w, values, ref, fp)
                    assert np.array_equal(fp, oracle.for_pack(values, ref, w)), (devs, n, w, "for_pack")
                    ctx.unfor_pack(w, fp, ref, out)
                    assert np.array_equal(out, oracle.unfor_pack(fp, ref, w, n_blocks=n)), (devs, n, w, "unfor_pack")
                    ctx.undelta_pack(w, packed, base, out)
                    want = oracle.undelta_pack(packed, base, w, n_blocks=n)
                    assert np.array_equal(out, want), (devs, n, w, "undelta_pack")
                    ctx.undelta_pack_untranspose(w, packed, base, out)
                    assert np.array_equal(out, oracle.untranspose(want)), (devs, n, w, "undelta_pack_untranspose")
                    tdp = np.zeros_like(packed)
                    ctx.transpose_delta_pack(w, values, base, tdp)
                    assert np.array_equal(tdp, oracle.pack(oracle.delta(oracle.transpose(values), base), w)), (devs, n, w, "tdp")
                    # fused scans: bitmap + counts shard on the block index too
                    full = (1 << tb) - 1
                    lo, hi = full // 4, full // 4 * 3
                    bitmap = np.zeros(n * 128, dtype=np.uint8)
                    counts = np.zeros(n, dtype=np.uint32)
                    ctx.filter_range(w, packed, ref, lo, hi, bitmap, counts)
                    got = oracle.unfor_pack(packed, ref, w, n_blocks=n)
                    sel = (got >= dt(lo)) & (got <= dt(hi))
                    assert np.array_equal(bitmap, np.packbits(sel, bitorder="little")), (devs, n, w, "filter")
                    assert np.array_equal(counts, sel.reshape(n, 1024).sum(1).astype(np.uint32)), (devs, n, w, "counts")
                    ctx.filter_range_delta(w, packed, base, lo, hi, bitmap, counts)
                    orig = oracle.untranspose(want)
                    sel = (orig >= dt(lo)) & (orig <= dt(hi))
                    assert np.array_equal(bitmap, np.packbits(sel, bitorder="little")), (devs, n, w, "delta filter")
                values = rand_bytes(rng, n * 128 * tb, tb)
                base = rand_bytes(rng, n * 128, tb)
                out = np.zeros_like(values)
                ctx.delta(values, base, out)
                assert np.array_equal(out, oracle.delta(values, base)), (devs, n, "delta")
                ctx.undelta(values, base, out)
                assert np.array_equal(out, oracle.undelta(values, base)), (devs, n, "undelta")
                ctx.transpose(values, out)
                assert np.array_equal(out, oracle.transpose(values)), (devs, n, "transpose")
                ctx.untranspose(values, out)
                assert np.array_equal(out, oracle.untranspose(values)), (devs, n, "untranspose")
                mins, maxs = np.zeros(n, dtype=dt), np.zeros(n, dtype=dt)
                ctx.block_minmax(values, mins, maxs)
                assert np.array_equal(mins, values.reshape(n, 1024).min(1)) and np.array_equal(maxs, values.reshape(n, 1024).max(1))
    assert lanes * tb == 1024


def test_ctx_large_batch_pinned_sharded_buffers(fl, oracle):
    """Bulk path (many pipeline chunks per shard) through shard-placed page-locked buffers."""
    rng = np.random.default_rng(77)
    n, w = 40_000 + 3, 13
    for devs in device_lists(fl)[:1] + device_lists(fl)[2:3]:
        with fl.Context(devs) as ctx:
            packed = ctx.pinned_empty(n, 32 * w, np.uint32)
            out = ctx.pinned_empty(n, 1024, np.uint32)
            packed[:] = rng.integers(0, 1 << 32, size=packed.size, dtype=np.uint32)
            out.fill(0)
            ctx.unpack(w, packed, out)
            assert np.array_equal(out, oracle.unpack(packed, w, n_blocks=n, threads=8))
            back = np.zeros_like(packed)
            ctx.pack(w, out, back)
            assert np.array_equal(back, packed)
            # the placement query works on both ends of the buffer (-1 = platform does not say)
            assert fl.buffer_node(out) >= -1 and fl.buffer_node(out, out.nbytes - 1) >= -1


def test_ctx_errors_propagate(fl):
    with fl.Context([0, 0]) as ctx:
        values = np.zeros(2048, dtype=np.uint16)
        with pytest.raises(fl.FastLanesError) as e:
            ctx.pack(17, values, np.zeros(10, dtype=np.uint16))
        assert e.value.status == 1  # FL_ERR_WIDTH, raised by the mirror like the reference's unreachable!()
        # through the raw ABI: width beyond T is refused before any worker runs
        from fastlanes_b200 import _lib

        st = _lib.fn("fl_ctx_host_unpack", 16)(ctx._h, 17, 2, values.ctypes.data, values.ctypes.data)
        assert st == 1
        st = _lib.fn("fl_ctx_host_unpack", 16)(ctx._h, 3, 2, None, values.ctypes.data)
        assert st == 6 and b"device 0" in _lib.lib().fl_last_error_string()  # FL_ERR_NULL from a worker, message carried over
        # zero blocks is a no-op
        assert _lib.fn("fl_ctx_host_unpack", 16)(ctx._h, 3, 0, None, None) == 0
    with pytest.raises(fl.FastLanesError):
        fl.Context([fl.device_count()])  # no such device


def test_ctx_scatter_gather_blocks(fl, oracle):
    """The 'trivial block shard/gather' of north_star between DEVICE buffers: scatter a packed column, decode every
    shard on its own device, gather the decoded shards back, compare with the oracle."""
    import torch

    rng = np.random.default_rng(5)
    n, w = 1001, 11
    packed = rng.integers(0, 1 << 32, size=n * 32 * w, dtype=np.uint32)
    for devs in device_lists(fl):
        with fl.Context(devs) as ctx:
            src = torch.from_numpy(packed.view(np.int32)).to(f"cuda:{devs[0]}")
            shards = ctx.scatter_blocks(src, 32 * w, root=0)
            assert [s.device.index for s in shards] == devs
            outs = []
            for i, s in enumerate(shards):
                b0, b1 = ctx.block_range(n, i)
                assert s.numel() == (b1 - b0) * 32 * w
                o = torch.empty((b1 - b0) * 1024, dtype=torch.int32, device=s.device)
                fl.BitPacking.unpack(w, s, o)
                outs.append(o)
            whole = ctx.gather_blocks(outs, 1024, root=0)
            assert np.array_equal(whole.cpu().numpy().view(np.uint32), oracle.unpack(packed, w, n_blocks=n))


def test_second_device_in_one_process_u64(fl, oracle):
    """the > 48 KiB dynamic shared-memory opt-in used to be cached once per process, so the u64 pack /
    delta / transpose / fused kernels failed on every device but the first.  Run them on device 1 AFTER device 0."""
    import torch

    if fl.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    rng = np.random.default_rng(64)
    n = 37
    values = rand_bytes(rng, n * 128 * 64, 64)
    base = rand_bytes(rng, n * 128, 64)
    for dev in (0, 1, 0):
        d = f"cuda:{dev}"
        v = torch.from_numpy(values.view(np.int64)).to(d)
        b = torch.from_numpy(base.view(np.int64)).to(d)
        for w in (1, 17, 33, 48, 64):
            p = torch.empty(n * 16 * w, dtype=torch.int64, device=d)
            fl.BitPacking.pack(w, v, p)
            assert np.array_equal(p.cpu().numpy().view(np.uint64), oracle.pack(values, w)), (dev, w, "pack")
            o = torch.empty_like(v)
            fl.Delta.undelta_pack_untranspose(w, p, b, o)
            want = oracle.untranspose(oracle.undelta_pack(oracle.pack(values, w), base, w, n_blocks=n))
            assert np.array_equal(o.cpu().numpy().view(np.uint64), want), (dev, w, "undelta_pack_untranspose")
            fl.Delta.transpose_delta_pack(w, v, b, p)
            assert np.array_equal(p.cpu().numpy().view(np.uint64), oracle.pack(oracle.delta(oracle.transpose(values), base), w)), (dev, w)
        o = torch.empty_like(v)
        fl.Delta.delta(v, b, o)
        assert np.array_equal(o.cpu().numpy().view(np.uint64), oracle.delta(values, base)), (dev, "delta")
        fl.Delta.undelta(v, b, o)
        assert np.array_equal(o.cpu().numpy().view(np.uint64), oracle.undelta(values, base)), (dev, "undelta")
        fl.Transpose.transpose(v, o)
        assert np.array_equal(o.cpu().numpy().view(np.uint64), oracle.transpose(values)), (dev, "transpose")
        fl.Transpose.untranspose(v, o)
        assert np.array_equal(o.cpu().numpy().view(np.uint64), oracle.untranspose(values)), (dev, "untranspose")


def test_mirror_launches_on_the_tensors_device(fl, oracle):
    """tensors on cuda:1 while cuda:0 is current must launch on device 1 (its current stream), and
    tensors of two devices in one call are refused."""
    import torch

    if fl.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    rng = np.random.default_rng(3)
    packed = rng.integers(0, 1 << 32, size=9 * 32 * 7, dtype=np.uint32)
    torch.cuda.set_device(0)
    p1 = torch.from_numpy(packed.view(np.int32)).to("cuda:1")
    o1 = torch.empty(9 * 1024, dtype=torch.int32, device="cuda:1")
    fl.BitPacking.unpack(7, p1, o1)
    assert torch.cuda.current_device() == 0
    assert np.array_equal(o1.cpu().numpy().view(np.uint32), oracle.unpack(packed, 7, n_blocks=9))
    o0 = torch.empty(9 * 1024, dtype=torch.int32, device="cuda:0")
    with pytest.raises(fl.FastLanesError):
        fl.BitPacking.unpack(7, p1, o0)
