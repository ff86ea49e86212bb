"""-m gpu: bit-exact parity of the sm_100a kernels (through the C-ABI) against the CPU oracle.

Small/medium sizes compare every byte with the oracle; BASELINE.json's full sizes are covered in
test_gpu_configs.py.  Device path = torch CUDA tensors -> fl_<op>_<T>; host path = numpy -> fl_host_<op>_<T>.
"""
import numpy as np
import pytest

from gpu_util import DT, dev_empty, mask, rand_bytes, to_dev, to_host

pytestmark = pytest.mark.gpu

ALL_TW = [(tb, w) for tb in (8, 16, 32, 64) for w in range(tb + 1)]
N_BLOCKS = 37  # ragged against the 32-blocks-per-CTA tiling


@pytest.fixture(scope="module")
def fl():
    import torch

    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import fastlanes_b200

    assert fastlanes_b200.device_count() >= 1
    return fastlanes_b200


@pytest.mark.parametrize("tb", [8, 16, 32, 64])
def test_unpack_every_width_device(fl, oracle, tb):
    rng = np.random.default_rng(100 + tb)
    for w in range(tb + 1):
        packed = rand_bytes(rng, N_BLOCKS * 128 * w, tb)  # any bit pattern is a valid packing
        out = dev_empty(N_BLOCKS * 1024, tb)
        out.fill_(0x5A if tb == 8 else 0x5A5A)
        fl.BitPacking.unpack(w, to_dev(packed), out)
        assert np.array_equal(to_host(out, tb), oracle.unpack(packed, w, n_blocks=N_BLOCKS)), (tb, w)


@pytest.mark.parametrize("tb", [8, 16, 32, 64])
def test_pack_every_width_device(fl, oracle, tb):
    rng = np.random.default_rng(200 + tb)
    for w in range(tb + 1):
        values = rand_bytes(rng, N_BLOCKS * 128 * tb, tb)  # full-range: exercises the & mask truncation (macros.rs:73)
        out = dev_empty(N_BLOCKS * 1024 * w // tb, tb)
        fl.BitPacking.pack(w, to_dev(values), out)
        assert np.array_equal(to_host(out, tb), oracle.pack(values, w)), (tb, w)


@pytest.mark.parametrize("tb", [8, 16, 32, 64])
def test_for_family_every_width_device(fl, oracle, tb):
    rng = np.random.default_rng(300 + tb)
    for w in range(tb + 1):
        values = rand_bytes(rng, N_BLOCKS * 128 * tb, tb)
        ref = int(rand_bytes(rng, tb // 8, tb)[0])
        refs = rand_bytes(rng, N_BLOCKS * (tb // 8), tb)
        for r_host, r_dev in ((ref, ref), (refs, to_dev(refs))):
            p = dev_empty(N_BLOCKS * 1024 * w // tb, tb)
            fl.FoR.for_pack(w, to_dev(values), r_dev, p)
            expect_p = oracle.for_pack(values, r_host, w)
            assert np.array_equal(to_host(p, tb), expect_p), (tb, w, "for_pack")
            out = dev_empty(N_BLOCKS * 1024, tb)
            fl.FoR.unfor_pack(w, p, r_dev, out)
            assert np.array_equal(to_host(out, tb), oracle.unfor_pack(expect_p, r_host, w, n_blocks=N_BLOCKS)), (tb, w, "unfor_pack")


@pytest.mark.parametrize("tb", [8, 16, 32, 64])
def test_delta_family_device(fl, oracle, tb):
    rng = np.random.default_rng(400 + tb)
    L = 1024 // tb
    values = rand_bytes(rng, N_BLOCKS * 128 * tb, tb)
    base = rand_bytes(rng, N_BLOCKS * 128, tb)
    d = dev_empty(N_BLOCKS * 1024, tb)
    fl.Delta.delta(to_dev(values), to_dev(base), d)
    expect_d = oracle.delta(values, base)
    assert np.array_equal(to_host(d, tb), expect_d)
    u = dev_empty(N_BLOCKS * 1024, tb)
    fl.Delta.undelta(d, to_dev(base), u)
    assert np.array_equal(to_host(u, tb), values)
    assert base.size == N_BLOCKS * L
    for w in range(tb + 1):
        packed = rand_bytes(rng, N_BLOCKS * 128 * w, tb)
        out = dev_empty(N_BLOCKS * 1024, tb)
        fl.Delta.undelta_pack(w, to_dev(packed), to_dev(base), out)
        assert np.array_equal(to_host(out, tb), oracle.undelta_pack(packed, base, w, n_blocks=N_BLOCKS)), (tb, w)


@pytest.mark.parametrize("tb", [8, 16, 32, 64])
def test_transpose_device(fl, oracle, tb):
    rng = np.random.default_rng(500 + tb)
    for n in (1, 2, 7, N_BLOCKS):
        values = rand_bytes(rng, n * 128 * tb, tb)
        t = dev_empty(n * 1024, tb)
        fl.Transpose.transpose(to_dev(values), t)
        assert np.array_equal(to_host(t, tb), oracle.transpose(values)), (tb, n)
        u = dev_empty(n * 1024, tb)
        fl.Transpose.untranspose(t, u)
        assert np.array_equal(to_host(u, tb), values), (tb, n)
        # untranspose of arbitrary data also matches the oracle (never tested by the reference, SURVEY §4 gap 3)
        fl.Transpose.untranspose(to_dev(values), u)
        assert np.array_equal(to_host(u, tb), oracle.untranspose(values)), (tb, n)


@pytest.mark.parametrize("tb", [8, 16, 32, 64])
def test_unpack_gather_device(fl, oracle, tb):
    import torch

    rng = np.random.default_rng(600 + tb)
    for w in sorted({0, 1, 3, tb // 2, tb - 1, tb}):
        packed = rand_bytes(rng, N_BLOCKS * 128 * w, tb) if w else np.zeros(0, DT[tb])
        gi = rng.integers(0, N_BLOCKS * 1024, size=5000, dtype=np.uint64)
        out = dev_empty(gi.size, tb)
        if w == 0:
            continue  # a zero-width packed tensor has no blocks to index on the device path
        fl.BitPacking.unpack_gather(w, to_dev(packed), torch.from_numpy(gi.view(np.int64)).cuda(), out)
        assert np.array_equal(to_host(out, tb), oracle.unpack_gather(packed, w, gi)), (tb, w)
    # out-of-range index -> the reference's assert!(index < 1024) (bitpacking.rs:152)
    packed = rand_bytes(rng, 2 * 128 * 5, tb)
    bad = torch.tensor([0, 2 * 1024], dtype=torch.int64, device="cuda")
    with pytest.raises(fl.FastLanesError) as e:
        fl.BitPacking.unpack_gather(5, to_dev(packed), bad, dev_empty(2, tb))
    assert e.value.status == 3


def test_ragged_and_empty_batches(fl, oracle):
    rng = np.random.default_rng(7)
    for n in (0, 1, 3, 31, 32, 33, 255, 257):
        packed = rand_bytes(rng, n * 128 * 11, 32)
        out = dev_empty(n * 1024, 32)
        fl.BitPacking.unpack(11, to_dev(packed), out)
        assert np.array_equal(to_host(out, 32), oracle.unpack(packed, 11, n_blocks=n))
        values = to_host(out, 32)
        p2 = dev_empty(n * 32 * 11, 32)
        fl.BitPacking.pack(11, out, p2)
        assert np.array_equal(to_host(p2, 32), packed)
        assert values.size == n * 1024


def test_error_behaviour(fl):
    import torch

    v = torch.zeros(1024, dtype=torch.int16, device="cuda")
    with pytest.raises(fl.FastLanesError) as e:  # unreachable!("Unsupported width") bitpacking.rs:93
        fl.BitPacking.pack(17, v, torch.zeros(1088, dtype=torch.int16, device="cuda"))
    assert e.value.status == 1
    with pytest.raises(fl.FastLanesError) as e:  # debug_assert on lengths bitpacking.rs:78-80
        fl.BitPacking.pack(3, v, torch.zeros(191, dtype=torch.int16, device="cuda"))
    assert e.value.status == 2
    # the raw C ABI also rejects width > T and misaligned device pointers
    from fastlanes_b200 import _lib

    buf = torch.zeros(4096, dtype=torch.int32, device="cuda")
    st = _lib.fn("fl_unpack", 32)(33, 1, buf.data_ptr(), buf.data_ptr() + 8192, None)
    assert st == _lib.FL_ERR_WIDTH
    st = _lib.fn("fl_unpack", 32)(5, 1, buf.data_ptr() + 4, buf.data_ptr() + 8192, None)
    assert st == _lib.FL_ERR_ALIGN
    st = _lib.fn("fl_unpack", 32)(5, 1, None, buf.data_ptr(), None)
    assert st == _lib.FL_ERR_NULL
    assert b"" != _lib.lib().fl_last_error_string()


@pytest.mark.parametrize("tb", [8, 16, 32, 64])
def test_fused_original_order_chains(fl, oracle, tb):
    """decode straight to original order / encode straight from it, against the oracle's
    composition of the reference methods (src/delta.rs:88-99, src/transpose.rs:11-22)."""
    rng = np.random.default_rng(700 + tb)
    n = N_BLOCKS
    base = rand_bytes(rng, n * 128, tb)
    for w in range(tb + 1):
        packed = rand_bytes(rng, n * 128 * w, tb)
        out = dev_empty(n * 1024, tb)
        fl.Delta.undelta_pack_untranspose(w, to_dev(packed), to_dev(base), out)
        expect = oracle.untranspose(oracle.undelta_pack(packed, base, w, n_blocks=n))
        assert np.array_equal(to_host(out, tb), expect), (tb, w, "undelta_pack_untranspose")
        values = rand_bytes(rng, n * 128 * tb, tb)
        p = dev_empty(n * 1024 * w // tb, tb)
        fl.Delta.transpose_delta_pack(w, to_dev(values), to_dev(base), p)
        expect_p = oracle.pack(oracle.delta(oracle.transpose(values), base), w)
        assert np.array_equal(to_host(p, tb), expect_p), (tb, w, "transpose_delta_pack")
    # the README-style pipeline: encode then decode returns the input when deltas fit the width
    w = tb // 2
    if tb < 32:
        return
    values = (np.arange(n * 1024, dtype=np.uint64) * 3 + 7).astype(DT[tb])
    zero_base = np.zeros(n * (1024 // tb), dtype=DT[tb])
    p = dev_empty(n * 1024 * w // tb, tb)
    fl.Delta.transpose_delta_pack(w, to_dev(values), to_dev(zero_base), p)
    back = dev_empty(n * 1024, tb)
    fl.Delta.undelta_pack_untranspose(w, p, to_dev(zero_base), back)
    # per-block: values restart, so the first delta of each lane is the value itself; only check it round-trips
    # when every delta (incl. the first, relative to base 0) fits in w bits: true for block 0 of u32/u64 here
    assert np.array_equal(to_host(back, tb)[:1024], values[:1024])


@pytest.mark.parametrize("tb", [8, 16, 32, 64])
def test_block_minmax_and_for_pipeline(fl, oracle, tb):
    """per-block statistics -> reference / width -> for_pack_refs -> unfor_pack_refs round trip."""
    import torch

    rng = np.random.default_rng(800 + tb)
    n = 133
    span_bits = tb // 2 - 1
    lo = rand_bytes(rng, n * (tb // 8), tb) >> DT[tb](1)                      # per-block offsets (no overflow)
    values = (np.repeat(lo, 1024) + (rand_bytes(rng, n * 128 * tb, tb) & DT[tb](mask(span_bits)))).astype(DT[tb])
    d_values = to_dev(values)
    d_min, d_max = dev_empty(n, tb), dev_empty(n, tb)
    fl.FoR.block_minmax(d_values, d_min, d_max)
    v2 = values.reshape(n, 1024)
    assert np.array_equal(to_host(d_min, tb), v2.min(axis=1))
    assert np.array_equal(to_host(d_max, tb), v2.max(axis=1))
    h_min, h_max = np.zeros(n, DT[tb]), np.zeros(n, DT[tb])
    fl.FoR.block_minmax(values, h_min, h_max)                                 # host family
    assert np.array_equal(h_min, v2.min(axis=1)) and np.array_equal(h_max, v2.max(axis=1))
    refs, w = fl.FoR.choose(h_min, h_max, tb)
    assert w <= span_bits
    packed = dev_empty(n * 1024 * w // tb, tb)
    fl.FoR.for_pack(w, d_values, d_min, packed)
    assert np.array_equal(to_host(packed, tb), oracle.for_pack(values, refs, w))
    back = dev_empty(n * 1024, tb)
    fl.FoR.unfor_pack(w, packed, d_min, back)
    assert torch.equal(back, d_values)
