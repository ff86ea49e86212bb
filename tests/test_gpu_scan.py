"""-m gpu: the fused scan kernels (fl_unpack_filter / fl_unpack_select) against the
composition they replace: oracle `unfor_pack` (src/ffor.rs:38-50) / `unpack` (src/bitpacking.rs:98-107) followed by
the caller-side loop over the 1024 values (README.md:40-41) — here numpy compare / packbits / boolean take.
Bit-exact, every element type and every width, through the C ABI."""
import numpy as np
import pytest

from gpu_util import DT, dev_empty, mask, rand_bytes, to_dev, to_host

pytestmark = pytest.mark.gpu

N_BLOCKS = 37  # ragged against the 8-blocks-per-CTA tiling


@pytest.fixture(scope="module")
def fl():
    import torch

    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import fastlanes_b200

    return fastlanes_b200


def expected_bitmap(values: np.ndarray, lo: int, hi: int) -> np.ndarray:
    sel = (values >= values.dtype.type(lo)) & (values <= values.dtype.type(hi)) if lo <= hi else np.zeros(values.shape, bool)
    return np.packbits(sel, bitorder="little"), sel


def ranges_for(rng, tb: int, w: int):
    """A few (reference, lo, hi) triples: mid-range band, equality, empty (hi < lo), everything, wrap-around ref."""
    full = mask(tb)
    m = mask(w) if w else 0
    mid_lo, mid_hi = m // 4, m // 2 + 1
    r = int(rng.integers(0, 1 << min(tb, 62)))
    return [
        (0, mid_lo, mid_hi),
        (0, m // 3, m // 3),
        (0, 5, 4),
        (0, 0, full),
        (r, (r + mid_lo) & full, (r + mid_hi) & full),  # may wrap to hi < lo: then nothing is selected
        (full, 0, mid_hi),  # reference = -1: values wrap
    ]


@pytest.mark.parametrize("tb", [8, 16, 32, 64])
def test_filter_every_width_device(fl, oracle, tb):
    import torch

    rng = np.random.default_rng(900 + tb)
    for w in range(tb + 1):
        packed = rand_bytes(rng, N_BLOCKS * 128 * w, tb)
        d_packed = to_dev(packed)
        for ref, lo, hi in ranges_for(rng, tb, w):
            values = oracle.unfor_pack(packed, ref, w, n_blocks=N_BLOCKS)
            want, sel = expected_bitmap(values, lo, hi)
            bitmap = torch.full((N_BLOCKS * 128,), 0xA5, dtype=torch.uint8, device="cuda")
            counts = torch.full((N_BLOCKS,), -1, dtype=torch.int32, device="cuda")
            fl.Scan.filter_range(w, d_packed, ref, lo, hi, bitmap, counts)
            assert np.array_equal(bitmap.cpu().numpy(), want), (tb, w, ref, lo, hi)
            assert np.array_equal(counts.cpu().numpy().view(np.uint32), sel.reshape(N_BLOCKS, 1024).sum(1).astype(np.uint32)), (tb, w, "counts")


@pytest.mark.parametrize("tb", [8, 16, 32, 64])
def test_filter_per_block_references(fl, oracle, tb):
    import torch

    rng = np.random.default_rng(950 + tb)
    for w in (0, 1, tb // 2 - 1, tb // 2, tb - 1, tb):
        packed = rand_bytes(rng, N_BLOCKS * 128 * w, tb)
        refs = rand_bytes(rng, N_BLOCKS * (tb // 8), tb)
        values = oracle.unfor_pack(packed, refs, w, n_blocks=N_BLOCKS)
        lo, hi = int(mask(tb) // 3), int(mask(tb) // 3 * 2)
        want, _ = expected_bitmap(values, lo, hi)
        bitmap = torch.zeros(N_BLOCKS * 128, dtype=torch.uint8, device="cuda")
        fl.Scan.filter_range(w, to_dev(packed), to_dev(refs), lo, hi, bitmap)
        assert np.array_equal(bitmap.cpu().numpy(), want), (tb, w)


@pytest.mark.parametrize("tb", [8, 16, 32, 64])
def test_select_every_width_device(fl, oracle, tb):
    import torch

    rng = np.random.default_rng(1000 + tb)
    for w in range(tb + 1):
        packed = rand_bytes(rng, N_BLOCKS * 128 * w, tb)
        ref = int(rng.integers(0, 1 << min(tb, 62)))
        values = oracle.unfor_pack(packed, ref, w, n_blocks=N_BLOCKS)
        # selection bitmaps independent of the values: random density per block incl. empty and full blocks
        sel = rng.random(N_BLOCKS * 1024) < np.repeat(rng.choice([0.0, 0.02, 0.5, 0.97, 1.0], N_BLOCKS), 1024)
        bitmap = np.packbits(sel, bitorder="little")
        counts = sel.reshape(N_BLOCKS, 1024).sum(1).astype(np.int64)
        offsets = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int64)
        total = int(counts.sum())
        out = dev_empty(total + 16, tb)
        out.fill_(0x3C if tb == 8 else 0x3C3C)
        fl.Scan.select(w, to_dev(packed), ref, torch.from_numpy(bitmap).cuda(), torch.from_numpy(offsets).cuda(), out)
        got = to_host(out, tb)
        assert np.array_equal(got[:total], values[sel]), (tb, w)
        assert np.all(got[total:] == DT[tb](0x3C if tb == 8 else 0x3C3C)), (tb, w, "wrote past the selected count")


@pytest.mark.parametrize("tb", [8, 16, 32, 64])
def test_select_per_block_references_ragged(fl, oracle, tb):
    """Per-block FoR references (added on the selected values only, src/ffor.rs:47) on a batch that is ragged against the
    blocks-per-warp loop of the kernel (1 and 8 per warp depending on type / width) and whose output runs start at every
    phase of a 16-byte vector; sparse bitmaps with many empty blocks exercise the early exit inside that loop."""
    import torch

    rng = np.random.default_rng(1300 + tb)
    n = 1031
    for w in sorted({0, 1, 4, tb // 4, tb // 2 + 1, tb - 3, tb}):
        packed = rand_bytes(rng, n * 128 * w, tb)
        refs = rand_bytes(rng, n * (tb // 8), tb)
        values = oracle.unfor_pack(packed, refs, w, n_blocks=n)
        sel = rng.random(n * 1024) < np.repeat(rng.choice([0.0, 0.0, 0.001, 0.25, 1.0], n), 1024)
        bitmap = np.packbits(sel, bitorder="little")
        counts = sel.reshape(n, 1024).sum(1).astype(np.int64)
        offsets = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int64)
        total = int(counts.sum())
        out = dev_empty(total + 16, tb)
        out.fill_(0x3C if tb == 8 else 0x3C3C)
        fl.Scan.select(w, to_dev(packed), to_dev(refs), torch.from_numpy(bitmap).cuda(), torch.from_numpy(offsets).cuda(), out)
        got = to_host(out, tb)
        assert np.array_equal(got[:total], values[sel]), (tb, w)
        assert np.all(got[total:] == DT[tb](0x3C if tb == 8 else 0x3C3C)), (tb, w, "wrote past the selected count")


def test_filter_then_select_pipeline_u32(fl, oracle):
    """filter -> exclusive scan of the counts (torch.cumsum, plumbing) -> select == values[lo <= values <= hi]."""
    import torch

    rng = np.random.default_rng(77)
    n, w, ref = 4099, 13, 1000
    packed = rand_bytes(rng, n * 128 * w, 32)
    d_packed = to_dev(packed)
    lo, hi = 2000, 4000
    bitmap = torch.empty(n * 128, dtype=torch.uint8, device="cuda")
    counts = torch.empty(n, dtype=torch.int32, device="cuda")
    fl.Scan.filter_range(w, d_packed, ref, lo, hi, bitmap, counts)
    c64 = counts.to(torch.int64)
    offsets = torch.cumsum(c64, 0) - c64
    total = int(c64.sum().item())
    out = dev_empty(max(total, 1), 32)
    fl.Scan.select(w, d_packed, ref, bitmap, offsets, out)
    values = oracle.unfor_pack(packed, ref, w, n_blocks=n)
    want = values[(values >= lo) & (values <= hi)]
    assert total == want.size
    assert np.array_equal(to_host(out, 32)[:total], want)


@pytest.mark.parametrize("tb", [8, 16, 32, 64])
def test_filter_host_path(fl, oracle, tb):
    rng = np.random.default_rng(1100 + tb)
    n = 300
    fl.host_configure(chunk_blocks=64, n_streams=3)  # 5 chunks over 3 slots, last one ragged
    try:
        for w in (0, 3, tb // 2 + 1, tb):
            packed = rand_bytes(rng, n * 128 * w, tb)
            ref = int(rng.integers(0, 1 << min(tb, 62)))
            values = oracle.unfor_pack(packed, ref, w, n_blocks=n)
            lo, hi = int(mask(tb) // 5), int(mask(tb) // 5 * 3)
            want, sel = expected_bitmap(values, lo, hi)
            bitmap = np.zeros(n * 128, dtype=np.uint8)
            counts = np.zeros(n, dtype=np.uint32)
            fl.Scan.filter_range(w, packed, ref, lo, hi, bitmap, counts)
            assert np.array_equal(bitmap, want), (tb, w)
            assert np.array_equal(counts, sel.reshape(n, 1024).sum(1).astype(np.uint32))
            # the same call with every buffer page-locked: one launch on the caller's memory (direct path), same bytes
            p_packed = fl.pinned_empty(max(1, packed.size), DT[tb])[: packed.size]; p_packed[:] = packed
            p_bitmap = fl.pinned_empty(n * 128, np.uint8); p_bitmap[:] = 0
            p_counts = fl.pinned_empty(n, np.uint32); p_counts[:] = 0
            fl.Scan.filter_range(w, p_packed, ref, lo, hi, p_bitmap, p_counts)
            assert np.array_equal(p_bitmap, want), (tb, w, "direct")
            assert np.array_equal(p_counts, counts), (tb, w, "direct")
    finally:
        fl.host_configure(0, 0)


def test_scan_errors(fl):
    import torch

    p = torch.zeros(32 * 33, dtype=torch.int32, device="cuda")
    b = torch.zeros(128, dtype=torch.uint8, device="cuda")
    with pytest.raises(fl.FastLanesError) as e:
        fl.Scan.filter_range(33, p, 0, 0, 1, b)
    assert e.value.status == 1  # FL_ERR_WIDTH
    with pytest.raises(fl.FastLanesError):
        fl.Scan.filter_range(8, p[: 32 * 8], 0, 0, 1, b[:100])
    # empty batch is a no-op
    fl.Scan.filter_range(8, p[:0], 0, 0, 1, b[:0])


@pytest.mark.parametrize("tb", [8, 16, 32, 64])
def test_for_pack_auto_every_width(fl, oracle, tb):
    """Fused statistics + for_pack: refs == per-block minima, spans == max - min, packed ==
    oracle for_pack with those references (src/ffor.rs:24-36); blocks whose span fits W bits round-trip exactly."""
    rng = np.random.default_rng(1200 + tb)
    n = N_BLOCKS
    for w in range(tb + 1):
        # per-block windows [base, base + 2^w') with w' = w (lossless) or wider (truncating), incl. wrap-free extremes
        base = rng.integers(0, 1 << (tb - 1), size=n, dtype=np.uint64)
        wide = rng.random(n) < 0.3
        width_of_block = np.where(wide, min(tb, w + 3), w)
        span_cap = np.array([(1 << int(x)) - 1 for x in width_of_block], dtype=np.uint64)
        vals = (base[:, None] + (rng.integers(0, 1 << 62, size=(n, 1024), dtype=np.uint64) & span_cap[:, None])) & np.uint64(mask(tb))
        values = vals.astype(DT[tb]).reshape(-1)
        refs = dev_empty(n, tb)
        spans = dev_empty(n, tb)
        packed = dev_empty(n * 1024 * w // tb, tb)
        fl.FoR.for_pack_auto(w, to_dev(values), refs, packed, spans)
        v2 = values.reshape(n, 1024)
        assert np.array_equal(to_host(refs, tb), v2.min(1)), (tb, w, "refs")
        assert np.array_equal(to_host(spans, tb), v2.max(1) - v2.min(1)), (tb, w, "spans")
        assert np.array_equal(to_host(packed, tb), oracle.for_pack(values, v2.min(1), w)), (tb, w, "packed")
        out = dev_empty(n * 1024, tb)
        fl.FoR.unfor_pack(w, packed, refs, out)
        fits = (v2.max(1) - v2.min(1)).astype(np.uint64) <= np.uint64(mask(w))
        got = to_host(out, tb).reshape(n, 1024)
        assert np.array_equal(got[fits], v2[fits]), (tb, w, "round trip of the lossless blocks")
        if w in (0, 5, tb):  # spans are optional
            refs2 = dev_empty(n, tb)
            packed2 = dev_empty(n * 1024 * w // tb, tb)
            fl.FoR.for_pack_auto(w, to_dev(values), refs2, packed2)
            assert np.array_equal(to_host(refs2, tb), v2.min(1)) and np.array_equal(to_host(packed2, tb), to_host(packed, tb))


@pytest.mark.parametrize("tb", [8, 16, 32, 64])
def test_delta_filter_every_width(fl, oracle, tb):
    """Delta scan: bitmap of lo <= untranspose(undelta_pack(packed, base)) <= hi in ORIGINAL order (src/delta.rs:48-63,
    src/transpose.rs:18-22), device and host families, against the oracle composition."""
    import torch

    rng = np.random.default_rng(1300 + tb)
    n = N_BLOCKS
    full = mask(tb)
    for w in range(tb + 1):
        packed = rand_bytes(rng, n * 128 * w, tb)
        base = rand_bytes(rng, n * 128, tb)
        values = oracle.untranspose(oracle.undelta_pack(packed, base, w, n_blocks=n))
        v0 = int(values[rng.integers(0, values.size)])
        for lo, hi in ((full // 4, full // 4 * 3), (v0, v0), (9, 3), (0, full), (full - full // 7, full)):
            want, sel = expected_bitmap(values, lo, hi)
            bitmap = torch.full((n * 128,), 0x5A, dtype=torch.uint8, device="cuda")
            counts = torch.full((n,), -1, dtype=torch.int32, device="cuda")
            fl.Scan.filter_range_delta(w, to_dev(packed), to_dev(base), lo, hi, bitmap, counts)
            assert np.array_equal(bitmap.cpu().numpy(), want), (tb, w, lo, hi)
            assert np.array_equal(counts.cpu().numpy().view(np.uint32), sel.reshape(n, 1024).sum(1).astype(np.uint32)), (tb, w)
        if w in (0, 1, tb // 2, tb):  # host family
            h_bitmap = np.zeros(n * 128, dtype=np.uint8)
            h_counts = np.zeros(n, dtype=np.uint32)
            fl.Scan.filter_range_delta(w, packed, base, full // 3, full // 3 * 2, h_bitmap, h_counts)
            want, sel = expected_bitmap(values, full // 3, full // 3 * 2)
            assert np.array_equal(h_bitmap, want), (tb, w, "host")
            assert np.array_equal(h_counts, sel.reshape(n, 1024).sum(1).astype(np.uint32))


def test_delta_filter_sorted_column_u64(fl, oracle):
    """The use case: a sorted u64 column (timestamps) delta-encoded block by block with the reference's own chain
    transpose -> delta -> pack (src/delta.rs:88-95, here the fused encoder), scanned for a time range."""
    import torch

    rng = np.random.default_rng(99)
    n, w = 513, 20
    steps = rng.integers(0, 1 << w, size=n * 1024, dtype=np.uint64)
    ts = np.uint64(1_700_000_000_000) + np.cumsum(steps, dtype=np.uint64)
    # base[lane] = the value preceding the lane's run: u64 lane l walks originals 64*l .. 64*l+63, so
    # its base is the column value just before original 64*l (the running predecessor across block boundaries)
    prev = np.concatenate([[ts[0] - steps[0]], ts[:-1]]).reshape(n, 1024)
    base = prev[:, ::64].reshape(-1).copy()
    d_packed = dev_empty(n * 16 * w, 64)
    fl.Delta.transpose_delta_pack(w, to_dev(ts), to_dev(base), d_packed)
    back = dev_empty(n * 1024, 64)
    fl.Delta.undelta_pack_untranspose(w, d_packed, to_dev(base), back)
    assert np.array_equal(to_host(back, 64), ts), "encode/decode chain does not round-trip"
    lo, hi = int(ts[n * 300]), int(ts[n * 700])
    bitmap = torch.empty(n * 128, dtype=torch.uint8, device="cuda")
    counts = torch.empty(n, dtype=torch.int32, device="cuda")
    fl.Scan.filter_range_delta(w, d_packed, to_dev(base), lo, hi, bitmap, counts)
    want, sel = expected_bitmap(ts, lo, hi)
    assert np.array_equal(bitmap.cpu().numpy(), want)
    assert int(counts.sum().item()) == int(sel.sum())


@pytest.mark.parametrize("tb", [8, 16, 32, 64])
def test_cwida_row_order_every_width(fl, tb):
    """SURVEY.md §8f rank 4 (PARITY UNPINNED, see oracle/cwida.py): pack / unpack / for_pack / unfor_pack in the linear row
    order of the original FastLanes layout, against the closed-form cwida oracle."""
    from oracle import cwida

    rng = np.random.default_rng(1400 + tb)
    n = 5
    for w in range(tb + 1):
        values = rand_bytes(rng, n * 128 * tb, tb)
        ref = int(rand_bytes(rng, tb // 8, tb)[0])
        p = dev_empty(n * 1024 * w // tb, tb)
        fl.Cwida.pack(w, to_dev(values), p)
        want_p = cwida.pack(values, w)
        assert np.array_equal(to_host(p, tb), want_p), (tb, w, "pack")
        out = dev_empty(n * 1024, tb)
        fl.Cwida.unpack(w, p, out)
        assert np.array_equal(to_host(out, tb), cwida.unpack(want_p, w, n)), (tb, w, "unpack")
        fl.Cwida.for_pack(w, to_dev(values), ref, p)
        want_fp = cwida.for_pack(values, ref, w)
        assert np.array_equal(to_host(p, tb), want_fp), (tb, w, "for_pack")
        fl.Cwida.unfor_pack(w, p, ref, out)
        assert np.array_equal(to_host(out, tb), cwida.unfor_pack(want_fp, ref, w, n)), (tb, w, "unfor_pack")


def test_example_column_scan():
    """examples/column_scan.py: encode (fused delta chain, fused statistics + FoR) -> two fused scans -> select -> numpy."""
    import importlib.util
    import os

    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "column_scan.py")
    spec = importlib.util.spec_from_file_location("column_scan", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.main(257) == 0
