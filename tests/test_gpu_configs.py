"""-m gpu: BASELINE.json's configs at their FULL sizes, through size-independent properties plus oracle
samples (the oracle cannot process 2^30 values per width in test time, so full-buffer identities run on
the device and oracle comparisons run on sampled blocks).

  configs[1]  u32 unpack, W = 1..32, 2^20 blocks: pack(unpack(bits)) == bits on the whole buffer (every bit
              pattern is a valid packing, and pack is the independent inverse), + 256 sampled blocks vs oracle
  configs[2]  u64 pack+unpack, W in {1,17,33,48,64}, 2^16 blocks: GPU pack bytes == oracle pack bytes (full),
              GPU unpack == input (full)
  configs[3]  fused Delta+BitPack decode u32 W=8, 2^20 blocks: fused == unfused on the device (full),
              sampled blocks == oracle undelta_pack (src/delta.rs:48-63)
  configs[4]  batched u32 W=16 (the multi-GPU shard shape): one 2^22-block wave round-trips
"""
import numpy as np
import pytest

from conftest import splitmix64
from gpu_util import dev_empty, to_dev, to_host

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fl():
    import fastlanes_b200

    return fastlanes_b200


def device_random_i32(n, seed):
    import torch

    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    t = torch.empty(n, dtype=torch.int32, device="cuda")
    step = 1 << 26
    for i in range(0, n, step):
        t[i:i + step].random_(-(1 << 31), (1 << 31) - 1, generator=g)
    return t


def test_config2_u32_width_sweep_full_size(fl, oracle):
    import torch

    n = 1 << 20
    bits = device_random_i32(n * 32 * 32, 42)  # sized for W = 32; width W uses the first n*32*W words
    out = torch.empty(n * 1024, dtype=torch.int32, device="cuda")
    back = torch.empty(n * 32 * 32, dtype=torch.int32, device="cuda")
    rng = np.random.default_rng(2)
    sample = np.unique(np.concatenate([[0, 1, n - 2, n - 1], rng.integers(0, n, size=252)]))
    for w in range(1, 33):
        p = bits[: n * 32 * w]
        fl.BitPacking.unpack(w, p, out)
        if w < 32:  # no value may exceed W bits
            assert int(out.view(torch.int64).bitwise_and(~((((1 << w) - 1) << 32) | ((1 << w) - 1))).count_nonzero()) == 0, w
        b = back[: n * 32 * w]
        fl.BitPacking.pack(w, out, b)
        assert torch.equal(b, p), f"pack(unpack(x)) != x at W={w}"
        # sampled blocks against the oracle
        pv = p.view(n, 32 * w)[torch.from_numpy(sample).cuda()].cpu().numpy().view(np.uint32).reshape(-1)
        ov = out.view(n, 1024)[torch.from_numpy(sample).cuda()].cpu().numpy().view(np.uint32).reshape(-1)
        assert np.array_equal(ov, oracle.unpack(pv, w, threads=4)), w


def test_config3_u64_pack_unpack(fl, oracle):
    n = 1 << 16
    idx = np.arange(n * 1024, dtype=np.uint64)
    for w in (1, 17, 33, 48, 64):
        values = splitmix64(np.uint64(42) * np.uint64(1 << 40) + idx)
        if w < 64:
            values &= np.uint64((1 << w) - 1)
        d_values = to_dev(values)
        d_packed = dev_empty(n * 16 * w, 64)
        fl.BitPacking.pack(w, d_values, d_packed)
        assert np.array_equal(to_host(d_packed, 64), oracle.pack(values, w, threads=8)), f"u64 pack bytes differ at W={w}"
        d_out = dev_empty(n * 1024, 64)
        fl.BitPacking.unpack(w, d_packed, d_out)
        assert np.array_equal(to_host(d_out, 64), values), f"u64 unpack != input at W={w}"


def test_config4_fused_delta_u32_w8(fl, oracle):
    import torch

    n, w = 1 << 20, 8
    packed = device_random_i32(n * 32 * w, 7)          # deltas uniform in [0, 255]
    base = device_random_i32(n * 32, 99)               # uniform u32 bases, 32 per block
    fused = torch.empty(n * 1024, dtype=torch.int32, device="cuda")
    fl.Delta.undelta_pack(w, packed, base, fused)
    unpacked = torch.empty(n * 1024, dtype=torch.int32, device="cuda")
    fl.BitPacking.unpack(w, packed, unpacked)
    unfused = torch.empty(n * 1024, dtype=torch.int32, device="cuda")
    fl.Delta.undelta(unpacked, base, unfused)
    assert torch.equal(fused, unfused), "fused undelta_pack != unpack + undelta"
    # and delta() inverts it
    d = torch.empty(n * 1024, dtype=torch.int32, device="cuda")
    fl.Delta.delta(fused, base, d)
    assert torch.equal(d, unpacked)
    rng = np.random.default_rng(4)
    sample = torch.from_numpy(np.unique(np.concatenate([[0, n - 1], rng.integers(0, n, size=1022)]))).cuda()
    pv = packed.view(n, 32 * w)[sample].cpu().numpy().view(np.uint32).reshape(-1)
    bv = base.view(n, 32)[sample].cpu().numpy().view(np.uint32).reshape(-1)
    fv = fused.view(n, 1024)[sample].cpu().numpy().view(np.uint32).reshape(-1)
    assert np.array_equal(fv, oracle.undelta_pack(pv, bv, w, threads=4))


def test_config5_shard_wave_u32_w16(fl):
    import torch

    n, w = 1 << 22, 16  # one wave of the 64M-block sharded workload: 8 GiB packed, 16 GiB unpacked
    packed = device_random_i32(n * 32 * w, 5)
    out = torch.empty(n * 1024, dtype=torch.int32, device="cuda")
    fl.BitPacking.unpack(w, packed, out)
    back = torch.empty_like(packed)
    fl.BitPacking.pack(w, out, back)
    assert torch.equal(back, packed)
    # W = 16: each u32 lane word holds rows 2k (low half) and 2k+1 (high half); check against torch ops
    words = packed.view(n, 16, 32)[:4].cpu().numpy().view(np.uint32)
    got = out.view(n, 1024)[:4].cpu().numpy().view(np.uint32)
    order = [0, 4, 2, 6, 1, 5, 3, 7]
    for row in range(32):
        expect = (words[:, row // 2, :] >> (16 * (row % 2))) & 0xFFFF
        start = order[row // 8] * 16 + (row % 8) * 128
        assert np.array_equal(got[:, start:start + 32], expect)


def test_scan_full_size_u32(fl):
    """Fused scan at configs[1] size (2^20 blocks, a few widths): bitmap popcounts == counts == the number of values of
    the materialised unpack that satisfy the predicate (computed with torch on the device), per block; select returns
    exactly those values in order."""
    import torch

    n = 1 << 20
    bits = device_random_i32(n * 32 * 32, 42)
    out = torch.empty(n * 1024, dtype=torch.int32, device="cuda")
    bitmap = torch.empty(n * 128, dtype=torch.uint8, device="cuda")
    counts = torch.empty(n, dtype=torch.int32, device="cuda")
    for w, ref in ((3, 0), (13, 1000), (24, 0x7FFFFF00), (32, 12345)):
        p = bits[: n * 32 * w]
        fl.FoR.unfor_pack(w, p, ref, out)
        lo = (ref + ((1 << w) - 1) // 5) & 0xFFFFFFFF
        hi = (lo + ((1 << w) - 1) // 3) & 0xFFFFFFFF
        fl.Scan.filter_range(w, p, ref, lo, hi, bitmap, counts)
        v = out.to(torch.int64) & 0xFFFFFFFF
        sel = (v >= lo) & (v <= hi) if lo <= hi else torch.zeros_like(v, dtype=torch.bool)
        want = sel.view(n, 1024).sum(1).to(torch.int32)
        assert torch.equal(counts, want), (w, "counts")
        # bitmap bit i of block b <-> sel[b, i] (little-endian bit order), checked on the whole buffer
        weights = (1 << torch.arange(8, device="cuda", dtype=torch.int32))
        packed_bits = (sel.view(-1, 8).to(torch.int32) * weights).sum(1).to(torch.uint8)
        assert torch.equal(bitmap, packed_bits), (w, "bitmap")
        c64 = counts.to(torch.int64)
        offsets = torch.cumsum(c64, 0) - c64
        total = int(c64.sum().item())
        dense = torch.empty(max(total, 1), dtype=torch.int32, device="cuda")
        fl.Scan.select(w, p, ref, bitmap, offsets, dense)
        assert torch.equal(dense[:total], out[sel.view(-1)]), (w, "select")
        del v, sel, packed_bits, dense
