"""Known-answer and cross-implementation tests of the CPU oracle (not gpu).

* SURVEY.md Appendix B KATs: digests produced by an INDEPENDENT restatement written during the
  survey (a different session, since discarded).  They are not crate outputs; agreement of two
  independent restatements plus the reference's own tests is the strongest pin available without
  a Rust toolchain.
* The numpy closed-form oracle (oracle/np_closed_form.py, written from unpack_single's arithmetic)
  must agree with the streaming C++ oracle (written from pack!/unpack!) on random data, for every
  type and width, at every x86-64 ISA level the host supports.
* Gap-closing cases the reference's tests never exercise (SURVEY.md §4 gaps 2-3): W==T with
  non-zero data, values with bits above W (mask truncation), untranspose, unfor_pack, Delta for
  every type.
"""
import hashlib

import numpy as np
import pytest

from conftest import splitmix64
from oracle import np_closed_form as cf

DT = {8: np.uint8, 16: np.uint16, 32: np.uint32, 64: np.uint64}


def sha16(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).astype(a.dtype.newbyteorder("<")).tobytes()).hexdigest()[:16]


def mask(w):
    return (1 << w) - 1


# ---- Appendix B: closed inputs from the reference's tests/docs ---------------------------------

def test_kat_readme_u16_w3(oracle):
    values = (np.arange(1024) % 8).astype(np.uint16)
    p = oracle.pack(values, 3)
    assert p.size == 192
    assert [int(x) for x in p[:8]] == [0x0000, 0x9249, 0x2492, 0xB6DB, 0x4924, 0xDB6D, 0x6DB6, 0xFFFF]
    assert p[64] == 0 and p[128] == 0
    assert sha16(p) == "f949547d2b920f40"


def test_kat_u32_iota_w10(oracle):
    p = oracle.pack(np.arange(1024, dtype=np.uint32), 10)
    assert p.size == 320
    assert [int(x) for x in p[:4]] == [0x10020000, 0x50120401, 0x90220802, 0xD0320C03]
    assert [int(x) for x in p[32:36]] == [0x0A020060, 0x1A060160, 0x2A0A0260, 0x3A0E0360]
    assert sha16(p) == "fded69a758643dbc"


def test_kat_u32_iota_w16(oracle):
    p = oracle.pack(np.arange(1024, dtype=np.uint32), 16)
    assert [int(x) for x in p[:4]] == [0x00800000, 0x00810001, 0x00820002, 0x00830003]
    assert sha16(p) == "50608d099e6729ba"


def test_kat_transpose_iota(oracle):
    t = oracle.transpose(np.arange(1024, dtype=np.uint16))
    assert [int(x) for x in t[:4]] == [0, 64, 128, 192]
    assert int(t[15]) == 960 and [int(x) for x in t[16:20]] == [32, 96, 160, 224]
    assert sha16(t) == "6eaa6b0bd018e4a0"


# ---- Appendix B: seeded inputs -------------------------------------------------------------------

SEEDED = [
    (8, 1, 128, "035e447272a8405d", "667f4522676eddfe", 0x8E),
    (8, 3, 384, "dd71847ceb305e83", "51fc72990136c1e1", 0x68),
    (8, 7, 896, "a62c3a94cdf0a29f", "8865d844c80590ab", 0xE8),
    (8, 8, 1024, "7ff42523246338f6", "7ff42523246338f6", 0xE8),
    (16, 3, 192, "0eb6c0ab8ca78ea1", "7cdd5ea4fe526389", 0x2B68),
    (16, 9, 576, "700136a2770e1204", "72306ad03bff2934", 0xBAE8),
    (16, 15, 960, "e3e34e1be3d41ef0", "5f62d716a702abaa", 0x8CE8),
    (16, 16, 1024, "66df14a74724916d", "67206438fc852069", 0x8CE8),
    (32, 1, 32, "ab3555c37ca735d4", "2f142ce19bd3a208", 0xB463688E),
    (32, 8, 256, "6552b4c23045e749", "f3dce22e24515cbc", 0x7DBD5DE8),
    (32, 10, 320, "3476560f38bf166a", "1e079582293c592b", 0x7BD574E8),
    (32, 16, 512, "ccaac2703328e3bd", "58f96bb5a0831a26", 0xBD5D8CE8),
    (32, 31, 992, "748440de38ded265", "e221885852122a54", 0xA25F8CE8),
    (32, 32, 1024, "d01e596f1291a913", "c2ead754f334a735", 0x225F8CE8),
    (64, 1, 16, "cf1fc70bab008a4e", "8d0c63fa5f91f720", 0xD01355C9B463688E),
    (64, 17, 272, "39910dbddcf01b45", "4fee59ce85ea59a2", 0x93E8AEF57ABB8CE8),
    (64, 33, 528, "56008cc68339b563", "dbbfcbb7ef9f3841", 0x46D97ABA225F8CE8),
    (64, 48, 768, "cd5ee1251fb242a6", "97acccdf3aa27073", 0xBD5DD1BC225F8CE8),
    (64, 64, 1024, "e810246c70cebd95", "cafb9c55d201a7b4", 0x8B1FD1BC225F8CE8),
]


def seeded_values(tb, w, seed=42):
    v = splitmix64(np.uint64(seed * 1024) + np.arange(1024, dtype=np.uint64))
    if w < 64:
        v = v & np.uint64(mask(w))
    return v.astype(DT[tb])


@pytest.mark.parametrize("tb,w,plen,sha_v,sha_p,p0", SEEDED, ids=[f"u{r[0]}_{r[1]}" for r in SEEDED])
def test_kat_seeded(oracle, tb, w, plen, sha_v, sha_p, p0):
    v = seeded_values(tb, w)
    assert sha16(v) == sha_v
    p = oracle.pack(v, w)
    assert p.size == plen
    assert int(p[0]) == p0
    assert sha16(p) == sha_p
    assert np.array_equal(oracle.unpack(p, w), v)


def test_kat_fused_delta_u32_w8(oracle):
    deltas = (splitmix64(np.uint64(7 * 1024) + np.arange(1024, dtype=np.uint64)) & np.uint64(0xFF)).astype(np.uint32)
    base = (splitmix64(np.uint64(99) + np.arange(32, dtype=np.uint64)) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    packed = oracle.pack(deltas, 8)
    assert sha16(packed) == "81af657871912b2c"
    assert sha16(base) == "573aebde8e6a3508"
    out = oracle.undelta_pack(packed, base, 8)
    assert sha16(out) == "0c24fa62122d5566"
    assert np.array_equal(out, oracle.undelta(oracle.unpack(packed, 8), base))


def test_kat_truncation(oracle):
    # pack masks inputs to W bits (src/macros.rs:73): all-ones u32 at W=5 unpacks to all-31
    v = np.full(1024, 0xFFFFFFFF, dtype=np.uint32)
    assert np.all(oracle.unpack(oracle.pack(v, 5), 5) == 31)


# ---- streaming C++ oracle == closed-form numpy oracle, all types × widths × ISA levels ----------

ALL_TW = [(tb, w) for tb in (8, 16, 32, 64) for w in range(tb + 1)]


@pytest.mark.parametrize("level", [2, 3, 4])
def test_streaming_equals_closed_form_every_width(oracle, level):
    got = oracle.set_isa_level(level)
    try:
        if got != level:
            pytest.skip(f"host supports only x86-64-v{got}")
        rng = np.random.default_rng(1234 + level)
        for tb, w in ALL_TW:
            n = 3
            raw = rng.integers(0, 1 << 63, size=n * 1024, dtype=np.uint64) * 2 + rng.integers(0, 2, size=n * 1024, dtype=np.uint64)
            values = raw.astype(DT[tb])  # full-range bits: exercises the & mask truncation
            p = oracle.pack(values, w)
            assert np.array_equal(p, cf.pack(values, w)), (tb, w, "pack")
            expect = values if w == tb else (values & DT[tb](mask(w)))
            u = oracle.unpack(p, w, n_blocks=n)
            assert np.array_equal(u, expect), (tb, w, "unpack")
            assert np.array_equal(cf.unpack(p, w, n_blocks=n), expect), (tb, w, "cf.unpack")
            # any bit pattern is a valid packing: unpack random bytes both ways
            if w:
                rp = rng.integers(0, 256, size=n * 128 * w, dtype=np.uint8).view(DT[tb])
                assert np.array_equal(oracle.unpack(rp, w), cf.unpack(rp, w)), (tb, w, "unpack-random")
                assert np.array_equal(oracle.pack(oracle.unpack(rp, w), w), rp), (tb, w, "repack")
    finally:
        oracle.set_isa_level(4)


@pytest.mark.parametrize("tb", [8, 16, 32, 64])
def test_transpose_delta_for_every_type(oracle, tb):
    rng = np.random.default_rng(tb)
    n = 4
    L = 1024 // tb
    values = rng.integers(0, 256, size=n * 1024 * (tb // 8), dtype=np.uint8).view(DT[tb])
    base = rng.integers(0, 256, size=n * L * (tb // 8), dtype=np.uint8).view(DT[tb])
    t = oracle.transpose(values)
    assert np.array_equal(t, cf.transpose(values))
    assert np.array_equal(oracle.untranspose(t), values)
    assert np.array_equal(cf.untranspose(t), values)
    d = oracle.delta(t, base)
    assert np.array_equal(d, cf.delta(t, base))
    assert np.array_equal(oracle.undelta(d, base), t)
    assert np.array_equal(cf.undelta(d, base), t)
    for w in sorted({0, 1, tb // 2 - 1, tb // 2, tb - 1, tb}):
        p = oracle.pack(d, w)
        dm = d if w == tb else d & DT[tb](mask(w))
        assert np.array_equal(oracle.undelta_pack(p, base, w, n_blocks=n), cf.undelta(dm, base)), (tb, w)
        ref = DT[tb](rng.integers(0, 1 << min(tb, 62)))
        refs = rng.integers(0, 256, size=n * (tb // 8), dtype=np.uint8).view(DT[tb])
        for r in (ref, refs):
            rr = np.repeat(r, 1024) if np.ndim(r) else r
            fp = oracle.for_pack(values, r, w)
            assert np.array_equal(fp, cf.pack((values - rr).astype(DT[tb]), w)), (tb, w, "for_pack")
            got = oracle.unfor_pack(fp, r, w, n_blocks=n)
            diff = (values - rr).astype(DT[tb])
            dmask = diff if w == tb else diff & DT[tb](mask(w))
            assert np.array_equal(got, (dmask + rr).astype(DT[tb])), (tb, w, "unfor_pack")


def test_lane_runs_are_consecutive_originals(oracle):
    # in the transposed vector lane l walks consecutive originals → per-lane delta
    # equals an ordinary delta inside runs of T values.
    for tb in (8, 16, 32, 64):
        idx = cf.index_table(tb)
        t = cf.transpose_table()
        orig = t[idx]  # [row, lane] original position
        assert np.all(np.diff(orig, axis=0) == 1)


def test_gather_matches_unpack(oracle):
    rng = np.random.default_rng(5)
    for tb in (8, 16, 32, 64):
        for w in (0, 1, 5, tb - 1, tb):
            n = 5
            rp = rng.integers(0, 256, size=n * 128 * w, dtype=np.uint8).view(DT[tb]) if w else np.zeros(0, DT[tb])
            full = oracle.unpack(rp, w, n_blocks=n)
            gi = rng.integers(0, n * 1024, size=300, dtype=np.uint64)
            assert np.array_equal(oracle.unpack_gather(rp, w, gi), full[gi.astype(np.int64)])


def test_threads_agree(oracle):
    rng = np.random.default_rng(9)
    rp = rng.integers(0, 1 << 32, size=257 * 32 * 13, dtype=np.uint32)
    assert np.array_equal(oracle.unpack(rp, 13, threads=1), oracle.unpack(rp, 13, threads=5))
