"""The reference crate's own unit tests, restated against the CPU oracle (not gpu).

Each test cites the reference test it restates (paths relative to /root/reference).  These are the
assertions that PIN the oracle: the reference ships no golden bytes, but `unpack_single` is an
independent closed-form reader, so pack/unpack_single agreement for every (T, W, i) fixes the wire
format.
"""
import numpy as np
import pytest

from oracle import np_closed_form as cf

DT = {8: np.uint8, 16: np.uint16, 32: np.uint32, 64: np.uint64}
ALL_TW = [(tb, w) for tb in (8, 16, 32, 64) for w in range(tb + 1)]


def test_fl_order_is_own_inverse():
    # src/lib.rs:53-59
    for i in range(8):
        assert cf.FL_ORDER[cf.FL_ORDER[i]] == i


@pytest.mark.parametrize("tb,w", ALL_TW, ids=[f"u{tb}_{w}" for tb, w in ALL_TW])
def test_round_trip(oracle, tb, w):
    # src/bitpacking.rs:273-315 — 124 generated tests: values[i] = i % (1 << (W % T))
    values = (np.arange(1024, dtype=np.uint64) % np.uint64(1 << (w % tb))).astype(DT[tb])
    packed = oracle.pack(values, w)
    assert packed.size == 1024 * w // tb
    unpacked = oracle.unpack(packed, w, n_blocks=1)
    assert np.array_equal(unpacked, values)
    # unpack_single::<W> and unchecked_unpack_single(W) agree for every index (:292-298)
    singles = np.array([oracle.unpack_single(packed, w, i) for i in range(1024)], dtype=np.uint64)
    assert np.array_equal(singles, values.astype(np.uint64))


def test_unchecked_pack(oracle):
    # src/bitpacking.rs:249-256 — u32 iota, W=10
    values = np.arange(1024, dtype=np.uint32)
    packed = oracle.pack(values, 10)
    assert packed.size == 320
    assert np.array_equal(oracle.unpack(packed, 10), values)


def test_unpack_single(oracle):
    # src/bitpacking.rs:259-271 — u32 iota, W=16
    values = np.arange(1024, dtype=np.uint32)
    packed = oracle.pack(values, 16)
    assert packed.size == 512
    for i in range(1024):
        assert oracle.unpack_single(packed, 16, i) == values[i]


def test_macros_test_pack(oracle):
    # src/macros.rs:181-207 — u16 W=15, values[i] = i % (1<<15)
    values = (np.arange(1024) % (1 << 15)).astype(np.uint16)
    packed = oracle.pack(values, 15)
    assert packed.size == 960
    assert np.array_equal(oracle.unpack(packed, 15), values)


def test_delta(oracle):
    # src/delta.rs:81-107 — u16, values[i] = i/8 → transpose → delta(base 0) → pack 15
    W = 15
    values = (np.arange(1024) // 8).astype(np.uint16)
    transposed = oracle.transpose(values)
    base = np.zeros(64, dtype=np.uint16)
    deltas = oracle.delta(transposed, base)
    packed = oracle.pack(deltas, W)
    # fused kernel (:98-100)
    assert np.array_equal(oracle.undelta_pack(packed, base, W), transposed)
    # unfused (:103-106)
    unpacked = oracle.unpack(packed, W)
    assert np.array_equal(oracle.undelta(unpacked, base), transposed)
    # max delta after transpose+delta = 126 (fits the bench's W=9, benches/delta.rs:11)
    assert int(deltas.max()) == 126
    packed9 = oracle.pack(deltas, 9)
    assert np.array_equal(oracle.undelta_pack(packed9, base, 9), transposed)


def test_ffor(oracle):
    # src/ffor.rs:67-88 — u16 W=15 reference=10
    W = 15
    values = (np.arange(1024) % (1 << W)).astype(np.uint16)
    packed = oracle.for_pack(values, 10, W)
    unpacked = oracle.unpack(packed, W)
    expect = ((values.astype(np.int64) - 10) & ((1 << W) - 1)).astype(np.uint16)
    assert np.array_equal(unpacked, expect)


def test_readme_example(oracle):
    # README.md:14-47 / src/lib.rs:71-96 — BASELINE config 1: u16 W=3 values[i] = i % 8
    W = 3
    values = (np.arange(1024) % (1 << W)).astype(np.uint16)
    packed = oracle.pack(values, W)
    assert packed.size == 128 * W // 2
    assert np.array_equal(oracle.unpack(packed, W), values)
    for i in range(1024):
        assert oracle.unpack_single(packed, W, i) == values[i]


def test_error_behaviour(oracle):
    # width > T → unreachable!() (bitpacking.rs:93,126,197); index >= 1024 → assert! (:152)
    v = np.zeros(1024, dtype=np.uint16)
    with pytest.raises(oracle.OracleError) as e:
        oracle.pack(v, 17)
    assert e.value.code == oracle.FLO_ERR_WIDTH
    with pytest.raises(oracle.OracleError) as e:
        oracle.unpack_single(np.zeros(192, dtype=np.uint16), 3, 1024)
    assert e.value.code == oracle.FLO_ERR_INDEX
