"""not gpu: the cwida-order oracle (oracle/cwida.py, PARITY UNPINNED — see its header) is at least self-consistent and
consistent with the pinned oracle: closed form == reference pack of the row-permuted input, unpack inverts pack, and
the two layouts really differ (so the test would notice if the kernels ignored the flag)."""
import numpy as np
import pytest

from oracle import cwida

DT = {8: np.uint8, 16: np.uint16, 32: np.uint32, 64: np.uint64}


@pytest.mark.parametrize("tb", [8, 16, 32, 64])
def test_cwida_closed_form_vs_permuted_reference_pack(oracle, tb):
    rng = np.random.default_rng(700 + tb)
    n = 3
    values = rng.integers(0, 256, size=n * 128 * tb, dtype=np.uint8).view(DT[tb])
    for w in sorted({0, 1, 3, tb // 2, tb - 1, tb}):
        p = cwida.pack(values, w)
        assert p.size == n * 1024 * w // tb
        assert np.array_equal(p, oracle.pack(cwida.to_reference_order(values), w)), (tb, w)
        back = cwida.unpack(p, w, n)
        want = values if w == tb else values & DT[tb]((1 << w) - 1)
        assert np.array_equal(back, want), (tb, w)
        if 0 < w and tb > 8:
            assert not np.array_equal(p, oracle.pack(values, w)), "the two layouts should differ for T > 8"
    # u8: FL_ORDER[0] = 0 and T = 8 rows only -> the reference's order IS linear: identical bytes
    if tb == 8:
        assert np.array_equal(cwida.pack(values, 5), oracle.pack(values, 5))


def test_cwida_for_roundtrip():
    rng = np.random.default_rng(5)
    values = (rng.integers(0, 1 << 12, size=2048, dtype=np.uint64) + 1000).astype(np.uint32)
    p = cwida.for_pack(values, 1000, 12)
    assert np.array_equal(cwida.unfor_pack(p, 1000, 12, 2), values)
