"""Committed golden vectors (tests/golden/fastlanes_golden.npz; provenance in make_golden.py):
the oracle must still reproduce them (not gpu) and the CUDA kernels must reproduce them (-m gpu)."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fastlanes_golden.npz")
DT = {8: np.uint8, 16: np.uint16, 32: np.uint32, 64: np.uint64}


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(GOLD))


def cases(gold):
    for key in sorted(gold):
        if key.endswith("_packed") and "for_packed" not in key:
            tb, w = key.split("_")[0], key.split("_")[1]
            yield int(tb[1:]), int(w[1:])


def mask_to(values, w, tb):
    return values if w == tb else values & DT[tb]((1 << w) - 1)


def test_oracle_reproduces_golden(oracle, gold):
    n_cases = 0
    for tb, w in cases(gold):
        v, base, ref = gold[f"u{tb}_values"], gold[f"u{tb}_base"], int(gold[f"u{tb}_reference"][0])
        p = gold[f"u{tb}_w{w}_packed"]
        assert np.array_equal(oracle.pack(v, w), p)
        assert np.array_equal(oracle.unpack(p, w, n_blocks=2), mask_to(v, w, tb))
        assert np.array_equal(oracle.for_pack(v, ref, w), gold[f"u{tb}_w{w}_for_packed"])
        assert np.array_equal(oracle.undelta_pack(p, base, w, n_blocks=2), gold[f"u{tb}_w{w}_undelta_pack"])
        assert np.array_equal(oracle.unfor_pack(p, ref, w, n_blocks=2), gold[f"u{tb}_w{w}_unfor_pack"])
        n_cases += 1
    assert n_cases == 27
    for tb in (8, 16, 32, 64):
        assert np.array_equal(oracle.transpose(gold[f"u{tb}_values"]), gold[f"u{tb}_transposed"])
        assert np.array_equal(oracle.delta(gold[f"u{tb}_transposed"], gold[f"u{tb}_base"]), gold[f"u{tb}_delta_of_transposed"])


@pytest.mark.gpu
def test_gpu_reproduces_golden(gold):
    import fastlanes_b200 as fl

    for tb, w in cases(gold):
        v, base, ref = gold[f"u{tb}_values"], gold[f"u{tb}_base"], int(gold[f"u{tb}_reference"][0])
        p = gold[f"u{tb}_w{w}_packed"]
        got_p = np.zeros_like(p)
        fl.BitPacking.pack(w, v, got_p)
        assert np.array_equal(got_p, p), (tb, w, "pack")
        u = np.zeros_like(v)
        fl.BitPacking.unpack(w, p, u)
        assert np.array_equal(u, mask_to(v, w, tb)), (tb, w, "unpack")
        fp = np.zeros_like(p)
        fl.FoR.for_pack(w, v, ref, fp)
        assert np.array_equal(fp, gold[f"u{tb}_w{w}_for_packed"]), (tb, w, "for_pack")
        fl.Delta.undelta_pack(w, p, base, u)
        assert np.array_equal(u, gold[f"u{tb}_w{w}_undelta_pack"]), (tb, w, "undelta_pack")
        fl.FoR.unfor_pack(w, p, ref, u)
        assert np.array_equal(u, gold[f"u{tb}_w{w}_unfor_pack"]), (tb, w, "unfor_pack")
        # fused scans against the golden decoded arrays: FoR domain (index order) and Delta domain (original order)
        full = (1 << tb) - 1
        lo, hi = full // 4, full // 4 * 3
        bitmap = np.zeros(2 * 128, dtype=np.uint8)
        fl.Scan.filter_range(w, p, ref, lo, hi, bitmap)
        g = gold[f"u{tb}_w{w}_unfor_pack"]
        assert np.array_equal(bitmap, np.packbits((g >= DT[tb](lo)) & (g <= DT[tb](hi)), bitorder="little")), (tb, w, "filter")
        fl.Scan.filter_range_delta(w, p, base, lo, hi, bitmap)
        g = np.zeros_like(v)
        fl.Transpose.untranspose(gold[f"u{tb}_w{w}_undelta_pack"], g)
        assert np.array_equal(bitmap, np.packbits((g >= DT[tb](lo)) & (g <= DT[tb](hi)), bitorder="little")), (tb, w, "delta filter")
    for tb in (8, 16, 32, 64):
        v = gold[f"u{tb}_values"]
        t = np.zeros_like(v)
        fl.Transpose.transpose(v, t)
        assert np.array_equal(t, gold[f"u{tb}_transposed"])
        d = np.zeros_like(v)
        fl.Delta.delta(t, gold[f"u{tb}_base"], d)
        assert np.array_equal(d, gold[f"u{tb}_delta_of_transposed"])
        back = np.zeros_like(v)
        fl.Delta.undelta(d, gold[f"u{tb}_base"], back)
        fl.Transpose.untranspose(back, t)
        assert np.array_equal(t, v)
