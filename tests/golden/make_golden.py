"""Generates tests/golden/fastlanes_golden.npz — small committed input/output vectors for the hot path.

PROVENANCE (read this before trusting them): the reference crate cannot be built or run in this image
(Rust nightly-2024-06-19 required, no Rust toolchain), and it ships no golden bytes of its own.  These
vectors are therefore produced by the CPU ORACLE (oracle/fl_oracle_kernels.hpp, a line-by-line restatement
of the reference loops) AFTER that oracle passed (a) every assertion of the reference's own unit tests
restated in tests/test_oracle_reference_tests.py, (b) the digests of an independent restatement
(tests/test_oracle_kats.py) and (c) agreement with the closed-form numpy oracle.
They pin the wire format against drift of BOTH the oracle and the CUDA kernels; they are not crate outputs.

Run from the repository root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import splitmix64  # noqa: E402
from oracle import fl_oracle as oracle  # noqa: E402
from oracle import np_closed_form as cf  # noqa: E402

DT = {8: np.uint8, 16: np.uint16, 32: np.uint32, 64: np.uint64}
WIDTHS = {8: [0, 1, 3, 5, 7, 8], 16: [0, 1, 3, 9, 15, 16], 32: [0, 1, 8, 10, 16, 17, 31, 32], 64: [0, 1, 17, 33, 48, 63, 64]}
N_BLOCKS = 2


def seeded(tb, seed, n):
    return splitmix64(np.uint64(seed) * np.uint64(1 << 20) + np.arange(n, dtype=np.uint64)).astype(DT[tb])


def main():
    out = {}
    for tb in (8, 16, 32, 64):
        values = seeded(tb, 1000 + tb, N_BLOCKS * 1024)          # full-range bits (exercises truncation)
        base = seeded(tb, 2000 + tb, N_BLOCKS * (1024 // tb))
        ref = int(seeded(tb, 3000 + tb, 1)[0])
        out[f"u{tb}_values"] = values
        out[f"u{tb}_base"] = base
        out[f"u{tb}_reference"] = np.array([ref], dtype=DT[tb])
        t = oracle.transpose(values)
        assert np.array_equal(t, cf.transpose(values))
        out[f"u{tb}_transposed"] = t
        d = oracle.delta(t, base)
        assert np.array_equal(d, cf.delta(t, base))
        out[f"u{tb}_delta_of_transposed"] = d
        for w in WIDTHS[tb]:
            p = oracle.pack(values, w)
            assert np.array_equal(p, cf.pack(values, w))
            out[f"u{tb}_w{w}_packed"] = p
            out[f"u{tb}_w{w}_for_packed"] = oracle.for_pack(values, ref, w)
            out[f"u{tb}_w{w}_undelta_pack"] = oracle.undelta_pack(p, base, w, n_blocks=N_BLOCKS)
            out[f"u{tb}_w{w}_unfor_pack"] = oracle.unfor_pack(p, ref, w, n_blocks=N_BLOCKS)
    path = os.path.join(ROOT, "tests", "golden", "fastlanes_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
