#!/usr/bin/env python
"""tools/pageable_probe.py — fl_host_unpack_u32 on PAGEABLE (plain numpy) vs page-locked (fl_host_alloc) caller buffers:
what a caller who hands over ordinary heap memory gets.  Measurement tool."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import fastlanes_b200 as fl  # noqa: E402
from fastlanes_b200 import _lib  # noqa: E402


def run(packed, out, w, n, reps=3):
    f = _lib.fn("fl_host_unpack", 32)
    _lib.check(f(w, n, packed.ctypes.data, out.ctypes.data))
    best = 1e9
    for _ in range(reps):
        t = time.perf_counter()
        _lib.check(f(w, n, packed.ctypes.data, out.ctypes.data))
        best = min(best, time.perf_counter() - t)
    return best


def main():
    fl.init(0)
    n = 1 << 18
    rng = np.random.default_rng(1)
    for w in (4, 16, 32):
        src = rng.integers(0, 1 << 32, size=n * 32 * w, dtype=np.uint32)
        p_page, o_page = src.copy(), np.zeros(n * 1024, dtype=np.uint32)
        p_pin, o_pin = fl.pinned_empty(src.size, np.uint32), fl.pinned_empty(n * 1024, np.uint32)
        p_pin[:] = src
        o_pin.fill(0)
        tp, tn = run(p_page, o_page, w, n), run(p_pin, o_pin, w, n)
        assert np.array_equal(o_page, o_pin)
        gb = n * 128 * (w + 32) / 1e9
        print(f"u32 W={w:2d} 2^18 blocks: pageable {tp*1e3:8.1f} ms {gb/tp:6.1f} GB/s ({n*1024/tp/1e9:5.2f} Gint/s) | "
              f"page-locked {tn*1e3:8.1f} ms {gb/tn:6.1f} GB/s ({n*1024/tn/1e9:5.2f} Gint/s)", flush=True)


if __name__ == "__main__":
    main()
