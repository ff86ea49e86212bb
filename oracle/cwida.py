"""oracle/cwida.py — TEST INFRASTRUCTURE.  Closed-form numpy oracle for bit-packing / FoR in the row order of the
ORIGINAL cwida/FastLanes layout.

PARITY UNPINNED.  /root/reference (spiraldb/fastlanes v0.1.8) contains no code, test or vector for this layout; it only
says that its own layout differs from it: README.md:49-56 ("not binary compatible with original FastLanes ... reordered
vs the original") and src/macros.rs:1-9 ("It differs in that it iterates over the elements respecting the transposed
ordering").  What is restated here is therefore the layout of the FastLanes paper (Afroozeh & Boncz, VLDB 2023) as the
reference describes the difference: everything is as in the reference — LANES = 1024/T interleaved lanes, lane l's
bit-stream is packed[LANES*k + l] for k = 0..W-1, LSB first, row r at bits [r*W, r*W + W) (src/macros.rs:35-97) — EXCEPT
that rows are visited in plain order: row r of lane l is value r*LANES + l (the reference: index(row, lane) =
FL_ORDER[row/8]*16 + (row%8)*128 + lane, src/macros.rs:20-24).

Two independent formulations, cross-checked in tests/test_oracle_cwida.py:
  * closed form (this file's pack / unpack), written from the description above;
  * composition with the pinned oracle: cwida_pack(v) == reference_pack(v permuted so that the reference's row visit
    order reads the values in linear order).
"""
from __future__ import annotations

import numpy as np

FL_ORDER = (0, 4, 2, 6, 1, 5, 3, 7)


def _dims(dtype):
    tb = np.dtype(dtype).itemsize * 8
    return tb, 1024 // tb


def pack(values: np.ndarray, width: int) -> np.ndarray:
    """values: n*1024 unsigned ints -> n * (1024*W/T) packed words, linear row order.  Truncates to W bits (W < T)."""
    values = np.ascontiguousarray(values)
    tb, lanes = _dims(values.dtype)
    n = values.size // 1024
    v = values.reshape(n, tb, lanes).astype(object)            # [block, row, lane]: value r*LANES + l
    if width < tb:
        v = v & ((1 << width) - 1)
    stream = np.zeros((n, lanes), dtype=object)                 # lane bit-streams as Python ints
    for r in range(tb):
        stream = stream | (v[:, r, :] << (r * width))
    out = np.zeros((n, width, lanes), dtype=values.dtype)
    mask = (1 << tb) - 1
    for k in range(width):
        out[:, k, :] = ((stream >> (k * tb)) & mask).astype(values.dtype)
    return out.reshape(-1)


def unpack(packed: np.ndarray, width: int, n_blocks: int) -> np.ndarray:
    packed = np.ascontiguousarray(packed)
    tb, lanes = _dims(packed.dtype)
    words = packed.reshape(n_blocks, width, lanes).astype(object) if width else np.zeros((n_blocks, 0, lanes), dtype=object)
    stream = np.zeros((n_blocks, lanes), dtype=object)
    for k in range(width):
        stream = stream | (words[:, k, :] << (k * tb))
    out = np.zeros((n_blocks, tb, lanes), dtype=packed.dtype)
    mask = (1 << width) - 1
    for r in range(tb):
        out[:, r, :] = ((stream >> (r * width)) & mask).astype(packed.dtype)
    return out.reshape(-1)


def for_pack(values: np.ndarray, reference: int, width: int) -> np.ndarray:
    values = np.ascontiguousarray(values)
    with np.errstate(over="ignore"):
        return pack((values - values.dtype.type(reference)).astype(values.dtype), width)


def unfor_pack(packed: np.ndarray, reference: int, width: int, n_blocks: int) -> np.ndarray:
    out = unpack(packed, width, n_blocks)
    with np.errstate(over="ignore"):
        return (out + out.dtype.type(reference)).astype(out.dtype)


def to_reference_order(values: np.ndarray) -> np.ndarray:
    """Permutation p with reference_pack(p(values)) == cwida pack(values): put value r*LANES + l where the
    reference's index(r, l) reads it."""
    values = np.ascontiguousarray(values)
    tb, lanes = _dims(values.dtype)
    n = values.size // 1024
    r = np.arange(tb)[:, None]
    l = np.arange(lanes)[None, :]
    ref_index = (np.array(FL_ORDER)[r // 8] * 16 + (r % 8) * 128 + l).reshape(-1)   # index(row, lane), macros.rs:20-24
    out = np.empty_like(values).reshape(n, 1024)
    out[:, ref_index] = values.reshape(n, 1024)
    return out.reshape(-1)
