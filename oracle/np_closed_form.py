"""Second, INDEPENDENT oracle: the wire format in closed form, in numpy — TEST INFRASTRUCTURE.

Written from the format the reference's `unpack_single` (src/bitpacking.rs:132-179) and its
`lanes_by_index` / `rows_by_index` tables (:207-232) define, not from the streaming pack!/unpack!
macros, so that a bug shared between `fl_oracle_kernels.hpp`'s streaming loops and the CUDA
kernels (which also stream) would still be caught:

    lane(i) = i mod L                       (bitpacking.rs:210)
    row(i)  = FL_ORDER[f]*8 + s,  s = i div 128, f = ((i mod 128) - lane(i)) div 16   (:225-229)
    lane bit-stream S_lane = concat_k packed[L*k + lane]   (LSB first, k = 0..W-1)
    value(i) = (S_lane(i) >> (row(i)*W)) & (2^W - 1)

Everything is done on explicit 0/1 bit tensors; slow, but only used on small inputs.
"""
from __future__ import annotations

import numpy as np

FL_ORDER = np.array([0, 4, 2, 6, 1, 5, 3, 7])  # src/lib.rs:22


def lanes(tbits: int) -> int:
    return 1024 // tbits  # src/lib.rs:26


def index_table(tbits: int) -> np.ndarray:
    """idx[row, lane] = FL_ORDER[row/8]*16 + (row%8)*128 + lane  (src/macros.rs:20-24)."""
    row = np.arange(tbits)[:, None]
    lane = np.arange(lanes(tbits))[None, :]
    return FL_ORDER[row // 8] * 16 + (row % 8) * 128 + lane


def row_lane_of_index(tbits: int):
    """(row[i], lane[i]) by the formulas of rows_by_index / lanes_by_index (bitpacking.rs:207-232)."""
    i = np.arange(1024)
    L = lanes(tbits)
    lane = i % L
    s = i // 128
    f = (i - s * 128 - lane) // 16
    return FL_ORDER[f] * 8 + s, lane


def transpose_table() -> np.ndarray:
    """t[i] = (i%16)*64 + FL_ORDER[(i/16)%8]*8 + i/128  (src/transpose.rs:29-36)."""
    i = np.arange(1024)
    return (i % 16) * 64 + FL_ORDER[(i // 16) % 8] * 8 + i // 128


def _dtype(tbits):
    return {8: np.uint8, 16: np.uint16, 32: np.uint32, 64: np.uint64}[tbits]


def pack(values: np.ndarray, width: int) -> np.ndarray:
    tb = values.dtype.itemsize * 8
    L = lanes(tb)
    v = values.reshape(-1, 1024).astype(np.uint64)
    n = v.shape[0]
    if width == 0:
        return np.zeros(0, dtype=values.dtype)
    row, lane = row_lane_of_index(tb)
    # bits[n, i, b] = bit b of value i, truncated to `width` bits (macros.rs:73; W==T keeps all bits)
    bits = ((v[:, :, None] >> np.arange(width, dtype=np.uint64)[None, None, :]) & np.uint64(1)).astype(np.uint8)
    stream = np.zeros((n, L, tb * width), dtype=np.uint8)
    pos = row[:, None] * width + np.arange(width)[None, :]  # [i, b] position in the lane stream
    stream[:, lane[:, None].repeat(width, 1), pos] = bits
    words = stream.reshape(n, L, width, tb).astype(np.uint64)
    words = (words << np.arange(tb, dtype=np.uint64)[None, None, None, :]).sum(axis=3, dtype=np.uint64)
    # packed[L*k + lane]
    return np.ascontiguousarray(words.transpose(0, 2, 1)).reshape(-1).astype(values.dtype)


def unpack(packed: np.ndarray, width: int, n_blocks: int | None = None) -> np.ndarray:
    tb = packed.dtype.itemsize * 8
    L = lanes(tb)
    if width == 0:
        assert n_blocks is not None
        return np.zeros(n_blocks * 1024, dtype=packed.dtype)
    p = packed.reshape(-1, width, L).astype(np.uint64)  # [n, k, lane]
    n = p.shape[0]
    bits = ((p[:, :, :, None] >> np.arange(tb, dtype=np.uint64)[None, None, None, :]) & np.uint64(1)).astype(np.uint8)
    stream = bits.transpose(0, 2, 1, 3).reshape(n, L, width * tb)  # [n, lane, bit position]
    row, lane = row_lane_of_index(tb)
    pos = row[:, None] * width + np.arange(width)[None, :]
    vb = stream[:, lane[:, None].repeat(width, 1), pos].astype(np.uint64)  # [n, i, b]
    vals = (vb << np.arange(width, dtype=np.uint64)[None, None, :]).sum(axis=2, dtype=np.uint64)
    return vals.reshape(-1).astype(packed.dtype)


def transpose(values: np.ndarray) -> np.ndarray:
    t = transpose_table()
    return np.ascontiguousarray(values.reshape(-1, 1024)[:, t]).reshape(-1)


def untranspose(values: np.ndarray) -> np.ndarray:
    t = transpose_table()
    out = np.empty_like(values.reshape(-1, 1024))
    out[:, t] = values.reshape(-1, 1024)
    return out.reshape(-1)


def delta(values: np.ndarray, base: np.ndarray) -> np.ndarray:
    """Per lane, along rows, out = in - prev with prev0 = base[lane] (src/delta.rs:24-33)."""
    tb = values.dtype.itemsize * 8
    idx = index_table(tb)  # [row, lane]
    v = values.reshape(-1, 1024)
    g = v[:, idx]  # [n, row, lane]
    prev = np.concatenate([base.reshape(-1, 1, lanes(tb)), g[:, :-1, :]], axis=1)
    d = (g - prev).astype(values.dtype)
    out = np.empty_like(v)
    out[:, idx] = d
    return out.reshape(-1)


def undelta(deltas: np.ndarray, base: np.ndarray) -> np.ndarray:
    """Running wrapping sum along rows per lane (src/delta.rs:36-45)."""
    tb = deltas.dtype.itemsize * 8
    idx = index_table(tb)
    d = deltas.reshape(-1, 1024)[:, idx]
    acc = np.cumsum(d, axis=1, dtype=deltas.dtype) + base.reshape(-1, 1, lanes(tb))
    out = np.empty_like(deltas.reshape(-1, 1024))
    out[:, idx] = acc.astype(deltas.dtype)
    return out.reshape(-1)
