"""ctypes binding of the CPU oracle (oracle/_build/libfl_oracle.so) — TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (fastlanes_b200) never does.

The oracle restates spiraldb/fastlanes v0.1.8 (see fl_oracle_kernels.hpp for per-function
reference citations).  All arrays are numpy, host memory, contiguous arrays of blocks.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libfl_oracle.so")

OP_PACK, OP_UNPACK, OP_FOR_PACK, OP_UNFOR_PACK, OP_DELTA, OP_UNDELTA, OP_UNDELTA_PACK, OP_TRANSPOSE, OP_UNTRANSPOSE, \
    OP_UNFOR_FILTER = range(10)

FLO_OK, FLO_ERR_WIDTH, FLO_ERR_TYPE, FLO_ERR_INDEX, FLO_ERR_NULL = range(5)

DTYPES = {8: np.uint8, 16: np.uint16, 32: np.uint32, 64: np.uint64}


class OracleError(RuntimeError):
    def __init__(self, code: int, what: str):
        super().__init__(f"oracle {what} failed with status {code}")
        self.code = code


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only; no reference sources involved)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.run(["make", "-s", "-j8", "-C", _HERE], check=True)
    return _LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.flo_run.restype = ctypes.c_int
        L.flo_run.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_uint, ctypes.c_size_t, ctypes.c_void_p,
                              ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int]
        L.flo_unpack_single.restype = ctypes.c_int
        L.flo_unpack_single.argtypes = [ctypes.c_int, ctypes.c_uint, ctypes.c_void_p, ctypes.c_size_t,
                                        ctypes.POINTER(ctypes.c_uint64)]
        L.flo_unpack_gather.restype = ctypes.c_int
        L.flo_unpack_gather.argtypes = [ctypes.c_int, ctypes.c_uint, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_size_t, ctypes.c_void_p]
        L.flo_isa.restype = ctypes.c_char_p
        L.flo_set_isa_level.restype = ctypes.c_int
        L.flo_set_isa_level.argtypes = [ctypes.c_int]
        L.flo_hardware_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def isa() -> str:
    return lib().flo_isa().decode()


def set_isa_level(level: int) -> int:
    return lib().flo_set_isa_level(level)


def hardware_threads() -> int:
    return lib().flo_hardware_threads()


def tbits_of(a: np.ndarray) -> int:
    return a.dtype.itemsize * 8


def packed_len(tbits: int, width: int) -> int:
    """Elements per packed block: 1024*W/T (src/bitpacking.rs:19,77)."""
    return 1024 * width // tbits


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _run(tbits, op, width, n_blocks, inp, out, base=None, refs=None, ref_scalar=0, threads=1):
    rc = lib().flo_run(tbits, op, width, n_blocks, _ptr(inp), _ptr(out), _ptr(base), _ptr(refs),
                       int(ref_scalar) & 0xFFFFFFFFFFFFFFFF, threads)
    if rc != FLO_OK:
        raise OracleError(rc, f"op {op}")


def _blocks_of(a: np.ndarray, per_block: int) -> int:
    if per_block == 0:
        raise ValueError("cannot infer block count from an empty packed array; pass n_blocks")
    if a.size % per_block:
        raise ValueError(f"array of {a.size} elements is not a whole number of {per_block}-element blocks")
    return a.size // per_block


def pack(values: np.ndarray, width: int, threads: int = 1) -> np.ndarray:
    values = np.ascontiguousarray(values)
    tb = tbits_of(values)
    n = _blocks_of(values, 1024)
    out = np.zeros(n * packed_len(tb, min(width, tb)), dtype=values.dtype)
    _run(tb, OP_PACK, width, n, values, out, threads=threads)
    return out


def unpack(packed: np.ndarray, width: int, n_blocks: int | None = None, threads: int = 1) -> np.ndarray:
    packed = np.ascontiguousarray(packed)
    tb = tbits_of(packed)
    if n_blocks is None:
        n_blocks = _blocks_of(packed, packed_len(tb, width))
    out = np.empty(n_blocks * 1024, dtype=packed.dtype)
    _run(tb, OP_UNPACK, width, n_blocks, packed, out, threads=threads)
    return out


def _ref_args(reference, n, dtype):
    if np.ndim(reference) == 0:
        return None, int(reference)
    refs = np.ascontiguousarray(reference, dtype=dtype)
    assert refs.size == n
    return refs, 0


def for_pack(values: np.ndarray, reference, width: int, threads: int = 1) -> np.ndarray:
    values = np.ascontiguousarray(values)
    tb = tbits_of(values)
    n = _blocks_of(values, 1024)
    refs, scalar = _ref_args(reference, n, values.dtype)
    out = np.zeros(n * packed_len(tb, min(width, tb)), dtype=values.dtype)
    _run(tb, OP_FOR_PACK, width, n, values, out, refs=refs, ref_scalar=scalar, threads=threads)
    return out


def unfor_pack(packed: np.ndarray, reference, width: int, n_blocks: int | None = None, threads: int = 1) -> np.ndarray:
    packed = np.ascontiguousarray(packed)
    tb = tbits_of(packed)
    if n_blocks is None:
        n_blocks = _blocks_of(packed, packed_len(tb, width))
    refs, scalar = _ref_args(reference, n_blocks, packed.dtype)
    out = np.empty(n_blocks * 1024, dtype=packed.dtype)
    _run(tb, OP_UNFOR_PACK, width, n_blocks, packed, out, refs=refs, ref_scalar=scalar, threads=threads)
    return out


def _base_check(base, n, tb, dtype):
    base = np.ascontiguousarray(base, dtype=dtype)
    assert base.size == n * (1024 // tb), "base must hold LANES elements per block"
    return base


def delta(values: np.ndarray, base: np.ndarray, threads: int = 1) -> np.ndarray:
    values = np.ascontiguousarray(values)
    tb = tbits_of(values)
    n = _blocks_of(values, 1024)
    base = _base_check(base, n, tb, values.dtype)
    out = np.empty_like(values)
    _run(tb, OP_DELTA, 0, n, values, out, base=base, threads=threads)
    return out


def undelta(deltas: np.ndarray, base: np.ndarray, threads: int = 1) -> np.ndarray:
    deltas = np.ascontiguousarray(deltas)
    tb = tbits_of(deltas)
    n = _blocks_of(deltas, 1024)
    base = _base_check(base, n, tb, deltas.dtype)
    out = np.empty_like(deltas)
    _run(tb, OP_UNDELTA, 0, n, deltas, out, base=base, threads=threads)
    return out


def undelta_pack(packed: np.ndarray, base: np.ndarray, width: int, n_blocks: int | None = None,
                 threads: int = 1) -> np.ndarray:
    packed = np.ascontiguousarray(packed)
    tb = tbits_of(packed)
    if n_blocks is None:
        n_blocks = _blocks_of(packed, packed_len(tb, width))
    base = _base_check(base, n_blocks, tb, packed.dtype)
    out = np.empty(n_blocks * 1024, dtype=packed.dtype)
    _run(tb, OP_UNDELTA_PACK, width, n_blocks, packed, out, base=base, threads=threads)
    return out


def transpose(values: np.ndarray, threads: int = 1) -> np.ndarray:
    values = np.ascontiguousarray(values)
    out = np.empty_like(values)
    _run(tbits_of(values), OP_TRANSPOSE, 0, _blocks_of(values, 1024), values, out, threads=threads)
    return out


def untranspose(values: np.ndarray, threads: int = 1) -> np.ndarray:
    values = np.ascontiguousarray(values)
    out = np.empty_like(values)
    _run(tbits_of(values), OP_UNTRANSPOSE, 0, _blocks_of(values, 1024), values, out, threads=threads)
    return out


def unpack_single(packed: np.ndarray, width: int, index: int) -> int:
    packed = np.ascontiguousarray(packed)
    v = ctypes.c_uint64(0)
    rc = lib().flo_unpack_single(tbits_of(packed), width, _ptr(packed), index, ctypes.byref(v))
    if rc != FLO_OK:
        raise OracleError(rc, "unpack_single")
    return int(v.value)


def unpack_gather(packed: np.ndarray, width: int, global_index: np.ndarray) -> np.ndarray:
    packed = np.ascontiguousarray(packed)
    gi = np.ascontiguousarray(global_index, dtype=np.uint64)
    out = np.empty(gi.size, dtype=packed.dtype)
    rc = lib().flo_unpack_gather(tbits_of(packed), width, _ptr(packed), _ptr(gi), gi.size, _ptr(out))
    if rc != FLO_OK:
        raise OracleError(rc, "unpack_gather")
    return out


def unfor_filter(packed: np.ndarray, reference, width: int, lo: int, hi: int, n_blocks: int | None = None,
                 threads: int = 1, out: np.ndarray | None = None) -> np.ndarray:
    """NOT a reference function: `unfor_pack` (src/ffor.rs:38-50) into a per-block scratch buffer followed by the
    caller-side predicate loop the reference's README prescribes (README.md:40-41).  Returns the selection bitmap,
    128 bytes per block, bit i of a block = lo <= value[i] <= hi."""
    packed = np.ascontiguousarray(packed)
    tb = tbits_of(packed)
    if n_blocks is None:
        n_blocks = _blocks_of(packed, packed_len(tb, width))
    refs, scalar = _ref_args(reference, n_blocks, packed.dtype)
    bounds = np.array([lo, hi], dtype=packed.dtype)
    if out is None:
        out = np.empty(n_blocks * 128, dtype=np.uint8)
    _run(tb, OP_UNFOR_FILTER, width, n_blocks, packed, out, base=bounds, refs=refs, ref_scalar=scalar, threads=threads)
    return out


def run_raw(tbits, op, width, n_blocks, inp, out, base=None, refs=None, ref_scalar=0, threads=1):
    """Timing entry used by bench.py: no allocation, caller-provided numpy buffers."""
    _run(tbits, op, width, n_blocks, inp, out, base=base, refs=refs, ref_scalar=ref_scalar, threads=threads)
