"""fastlanes_b200 — host-side mirror of the spiraldb/fastlanes trait surface over the C-ABI library.

The reference's public API is four traits implemented for u8/u16/u32/u64 (src/lib.rs:17-32):

    BitPacking  pack / unchecked_pack / unpack / unchecked_unpack / unpack_single / unchecked_unpack_single
                (src/bitpacking.rs:16-59)
    FoR         for_pack / unfor_pack                      (src/ffor.rs:4-18)
    Delta       delta / undelta / undelta_pack             (src/delta.rs:6-17)
    Transpose   transpose / untranspose                    (src/transpose.rs:4-7)

This module keeps those names and argument orders, with the const-generic `W` as the leading `width`
argument (as in the reference's own `unchecked_*` family), batched over any whole number of 1024-value
blocks.  Where the reference panics (width > T, wrong slice lengths, index >= 1024) a FastLanesError is
raised.  Arguments are either

  * numpy arrays (host memory)  -> the `fl_host_*` entry points: H2D, sm_100a kernels, D2H; or
  * torch CUDA tensors          -> the stream-ordered device entry points on torch's current stream.

Every value is computed by the CUDA library; there is no CPU implementation in this package.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from ._lib import FastLanesError, LIB_PATH, exported_symbols  # noqa: F401

FL_ORDER = (0, 4, 2, 6, 1, 5, 3, 7)  # src/lib.rs:22

__all__ = ["BitPacking", "FoR", "Delta", "Transpose", "Scan", "Cwida", "Context", "FastLanes", "FastLanesError", "FL_ORDER",
           "packed_len", "version", "device_count", "init", "host_configure", "pinned_empty", "buffer_node", "shutdown"]


class FastLanes:
    """`trait FastLanes { const T; const LANES }` (src/lib.rs:24-27) for an element bit size."""

    def __init__(self, tbits: int):
        if tbits not in (8, 16, 32, 64):
            raise FastLanesError(_lib.FL_ERR_LEN, f"unsupported element size {tbits}")
        self.T = tbits
        self.LANES = 1024 // tbits


def packed_len(tbits: int, width: int) -> int:
    """Elements in one packed block: 1024 * W / T (src/bitpacking.rs:19)."""
    return 1024 * width // tbits


def version() -> str:
    return _lib.lib().fl_version().decode()


def device_count() -> int:
    return _lib.lib().fl_device_count()


def init(device: int = 0) -> None:
    """Optional warm-up: make `device` current and create the host-path streams (fl_init)."""
    _lib.check(_lib.lib().fl_init(device))


def host_configure(chunk_blocks: int = 0, n_streams: int = 0) -> None:
    _lib.check(_lib.lib().fl_host_configure(chunk_blocks, n_streams))


def shutdown() -> None:
    _lib.check(_lib.lib().fl_shutdown())


def _pinned_view(ptr: int, n: int, dt) -> np.ndarray:
    """numpy view of a library-owned page-locked allocation; fl_host_free runs when the last view dies."""
    buf = (ctypes.c_uint8 * (n * dt.itemsize)).from_address(ptr)
    arr = np.frombuffer(buf, dtype=dt, count=n)

    class _Owner:
        def __init__(self, p):
            self.ptr = p

        def __del__(self):
            try:
                _lib.lib().fl_host_free(self.ptr)
            except Exception:
                pass

    buf._fl_owner = _Owner(ptr)  # tie the owner to the base buffer so that views keep it alive
    return arr


def pinned_empty(n: int, dtype) -> np.ndarray:
    """A page-locked numpy array on the NUMA node of the current CUDA device (fl_host_alloc)."""
    dt = np.dtype(dtype)
    p = ctypes.c_void_p()
    _lib.check(_lib.lib().fl_host_alloc(ctypes.byref(p), max(1, n * dt.itemsize)))
    return _pinned_view(p.value, n, dt)


def buffer_node(a: np.ndarray, byte_offset: int = 0) -> int:
    """NUMA node holding the page at `byte_offset` of a host array (fl_host_buffer_node), -1 if unknown."""
    return _lib.lib().fl_host_buffer_node(a.ctypes.data + byte_offset)


# ---- argument plumbing ---------------------------------------------------------------------------

def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


class _Arg:
    __slots__ = ("ptr", "n", "tbits", "device")

    def __init__(self, x, what: str):
        if _is_torch(x):
            if not x.is_cuda:
                raise FastLanesError(_lib.FL_ERR_NULL, f"{what}: torch tensors must live on a CUDA device")
            if not x.is_contiguous():
                raise FastLanesError(_lib.FL_ERR_LEN, f"{what}: tensor must be contiguous")
            self.ptr, self.n, self.tbits, self.device = x.data_ptr(), x.numel(), x.element_size() * 8, x.device.index
        elif isinstance(x, np.ndarray):
            if not x.flags.c_contiguous:
                raise FastLanesError(_lib.FL_ERR_LEN, f"{what}: array must be C-contiguous")
            if x.dtype.kind != "u":
                raise FastLanesError(_lib.FL_ERR_LEN, f"{what}: unsigned integer dtype required")
            self.ptr, self.n, self.tbits, self.device = x.ctypes.data, x.size, x.dtype.itemsize * 8, None
        else:
            raise FastLanesError(_lib.FL_ERR_NULL, f"{what}: numpy array or torch CUDA tensor required")


class _Space:
    """Where a call's buffers live: host (falsy) or ONE CUDA device (truthy, `.index`)."""
    __slots__ = ("index",)

    def __init__(self, index):
        self.index = index

    def __bool__(self):
        return self.index is not None


def _same_space(*args: _Arg) -> _Space:
    """All buffers on the host, or all on the SAME CUDA device (a kernel launched on the current device with another
    device's pointers is an illegal address without peer access and silent NVLink traffic with it)."""
    devs = {a.device for a in args}
    if len({d is not None for d in devs}) != 1:
        raise FastLanesError(_lib.FL_ERR_NULL, "all buffers must be host arrays or all CUDA tensors")
    if len(devs) != 1:
        raise FastLanesError(_lib.FL_ERR_NULL, f"all CUDA tensors of one call must live on one device, got {sorted(devs)}")
    tb = {a.tbits for a in args}
    if len(tb) != 1:
        raise FastLanesError(_lib.FL_ERR_LEN, "all buffers must share one element type")
    return _Space(devs.pop())


def _join_space(space: _Space, *others: _Arg) -> None:
    """Buffers of a different element type (indices, bitmaps, counts, offsets) must share the call's memory space."""
    for a in others:
        if a.device != space.index:
            raise FastLanesError(_lib.FL_ERR_NULL, "all buffers must be host arrays or CUDA tensors on one device")


class _Launch:
    """`with _Launch(space) as stream:` — makes the buffers' device current for the call (the C ABI launches on the
    current device) and yields torch's current stream OF THAT DEVICE; a no-op (stream None) for host buffers."""

    def __init__(self, space: _Space):
        self.space, self.guard = space, None

    def __enter__(self):
        if not self.space:
            return None
        import torch

        self.guard = torch.cuda.device(self.space.index)
        self.guard.__enter__()
        return ctypes.c_void_p(torch.cuda.current_stream(self.space.index).cuda_stream)

    def __exit__(self, *exc):
        if self.guard is not None:
            self.guard.__exit__(*exc)
        return False


def _check_width(width: int, tbits: int):
    if width < 0 or width > tbits:
        # the reference: compile error for const W (bitpacking.rs:10-12), unreachable!() at run time (:93)
        raise FastLanesError(_lib.FL_ERR_WIDTH, f"Unsupported width: {width}")


def _n_blocks_unpacked(a: _Arg, what: str) -> int:
    if a.n % 1024:
        raise FastLanesError(_lib.FL_ERR_LEN, f"{what} buffer must be a whole number of 1024-element blocks")
    return a.n // 1024


def _expect(a: _Arg, n: int, what: str):
    if a.n != n:
        raise FastLanesError(_lib.FL_ERR_LEN, f"{what} buffer must hold {n} elements, got {a.n}")


def _call(base, tbits, space, *args):
    name = base if space else base.replace("fl_", "fl_host_", 1)
    with _Launch(space) as stream:
        if space:
            args = args + (stream,)
        _lib.check(_lib.fn(name, tbits)(*args))


class BitPacking:
    """src/bitpacking.rs:16-59."""

    @staticmethod
    def pack(width: int, input, output) -> None:
        """`pack::<W>(input: &[T; 1024], output: &mut [T; 1024*W/T])` (:19, impl :65-74), batched."""
        i, o = _Arg(input, "input"), _Arg(output, "output")
        dev = _same_space(i, o)
        _check_width(width, i.tbits)
        n = _n_blocks_unpacked(i, "Input")
        _expect(o, n * packed_len(i.tbits, width), "Output")  # :78
        _call("fl_pack", i.tbits, dev, width, n, i.ptr, o.ptr)

    unchecked_pack = pack  # :30 (impl :76-96): same runtime-width dispatch; lengths ARE checked here

    @staticmethod
    def unpack(width: int, input, output) -> None:
        """`unpack::<W>(input: &[T; 1024*W/T], output: &mut [T; 1024])` (:33, impl :98-107), batched."""
        i, o = _Arg(input, "input"), _Arg(output, "output")
        dev = _same_space(i, o)
        _check_width(width, o.tbits)
        n = _n_blocks_unpacked(o, "Output")
        _expect(i, n * packed_len(o.tbits, width), "Input")  # :111
        _call("fl_unpack", o.tbits, dev, width, n, i.ptr, o.ptr)

    unchecked_unpack = unpack  # :44 (impl :109-129)

    @staticmethod
    def unpack_single(width: int, packed, index: int):
        """`unpack_single::<W>(packed, index) -> T` (:47, impl :132-179) on ONE packed block (host array)."""
        p = _Arg(packed, "packed")
        if p.device is not None:
            raise FastLanesError(_lib.FL_ERR_NULL, "unpack_single takes a host array; use unpack_gather for tensors")
        _check_width(width, p.tbits)
        _expect(p, packed_len(p.tbits, width), "Input")  # :185
        if not 0 <= index < 1024:
            raise FastLanesError(_lib.FL_ERR_INDEX, f"Index must be less than 1024, got {index}")  # :152
        out = np.zeros(1, dtype=packed.dtype)
        _lib.check(_lib.fn("fl_host_unpack_single", p.tbits)(width, p.ptr, index, out.ctypes.data))
        return out[0]

    unchecked_unpack_single = unpack_single  # :58 (impl :181-200)

    @staticmethod
    def unpack_gather(width: int, packed, global_index, output) -> None:
        """Batched unpack_single: output[i] = value at block global_index[i]//1024, index global_index[i]%1024."""
        p, g, o = _Arg(packed, "packed"), _Arg(global_index, "global_index"), _Arg(output, "output")
        if g.tbits != 64:
            raise FastLanesError(_lib.FL_ERR_LEN, "global_index must be uint64")
        dev = _same_space(p, o)
        _join_space(dev, g)
        _check_width(width, p.tbits)
        per = packed_len(p.tbits, width)
        if per and p.n % per:
            raise FastLanesError(_lib.FL_ERR_LEN, "packed buffer must be a whole number of packed blocks")
        n_blocks = p.n // per if per else (1 << 40)
        _expect(o, g.n, "Output")
        if dev:
            import torch

            flag = torch.zeros(1, dtype=torch.int32, device=packed.device)
            with _Launch(dev) as stream:
                _lib.check(_lib.fn("fl_unpack_gather", p.tbits)(width, n_blocks, p.ptr, g.ptr, g.n, o.ptr,
                                                                flag.data_ptr(), stream))
            if int(flag.item()):
                raise FastLanesError(_lib.FL_ERR_INDEX, "index out of range")
        else:
            _lib.check(_lib.fn("fl_host_unpack_gather", p.tbits)(width, n_blocks, p.ptr, g.ptr, g.n, o.ptr))


def _ref_value(reference, tbits: int) -> int:
    return int(reference) & ((1 << tbits) - 1)


class FoR:
    """src/ffor.rs:4-18."""

    @staticmethod
    def for_pack(width: int, input, reference, output) -> None:
        """`for_pack::<W>(input, reference, output)` (:5-10, impl :24-36).  `reference`: a scalar, or (CUDA
        tensors only) one reference per block."""
        i, o = _Arg(input, "input"), _Arg(output, "output")
        dev = _same_space(i, o)
        _check_width(width, i.tbits)
        n = _n_blocks_unpacked(i, "Input")
        _expect(o, n * packed_len(i.tbits, width), "Output")
        if _is_torch(reference) and reference.dim() > 0:
            r = _Arg(reference, "reference")
            _same_space(i, r)
            _expect(r, n, "Reference")
            _call("fl_for_pack_refs", i.tbits, dev, width, n, i.ptr, r.ptr, o.ptr)
        else:
            _call("fl_for_pack", i.tbits, dev, width, n, i.ptr, _ref_value(reference, i.tbits), o.ptr)

    @staticmethod
    def unfor_pack(width: int, input, reference, output) -> None:
        """`unfor_pack::<W>(input, reference, output)` (:12-17, impl :38-50)."""
        i, o = _Arg(input, "input"), _Arg(output, "output")
        dev = _same_space(i, o)
        _check_width(width, o.tbits)
        n = _n_blocks_unpacked(o, "Output")
        _expect(i, n * packed_len(o.tbits, width), "Input")
        if _is_torch(reference) and reference.dim() > 0:
            r = _Arg(reference, "reference")
            _same_space(o, r)
            _expect(r, n, "Reference")
            _call("fl_unfor_pack_refs", o.tbits, dev, width, n, i.ptr, r.ptr, o.ptr)
        else:
            _call("fl_unfor_pack", o.tbits, dev, width, n, i.ptr, _ref_value(reference, o.tbits), o.ptr)


    @staticmethod
    def block_minmax(input, mins, maxs) -> None:
        """Per-block min and max (not in the reference: the statistics its callers compute before `for_pack`,
        which takes `reference` and W as givens, src/ffor.rs:5-10).  mins/maxs: one element per block."""
        i, lo, hi = _Arg(input, "input"), _Arg(mins, "mins"), _Arg(maxs, "maxs")
        dev = _same_space(i, lo, hi)
        n = _n_blocks_unpacked(i, "Input")
        _expect(lo, n, "Mins")
        _expect(hi, n, "Maxs")
        _call("fl_block_minmax", i.tbits, dev, n, i.ptr, lo.ptr, hi.ptr)

    @staticmethod
    def for_pack_auto(width: int, input, references, output, spans=None) -> None:
        """`for_pack::<W>` (src/ffor.rs:24-36) with reference = each block's own minimum, found in the same pass
        (CUDA tensors).  `references` (one per block) receives the minima; `spans` (optional) max - min per block:
        block b round-trips losslessly iff spans[b] < 2**width."""
        i, r, o = _Arg(input, "input"), _Arg(references, "references"), _Arg(output, "output")
        dev = _same_space(i, r, o)
        if not dev:
            raise FastLanesError(_lib.FL_ERR_NULL, "for_pack_auto takes CUDA tensors")
        _check_width(width, i.tbits)
        n = _n_blocks_unpacked(i, "Input")
        _expect(r, n, "References")
        _expect(o, n * packed_len(i.tbits, width), "Output")
        sptr = None
        if spans is not None:
            s = _Arg(spans, "spans")
            _same_space(i, s)
            _expect(s, n, "Spans")
            sptr = s.ptr
        with _Launch(dev) as stream:
            _lib.check(_lib.fn("fl_for_pack_auto", i.tbits)(width, n, i.ptr, r.ptr, sptr, o.ptr, stream))

    @staticmethod
    def choose(mins, maxs, tbits: int):
        """Host helper: (reference per block, one width for the batch) = (mins, bits(max over blocks of max - min))."""
        span = (np.asarray(maxs).astype(np.uint64) - np.asarray(mins).astype(np.uint64)) & np.uint64((1 << tbits) - 1 if tbits < 64 else 0xFFFFFFFFFFFFFFFF)
        return mins, int(span.max()).bit_length() if span.size else 0


class Delta:
    """src/delta.rs:6-17.  `base` holds LANES = 1024/T elements per block."""

    @staticmethod
    def _three(input, base, output, packed_width=None):
        i, b, o = _Arg(input, "input"), _Arg(base, "base"), _Arg(output, "output")
        dev = _same_space(i, b, o)
        n = _n_blocks_unpacked(o, "Output")
        _expect(b, n * (1024 // o.tbits), "Base")
        if packed_width is None:
            _expect(i, n * 1024, "Input")
        else:
            _check_width(packed_width, o.tbits)
            _expect(i, n * packed_len(o.tbits, packed_width), "Input")
        return i, b, o, dev, n

    @staticmethod
    def delta(input, base, output) -> None:
        """`delta(input, base, output)` (:7, impl :24-33)."""
        i, b, o, dev, n = Delta._three(input, base, output)
        _call("fl_delta", o.tbits, dev, n, i.ptr, b.ptr, o.ptr)

    @staticmethod
    def undelta(input, base, output) -> None:
        """`undelta(input, base, output)` (:9, impl :36-45)."""
        i, b, o, dev, n = Delta._three(input, base, output)
        _call("fl_undelta", o.tbits, dev, n, i.ptr, b.ptr, o.ptr)

    @staticmethod
    def undelta_pack(width: int, input, base, output) -> None:
        """`undelta_pack::<W>(input, base, output)` (:11-17, impl :48-63): fused unpack + prefix-add."""
        i, b, o, dev, n = Delta._three(input, base, output, packed_width=width)
        _call("fl_undelta_pack", o.tbits, dev, width, n, i.ptr, b.ptr, o.ptr)


    # ---- fused chains (not in the reference's trait: compositions of its methods) ----
    @staticmethod
    def undelta_pack_untranspose(width: int, input, base, output) -> None:
        """untranspose(undelta_pack::<W>(input, base)) in one pass: decode straight to ORIGINAL value order
        (src/delta.rs:99 then src/transpose.rs:18-22)."""
        i, b, o, dev, n = Delta._three(input, base, output, packed_width=width)
        _call("fl_undelta_pack_untranspose", o.tbits, dev, width, n, i.ptr, b.ptr, o.ptr)

    @staticmethod
    def transpose_delta_pack(width: int, input, base, output) -> None:
        """pack::<W>(delta(transpose(input), base)) in one pass: the encode chain of src/delta.rs:88-95."""
        i, b, o = _Arg(input, "input"), _Arg(base, "base"), _Arg(output, "output")
        dev = _same_space(i, b, o)
        _check_width(width, i.tbits)
        n = _n_blocks_unpacked(i, "Input")
        _expect(b, n * (1024 // i.tbits), "Base")
        _expect(o, n * packed_len(i.tbits, width), "Output")
        _call("fl_transpose_delta_pack", i.tbits, dev, width, n, i.ptr, b.ptr, o.ptr)


class Transpose:
    """src/transpose.rs:4-7."""

    @staticmethod
    def _two(input, output):
        i, o = _Arg(input, "input"), _Arg(output, "output")
        dev = _same_space(i, o)
        n = _n_blocks_unpacked(i, "Input")
        _expect(o, n * 1024, "Output")
        return i, o, dev, n

    @staticmethod
    def transpose(input, output) -> None:
        """`transpose(input, output)`: output[i] = input[transpose(i)] (:5, impl :11-15)."""
        i, o, dev, n = Transpose._two(input, output)
        _call("fl_transpose", i.tbits, dev, n, i.ptr, o.ptr)

    @staticmethod
    def untranspose(input, output) -> None:
        """`untranspose(input, output)`: output[transpose(i)] = input[i] (:6, impl :18-22)."""
        i, o, dev, n = Transpose._two(input, output)
        _call("fl_untranspose", i.tbits, dev, n, i.ptr, o.ptr)

    @staticmethod
    def transpose_index(idx: int) -> int:
        """`const fn transpose(idx)` (src/transpose.rs:29-36)."""
        return (idx % 16) * 64 + FL_ORDER[(idx // 16) % 8] * 8 + idx // 128


class Scan:
    """Fused decode + predicate.  Not a trait of the reference: its README (README.md:40-41)
    tells callers to unpack the whole block and loop over it.  Here the decoded block stays in registers.

    Value i of a block is `unfor_pack::<W>(packed, reference)[i]` (src/ffor.rs:38-50; reference 0 = plain
    `unpack`, src/bitpacking.rs:98-107).  Bitmaps are uint8, 128 bytes per block, bit i of a block at byte i//8,
    bit i%8 (numpy `packbits(bitorder="little")`)."""

    @staticmethod
    def _bitmap_args(width: int, packed, bitmap, counts):
        """Common plumbing: (packed arg, bitmap arg, on_device, n_blocks, counts pointer or None)."""
        p, b = _Arg(packed, "packed"), _Arg(bitmap, "bitmap")
        if b.tbits != 8:
            raise FastLanesError(_lib.FL_ERR_LEN, "bitmap must be uint8")
        dev = _Space(p.device)
        _join_space(dev, b)
        _check_width(width, p.tbits)
        if b.n % 128:
            raise FastLanesError(_lib.FL_ERR_LEN, "bitmap must hold 128 bytes per block")
        n = b.n // 128
        _expect(p, n * packed_len(p.tbits, width), "Input")
        cptr = None
        if counts is not None:
            c = _Arg(counts, "counts")
            if c.tbits != 32:
                raise FastLanesError(_lib.FL_ERR_LEN, "counts must be uint32")
            _join_space(dev, c)
            _expect(c, n, "Counts")
            cptr = c.ptr
        return p, b, dev, n, cptr

    @staticmethod
    def _reference_args(reference, p: _Arg, n: int):
        """(per-block references pointer or None, scalar reference)."""
        if _is_torch(reference) and reference.dim() > 0:
            r = _Arg(reference, "reference")
            if r.tbits != p.tbits or r.device is None or r.device != p.device:
                raise FastLanesError(_lib.FL_ERR_LEN, "per-block references: a CUDA tensor of the packed element type on the same device")
            _expect(r, n, "Reference")
            return r.ptr, 0
        return None, _ref_value(reference, p.tbits)

    @staticmethod
    def filter_range(width: int, packed, reference, lo, hi, bitmap, counts=None) -> None:
        """bitmap bit i = lo <= value[i] <= hi (unsigned, inclusive).  `reference`: scalar, or (CUDA tensors) one
        per block.  `counts` (optional, uint32 per block) receives the number of selected values."""
        p, b, dev, n, cptr = Scan._bitmap_args(width, packed, bitmap, counts)
        lo_v, hi_v = _ref_value(lo, p.tbits), _ref_value(hi, p.tbits)
        if dev:
            rptr, rval = Scan._reference_args(reference, p, n)
            with _Launch(dev) as stream:
                _lib.check(_lib.fn("fl_unpack_filter", p.tbits)(width, n, p.ptr, rptr, rval, lo_v, hi_v, b.ptr, cptr, stream))
        else:
            _lib.check(_lib.fn("fl_host_unpack_filter", p.tbits)(width, n, p.ptr, _ref_value(reference, p.tbits), lo_v,
                                                                 hi_v, b.ptr, cptr))

    @staticmethod
    def filter_range_delta(width: int, packed, base, lo, hi, bitmap, counts=None) -> None:
        """Range scan over a delta-encoded column: bitmap bit i = lo <= untranspose(undelta_pack::<W>(packed, base))[i]
        <= hi (src/delta.rs:48-63 then src/transpose.rs:18-22), i.e. in ORIGINAL value order.  `base`: LANES elements
        per block as for Delta.undelta_pack."""
        p, b, dev, n, cptr = Scan._bitmap_args(width, packed, bitmap, counts)
        bs = _Arg(base, "base")
        _same_space(p, bs)
        _expect(bs, n * (1024 // p.tbits), "Base")
        lo_v, hi_v = _ref_value(lo, p.tbits), _ref_value(hi, p.tbits)
        if dev:
            with _Launch(dev) as stream:
                _lib.check(_lib.fn("fl_undelta_pack_filter", p.tbits)(width, n, p.ptr, bs.ptr, lo_v, hi_v, b.ptr, cptr, stream))
        else:
            _lib.check(_lib.fn("fl_host_undelta_pack_filter", p.tbits)(width, n, p.ptr, bs.ptr, lo_v, hi_v, b.ptr, cptr))

    @staticmethod
    def select(width: int, packed, reference, bitmap, offsets, output) -> None:
        """Dense compaction (CUDA tensors): output[offsets[b] + k] = k-th selected value of block b in index order;
        `offsets` (uint64/int64 per block) = exclusive prefix sum of the per-block counts."""
        p, b, dev, n, _ = Scan._bitmap_args(width, packed, bitmap, None)
        f, o = _Arg(offsets, "offsets"), _Arg(output, "output")
        if not dev:
            raise FastLanesError(_lib.FL_ERR_NULL, "select takes CUDA tensors")
        _join_space(dev, f, o)
        if f.tbits != 64 or o.tbits != p.tbits:
            raise FastLanesError(_lib.FL_ERR_LEN, "offsets must be 64-bit, output of the packed element type")
        _expect(f, n, "Offsets")
        rptr, rval = Scan._reference_args(reference, p, n)
        with _Launch(dev) as stream:
            _lib.check(_lib.fn("fl_unpack_select", p.tbits)(width, n, p.ptr, rptr, rval, b.ptr, f.ptr, o.ptr, stream))


class Cwida:
    """Bit-packing and FoR in the row order of the ORIGINAL cwida/FastLanes layout: row r of a
    block holds values r*LANES .. r*LANES+LANES-1, where the reference crate visits rows in FL_ORDER-transposed order
    and is therefore "not binary compatible with original FastLanes" (README.md:49-56, src/macros.rs:1-9).  Same
    argument conventions as BitPacking / FoR; CUDA tensors only.  Parity unpinned: see oracle/cwida.py."""

    @staticmethod
    def _io(width, unpacked, packed, what_unpacked, what_packed):
        u, p = _Arg(unpacked, what_unpacked), _Arg(packed, what_packed)
        dev = _same_space(u, p)
        if not dev:
            raise FastLanesError(_lib.FL_ERR_NULL, "the cwida entry points take CUDA tensors")
        _check_width(width, u.tbits)
        n = _n_blocks_unpacked(u, what_unpacked.capitalize())
        _expect(p, n * packed_len(u.tbits, width), what_packed.capitalize())
        return u, p, n, dev

    @staticmethod
    def pack(width: int, input, output) -> None:
        u, p, n, dev = Cwida._io(width, input, output, "input", "output")
        _call("fl_pack_cwida", u.tbits, dev, width, n, u.ptr, p.ptr)

    @staticmethod
    def unpack(width: int, input, output) -> None:
        u, p, n, dev = Cwida._io(width, output, input, "output", "input")
        _call("fl_unpack_cwida", u.tbits, dev, width, n, p.ptr, u.ptr)

    @staticmethod
    def for_pack(width: int, input, reference, output) -> None:
        u, p, n, dev = Cwida._io(width, input, output, "input", "output")
        _call("fl_for_pack_cwida", u.tbits, dev, width, n, u.ptr, _ref_value(reference, u.tbits), p.ptr)

    @staticmethod
    def unfor_pack(width: int, input, reference, output) -> None:
        u, p, n, dev = Cwida._io(width, output, input, "output", "input")
        _call("fl_unfor_pack_cwida", u.tbits, dev, width, n, p.ptr, _ref_value(reference, u.tbits), u.ptr)


class Context:
    """`fl_ctx`: one process, several GPUs, contiguous block shards (SURVEY.md §8e behind the C ABI).  Methods mirror the
    trait calls above for HOST arrays; blocks [n*i/G, n*(i+1)/G) of every array run on device i through that device's
    own copy pipeline, concurrently.  `devices=None` = all visible devices; a device may be listed more than once."""

    def __init__(self, devices=None):
        self._h = ctypes.c_void_p()
        if devices is None:
            _lib.check(_lib.lib().fl_ctx_create(None, 0, ctypes.byref(self._h)))
        else:
            arr = (ctypes.c_int * len(devices))(*devices)
            _lib.check(_lib.lib().fl_ctx_create(arr, len(devices), ctypes.byref(self._h)))

    def close(self) -> None:
        if self._h:
            _lib.lib().fl_ctx_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    @property
    def devices(self) -> list:
        L = _lib.lib()
        return [L.fl_ctx_device(self._h, i) for i in range(L.fl_ctx_device_count(self._h))]

    def block_range(self, n_blocks: int, i: int) -> tuple:
        b0, b1 = ctypes.c_size_t(), ctypes.c_size_t()
        _lib.check(_lib.lib().fl_ctx_block_range(self._h, n_blocks, i, ctypes.byref(b0), ctypes.byref(b1)))
        return b0.value, b1.value

    def pinned_empty(self, n_blocks: int, elems_per_block: int, dtype) -> np.ndarray:
        """Page-locked array of n_blocks * elems_per_block elements whose pages follow the shards (fl_ctx_host_alloc)."""
        dt = np.dtype(dtype)
        p = ctypes.c_void_p()
        _lib.check(_lib.lib().fl_ctx_host_alloc(self._h, n_blocks, elems_per_block * dt.itemsize, ctypes.byref(p)))
        return _pinned_view(p.value, n_blocks * elems_per_block, dt)

    # ---- host arrays, sharded -----------------------------------------------------------------------
    @staticmethod
    def _host(*arrays):
        args = [_Arg(a, "buffer") for a in arrays]
        if _same_space(*args):
            raise FastLanesError(_lib.FL_ERR_NULL, "Context methods take host (numpy) arrays")
        return args

    def _call(self, name, tbits, *args):
        _lib.check(_lib.fn(name, tbits)(self._h, *args))

    def pack(self, width: int, input, output) -> None:
        i, o = self._host(input, output)
        _check_width(width, i.tbits)
        n = _n_blocks_unpacked(i, "Input")
        _expect(o, n * packed_len(i.tbits, width), "Output")
        self._call("fl_ctx_host_pack", i.tbits, width, n, i.ptr, o.ptr)

    def unpack(self, width: int, input, output) -> None:
        i, o = self._host(input, output)
        _check_width(width, o.tbits)
        n = _n_blocks_unpacked(o, "Output")
        _expect(i, n * packed_len(o.tbits, width), "Input")
        self._call("fl_ctx_host_unpack", o.tbits, width, n, i.ptr, o.ptr)

    def for_pack(self, width: int, input, reference, output) -> None:
        i, o = self._host(input, output)
        _check_width(width, i.tbits)
        n = _n_blocks_unpacked(i, "Input")
        _expect(o, n * packed_len(i.tbits, width), "Output")
        self._call("fl_ctx_host_for_pack", i.tbits, width, n, i.ptr, _ref_value(reference, i.tbits), o.ptr)

    def unfor_pack(self, width: int, input, reference, output) -> None:
        i, o = self._host(input, output)
        _check_width(width, o.tbits)
        n = _n_blocks_unpacked(o, "Output")
        _expect(i, n * packed_len(o.tbits, width), "Input")
        self._call("fl_ctx_host_unfor_pack", o.tbits, width, n, i.ptr, _ref_value(reference, o.tbits), o.ptr)

    def _delta3(self, name, input, base, output, width=None, packed_in=True):
        i, b, o = self._host(input, base, output)
        unp = o if packed_in else i
        n = _n_blocks_unpacked(unp, "Unpacked")
        _expect(b, n * (1024 // unp.tbits), "Base")
        if width is None:
            _expect(i, n * 1024, "Input")
            _expect(o, n * 1024, "Output")
            self._call(name, unp.tbits, n, i.ptr, b.ptr, o.ptr)
        else:
            _check_width(width, unp.tbits)
            _expect(i if packed_in else o, n * packed_len(unp.tbits, width), "Packed")
            self._call(name, unp.tbits, width, n, i.ptr, b.ptr, o.ptr)

    def delta(self, input, base, output) -> None:
        self._delta3("fl_ctx_host_delta", input, base, output)

    def undelta(self, input, base, output) -> None:
        self._delta3("fl_ctx_host_undelta", input, base, output)

    def undelta_pack(self, width: int, input, base, output) -> None:
        self._delta3("fl_ctx_host_undelta_pack", input, base, output, width)

    def undelta_pack_untranspose(self, width: int, input, base, output) -> None:
        self._delta3("fl_ctx_host_undelta_pack_untranspose", input, base, output, width)

    def transpose_delta_pack(self, width: int, input, base, output) -> None:
        self._delta3("fl_ctx_host_transpose_delta_pack", input, base, output, width, packed_in=False)

    def transpose(self, input, output) -> None:
        i, o = self._host(input, output)
        n = _n_blocks_unpacked(i, "Input")
        _expect(o, n * 1024, "Output")
        self._call("fl_ctx_host_transpose", i.tbits, n, i.ptr, o.ptr)

    def untranspose(self, input, output) -> None:
        i, o = self._host(input, output)
        n = _n_blocks_unpacked(i, "Input")
        _expect(o, n * 1024, "Output")
        self._call("fl_ctx_host_untranspose", i.tbits, n, i.ptr, o.ptr)

    def block_minmax(self, input, mins, maxs) -> None:
        i, lo, hi = self._host(input, mins, maxs)
        n = _n_blocks_unpacked(i, "Input")
        _expect(lo, n, "Mins")
        _expect(hi, n, "Maxs")
        self._call("fl_ctx_host_block_minmax", i.tbits, n, i.ptr, lo.ptr, hi.ptr)

    def filter_range(self, width: int, packed, reference, lo, hi, bitmap, counts=None) -> None:
        p, b, dev, n, cptr = Scan._bitmap_args(width, packed, bitmap, counts)
        if dev:
            raise FastLanesError(_lib.FL_ERR_NULL, "Context methods take host (numpy) arrays")
        self._call("fl_ctx_host_unpack_filter", p.tbits, width, n, p.ptr, _ref_value(reference, p.tbits),
                   _ref_value(lo, p.tbits), _ref_value(hi, p.tbits), b.ptr, cptr)

    def filter_range_delta(self, width: int, packed, base, lo, hi, bitmap, counts=None) -> None:
        p, b, dev, n, cptr = Scan._bitmap_args(width, packed, bitmap, counts)
        if dev:
            raise FastLanesError(_lib.FL_ERR_NULL, "Context methods take host (numpy) arrays")
        bs = _Arg(base, "base")
        _same_space(p, bs)
        _expect(bs, n * (1024 // p.tbits), "Base")
        self._call("fl_ctx_host_undelta_pack_filter", p.tbits, width, n, p.ptr, bs.ptr, _ref_value(lo, p.tbits),
                   _ref_value(hi, p.tbits), b.ptr, cptr)

    # ---- device tensors: the trivial block shard / gather (peer copies over NVLink) ------------------
    def scatter_blocks(self, src, elems_per_block: int, root: int = 0) -> list:
        """`src`: CUDA tensor on context device `root` holding whole blocks; returns one new tensor per shard, on that
        shard's device, holding blocks block_range(i) (fl_ctx_scatter_blocks)."""
        import torch

        n = src.numel() // elems_per_block
        devs = self.devices
        shards = []
        for i, d in enumerate(devs):
            b0, b1 = self.block_range(n, i)
            shards.append(torch.empty((b1 - b0) * elems_per_block, dtype=src.dtype, device=f"cuda:{d}"))
        torch.cuda.synchronize(src.device)
        ptrs = (ctypes.c_void_p * len(devs))(*[s.data_ptr() if s.numel() else None for s in shards])
        # empty shards are never dereferenced (their block range is empty)
        _lib.check(_lib.lib().fl_ctx_scatter_blocks(self._h, elems_per_block * src.element_size(), n, src.data_ptr(), root, ptrs))
        return shards

    def gather_blocks(self, shards: list, elems_per_block: int, root: int = 0):
        """Inverse of scatter_blocks: returns one tensor on context device `root` (fl_ctx_gather_blocks)."""
        import torch

        n = sum(s.numel() for s in shards) // elems_per_block
        devs = self.devices
        out = torch.empty(n * elems_per_block, dtype=shards[0].dtype, device=f"cuda:{devs[root]}")
        for s in shards:
            torch.cuda.synchronize(s.device)
        ptrs = (ctypes.c_void_p * len(devs))(*[s.data_ptr() if s.numel() else None for s in shards])
        _lib.check(_lib.lib().fl_ctx_gather_blocks(self._h, elems_per_block * out.element_size(), n, ptrs, root, out.data_ptr()))
        return out


_lib.lib()  # fail loudly at import time if the CUDA library is missing
