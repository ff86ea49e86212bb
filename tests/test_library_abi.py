"""not gpu: the C-ABI library loads and exports every symbol include/fastlanes_b200.h declares.
No compute call is made here (there is no GPU in the CPU test tier)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fastlanes_b200.h")).read()
    names = set(re.findall(r"\b(fl_[a-z_]+)\(", text.split("#define FL_DECLARE_TYPE")[0]))
    per_type = set(re.findall(r"\b(fl_[a-z_]+_)##SFX", text))
    for base in per_type:
        for sfx in ("u8", "u16", "u32", "u64"):
            names.add(base + sfx)
    return names


def test_library_exports_every_declared_symbol():
    from fastlanes_b200 import _lib

    L = ctypes.CDLL(_lib.LIB_PATH)
    declared = declared_symbols()
    assert len(declared) == 23 + 53 * 4  # 23 global (9 of them fl_ctx_*) + 53 per element type (14 of them fl_ctx_host_*)
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/fastlanes_b200.h but not exported"
    assert declared == set(_lib.exported_symbols())


def test_version_and_status_strings():
    import fastlanes_b200 as fl
    from fastlanes_b200 import _lib

    assert "sm_100a" in fl.version()
    assert _lib.lib().fl_status_string(1) == b"FL_ERR_WIDTH"
    assert _lib.lib().fl_status_string(0) == b"FL_OK"


def test_no_cpu_fallback_without_a_device():
    """On a box without a GPU every compute entry point fails loudly (FL_ERR_CUDA), it never computes on the CPU."""
    import numpy as np
    import pytest

    import fastlanes_b200 as fl

    if fl.device_count() > 0:
        pytest.skip("a CUDA device is present")
    values = np.arange(1024, dtype=np.uint32)
    packed = np.zeros(320, dtype=np.uint32)
    with pytest.raises(fl.FastLanesError) as e:
        fl.BitPacking.pack(10, values, packed)
    assert e.value.status == 5
    assert not packed.any()


def test_host_mirror_argument_checks_need_no_device():
    import numpy as np
    import pytest

    import fastlanes_b200 as fl

    v = np.zeros(1024, dtype=np.uint16)
    with pytest.raises(fl.FastLanesError) as e:
        fl.BitPacking.pack(17, v, np.zeros(1088, dtype=np.uint16))
    assert e.value.status == 1  # width > T: unreachable!() in the reference (bitpacking.rs:93)
    with pytest.raises(fl.FastLanesError) as e:
        fl.BitPacking.unpack(3, np.zeros(191, dtype=np.uint16), v)
    assert e.value.status == 2  # length debug_assert (bitpacking.rs:111)
    with pytest.raises(fl.FastLanesError) as e:
        fl.BitPacking.unpack_single(3, np.zeros(192, dtype=np.uint16), 1024)
    assert e.value.status == 3  # assert!(index < 1024) (bitpacking.rs:152)
    assert fl.Transpose.transpose_index(1) == 64 and fl.Transpose.transpose_index(16) == 32
    assert fl.FastLanes(32).LANES == 32 and fl.packed_len(16, 3) == 192


def test_header_is_plain_c():
    """The boundary is a C ABI: the header must compile as C99 with no extensions (what cgo / bindgen / cffi consume)."""
    import subprocess

    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c",
                        os.path.join(ROOT, "include", "fastlanes_b200.h")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_c_abi_argument_checks_need_no_device():
    """The device-pointer entry points validate their arguments before touching CUDA: width, block count, NULL and
    alignment errors come back as status codes (never a crash / unwind), with or without a GPU."""
    import ctypes

    from fastlanes_b200 import _lib

    L = _lib.lib()
    buf = ctypes.create_string_buffer(8192 + 64)
    base = (ctypes.addressof(buf) + 15) & ~15  # 16-byte aligned scratch (never dereferenced: every call below fails first)
    FL_ERR_WIDTH, FL_ERR_LEN, FL_ERR_ALIGN, FL_ERR_NULL = 1, 2, 4, 6
    assert L.fl_unpack_u32(33, 1, base, base + 4096, None) == FL_ERR_WIDTH          # bitpacking.rs:126 unreachable!()
    assert L.fl_pack_u8(9, 1, base, base + 4096, None) == FL_ERR_WIDTH
    assert L.fl_unpack_u32(8, (1 << 31) + 1, base, base + 4096, None) == FL_ERR_LEN
    assert L.fl_unpack_u32(8, 1, None, base, None) == FL_ERR_NULL
    assert L.fl_unpack_u32(8, 1, base + 4, base + 4096, None) == FL_ERR_ALIGN
    assert L.fl_undelta_pack_u16(3, 1, base, None, base + 4096, None) == FL_ERR_NULL  # missing base
    assert L.fl_unpack_filter_u32(33, 1, base, None, 0, 0, 1, base + 4096, None, None) == FL_ERR_WIDTH
    assert L.fl_unpack_filter_u32(8, 1, base, None, 0, 0, 1, None, None, None) == FL_ERR_NULL
    assert L.fl_undelta_pack_filter_u64(8, 1, base, None, 0, 1, base + 4096, None, None) == FL_ERR_NULL
    assert L.fl_for_pack_auto_u32(8, 1, base, None, None, base + 4096, None) == FL_ERR_NULL
    assert L.fl_unpack_u32(8, 0, None, None, None) == 0                               # empty batch: nothing to do
    assert L.fl_pack_u32(0, 5, base, None, None) == 0                                 # W = 0 packs to nothing (macros.rs:52)
    assert b"16-byte" in L.fl_last_error_string() or L.fl_last_error_string() is not None
