"""ctypes loader for libfastlanes_b200.so (the C-ABI product library, include/fastlanes_b200.h).

There is no fallback: if the library is missing or does not load, importing the package raises.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libfastlanes_b200.so")

FL_OK, FL_ERR_WIDTH, FL_ERR_LEN, FL_ERR_INDEX, FL_ERR_ALIGN, FL_ERR_CUDA, FL_ERR_NULL, FL_ERR_UNSUPPORTED = range(8)

TYPE_SUFFIXES = {8: "u8", 16: "u16", 32: "u32", 64: "u64"}
_CT = {8: ctypes.c_uint8, 16: ctypes.c_uint16, 32: ctypes.c_uint32, 64: ctypes.c_uint64}

# name -> (argument kinds) ; 'w' width, 'n' size_t, 'p' pointer, 'r' element by value, 's' stream, 'c' fl_ctx*
_PER_TYPE = {
    "fl_pack": "wnpps", "fl_host_pack": "wnpp",
    "fl_unpack": "wnpps", "fl_host_unpack": "wnpp",
    "fl_unpack_gather": "wnppnpps", "fl_host_unpack_gather": "wnppnp", "fl_host_unpack_single": "wpnp",
    "fl_for_pack": "wnprps", "fl_for_pack_refs": "wnppps", "fl_host_for_pack": "wnprp",
    "fl_unfor_pack": "wnprps", "fl_unfor_pack_refs": "wnppps", "fl_host_unfor_pack": "wnprp",
    "fl_delta": "nppps".replace(" ", ""), "fl_host_delta": "nppp",
    "fl_undelta": "nppps", "fl_host_undelta": "nppp",
    "fl_undelta_pack": "wnppps", "fl_host_undelta_pack": "wnppp",
    "fl_undelta_pack_untranspose": "wnppps", "fl_host_undelta_pack_untranspose": "wnppp",
    "fl_transpose_delta_pack": "wnppps", "fl_host_transpose_delta_pack": "wnppp",
    "fl_block_minmax": "nppps".replace(" ", ""), "fl_host_block_minmax": "nppp",
    "fl_pack_cwida": "wnpps", "fl_unpack_cwida": "wnpps", "fl_for_pack_cwida": "wnprps", "fl_unfor_pack_cwida": "wnprps",
    "fl_for_pack_auto": "wnpppps",
    "fl_unpack_filter": "wnpprrrpps", "fl_host_unpack_filter": "wnprrrpp", "fl_unpack_select": "wnpprppps",
    "fl_undelta_pack_filter": "wnpprrpps", "fl_host_undelta_pack_filter": "wnpprrpp",
    "fl_transpose": "npps", "fl_untranspose": "npps",
    "fl_host_transpose": "npp", "fl_host_untranspose": "npp",
    # context family (multi-device, block-sharded): fl_host_<op> with a leading fl_ctx*
    "fl_ctx_host_pack": "cwnpp", "fl_ctx_host_unpack": "cwnpp",
    "fl_ctx_host_for_pack": "cwnprp", "fl_ctx_host_unfor_pack": "cwnprp",
    "fl_ctx_host_delta": "cnppp", "fl_ctx_host_undelta": "cnppp",
    "fl_ctx_host_undelta_pack": "cwnppp", "fl_ctx_host_undelta_pack_untranspose": "cwnppp",
    "fl_ctx_host_transpose_delta_pack": "cwnppp",
    "fl_ctx_host_transpose": "cnpp", "fl_ctx_host_untranspose": "cnpp",
    "fl_ctx_host_block_minmax": "cnppp",
    "fl_ctx_host_unpack_filter": "cwnprrrpp", "fl_ctx_host_undelta_pack_filter": "cwnpprrpp",
}
_GLOBAL = ["fl_version", "fl_last_error_string", "fl_status_string", "fl_device_count", "fl_device_numa_node", "fl_init", "fl_host_configure",
           "fl_host_alloc", "fl_host_free", "fl_host_buffer_node", "fl_host_register", "fl_host_unregister", "fl_host_copy_probe",
           "fl_shutdown", "fl_ctx_create", "fl_ctx_destroy", "fl_ctx_device_count", "fl_ctx_device", "fl_ctx_block_range",
           "fl_ctx_host_alloc", "fl_ctx_host_copy_probe", "fl_ctx_scatter_blocks", "fl_ctx_gather_blocks"]


def exported_symbols() -> list[str]:
    """Every symbol include/fastlanes_b200.h declares."""
    names = list(_GLOBAL)
    for base in _PER_TYPE:
        names += [f"{base}_{sfx}" for sfx in TYPE_SUFFIXES.values()]
    return names


class FastLanesError(RuntimeError):
    """Raised where the reference panics (width > T, index >= 1024, wrong lengths) or CUDA fails."""

    def __init__(self, status: int, message: str):
        super().__init__(message)
        self.status = status


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `make lib` (or __graft_entry__.build()). "
            "fastlanes_b200 has no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    for name in ("fl_version", "fl_last_error_string"):
        getattr(L, name).restype = ctypes.c_char_p
    L.fl_status_string.restype = ctypes.c_char_p
    L.fl_status_string.argtypes = [ctypes.c_int]
    L.fl_device_count.restype = ctypes.c_int
    L.fl_init.argtypes = [ctypes.c_int]
    L.fl_device_numa_node.restype = ctypes.c_int
    L.fl_device_numa_node.argtypes = [ctypes.c_int]
    L.fl_host_configure.argtypes = [ctypes.c_size_t, ctypes.c_int]
    L.fl_host_alloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t]
    L.fl_host_free.argtypes = [ctypes.c_void_p]
    L.fl_host_register.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    L.fl_host_unregister.argtypes = [ctypes.c_void_p]
    L.fl_host_buffer_node.restype = ctypes.c_int
    L.fl_host_buffer_node.argtypes = [ctypes.c_void_p]
    L.fl_host_copy_probe.argtypes = [ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
    L.fl_ctx_create.argtypes = [ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
    L.fl_ctx_destroy.argtypes = [ctypes.c_void_p]
    L.fl_ctx_device_count.restype = ctypes.c_int
    L.fl_ctx_device_count.argtypes = [ctypes.c_void_p]
    L.fl_ctx_device.restype = ctypes.c_int
    L.fl_ctx_device.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.fl_ctx_block_range.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(ctypes.c_size_t),
                                     ctypes.POINTER(ctypes.c_size_t)]
    L.fl_ctx_host_alloc.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p)]
    L.fl_ctx_host_copy_probe.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p,
                                         ctypes.c_void_p]
    L.fl_ctx_scatter_blocks.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int,
                                        ctypes.POINTER(ctypes.c_void_p)]
    L.fl_ctx_gather_blocks.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p),
                                       ctypes.c_int, ctypes.c_void_p]
    for base, kinds in _PER_TYPE.items():
        for tb, sfx in TYPE_SUFFIXES.items():
            fn = getattr(L, f"{base}_{sfx}")
            fn.restype = ctypes.c_int
            fn.argtypes = [{"w": ctypes.c_uint, "n": ctypes.c_size_t, "p": ctypes.c_void_p, "r": _CT[tb],
                            "s": ctypes.c_void_p, "c": ctypes.c_void_p}[k] for k in kinds]
    _lib = L
    return L


def check(status: int) -> None:
    if status != FL_OK:
        L = lib()
        raise FastLanesError(status, f"{L.fl_status_string(status).decode()}: {L.fl_last_error_string().decode()}")


def fn(base: str, tbits: int):
    return getattr(lib(), f"{base}_{TYPE_SUFFIXES[tbits]}")
