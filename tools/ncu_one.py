#!/usr/bin/env python
"""tools/ncu_one.py OP TBITS WIDTH [LOG2_BLOCKS] — launches one op of the shipped library a few times on device-resident
random data, for `ncu -k regex:<kernel> -s 2 -c 1 python tools/ncu_one.py ...` captures.  Measurement tool only."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from fastlanes_b200 import _lib  # noqa: E402

TDT = {8: torch.uint8, 16: torch.int16, 32: torch.int32, 64: torch.int64}


def main():
    op, tb, w = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    n = 1 << (int(sys.argv[4]) if len(sys.argv) > 4 else 20)
    unp = torch.empty(n * 1024, dtype=TDT[tb], device="cuda")
    pk = torch.empty(max(1, n * 1024 * w // tb), dtype=TDT[tb], device="cuda")
    for t in (unp, pk):
        v = t.view(torch.int32) if t.numel() * t.element_size() % 4 == 0 else t
        v.random_(-(1 << 31), (1 << 31) - 1) if v.dtype == torch.int32 else v.random_(0, 255)
    base = torch.zeros(max(n * (1024 // tb), n), dtype=TDT[tb], device="cuda")
    bm = torch.empty(n * 128, dtype=torch.uint8, device="cuda")
    cnt = torch.empty(n, dtype=torch.int32, device="cuda")
    sp = torch.cuda.current_stream().cuda_stream
    U, P, B = unp.data_ptr(), pk.data_ptr(), base.data_ptr()
    m = (1 << w) - 1
    calls = {
        "unpack": lambda: _lib.fn("fl_unpack", tb)(w, n, P, U, sp),
        "pack": lambda: _lib.fn("fl_pack", tb)(w, n, U, P, sp),
        "undelta_pack": lambda: _lib.fn("fl_undelta_pack", tb)(w, n, P, B, U, sp),
        "undelta_pack_untranspose": lambda: _lib.fn("fl_undelta_pack_untranspose", tb)(w, n, P, B, U, sp),
        "transpose_delta_pack": lambda: _lib.fn("fl_transpose_delta_pack", tb)(w, n, U, B, P, sp),
        "for_pack_auto": lambda: _lib.fn("fl_for_pack_auto", tb)(w, n, U, B, None, P, sp),
        "for_pack": lambda: _lib.fn("fl_for_pack", tb)(w, n, U, 12345 % (1 << tb), P, sp),
        "unfor_pack": lambda: _lib.fn("fl_unfor_pack", tb)(w, n, P, 12345 % (1 << tb), U, sp),
        "undelta_pack_filter": lambda: _lib.fn("fl_undelta_pack_filter", tb)(w, n, P, B, ((1 << tb) - 1) // 4, ((1 << tb) - 1) // 2, bm.data_ptr(), cnt.data_ptr(), sp),
        "unpack_filter": lambda: _lib.fn("fl_unpack_filter", tb)(w, n, P, None, 0, m // 4, m // 2, bm.data_ptr(), cnt.data_ptr(), sp),
    }
    if op == "unpack_select":  # value-independent bitmap, ~25 % selected, + the exclusive prefix of the block counts
        bm.view(torch.int32).random_(-(1 << 31), (1 << 31) - 1)
        bm2 = bm.clone(); bm2.view(torch.int32).random_(-(1 << 31), (1 << 31) - 1)
        bm &= bm2
        c64 = torch.tensor([bin(i).count("1") for i in range(256)], dtype=torch.int64, device="cuda")[bm.long()].view(n, 128).sum(1)
        offs = torch.cumsum(c64, 0) - c64
        sel_out = torch.empty(int(c64.sum().item()) + 16, dtype=TDT[tb], device="cuda")
        calls["unpack_select"] = lambda: _lib.fn("fl_unpack_select", tb)(w, n, P, None, 7, bm.data_ptr(), offs.data_ptr(), sel_out.data_ptr(), sp)
    for _ in range(5):
        assert calls[op]() == 0
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
