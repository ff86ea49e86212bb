#!/usr/bin/env python
"""tools/select_sweep.py [TBITS ...] — fl_unpack_select at 25 % selectivity, every width of the given types; prints
`T W us GB/s`.  profiles/select_sweep_r02.txt was produced with a development build in which FLB_SELECT / FLB_SELECT_NB chose
the kernel variant and the blocks per warp (tools/gpu_r02_l.sh); the shipped library instantiates only the chosen form per
type / width (select_nb() in fl_codec_inst.cu).  Measurement tool only."""
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from fastlanes_b200 import _lib  # noqa: E402

TDT = {8: torch.uint8, 16: torch.int16, 32: torch.int32, 64: torch.int64}


def main():
    types = [int(a) for a in sys.argv[1:]] or [8, 16, 32, 64]
    sp = torch.cuda.current_stream().cuda_stream
    for tb in types:
        n = (1 << 31) // (128 * tb)  # 2 GiB unpacked
        pk = torch.empty(n * 1024, dtype=TDT[tb], device="cuda")
        pk.view(torch.int32).random_(-(1 << 31), (1 << 31) - 1)
        bm = torch.empty(n * 128, dtype=torch.uint8, device="cuda")
        bm.view(torch.int32).random_(-(1 << 31), (1 << 31) - 1)
        bm2 = bm.clone(); bm2.view(torch.int32).random_(-(1 << 31), (1 << 31) - 1)
        bm &= bm2
        del bm2
        c64 = torch.tensor([bin(i).count("1") for i in range(256)], dtype=torch.int64, device="cuda")[bm.long()].view(n, 128).sum(1)
        offs = torch.cumsum(c64, 0) - c64
        total = int(c64.sum().item())
        out = torch.empty(total + 16, dtype=TDT[tb], device="cuda")
        fn = _lib.fn("fl_unpack_select", tb)
        for w in range(0, tb + 1):
            call = lambda: fn(w, n, pk.data_ptr(), None, 7, bm.data_ptr(), offs.data_ptr(), out.data_ptr(), sp)
            assert call() == 0
            for _ in range(2):
                call()
            torch.cuda.synchronize()
            ts = []
            for _ in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); call(); b.record(); b.synchronize()
                ts.append(a.elapsed_time(b))
            ms = statistics.median(ts)
            gb = (n * (128 * w + 128 + 8) + total * (tb // 8)) / 1e9
            print(f"{tb} {w} {ms * 1e3:.1f} {gb / (ms * 1e-3):.1f}", flush=True)
        del pk, bm, c64, offs, out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
