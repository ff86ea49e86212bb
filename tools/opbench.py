#!/usr/bin/env python
"""tools/opbench.py — every op x type through the shipped C ABI (device family), CUDA-event timed.
Prints a table (algorithmic GB/s, Gint/s) and writes gpurun_out/opbench.json.  Development/measurement tool."""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from fastlanes_b200 import _lib  # noqa: E402

TDT = {8: torch.uint8, 16: torch.int16, 32: torch.int32, 64: torch.int64}


def timeit(fn, iters=7):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    only = set(args[0].split(",")) if args else None  # e.g. "transpose,untranspose"
    types = (8, 16, 32, 64)
    if "--types" in sys.argv:  # e.g. --types 8,16
        types = tuple(int(t) for t in sys.argv[sys.argv.index("--types") + 1].split(","))
        only = set(sys.argv[1].split(",")) if not sys.argv[1].startswith("--") else None
    lg_bytes = 32  # 4 GiB unpacked per type
    rows = []
    sp = torch.cuda.current_stream().cuda_stream
    for tb in types:
        n = (1 << lg_bytes) // (128 * tb)
        unp = torch.empty(n * 1024, dtype=TDT[tb], device="cuda")
        unp.view(torch.int32).random_(-(1 << 31), (1 << 31) - 1)
        pk = torch.empty(n * 1024, dtype=TDT[tb], device="cuda")
        pk.view(torch.int32).random_(-(1 << 31), (1 << 31) - 1)
        base = torch.empty(n * (1024 // tb), dtype=TDT[tb], device="cuda")
        base.view(torch.int32).random_(-(1 << 31), (1 << 31) - 1)
        U, P, B = unp.data_ptr(), pk.data_ptr(), base.data_ptr()
        widths = sorted({1, tb // 4, tb // 2 + 1, tb - 3, tb})

        def rec(op, w, bytes_per_block, fn):
            if only and op not in only:
                return
            st = fn()
            assert st == 0, (op, tb, w, st)
            ms = timeit(fn)
            rows.append({"op": op, "T": tb, "W": w, "blocks": n, "us": round(ms * 1e3, 1),
                         "GBps": round(n * bytes_per_block / (ms * 1e-3) / 1e9, 1),
                         "Gints": round(n * 1024 / (ms * 1e-3) / 1e9, 1)})
            r = rows[-1]
            print(f"{op:26s} u{tb:<2d} W={w:<2d} {r['us']:9.1f} us {r['GBps']:8.1f} GB/s {r['Gints']:8.1f} Gint/s", flush=True)

        for w in widths:
            rec("unpack", w, 128 * (w + tb), lambda: _lib.fn("fl_unpack", tb)(w, n, P, U, sp))
            rec("pack", w, 128 * (w + tb), lambda: _lib.fn("fl_pack", tb)(w, n, U, P, sp))
            rec("unfor_pack", w, 128 * (w + tb), lambda: _lib.fn("fl_unfor_pack", tb)(w, n, P, 12345 % (1 << tb), U, sp))
            rec("for_pack", w, 128 * (w + tb), lambda: _lib.fn("fl_for_pack", tb)(w, n, U, 12345 % (1 << tb), P, sp))
            rec("unpack_cwida", w, 128 * (w + tb), lambda: _lib.fn("fl_unpack_cwida", tb)(w, n, P, U, sp))
            rec("pack_cwida", w, 128 * (w + tb), lambda: _lib.fn("fl_pack_cwida", tb)(w, n, U, P, sp))
            rec("for_pack_auto", w, 128 * (w + tb) + 2 * (tb // 8),
                lambda: _lib.fn("fl_for_pack_auto", tb)(w, n, U, base.data_ptr(), base.data_ptr() + n * (tb // 8), P, sp))
            rec("undelta_pack", w, 128 * (w + tb + 1), lambda: _lib.fn("fl_undelta_pack", tb)(w, n, P, B, U, sp))
            if True:  # fused chains: same algorithmic bytes as undelta_pack / pack + bases
                rec("undelta_pack_untranspose", w, 128 * (w + tb + 1), lambda: _lib.fn("fl_undelta_pack_untranspose", tb)(w, n, P, B, U, sp))
                rec("transpose_delta_pack", w, 128 * (w + tb + 1), lambda: _lib.fn("fl_transpose_delta_pack", tb)(w, n, U, B, P, sp))
        # fused scan: filter = decode + range predicate -> 128-byte bitmap + count per block;
        # select = decode + compaction of the selected values (here ~25 % selected by a value-independent bitmap)
        if not only or (only & {"unpack_filter", "unpack_select_25pct", "undelta_pack_filter"}):
            full = (1 << tb) - 1
            bm = torch.empty(n * 128, dtype=torch.uint8, device="cuda")
            cnt = torch.empty(n, dtype=torch.int32, device="cuda")
            for w in widths:
                lo, hi = ((1 << w) - 1) // 4, ((1 << w) - 1) // 2
                rec("unpack_filter", w, 128 * w + 128 + 4,
                    lambda: _lib.fn("fl_unpack_filter", tb)(w, n, P, None, 0, lo, hi, bm.data_ptr(), cnt.data_ptr(), sp))
            for w in widths:
                rec("undelta_pack_filter", w, 128 * w + 128 + 128 + 4,
                    lambda: _lib.fn("fl_undelta_pack_filter", tb)(w, n, P, B, full // 4, full // 2, bm.data_ptr(), cnt.data_ptr(), sp))
            bm.view(torch.int32).random_(-(1 << 31), (1 << 31) - 1)
            bm2 = bm.clone(); bm2.view(torch.int32).random_(-(1 << 31), (1 << 31) - 1)
            bm &= bm2  # density 1/4
            del bm2
            c64 = torch.tensor([bin(i).count("1") for i in range(256)], dtype=torch.int64, device="cuda")[bm.long()].view(n, 128).sum(1)
            offs = torch.cumsum(c64, 0) - c64
            total = int(c64.sum().item())
            sel_out = torch.empty(total + 16, dtype=TDT[tb], device="cuda")
            for w in widths:
                rec("unpack_select_25pct", w, 128 * w + 128 + 8 + (tb // 8) * 256,
                    lambda: _lib.fn("fl_unpack_select", tb)(w, n, P, None, 7, bm.data_ptr(), offs.data_ptr(), sel_out.data_ptr(), sp))
            del bm, cnt, c64, offs, sel_out
        rec("delta", 0, 128 * (2 * tb + 1), lambda: _lib.fn("fl_delta", tb)(n, U, B, P, sp))
        rec("undelta", 0, 128 * (2 * tb + 1), lambda: _lib.fn("fl_undelta", tb)(n, U, B, P, sp))
        mn = torch.empty(n, dtype=TDT[tb], device="cuda"); mx = torch.empty(n, dtype=TDT[tb], device="cuda")
        rec("block_minmax", 0, 128 * tb, lambda: _lib.fn("fl_block_minmax", tb)(n, U, mn.data_ptr(), mx.data_ptr(), sp))
        rec("transpose", 0, 256 * tb, lambda: _lib.fn("fl_transpose", tb)(n, U, P, sp))
        rec("untranspose", 0, 256 * tb, lambda: _lib.fn("fl_untranspose", tb)(n, U, P, sp))
        if only and "unpack_gather" not in only:
            continue
        # batched unpack_single: 2^24 random queries
        nq = 1 << 24
        gi = torch.randint(0, n * 1024, (nq,), dtype=torch.int64, device="cuda")
        qo = torch.empty(nq, dtype=TDT[tb], device="cuda")
        w = tb // 2 + 1
        fn = lambda: _lib.fn("fl_unpack_gather", tb)(w, n, P, gi.data_ptr(), nq, qo.data_ptr(), None, sp)
        assert fn() == 0
        ms = timeit(fn)
        rows.append({"op": "unpack_gather", "T": tb, "W": w, "queries": nq, "us": round(ms * 1e3, 1), "Gq_per_s": round(nq / (ms * 1e-3) / 1e9, 2)})
        print(f"unpack_gather  u{tb:<2d} W={w:<2d} {ms * 1e3:9.1f} us {nq / (ms * 1e-3) / 1e9:8.2f} Gqueries/s", flush=True)
        del unp, pk, base, gi, qo
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "opbench.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
