#!/usr/bin/env python
"""tools/refbench.py — the reference's own criterion bench SHAPES (benches/bitpacking.rs, benches/delta.rs,
benches/transpose.rs; they ship no recorded results) run through this library and through the CPU oracle.
These workloads are tiny (one block, or 1024 blocks = 2 MiB): on a GPU they are launch-/PCIe-latency bound and are
reported for comparability with anyone's `cargo bench`, not as a throughput claim."""
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import fastlanes_b200 as fl  # noqa: E402
from oracle import fl_oracle as oracle  # noqa: E402


def wall(fn, iters=200):
    for _ in range(10):
        fn()
    ts = []
    for _ in range(iters):
        t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
    return statistics.median(ts)


def gpu(fn, iters=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / iters * 1e-3


def line(name, secs, nbytes=None):
    extra = f"  {nbytes / secs / 1e9:8.2f} GB/s (unpacked bytes, criterion's Throughput::Bytes)" if nbytes else ""
    print(f"{name:62s} {secs * 1e6:10.2f} us{extra}", flush=True)


def main():
    fl.init(0)
    # benches/bitpacking.rs:10-65 — u16 W=3 single block
    W = 3
    v1 = np.full(1024, 3, dtype=np.uint16); p1 = np.zeros(192, dtype=np.uint16); u1 = np.zeros(1024, dtype=np.uint16)
    dv1 = torch.from_numpy(v1.view(np.int16)).cuda(); dp1 = torch.zeros(192, dtype=torch.int16, device="cuda"); du1 = torch.zeros(1024, dtype=torch.int16, device="cuda")
    line("pack 16->3, 1 block: CPU oracle (1 thread)", wall(lambda: oracle.run_raw(16, oracle.OP_PACK, W, 1, v1, p1)))
    line("pack 16->3, 1 block: fl_host_pack_u16 (H2D+kernel+D2H)", wall(lambda: fl.BitPacking.pack(W, v1, p1)))
    line("pack 16->3, 1 block: fl_pack_u16 (device, back-to-back launches)", gpu(lambda: fl.BitPacking.pack(W, dv1, dp1)))
    line("unpack 16<-3, 1 block: CPU oracle (1 thread)", wall(lambda: oracle.run_raw(16, oracle.OP_UNPACK, W, 1, p1, u1)))
    line("unpack 16<-3, 1 block: fl_host_unpack_u16", wall(lambda: fl.BitPacking.unpack(W, p1, u1)))
    line("unpack 16<-3, 1 block: fl_unpack_u16 (device)", gpu(lambda: fl.BitPacking.unpack(W, dp1, du1)))
    fl.BitPacking.pack(W, v1, p1)
    line("unpack-single 16<-3 x1024: CPU oracle", wall(lambda: [oracle.unpack_single(p1, W, i) for i in range(0, 1024, 64)], 50) * 64)
    gi = np.arange(1024, dtype=np.uint64); s1 = np.zeros(1024, dtype=np.uint16)
    line("unpack-single 16<-3 x1024: fl_host_unpack_gather_u16 (one call)", wall(lambda: fl.BitPacking.unpack_gather(W, p1, gi, s1)))
    # benches/bitpacking.rs:67-98 — throughput: 1024 blocks, values i % 8
    N = 1024
    vN = (np.arange(N * 1024) % 8).astype(np.uint16); pN = np.zeros(N * 192, dtype=np.uint16); uN = np.zeros(N * 1024, dtype=np.uint16)
    dvN = torch.from_numpy(vN.view(np.int16)).cuda(); dpN = torch.zeros(N * 192, dtype=torch.int16, device="cuda"); duN = torch.zeros(N * 1024, dtype=torch.int16, device="cuda")
    nb = N * 1024 * 2
    line("throughput/compress 1024 blocks: CPU oracle (1 thread)", wall(lambda: oracle.run_raw(16, oracle.OP_PACK, W, N, vN, pN), 50), nb)
    line("throughput/compress 1024 blocks: fl_host_pack_u16", wall(lambda: fl.BitPacking.pack(W, vN, pN), 50), nb)
    line("throughput/compress 1024 blocks: fl_pack_u16 (device)", gpu(lambda: fl.BitPacking.pack(W, dvN, dpN)), nb)
    line("throughput/decompress 1024 blocks: CPU oracle (1 thread)", wall(lambda: oracle.run_raw(16, oracle.OP_UNPACK, W, N, pN, uN), 50), nb)
    line("throughput/decompress 1024 blocks: fl_host_unpack_u16", wall(lambda: fl.BitPacking.unpack(W, pN, uN), 50), nb)
    line("throughput/decompress 1024 blocks: fl_unpack_u16 (device)", gpu(lambda: fl.BitPacking.unpack(W, dpN, duN)), nb)
    # benches/delta.rs:10-44 — u16 W=9, i/8 -> transpose -> delta -> pack; decode fused vs unfused
    W9 = 9
    vals = (np.arange(1024) // 8).astype(np.uint16)
    tr = oracle.transpose(vals); base = np.zeros(64, dtype=np.uint16); d = oracle.delta(tr, base); pk = oracle.pack(d, W9)
    out = np.zeros(1024, dtype=np.uint16); tmp = np.zeros(1024, dtype=np.uint16)
    dpk = torch.from_numpy(pk.view(np.int16)).cuda(); dbase = torch.zeros(64, dtype=torch.int16, device="cuda")
    dout = torch.zeros(1024, dtype=torch.int16, device="cuda"); dtmp = torch.zeros(1024, dtype=torch.int16, device="cuda")
    line("delta u16 fused, 1 block: CPU oracle", wall(lambda: oracle.run_raw(16, oracle.OP_UNDELTA_PACK, W9, 1, pk, out, base=base)), 2048)
    line("delta u16 unfused, 1 block: CPU oracle", wall(lambda: (oracle.run_raw(16, oracle.OP_UNPACK, W9, 1, pk, tmp), oracle.run_raw(16, oracle.OP_UNDELTA, 0, 1, tmp, out, base=base))), 2048)
    line("delta u16 fused, 1 block: fl_undelta_pack_u16 (device)", gpu(lambda: fl.Delta.undelta_pack(W9, dpk, dbase, dout)), 2048)
    line("delta u16 unfused, 1 block: fl_unpack_u16 + fl_undelta_u16 (device)", gpu(lambda: (fl.BitPacking.unpack(W9, dpk, dtmp), fl.Delta.undelta(dtmp, dbase, dout))), 2048)
    fl.Delta.undelta_pack(W9, dpk, dbase, dout)
    assert np.array_equal(dout.cpu().numpy().view(np.uint16), tr)
    # benches/transpose.rs:8-19 — transpose u16, one block
    dvals = torch.from_numpy(vals.view(np.int16)).cuda()
    line("transpose u16, 1 block: CPU oracle", wall(lambda: oracle.run_raw(16, oracle.OP_TRANSPOSE, 0, 1, vals, out)))
    line("transpose u16, 1 block: fl_transpose_u16 (device)", gpu(lambda: fl.Transpose.transpose(dvals, dout)))


if __name__ == "__main__":
    main()
