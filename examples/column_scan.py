#!/usr/bin/env python
"""examples/column_scan.py — the path a columnar engine takes through the library, end to end on one GPU:

  encode   a sorted u64 timestamp column with the reference's chain transpose -> delta -> pack (src/delta.rs:88-95),
           and a u32 measurement column with FoR (reference = each block's own minimum, src/ffor.rs:24-36), one pass each
  scan     WHERE t0 <= ts <= t1      : fused undelta_pack + predicate, bitmap in original row order
           AND   lo <= value <= hi   : fused unfor_pack + predicate on the second column
  take     SELECT value              : compaction of the rows both predicates keep
  check    against numpy on the raw columns

Run on a B200: python examples/column_scan.py   (prints one summary line; exits non-zero on a mismatch)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import fastlanes_b200 as fl  # noqa: E402


def main(n_blocks: int = 1 << 12) -> int:
    dev = torch.device("cuda")
    rng = np.random.default_rng(1)
    n = n_blocks * 1024
    # raw columns
    steps = rng.integers(0, 1 << 18, size=n, dtype=np.uint64)
    ts = np.uint64(1_700_000_000_000) + np.cumsum(steps, dtype=np.uint64)
    value = (rng.integers(0, 1 << 14, size=n, dtype=np.uint64) + 50_000).astype(np.uint32)

    # ---- encode -----------------------------------------------------------------------------------------------
    W_TS, W_VAL = 18, 14
    prev = np.concatenate([[ts[0] - steps[0]], ts[:-1]])
    base = np.ascontiguousarray(prev.reshape(n_blocks, 1024)[:, ::64].reshape(-1))   # u64: lane l's run starts at row 64*l
    d_ts, d_base = torch.from_numpy(ts.view(np.int64)).to(dev), torch.from_numpy(base.view(np.int64)).to(dev)
    ts_packed = torch.empty(n_blocks * 16 * W_TS, dtype=torch.int64, device=dev)
    fl.Delta.transpose_delta_pack(W_TS, d_ts, d_base, ts_packed)
    d_val = torch.from_numpy(value.view(np.int32)).to(dev)
    val_packed = torch.empty(n_blocks * 32 * W_VAL, dtype=torch.int32, device=dev)
    val_refs = torch.empty(n_blocks, dtype=torch.int32, device=dev)
    val_spans = torch.empty(n_blocks, dtype=torch.int32, device=dev)
    fl.FoR.for_pack_auto(W_VAL, d_val, val_refs, val_packed, val_spans)
    assert int(val_spans.max().item()) < (1 << W_VAL), "W_VAL too small for this column"
    packed_bytes = ts_packed.numel() * 8 + d_base.numel() * 8 + val_packed.numel() * 4 + val_refs.numel() * 4

    # ---- scan -------------------------------------------------------------------------------------------------
    t0, t1 = int(ts[n // 4]), int(ts[n // 2])
    lo, hi = 52_000, 60_000
    bm_ts = torch.empty(n_blocks * 128, dtype=torch.uint8, device=dev)
    bm_val = torch.empty(n_blocks * 128, dtype=torch.uint8, device=dev)
    fl.Scan.filter_range_delta(W_TS, ts_packed, d_base, t0, t1, bm_ts)          # bitmap in ORIGINAL row order
    fl.Scan.filter_range(W_VAL, val_packed, val_refs, lo, hi, bm_val)            # per-block FoR references
    keep = bm_ts & bm_val

    # ---- take -------------------------------------------------------------------------------------------------
    popc = torch.tensor([bin(i).count("1") for i in range(256)], dtype=torch.int64, device=dev)
    counts = popc[keep.long()].view(n_blocks, 128).sum(1)
    offsets = torch.cumsum(counts, 0) - counts
    total = int(counts.sum().item())
    out = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
    fl.Scan.select(W_VAL, val_packed, val_refs, keep, offsets, out)

    # ---- check ------------------------------------------------------------------------------------------------
    sel = (ts >= np.uint64(t0)) & (ts <= np.uint64(t1)) & (value >= lo) & (value <= hi)
    ok = total == int(sel.sum()) and np.array_equal(out[:total].cpu().numpy().view(np.uint32), value[sel])
    print(f"column_scan: {n} rows, {packed_bytes / (n * 12):.3f} of the raw bytes, {total} rows selected, "
          f"{'OK' if ok else 'MISMATCH'}")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
