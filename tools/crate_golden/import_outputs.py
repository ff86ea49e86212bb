"""Step 3: collect the files the Rust program wrote into tests/golden/crate_outputs.npz (commit that file).
tests/test_golden.py::test_crate_outputs_pin_the_oracle consumes it when present and prints "parity pinned".
   python tools/crate_golden/import_outputs.py <out_dir>"""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main(out_dir):
    arrays = {}
    for name in sorted(os.listdir(out_dir)):
        m = re.match(r"u(8|16|32|64)_(.+)\.bin$", name)
        if not m:
            continue
        tb = int(m.group(1))
        arrays[f"u{tb}_{m.group(2)}"] = np.fromfile(os.path.join(out_dir, name), dtype=f"<u{tb // 8}")
    assert len(arrays) == 4 * 4 + 7 * (9 + 17 + 33 + 65), f"unexpected number of output files: {len(arrays)}"
    path = os.path.join(ROOT, "tests", "golden", "crate_outputs.npz")
    np.savez_compressed(path, **arrays)
    print("wrote", path, len(arrays), "arrays")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tools", "crate_golden", "_out"))
