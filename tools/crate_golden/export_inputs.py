"""Step 1: write the golden INPUTS (tests/golden/fastlanes_golden.npz: values, base, reference per type) as raw
little-endian files for the Rust program.   python tools/crate_golden/export_inputs.py <in_dir>"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "fastlanes_golden.npz"))
    for tb in (8, 16, 32, 64):
        for what in ("values", "base", "reference"):
            a = np.ascontiguousarray(gold[f"u{tb}_{what}"]).astype(f"<u{tb // 8}")
            a.tofile(os.path.join(out_dir, f"u{tb}_{what}.bin"))
    print("wrote inputs to", out_dir)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tools", "crate_golden", "_in"))
