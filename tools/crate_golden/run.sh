#!/bin/bash
# Regenerate tests/golden/crate_outputs.npz from the REAL crate.  Needs: rustup (nightly-2024-06-19 is selected by
# rust-toolchain.toml), network access to crates.io (or a vendored `fastlanes = { path = ... }`), python + numpy.
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
python "$here/export_inputs.py" "$here/_in"
(cd "$here" && cargo run --release -- "$here/_in" "$here/_out")
python "$here/import_outputs.py" "$here/_out"
(cd "$here/../.." && python -m pytest tests/test_golden.py -q -m "not gpu" -s)
