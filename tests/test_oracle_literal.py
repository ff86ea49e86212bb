"""not gpu: differential test of the C++ oracle (oracle/fl_oracle_kernels.hpp — the checker every GPU parity test uses)
against oracle/literal_rs.py, a statement-by-statement Python transcription of the reference's Rust macros
(src/macros.rs:12-173, src/bitpacking.rs:65-232, src/ffor.rs:24-50, src/delta.rs:24-63, src/transpose.rs:11-36).

Every (T, W) pair is checked on full-range random data (so the mask truncation of macros.rs:73 is exercised), plus a
hypothesis sweep over seeds and special bit patterns.  Three independently written restatements now agree (C++ streaming,
numpy closed form, literal Python); none has been diffed against crate-executed bytes — see tools/crate_golden/."""
import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import literal_rs as rs

DT = {8: np.uint8, 16: np.uint16, 32: np.uint32, 64: np.uint64}
ALL_TW = [(tb, w) for tb in (8, 16, 32, 64) for w in range(tb + 1)]


def rand(tb, n, seed):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=n * (tb // 8), dtype=np.uint8).view(DT[tb])


def as_list(a):
    return [int(x) for x in a]


def check_all_ops(oracle, tb, w, values, packed_bits, base, reference):
    """One block; values: 1024 elements, packed_bits: 1024*w/tb elements (arbitrary bit patterns), base: LANES."""
    dt = DT[tb]
    v, pb, bs = as_list(values), as_list(packed_bits), as_list(base)
    assert as_list(oracle.pack(values, w)) == rs.pack(tb, w, v), "pack"
    assert as_list(oracle.unpack(packed_bits, w, n_blocks=1)) == rs.unpack(tb, w, pb), "unpack"
    assert as_list(oracle.for_pack(values, reference, w)) == rs.for_pack(tb, w, v, reference), "for_pack"
    assert as_list(oracle.unfor_pack(packed_bits, reference, w, n_blocks=1)) == rs.unfor_pack(tb, w, pb, reference), "unfor_pack"
    assert as_list(oracle.undelta_pack(packed_bits, base, w, n_blocks=1)) == rs.undelta_pack(tb, w, pb, bs), "undelta_pack"
    # unpack_single on a spread of indices (all 1024 in the exhaustive test below for a few widths)
    for idx in (0, 1, 15, 16, 127, 128, 511, 512, 1000, 1023):
        assert oracle.unpack_single(packed_bits, w, idx) == rs.unpack_single(tb, w, pb, idx), f"unpack_single {idx}"
    del dt


@pytest.mark.parametrize("tb,w", ALL_TW, ids=[f"u{t}w{w}" for t, w in ALL_TW])
def test_every_type_and_width_matches_literal_transcription(oracle, tb, w):
    seed = tb * 1000 + w
    values = rand(tb, 1024, seed)
    packed_bits = rand(tb, 1024 * w // tb, seed + 1) if w else np.zeros(0, dtype=DT[tb])
    base = rand(tb, 1024 // tb, seed + 2)
    reference = int(rand(tb, 1, seed + 3)[0])
    check_all_ops(oracle, tb, w, values, packed_bits, base, reference)


@pytest.mark.parametrize("tb", [8, 16, 32, 64])
def test_delta_transpose_match_literal_transcription(oracle, tb):
    values, base = rand(tb, 1024, tb), rand(tb, 1024 // tb, tb + 7)
    v, bs = as_list(values), as_list(base)
    assert as_list(oracle.delta(values, base)) == rs.delta(tb, v, bs)
    assert as_list(oracle.undelta(values, base)) == rs.undelta(tb, v, bs)
    assert as_list(oracle.transpose(values)) == rs.transpose(v)
    assert as_list(oracle.untranspose(values)) == rs.untranspose(v)


@pytest.mark.parametrize("tb,w", [(8, 3), (16, 15), (32, 10), (32, 31), (64, 33), (64, 64)])
def test_unpack_single_all_indices(oracle, tb, w):
    packed_bits = rand(tb, 1024 * w // tb, 99 + w)
    pb = as_list(packed_bits)
    whole = rs.unpack(tb, w, pb)
    for idx in range(1024):
        got = oracle.unpack_single(packed_bits, w, idx)
        assert got == rs.unpack_single(tb, w, pb, idx) == whole[idx]


PATTERNS = ("random", "zeros", "ones", "alternating", "msb", "lsb")


def patterned(tb, n, seed, pattern):
    dt = DT[tb]
    full = (1 << tb) - 1
    if pattern == "random":
        return rand(tb, n, seed)
    if pattern == "zeros":
        return np.zeros(n, dtype=dt)
    if pattern == "ones":
        return np.full(n, full, dtype=dt)
    if pattern == "alternating":
        return np.full(n, int("10" * (tb // 2), 2), dtype=dt)
    if pattern == "msb":
        return np.full(n, 1 << (tb - 1), dtype=dt)
    return np.ones(n, dtype=dt)


@st.composite
def case(draw):
    tb = draw(st.sampled_from([8, 16, 32, 64]))
    w = draw(st.integers(0, tb))
    seed = draw(st.integers(0, 2**31 - 1))
    pv = draw(st.sampled_from(PATTERNS))
    pp = draw(st.sampled_from(PATTERNS))
    ref = draw(st.integers(0, (1 << tb) - 1))
    return tb, w, seed, pv, pp, ref


@settings(max_examples=40, deadline=None)
@given(case())
def test_hypothesis_differential(oracle, c):
    tb, w, seed, pv, pp, ref = c
    values = patterned(tb, 1024, seed, pv)
    packed_bits = patterned(tb, 1024 * w // tb, seed + 1, pp)
    base = patterned(tb, 1024 // tb, seed + 2, pv)
    check_all_ops(oracle, tb, w, values, packed_bits, base, ref)
