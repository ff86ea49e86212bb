"""not gpu: property tests (hypothesis) of the CPU oracle — algebraic identities the reference's design
implies (src/macros.rs:5-9: BitPack(Delta(Transpose(V))) == Delta+BitPack(Transpose(V))) on arbitrary data."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import np_closed_form as cf

DT = {8: np.uint8, 16: np.uint16, 32: np.uint32, 64: np.uint64}


@st.composite
def typed_case(draw):
    tb = draw(st.sampled_from([8, 16, 32, 64]))
    w = draw(st.integers(0, tb))
    n = draw(st.integers(1, 3))
    seed = draw(st.integers(0, 2**32 - 1))
    return tb, w, n, seed


def data(tb, n_elems, seed):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=n_elems * (tb // 8), dtype=np.uint8).view(DT[tb])


@settings(max_examples=60, deadline=None)
@given(typed_case())
def test_pack_unpack_identities(oracle, case):
    tb, w, n, seed = case
    v = data(tb, n * 1024, seed)
    m = DT[tb]((1 << w) - 1) if w < tb else DT[tb](~DT[tb](0))
    p = oracle.pack(v, w)
    assert p.size == n * 1024 * w // tb
    assert np.array_equal(oracle.unpack(p, w, n_blocks=n), v & m)      # unpack o pack = mask
    assert np.array_equal(oracle.pack(v & m, w), p)                    # pack ignores bits above W (macros.rs:73)
    assert np.array_equal(p, cf.pack(v, w))                            # streaming == closed form
    if w:
        bits = data(tb, n * 1024 * w // tb, seed ^ 0x5555)
        assert np.array_equal(oracle.pack(oracle.unpack(bits, w), w), bits)  # pack o unpack = id on packed bytes


@settings(max_examples=40, deadline=None)
@given(typed_case())
def test_for_and_delta_identities(oracle, case):
    tb, w, n, seed = case
    v = data(tb, n * 1024, seed)
    base = data(tb, n * (1024 // tb), seed + 1)
    ref = int(data(tb, 1, seed + 2)[0])
    # for_pack(v, r) == pack(v - r)  (ffor.rs:32-34);  unfor_pack(p, r) == unpack(p) + r  (ffor.rs:46-48)
    assert np.array_equal(oracle.for_pack(v, ref, w), oracle.pack((v - DT[tb](ref)).astype(DT[tb]), w))
    p = oracle.pack(v, w)
    assert np.array_equal(oracle.unfor_pack(p, ref, w, n_blocks=n), (oracle.unpack(p, w, n_blocks=n) + DT[tb](ref)).astype(DT[tb]))
    # undelta o delta = id; fused undelta_pack == undelta o unpack  (delta.rs:99-106)
    t = oracle.transpose(v)
    d = oracle.delta(t, base)
    assert np.array_equal(oracle.undelta(d, base), t)
    assert np.array_equal(oracle.untranspose(t), v)
    assert np.array_equal(oracle.undelta_pack(p, base, w, n_blocks=n), oracle.undelta(oracle.unpack(p, w, n_blocks=n), base))
    # linearity of undelta in the base: shifting every base by c shifts every output by c
    c = DT[tb](12345 % (1 << tb))
    assert np.array_equal(oracle.undelta(d, (base + c).astype(DT[tb])), (t + c).astype(DT[tb]))
