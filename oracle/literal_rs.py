"""literal_rs.py — a THIRD, statement-by-statement transcription of the reference's Rust into pure Python.
TEST INFRASTRUCTURE (only tests/ may import it); never on a product path.

Purpose (VERDICT r01 "next round" item 7): the C++ oracle (fl_oracle_kernels.hpp) is a restatement and the numpy
oracle (np_closed_form.py) is derived from `unpack_single`'s closed form.  This file is deliberately dumb: every Rust
statement of the macros becomes one Python statement, in the same order, with the same names, operating on Python
ints that are truncated to T bits exactly where Rust's fixed-width integer would truncate.  Rust semantics relied on:

  * operator precedence:  `*  /  %`  bind tighter than  `+  -`,  which bind tighter than  `<<  >>`.  Hence
        `src << (row * $W) % T`     is  src << ((row * W) % T)        (src/macros.rs:79)
        `src >> $W - remaining_bits` is  src >> (W - remaining_bits)   (src/macros.rs:92)
  * `<<` on an unsigned T-bit integer discards the bits shifted out (no wrap, no panic while the amount < T);
  * `wrapping_add` / `wrapping_sub` are mod 2^T;
  * `(1 << W) - 1` is evaluated in T (only reached with 0 < W < T, so it never overflows).

File:line citations are relative to /root/reference.  PARITY NOTE: like the other oracles this has never been diffed
against crate-EXECUTED output (no Rust toolchain in this image) — tools/crate_golden/ is the recipe that closes that.
"""
from __future__ import annotations

FL_ORDER = [0, 4, 2, 6, 1, 5, 3, 7]  # src/lib.rs:22


def _lanes(T: int) -> int:
    return 1024 // T  # src/lib.rs:26


def index(row: int, lane: int) -> int:
    """src/macros.rs:20-24 (repeated :46-50, :112-116)."""
    o = row // 8
    s = row % 8
    return (FL_ORDER[o] * 16) + (s * 128) + lane


def pack_lane(T: int, W: int, packed: list, lane: int, kernel) -> None:
    """`pack!` (src/macros.rs:35-97) for one lane; `kernel(idx)` is the spliced closure."""
    M = (1 << T) - 1  # truncation to $T
    LANES = _lanes(T)
    if W == 0:  # :52
        pass
    elif W == T:  # :54
        for row in range(T):  # seq_t!(row in $T ...)
            idx = index(row, lane)
            packed[LANES * row + lane] = kernel(idx) & M
    else:
        mask = ((1 << W) - 1) & M  # :62
        tmp = 0  # :65
        for row in range(T):  # :70
            idx = index(row, lane)
            src = kernel(idx) & M
            src = src & mask  # :73
            if row == 0:  # :76
                tmp = src
            else:
                tmp |= (src << ((row * W) % T)) & M  # :79
            curr_word = (row * W) // T  # :84
            next_word = ((row + 1) * W) // T  # :85
            if next_word > curr_word:  # :88
                packed[LANES * curr_word + lane] = tmp
                remaining_bits = ((row + 1) * W) % T  # :90
                tmp = src >> (W - remaining_bits)  # :92


def unpack_lane(T: int, W: int, packed: list, lane: int, kernel) -> None:
    """`unpack!` (src/macros.rs:101-173) for one lane; `kernel(idx, elem)` is the spliced closure."""
    M = (1 << T) - 1
    LANES = _lanes(T)
    if W == 0:  # :118
        for row in range(T):
            idx = index(row, lane)
            zero = 0
            kernel(idx, zero)
    elif W == T:  # :126
        for row in range(T):
            idx = index(row, lane)
            src = packed[LANES * row + lane]
            kernel(idx, src)
    else:
        def mask(width: int) -> int:  # :134-137
            return M if width == T else (1 << (width % T)) - 1

        src = packed[lane]  # :139
        for row in range(T):  # :142
            curr_word = (row * W) // T
            next_word = ((row + 1) * W) // T
            shift = (row * W) % T  # :147
            if next_word > curr_word:  # :149
                remaining_bits = ((row + 1) * W) % T
                current_bits = W - remaining_bits
                tmp = (src >> shift) & mask(current_bits)  # :154
                if next_word < W:  # :156
                    src = packed[LANES * next_word + lane]  # :158
                    tmp |= ((src & mask(remaining_bits)) << current_bits) & M  # :160
            else:
                tmp = (src >> shift) & mask(W)  # :164
            idx = index(row, lane)  # :168
            kernel(idx, tmp)


def iterate_lane(T: int, lane: int, kernel) -> None:
    """`iterate!` (src/macros.rs:12-31)."""
    for row in range(T):
        idx = index(row, lane)
        kernel(idx)


# ---- src/bitpacking.rs ---------------------------------------------------------------------------------------------

def pack(T: int, W: int, input: list) -> list:
    """`BitPacking::pack::<W>` (src/bitpacking.rs:65-74).  Returns the 1024*W/T packed words."""
    output = [0] * (1024 * W // T)
    for lane in range(_lanes(T)):
        pack_lane(T, W, output, lane, lambda idx: input[idx])
    return output


def unpack(T: int, W: int, input: list) -> list:
    """`BitPacking::unpack::<W>` (src/bitpacking.rs:98-107)."""
    output = [0] * 1024

    for lane in range(_lanes(T)):
        def k(idx, elem):
            output[idx] = elem

        unpack_lane(T, W, input, lane, k)
    return output


def lanes_by_index(T: int) -> list:
    """src/bitpacking.rs:207-213."""
    return [i % _lanes(T) for i in range(1024)]


def rows_by_index(T: int) -> list:
    """src/bitpacking.rs:216-232."""
    rows = [0] * 1024
    for i in range(1024):
        lane = i % _lanes(T)
        s = i // 128
        fl_order = (i - s * 128 - lane) // 16
        o = FL_ORDER[fl_order]
        rows[i] = o * 8 + s
    return rows


def unpack_single(T: int, W: int, packed: list, idx: int) -> int:
    """`BitPacking::unpack_single::<W>` (src/bitpacking.rs:132-179)."""
    M = (1 << T) - 1
    LANES = _lanes(T)
    if W == 0:  # :136
        return 0
    assert idx < 1024, f"Index must be less than 1024, got {idx}"  # :152
    lane, row = lanes_by_index(T)[idx], rows_by_index(T)[idx]
    if W == T:  # :159
        return packed[LANES * row + lane]
    mask = (1 << (W % T)) - 1  # :164
    start_bit = row * W
    start_word = start_bit // T
    lo_shift = start_bit % T
    remaining_bits = T - lo_shift
    lo = packed[LANES * start_word + lane] >> lo_shift  # :170
    if remaining_bits >= W:
        return lo & mask
    hi = (packed[LANES * (start_word + 1) + lane] << remaining_bits) & M  # :176
    return (lo | hi) & mask


# ---- src/ffor.rs ---------------------------------------------------------------------------------------------------

def for_pack(T: int, W: int, input: list, reference: int) -> list:
    """`FoR::for_pack::<W>` (src/ffor.rs:24-36)."""
    M = (1 << T) - 1
    output = [0] * (1024 * W // T)
    for lane in range(_lanes(T)):
        pack_lane(T, W, output, lane, lambda idx: (input[idx] - reference) & M)  # wrapping_sub
    return output


def unfor_pack(T: int, W: int, input: list, reference: int) -> list:
    """`FoR::unfor_pack::<W>` (src/ffor.rs:38-50)."""
    M = (1 << T) - 1
    output = [0] * 1024
    for lane in range(_lanes(T)):
        def k(idx, elem):
            output[idx] = (elem + reference) & M  # wrapping_add

        unpack_lane(T, W, input, lane, k)
    return output


# ---- src/delta.rs --------------------------------------------------------------------------------------------------

def delta(T: int, input: list, base: list) -> list:
    """`Delta::delta` (src/delta.rs:24-33)."""
    M = (1 << T) - 1
    output = [0] * 1024
    for lane in range(_lanes(T)):
        prev = [base[lane]]

        def k(idx):
            nxt = input[idx]
            output[idx] = (nxt - prev[0]) & M
            prev[0] = nxt

        iterate_lane(T, lane, k)
    return output


def undelta(T: int, input: list, base: list) -> list:
    """`Delta::undelta` (src/delta.rs:36-45)."""
    M = (1 << T) - 1
    output = [0] * 1024
    for lane in range(_lanes(T)):
        prev = [base[lane]]

        def k(idx):
            nxt = (input[idx] + prev[0]) & M
            output[idx] = nxt
            prev[0] = nxt

        iterate_lane(T, lane, k)
    return output


def undelta_pack(T: int, W: int, input: list, base: list) -> list:
    """`Delta::undelta_pack::<W>` (src/delta.rs:48-63)."""
    M = (1 << T) - 1
    output = [0] * 1024
    for lane in range(_lanes(T)):
        prev = [base[lane]]

        def k(idx, elem):
            nxt = (elem + prev[0]) & M
            output[idx] = nxt
            prev[0] = nxt

        unpack_lane(T, W, input, lane, k)
    return output


# ---- src/transpose.rs ----------------------------------------------------------------------------------------------

def transpose_index(idx: int) -> int:
    """`const fn transpose` (src/transpose.rs:29-36)."""
    lane = idx % 16
    order = (idx // 16) % 8
    row = idx // 128
    return (lane * 64) + (FL_ORDER[order] * 8) + row


def transpose(input: list) -> list:
    """src/transpose.rs:11-15."""
    output = [0] * 1024
    for i in range(1024):
        output[i] = input[transpose_index(i)]
    return output


def untranspose(input: list) -> list:
    """src/transpose.rs:18-22."""
    output = [0] * 1024
    for i in range(1024):
        output[transpose_index(i)] = input[i]
    return output
