// fl_oracle_kernels.hpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the spiraldb/fastlanes v0.1.8 hot path (the parity oracle and the CPU
// baseline).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may use anything under oracle/.  The product (fastlanes_b200/) never links or calls it.
//
// Parity pin: the reference crate cannot be compiled in this environment (no Rust toolchain) and
// ships no golden byte vectors.  This restatement is pinned by restating EVERY assertion of the
// reference's own unit tests (tests/test_oracle_reference_tests.py): all 124 generated round-trips
// with unpack_single agreement (src/bitpacking.rs:273-315), test_unchecked_pack (:249-256),
// test_unpack_single (:259-271), macros::test_pack (src/macros.rs:181-207), test_delta
// (src/delta.rs:81-107), test_ffor (src/ffor.rs:67-88), FL_ORDER involution (src/lib.rs:53-59) and
// the README example (README.md:14-47).  unpack_single is the reference's independent closed-form
// reader, so agreement of pack with it for every (T, W, i) fixes the wire format uniquely.
// No crate-executed outputs exist here; DESIGN.md says so too.
//
// This header is compiled once per x86-64 ISA level (see Makefile) with -DFLO_NS=<namespace>;
// the lane loop is left to the auto-vectoriser with the T rows fully unrolled at compile time,
// which is the recipe the crate itself relies on (README.md:9-10, src/macros.rs:68-70).
//
// Every function cites the reference lines it follows.  Citations are relative to /root/reference.
#pragma once
#include <cstddef>
#include <cstdint>
#include <type_traits>
#include <utility>

#ifndef FLO_NS
#error "define FLO_NS"
#endif

namespace FLO_NS {

// src/lib.rs:22 — FL_ORDER
constexpr int kOrder[8] = {0, 4, 2, 6, 1, 5, 3, 7};

// src/lib.rs:24-27 — FastLanes::{T, LANES}
template <class T>
struct Lay {
    static constexpr int TB = int(sizeof(T)) * 8;
    static constexpr int L = 1024 / TB;
};

// src/macros.rs:20-24 (dup :46-50, :112-116) — index(row, lane)
constexpr int index_of(int row, int lane) {
    return kOrder[row / 8] * 16 + (row % 8) * 128 + lane;
}

// src/lib.rs:41-47 — seq_t!: unroll `row in 0..N` at compile time.
template <int N, class F>
inline __attribute__((always_inline)) void seq_rows(F&& f) {
    [&]<int... R>(std::integer_sequence<int, R...>) __attribute__((always_inline)) {
        (f(std::integral_constant<int, R>{}), ...);
    }(std::make_integer_sequence<int, N>{});
}

// src/macros.rs:12-31 — iterate!: visit index(row, lane) for row in 0..T
template <class T, class K>
inline __attribute__((always_inline)) void iterate_lane(int lane, K&& kernel) {
    seq_rows<Lay<T>::TB>([&](auto rc) __attribute__((always_inline)) {
        kernel(index_of(decltype(rc)::value, lane));
    });
}

// src/macros.rs:35-97 — pack!
template <class T, int W, class K>
inline __attribute__((always_inline)) void pack_lane(T* __restrict packed, int lane, K&& kernel) {
    constexpr int TB = Lay<T>::TB;
    constexpr int L = Lay<T>::L;
    if constexpr (W == 0) {
        // :52 — nothing to write, the packed array is zero bytes (kernel is not evaluated)
    } else if constexpr (W == TB) {
        // :54-59 — verbatim copy in row order (no mask)
        seq_rows<TB>([&](auto rc) __attribute__((always_inline)) {
            constexpr int row = decltype(rc)::value;
            packed[L * row + lane] = kernel(index_of(row, lane));
        });
    } else {
        constexpr T mask = T((T(1) << W) - 1);  // :62
        T tmp = 0;                               // :65
        seq_rows<TB>([&](auto rc) __attribute__((always_inline)) {
            constexpr int row = decltype(rc)::value;
            T src = T(kernel(index_of(row, lane)) & mask);  // :72-73
            if constexpr (row == 0) {
                tmp = src;  // :76-77
            } else {
                tmp = T(tmp | T(src << ((row * W) % TB)));  // :79
            }
            constexpr int curr_word = (row * W) / TB;        // :84
            constexpr int next_word = ((row + 1) * W) / TB;  // :85
            if constexpr (next_word > curr_word) {           // :88
                packed[L * curr_word + lane] = tmp;          // :89
                constexpr int remaining_bits = ((row + 1) * W) % TB;  // :90
                tmp = T(src >> (W - remaining_bits));                 // :92
            }
        });
    }
}

// src/macros.rs:135-137 — mask(width)
template <class T>
constexpr T mask_of(int width) {
    constexpr int TB = Lay<T>::TB;
    return width == TB ? T(~T(0)) : T((T(1) << (width % TB)) - 1);
}

// src/macros.rs:101-173 — unpack!
template <class T, int W, class K>
inline __attribute__((always_inline)) void unpack_lane(const T* __restrict packed, int lane, K&& kernel) {
    constexpr int TB = Lay<T>::TB;
    constexpr int L = Lay<T>::L;
    if constexpr (W == 0) {
        // :118-125 — zeros, still visiting every index in order
        seq_rows<TB>([&](auto rc) __attribute__((always_inline)) {
            kernel(index_of(decltype(rc)::value, lane), T(0));
        });
    } else if constexpr (W == TB) {
        // :126-132
        seq_rows<TB>([&](auto rc) __attribute__((always_inline)) {
            constexpr int row = decltype(rc)::value;
            kernel(index_of(row, lane), packed[L * row + lane]);
        });
    } else {
        T src = packed[lane];  // :139
        seq_rows<TB>([&](auto rc) __attribute__((always_inline)) {
            constexpr int row = decltype(rc)::value;
            constexpr int curr_word = (row * W) / TB;        // :144
            constexpr int next_word = ((row + 1) * W) / TB;  // :145
            constexpr int shift = (row * W) % TB;            // :147
            T tmp;
            if constexpr (next_word > curr_word) {  // :149
                constexpr int remaining_bits = ((row + 1) * W) % TB;  // :152
                constexpr int current_bits = W - remaining_bits;      // :153
                tmp = T(T(src >> shift) & mask_of<T>(current_bits));  // :154
                if constexpr (next_word < W) {                        // :156
                    src = packed[L * next_word + lane];               // :158
                    tmp = T(tmp | T(T(src & mask_of<T>(remaining_bits)) << current_bits));  // :160
                }
            } else {
                tmp = T(T(src >> shift) & mask_of<T>(W));  // :164
            }
            kernel(index_of(row, lane), tmp);  // :168-169
        });
    }
}

// ---- trait level: one 1024-element block ---------------------------------------------------

// src/bitpacking.rs:65-74 — BitPacking::pack<W>
template <class T, int W>
void pack_block(const T* __restrict in, T* __restrict out) {
    for (int lane = 0; lane < Lay<T>::L; ++lane) {
        pack_lane<T, W>(out, lane, [&](int idx) __attribute__((always_inline)) { return in[idx]; });
    }
}

// src/bitpacking.rs:98-107 — BitPacking::unpack<W>
template <class T, int W>
void unpack_block(const T* __restrict in, T* __restrict out) {
    for (int lane = 0; lane < Lay<T>::L; ++lane) {
        unpack_lane<T, W>(in, lane, [&](int idx, T elem) __attribute__((always_inline)) { out[idx] = elem; });
    }
}

// src/ffor.rs:24-36 — FoR::for_pack<W>
template <class T, int W>
void for_pack_block(const T* __restrict in, T reference, T* __restrict out) {
    for (int lane = 0; lane < Lay<T>::L; ++lane) {
        pack_lane<T, W>(out, lane, [&](int idx) __attribute__((always_inline)) { return T(in[idx] - reference); });
    }
}

// src/ffor.rs:38-50 — FoR::unfor_pack<W>
template <class T, int W>
void unfor_pack_block(const T* __restrict in, T reference, T* __restrict out) {
    for (int lane = 0; lane < Lay<T>::L; ++lane) {
        unpack_lane<T, W>(in, lane,
                          [&](int idx, T elem) __attribute__((always_inline)) { out[idx] = T(elem + reference); });
    }
}

// src/delta.rs:24-33 — Delta::delta
template <class T>
void delta_block(const T* __restrict in, const T* __restrict base, T* __restrict out) {
    for (int lane = 0; lane < Lay<T>::L; ++lane) {
        T prev = base[lane];
        iterate_lane<T>(lane, [&](int idx) __attribute__((always_inline)) {
            T next = in[idx];
            out[idx] = T(next - prev);
            prev = next;
        });
    }
}

// src/delta.rs:36-45 — Delta::undelta
template <class T>
void undelta_block(const T* __restrict in, const T* __restrict base, T* __restrict out) {
    for (int lane = 0; lane < Lay<T>::L; ++lane) {
        T prev = base[lane];
        iterate_lane<T>(lane, [&](int idx) __attribute__((always_inline)) {
            T next = T(in[idx] + prev);
            out[idx] = next;
            prev = next;
        });
    }
}

// src/delta.rs:48-63 — Delta::undelta_pack<W>
template <class T, int W>
void undelta_pack_block(const T* __restrict in, const T* __restrict base, T* __restrict out) {
    for (int lane = 0; lane < Lay<T>::L; ++lane) {
        T prev = base[lane];
        unpack_lane<T, W>(in, lane, [&](int idx, T elem) __attribute__((always_inline)) {
            T next = T(elem + prev);
            out[idx] = next;
            prev = next;
        });
    }
}

// src/transpose.rs:29-36 — const fn transpose(idx)
constexpr int transpose_index(int idx) {
    const int lane = idx % 16;
    const int order = (idx / 16) % 8;
    const int row = idx / 128;
    return lane * 64 + kOrder[order] * 8 + row;
}

// src/transpose.rs:11-15 — Transpose::transpose
template <class T>
void transpose_block(const T* __restrict in, T* __restrict out) {
    for (int i = 0; i < 1024; ++i) out[i] = in[transpose_index(i)];
}

// src/transpose.rs:18-22 — Transpose::untranspose
template <class T>
void untranspose_block(const T* __restrict in, T* __restrict out) {
    for (int i = 0; i < 1024; ++i) out[transpose_index(i)] = in[i];
}

// src/bitpacking.rs:207-232 — lanes_by_index / rows_by_index, evaluated inline
template <class T>
constexpr int lane_by_index(int i) { return i % Lay<T>::L; }
template <class T>
constexpr int row_by_index(int i) {
    const int lane = i % Lay<T>::L;
    const int s = i / 128;
    const int fl_order = (i - s * 128 - lane) / 16;
    const int o = kOrder[fl_order];
    return o * 8 + s;
}

// src/bitpacking.rs:132-179 — BitPacking::unpack_single<W> (runtime W form of the same arithmetic)
template <class T>
T unpack_single_rt(unsigned W, const T* packed, size_t index) {
    constexpr int TB = Lay<T>::TB;
    constexpr int L = Lay<T>::L;
    if (W == 0) return T(0);  // :137-140
    const int lane = lane_by_index<T>(int(index));
    const int row = row_by_index<T>(int(index));
    if (W == unsigned(TB)) return packed[L * row + lane];  // :159-162
    const T mask = T((T(1) << (W % TB)) - 1);              // :164
    const unsigned start_bit = unsigned(row) * W;          // :165
    const unsigned start_word = start_bit / TB;            // :166
    const unsigned lo_shift = start_bit % TB;              // :167
    const unsigned remaining_bits = TB - lo_shift;         // :168
    const T lo = T(packed[L * start_word + lane] >> lo_shift);  // :170
    if (remaining_bits >= W) return T(lo & mask);               // :171-173
    const T hi = T(packed[L * (start_word + 1) + lane] << remaining_bits);  // :176
    return T(T(lo | hi) & mask);                                           // :177
}

// ---- runtime-width dispatch tables: the `match width` of src/bitpacking.rs:82-95, :115-128 ----

template <class T> using fn_pk_t = void (*)(const T*, T*);
template <class T> using fn_ref_t = void (*)(const T*, T, T*);
template <class T> using fn_base_t = void (*)(const T*, const T*, T*);

template <class T>
struct Tables {
    static constexpr int N = Lay<T>::TB + 1;
    fn_pk_t<T> pack[N];
    fn_pk_t<T> unpack[N];
    fn_ref_t<T> for_pack[N];
    fn_ref_t<T> unfor_pack[N];
    fn_base_t<T> undelta_pack[N];
};

template <class T, int... W>
constexpr Tables<T> make_tables(std::integer_sequence<int, W...>) {
    return Tables<T>{{&pack_block<T, W>...},
                     {&unpack_block<T, W>...},
                     {&for_pack_block<T, W>...},
                     {&unfor_pack_block<T, W>...},
                     {&undelta_pack_block<T, W>...}};
}

template <class T>
inline const Tables<T>& tables() {
    static constexpr Tables<T> t = make_tables<T>(std::make_integer_sequence<int, Lay<T>::TB + 1>{});
    return t;
}

// Batched entry points over contiguous arrays of blocks (block b at in + b*in_stride).
// op codes shared with fl_oracle.cpp
enum Op : int {
    OP_PACK = 0,
    OP_UNPACK = 1,
    OP_FOR_PACK = 2,
    OP_UNFOR_PACK = 3,
    OP_DELTA = 4,
    OP_UNDELTA = 5,
    OP_UNDELTA_PACK = 6,
    OP_TRANSPOSE = 7,
    OP_UNTRANSPOSE = 8,
    OP_UNFOR_FILTER = 9,  // not a reference function: unfor_pack + the caller-side predicate loop (README.md:40-41)
};

// What a user of the reference writes for a range scan (README.md:40-41: "unpack all values and then access the
// desired one"): FoR::unfor_pack (src/ffor.rs:38-50) into a cache-resident 1024-value buffer, then one pass over
// it.  bit i of the 128-byte block bitmap = lo <= value[i] <= hi (little-endian bit order).  The baseline the fused
// GPU scan kernels are timed against; also their test oracle's cross-check.
template <class T>
inline void filter_block(const T* values, T lo, T hi, uint8_t* bitmap) {
    const T span = T(hi - lo);
    const bool empty = hi < lo;
    alignas(64) uint8_t pass[1024];
    for (int i = 0; i < 1024; ++i) pass[i] = uint8_t(T(values[i] - lo) <= span);  // vectorises: sub, cmp, narrow
    for (int k = 0; k < 128; ++k) {
        uint64_t x;
        __builtin_memcpy(&x, pass + 8 * k, 8);
        bitmap[k] = empty ? uint8_t(0) : uint8_t((x * 0x0102040810204080ull) >> 56);  // 8 bool bytes -> 8 bits
    }
}

// `refs`: per-block reference array (may be null → `ref_scalar`); `base`: n_blocks × LANES.
template <class T>
void run_blocks(int op, unsigned width, size_t b0, size_t b1, const T* in, T* out, const T* base,
                const T* refs, T ref_scalar) {
    constexpr int TB = Lay<T>::TB;
    constexpr int L = Lay<T>::L;
    const size_t pw = size_t(1024) * width / TB;  // packed elements per block (bitpacking.rs:77)
    const Tables<T>& t = tables<T>();
    for (size_t b = b0; b < b1; ++b) {
        switch (op) {
            case OP_PACK: t.pack[width](in + b * 1024, out + b * pw); break;
            case OP_UNPACK: t.unpack[width](in + b * pw, out + b * 1024); break;
            case OP_FOR_PACK: t.for_pack[width](in + b * 1024, refs ? refs[b] : ref_scalar, out + b * pw); break;
            case OP_UNFOR_PACK: t.unfor_pack[width](in + b * pw, refs ? refs[b] : ref_scalar, out + b * 1024); break;
            case OP_DELTA: delta_block<T>(in + b * 1024, base + b * L, out + b * 1024); break;
            case OP_UNDELTA: undelta_block<T>(in + b * 1024, base + b * L, out + b * 1024); break;
            case OP_UNDELTA_PACK: t.undelta_pack[width](in + b * pw, base + b * L, out + b * 1024); break;
            case OP_TRANSPOSE: transpose_block<T>(in + b * 1024, out + b * 1024); break;
            case OP_UNTRANSPOSE: untranspose_block<T>(in + b * 1024, out + b * 1024); break;
            case OP_UNFOR_FILTER: {  // base = {lo, hi}; out = bitmap bytes (128 per block)
                alignas(64) T tmp[1024];
                t.unfor_pack[width](in + b * pw, refs ? refs[b] : ref_scalar, tmp);
                filter_block<T>(tmp, base[0], base[1], reinterpret_cast<uint8_t*>(out) + b * 128);
                break;
            }
            default: break;
        }
    }
}

}  // namespace FLO_NS
