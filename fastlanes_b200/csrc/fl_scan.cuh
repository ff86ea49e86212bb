// fl_scan.cuh — fused decode + predicate kernels ("unpack then filter / take").
//
// The reference has no such operator: its README tells callers that more than ~10 random accesses should unpack
// the whole block (README.md:40-41), i.e. a scan is `unpack` / `unfor_pack` (src/bitpacking.rs:98-107,
// src/ffor.rs:38-50) followed by the caller's own loop over the 1024 values.  On the GPU the materialised block
// is the expensive part (128*T bytes written per 128*W read), so these kernels keep the decoded tile in registers:
//
//   filter_warp_kernel : bit i of the block's 1024-bit bitmap = lo <= (unpack(packed)[i] + reference) <= hi
//                        reads 128*W, writes 128 (+4 with counts) bytes per block
//   select_warp_kernel : out[offsets[b] + rank_b(i)] = unpack(packed)[i] + reference for every set bit i of block b,
//                        in index order (stream compaction given the exclusive prefix of the block counts)
//                        reads 128*W + 128 + 8, writes sizeof(T) * selected bytes per block
//
// Both reuse the decode core of unpack_warp_kernel (warp_run_from / warp_extract_rows), one warp = one block at a time, 1-8
// consecutive blocks per warp with the loads of the next block issued ahead (RunLoads below).  Bit i of a block
// refers to the ORIGINAL value index i of the unpacked vector, i.e. exactly the element `output[i]` that
// BitPacking::unpack would have produced (index(row,lane), src/macros.rs:20-24); the bitmap is little-endian
// (byte i/8, bit i%8), the layout of an Arrow validity/selection buffer.
#pragma once
#include "fl_kernels.cuh"
#include "fl_scan_bits.h"

namespace flb {

// ---- one-block-ahead loads ---------------------------------------------------------------------------------------
// The scan kernels run NB consecutive blocks per warp.  Their per-block work is a few hundred instructions behind one DRAM
// round trip, so a warp that loads, waits and computes block after block idles for most of its life (ncu, select with one
// block per warp: 29 % of all warp samples on the first use of the loaded words, profiles/ncu_r02_select.md).  RunLoads
// holds a thread's slices of the packed word-rows of its run as raw registers: issue() for block b + 1 is called before
// block b is processed, assemble() applies the alignment shifts of warp_run_from at the point of use.
template <class T, int W>
__host__ __device__ constexpr int run_loads() {  // word-row loads warp_run_from issues for one thread
    constexpr int TB = Lay<T>::TB, RPG = TB / 4;
    return W == 0 ? 0 : (W == TB ? RPG : ((W % 4) == 0 ? run_words<T, W>() : run_words<T, W>() + 1));
}
template <class T, int W>
struct RunLoads {
    uint4 raw[run_loads<T, W>() > 0 ? run_loads<T, W>() : 1];
    __device__ __forceinline__ void issue(const char* __restrict__ blk_packed, int q, int j) {
        if constexpr (W > 0) {
            const char* pk = blk_packed + j * 16;
            int n = 0;
            Slice<T> unused[run_words<T, W>()];  // the shifts are dead here
            warp_run_from<T, W>([&](unsigned k) -> Slice<T> { raw[n] = ldg128_stream(pk + k * 128); return to_slice<T>(raw[n++]); }, q, unused);
        }
    }
    __device__ __forceinline__ void assemble(int q, Slice<T> (&a)[run_words<T, W>()]) const {
        int n = 0;
        warp_run_from<T, W>([&](unsigned) -> Slice<T> { return to_slice<T>(raw[n++]); }, q, a);
    }
    __device__ __forceinline__ uint32_t any_word() const {  // a value that depends on every load
        uint32_t d = 0;
#pragma unroll
        for (int n = 0; n < run_loads<T, W>(); ++n) d |= raw[n].x;
        return d;
    }
};

// Predicate bits of one 16-byte slice, ORed into `acc` at bit position POS: bit POS+k = lane k passes
// (v - c) <= span  (lane-wise, wrapping, unsigned).  spanH = span | H for the SWAR types.
template <class T, int POS>
__device__ __forceinline__ void slice_range_bits(uint32_t& acc, const Slice<T>& v, typename Lay<T>::R c,
                                                 typename Lay<T>::R span, typename Lay<T>::R spanH) {
    if constexpr (sizeof(T) >= 4) {
        // one lane per register: handled by shift_in_fail() in the kernel (carry chain), not here
        static_assert(sizeof(T) < 4, "u32/u64 use shift_in_fail");
    } else if constexpr (sizeof(T) == 2) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            acc |= top_bits_u16(swar_leu_top<16>(lane_sub<T>(v.r[r], c), span, spanH)) << (POS + 2 * r);
        }
    } else {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            acc |= top_bits_u8(swar_leu_top<8>(lane_sub<T>(v.r[r], c), span, spanH)) << (POS + 4 * r);
        }
    }
}

// The values of a block are v + ref with 0 <= v <= maxv = 2^W - 1.  The predicate (v - c) mod 2^T <= span (c = lo - ref)
// selects the cyclic interval [c, c + span]; intersected with [0, maxv] that is one interval [a, b], or everything
// but one interval (invert = ~0), with 0 <= a <= b <= maxv.  Uniform per block: ~15 scalar instructions.
template <class R>
struct FieldRange {
    R a, b;
    uint32_t invert;  // 0: pass = a <= v <= b;  ~0: pass = !(a <= v <= b)
};
template <class R>
__device__ __forceinline__ FieldRange<R> range_in_field(R c, R span, bool empty, R maxv) {
    FieldRange<R> f{R(0), maxv, 0u};  // everything passes
    const R e = R(c + span);
    if (empty) {
        f.invert = ~0u;  // nothing passes
    } else if (e >= c) {  // the cyclic interval does not wrap
        if (c > maxv) f.invert = ~0u;
        else { f.a = c; f.b = e < maxv ? e : maxv; }
    } else if (e < maxv) {  // wraps: [c, 2^T) u [0, e]; e >= maxv covers every field value
        if (c > maxv) { f.b = e; }                                          // only [0, e] reaches the field
        else if (R(e + 1) != c) { f.a = R(e + 1); f.b = R(c - 1); f.invert = ~0u; }  // fails exactly on (e, c)
    }
    return f;
}

// u32 / u64 (one lane per register): acc = 2*acc + (t > span) as a carry chain.  t > span  <=>  t + ~span carries out
// of the register, so ADD.CC sets the carry flag to "fail" and ADDC shifts it into acc: two instructions per value,
// no predicate / select / shift-or.  (ADD, not SUB: the carry of an addition is unambiguous, whereas after sub.cc the
// flag read by addc is the hardware carry = NOT borrow.)
__device__ __forceinline__ void shift_in_fail(uint32_t& acc, uint32_t t, uint32_t not_span) {
    asm("{\n .reg .u32 d;\n add.cc.u32 d, %1, %2;\n addc.u32 %0, %0, %0;\n}" : "+r"(acc) : "r"(t), "r"(not_span));
}
__device__ __forceinline__ void shift_in_fail(uint32_t& acc, uint64_t t, uint64_t not_span) {
    asm("{\n .reg .u64 d;\n add.cc.u64 d, %1, %2;\n addc.u32 %0, %0, %0;\n}" : "+r"(acc) : "l"(t), "l"(not_span));
}

// Block-invariant part of the predicate for one FoR reference (per block only when `refs` is given).  Three modes:
//   IN_PLACE   u32/u64, W > 0     field compared inside the packed word (see the kernel); p0 = -(A << k), p1 = ~bound
//   SWAR_FIELD u8/u16, 0 < W < T  the extracted values are < 2^W <= H = 2^(T-1), so the top bit of every SWAR lane is
//                                  free and A <= v <= B needs no carry isolation: top(v + (H - A)) & top((H | B) - v);
//                                  p0 = splat(H - A), p1 = splat(H | B)
//   GENERIC    W == 0, or u8/u16 at W == T: (v - c) <= span lane-wise; p0 = c, p1 = span, p2 = span | H
// `invert`: IN_PLACE / SWAR_FIELD: XOR mask of the result (range_in_field); GENERIC: ~0 when the range is empty.
template <class T, int W>
struct FilterPred {
    using R = typename Lay<T>::R;
    static constexpr bool IN_PLACE = sizeof(T) >= 4 && W > 0;
    static constexpr bool SWAR_FIELD = sizeof(T) <= 2 && W > 0 && W < Lay<T>::TB;
    R p0, p1, p2;
    uint32_t invert;
    __device__ __forceinline__ FilterPred(T ref, T lo, T hi) {
        constexpr int TB = Lay<T>::TB;
        if constexpr (IN_PLACE) {
            const FieldRange<R> fr = range_in_field<R>(R(T(lo - ref)), R(T(hi - lo)), hi < lo, rep_mask<T>(W));
            constexpr int k = TB - W;
            p0 = R(0) - R(fr.a << k);
            p1 = ~R(R(R(fr.b - fr.a) << k) | R((R(1) << k) - 1));
            p2 = 0;
            invert = fr.invert;
        } else if constexpr (SWAR_FIELD) {
            constexpr T H = T(T(1) << (TB - 1));
            const FieldRange<T> fr = range_in_field<T>(T(lo - ref), T(hi - lo), hi < lo, T((T(1) << W) - 1));  // mod 2^T
            p0 = slice_splat<T>(T(H - fr.a)).r[0];
            p1 = slice_splat<T>(T(H | fr.b)).r[0];
            p2 = 0;
            invert = fr.invert;
        } else {
            p0 = slice_splat<T>(T(lo - ref)).r[0];
            p1 = slice_splat<T>(T(hi - lo)).r[0];
            p2 = p1;
            if constexpr (sizeof(T) <= 2) p2 = p1 | rep_value<T>(T(T(1) << (TB - 1)));
            invert = (hi < lo) ? ~0u : 0u;
        }
    }
};

// NB consecutive blocks per warp: the per-warp set-up (thread mapping, tile address, and — with a scalar reference —
// the whole FilterPred) is paid once per NB blocks; the filter is issue-bound below W ~ 3T/4, so this is throughput.
template <class T, int W, bool TMA, int NB, bool PIPE = false>
__global__ void __launch_bounds__(kThreads)
filter_warp_kernel(const char* __restrict__ packed, unsigned char* __restrict__ bitmap, uint32_t* __restrict__ counts,
                   size_t n_blocks, const T* __restrict__ refs, T ref_scalar, T lo, T hi) {
    using R = typename Lay<T>::R;
    using WL = WarpLay<T>;
    constexpr int TB = Lay<T>::TB;
    constexpr int RPG = WL::RPG;
    constexpr int BPT = 128 / TB;  // predicate bits per thread per row
    const size_t blk0 = ((size_t(blockIdx.x) * kThreads + threadIdx.x) >> 5) * NB;
    if (blk0 >= n_blocks) return;  // warp-uniform
    const int lane = threadIdx.x & 31;
    const int g = lane >> 3, j = lane & 7;
    const int q = WL::rank_of_group(g);
    __shared__ __align__(16) unsigned char scan_tile[kThreads / 32][128];
    unsigned char* tile = scan_tile[threadIdx.x >> 5];
    FilterPred<T, W> pred(ref_scalar, lo, hi);

    // one block; with PIPE the packed words of block blk + 1 are requested before block blk is evaluated (RunLoads)
    auto one_block = [&](size_t blk, bool more, const RunLoads<T, W>& cur, RunLoads<T, W>& nxt, T ref_cur, T& ref_nxt) {
        if constexpr (PIPE) {
            if (more) {
                nxt.issue(packed + (blk + 1) * (size_t(128) * W), q, j);
                if (refs != nullptr) ref_nxt = refs[blk + 1];
            }
        }
        if (refs != nullptr) pred = FilterPred<T, W>(PIPE ? ref_cur : refs[blk], lo, hi);
        Slice<T> a[run_words<T, W>()];
        if constexpr (PIPE) cur.assemble(q, a);
        else warp_load_run<T, W, TMA, (kThreads / 32) * 128>(packed + blk * (size_t(128) * W), lane, q, j, a);

        // value = v + ref (ffor.rs:47, wrapping).  lo <= value <= hi  <=>  (v - c) mod 2^T <= span, c = lo - ref, span = hi - lo
        uint32_t x = 0;
        if constexpr (FilterPred<T, W>::IN_PLACE) {
            // One lane per register: compare the W-bit field IN PLACE, without extracting it.  Because 0 <= v < 2^W the
            // cyclic interval [c, c + span] restricted to [0, 2^W) is a plain interval [A, B] or the complement of one
            // (range_in_field), so with k = T - W and the field moved to the TOP of the register (one shift-add, or one
            // funnel shift when it straddles two words; the low k bits are garbage g < 2^k):
            //     X - (A << k)  <=  ((B - A) << k) | (2^k - 1)      <=>      A <= v <= B
            // -> per value: LEA/IMAD (align and subtract), ADD.CC (compare as a carry), ADDC (shift the bit in).
            const R neg_a = pred.p0, not_bound = pred.p1;
            // values in DESCENDING bit position (row RPG-1 first): the last one shifted in lands at bit 0
            seq_rows<RPG>([&](auto ic) {
                constexpr int i = RPG - 1 - decltype(ic)::value;
                constexpr int idx = (i * W) / TB;
                constexpr int sh = (i * W) % TB;
#pragma unroll
                for (int r = Lay<T>::NR - 1; r >= 0; --r) {
                    R top;
                    if constexpr (sh + W <= TB) top = R(a[idx].r[r] << (TB - sh - W));
                    else top = R(a[idx].r[r] >> (sh + W - TB)) | R(a[idx + 1].r[r] << (2 * TB - sh - W));  // funnel shift
                    shift_in_fail(x, R(top + neg_a), not_bound);
                }
            });
            x = ~x ^ pred.invert;  // fail bits -> pass bits (RPG * NR == 32 values: every bit of x is one value)
        } else {
            Slice<T> v[RPG];
            warp_extract_rows<T, W>(a, v);
            const R c = pred.p0, span = pred.p1;
            if constexpr (sizeof(T) >= 4) {  // W == 0: every value is 0
                x = (R(R(0) - c) <= span) ? 0xffffffffu : 0u;
                x &= ~pred.invert;  // empty range
            } else if constexpr (FilterPred<T, W>::SWAR_FIELD) {
                constexpr R H = rep_value<T>(T(T(1) << (TB - 1)));
                seq_rows<RPG>([&](auto ic) {
                    constexpr int i = decltype(ic)::value;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const uint32_t top = (v[i].r[r] + c) & (span - v[i].r[r]) & H;  // c = splat(H - A), span = splat(H | B)
                        if constexpr (sizeof(T) == 2) x |= top_bits_u16(top) << (i * BPT + 2 * r);
                        else x |= top_bits_u8(top) << (i * BPT + 4 * r);
                    }
                });
                x ^= pred.invert;
            } else {
                const R spanH = pred.p2;
                seq_rows<RPG>([&](auto ic) {
                    constexpr int i = decltype(ic)::value;
                    slice_range_bits<T, i * BPT>(x, v[i], c, span, spanH);
                });
                x &= ~pred.invert;  // empty range
            }
        }

        if constexpr (sizeof(T) == 4) {
            x = merge_pair_bpt4(x, __shfl_xor_sync(0xffffffffu, x, 1), j);
        } else if constexpr (sizeof(T) == 8) {
            x = merge_pair_bpt2(x, __shfl_xor_sync(0xffffffffu, x, 1), j);
            x = merge_quad_bpt2(x, __shfl_xor_sync(0xffffffffu, x, 2), j);
        }
        scan_store<TB>(tile, q, j, x);
        __syncwarp();
        const uint32_t word = reinterpret_cast<const uint32_t*>(tile)[lane];
        reinterpret_cast<uint32_t*>(bitmap + blk * 128)[lane] = word;  // one 128-byte line per warp
        if (counts != nullptr) {
            const uint32_t n = __reduce_add_sync(0xffffffffu, uint32_t(__popc(word)));
            if (lane == 0) counts[blk] = n;
        }
        __syncwarp();  // the tile is rewritten by the next block
    };

    const size_t blk_end = (n_blocks - blk0 < size_t(NB)) ? n_blocks : blk0 + NB;
    RunLoads<T, W> ld0, ld1;
    T r0 = ref_scalar, r1 = ref_scalar;
    if constexpr (PIPE) {
        static_assert(!TMA && (NB == 1 || NB % 2 == 0), "pipelined loads: direct loads, block loop unrolled by two");
        ld0.issue(packed + blk0 * (size_t(128) * W), q, j);
        if (refs != nullptr) r0 = refs[blk0];
#pragma unroll 1
        for (size_t blk = blk0; blk < blk_end; blk += 2) {
            one_block(blk, blk + 1 < blk_end, ld0, ld1, r0, r1);
            if (blk + 1 >= blk_end) break;
            one_block(blk + 1, blk + 2 < blk_end, ld1, ld0, r1, r0);
        }
    } else {
#pragma unroll 1
        for (size_t blk = blk0; blk < blk_end; ++blk) one_block(blk, false, ld0, ld1, r0, r1);
    }
}

// ---------------------------------------------------------------------------------------------------
// u8 filter, ROW-SLICE mapping (8 threads = one block, a warp = 4 blocks; same reasoning as the u8 fused chains in
// fl_kernels.cuh): a 1 KiB block gives a warp too little work to amortise the per-block set-up and the cross-thread
// byte assembly of filter_warp_kernel.  Here thread j holds ALL 8 rows of lanes 16j .. 16j+15; for u8 index(r, lane) =
// r*128 + lane (FL_ORDER[0] = 0), so its 16 predicate bits of row r ARE halfword r*8 + j of the block bitmap: no
// shuffles, no shared memory, one 16-bit store per row (the 8 threads of a block write 16 contiguous bytes).
// ---------------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(kThreads)
filter_u8_slice_kernel(const char* __restrict__ packed, unsigned char* __restrict__ bitmap, uint32_t* __restrict__ counts,
                       size_t n_blocks, const uint8_t* __restrict__ refs, uint8_t ref_scalar, uint8_t lo, uint8_t hi) {
    using T = uint8_t;
    using R = uint32_t;
    const size_t tid = size_t(blockIdx.x) * kThreads + threadIdx.x;
    const size_t blk_raw = tid / kSlicesPerBlock;
    const int j = int(tid % kSlicesPerBlock);
    const bool active = blk_raw < n_blocks;  // whole 8-thread groups are active or not; shuffles below stay inside a group
    const size_t blk = active ? blk_raw : 0;
    const FilterPred<T, W> pred(refs ? refs[blk] : ref_scalar, lo, hi);
    const char* pk = packed + blk * (size_t(128) * W) + j * 16;
    Slice<T> w[W > 0 ? W : 1];
    seq_rows<W>([&](auto kc) {
        constexpr int k = decltype(kc)::value;
        w[k] = load_slice<T>(pk + k * 128);
    });
    uint16_t* bm = reinterpret_cast<uint16_t*>(bitmap + blk * 128) + j;
    uint32_t cnt = 0;
    seq_rows<8>([&](auto rc) {
        constexpr int row = decltype(rc)::value;
        Slice<T> v;
        if constexpr (W == 0) v = slice_zero<T>();
        else if constexpr (W == 8) v = w[row];
        else {
            constexpr int curr = (row * W) / 8;
            constexpr int nxt = (curr + 1 < W) ? curr + 1 : curr;
            v = extract_row<T, W, row>(w[curr], w[nxt]);
        }
        uint32_t bits = 0;
        if constexpr (FilterPred<T, W>::SWAR_FIELD) {
            // pass bits at 7, 15, 23, 31 of each register.  u * 0x00204081 moves them to bits 28..31 (the partial products
            // land on distinct positions, the unwanted ones at <= 23 or beyond bit 31: no carries), so a register's nibble
            // needs no shift-and-mask before the multiply and no mask after it; the four nibbles are chained with
            // acc = (acc >> 4) | (p & 0xF0000000): 6 instructions per register instead of ~10 (top_bits_u8 + shift-or).
            constexpr R H = 0x80808080u;
            uint32_t acc = 0;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const uint32_t u = (v.r[r] + pred.p0) & (pred.p1 - v.r[r]) & H;
                acc = (acc >> 4) | ((u * 0x00204081u) & 0xF0000000u);
            }
            bits = ((acc >> 16) ^ pred.invert) & 0xFFFFu;
        } else {
            slice_range_bits<T, 0>(bits, v, pred.p0, pred.p1, pred.p2);
            bits &= ~pred.invert;  // empty range
        }
        if (active) bm[row * 8] = uint16_t(bits);
        cnt += uint32_t(__popc(bits));
    });
    if (counts != nullptr) {
#pragma unroll
        for (int d = 4; d >= 1; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
        if (active && j == 0) counts[blk] = cnt;
    }
}

// ---------------------------------------------------------------------------------------------------
// u16 filter, ROW-SLICE mapping (0 < W < 16).  The warp-block filter spends, per u16 block, 12-16 run-time lane funnel
// shifts (4 instructions each: u16 has no native funnel) to align the four groups' runs, ~3 instructions per register to
// extract, and 5 per register to move two pass bits into place: issue-bound at 0.39-0.72 of the HBM roofline (W = 4..13,
// profiles/opbench_select_r02_staged_v1.txt).  With 8 threads per block a thread owns the 16-byte column slice of ALL 16
// rows: every shift is a compile-time constant again (extract_row<W, row>, as in the reference's seq_t! unrolling,
// src/lib.rs:41-47), nothing crosses threads, and the pass bits of a row's four registers are merged by a shift-or chain
//     acc = (acc >> 2) | (t1 & t2 & H)        ->  lanes 0,2,4,6 at bits 9,11,13,15; lanes 1,3,5,7 at bits 25,27,29,31
// so a row's 8 predicate bits — one whole byte of the bitmap, index(r, 8j) / 8 = FL_ORDER[r/8]*2 + (r%8)*16 + j — cost
// 4 x (2 IADD + LOP3 + SHF + LOP3) + 4 instructions instead of 4 x 10.
// ---------------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(kThreads)
filter_u16_slice_kernel(const char* __restrict__ packed, unsigned char* __restrict__ bitmap, uint32_t* __restrict__ counts,
                        size_t n_blocks, const uint16_t* __restrict__ refs, uint16_t ref_scalar, uint16_t lo, uint16_t hi) {
    using T = uint16_t;
    using R = uint32_t;
    static_assert(W > 0 && W < 16, "row-slice u16 filter: widths with a free top bit per lane");
    constexpr int TB = 16;
    const size_t tid = size_t(blockIdx.x) * kThreads + threadIdx.x;
    const size_t blk_raw = tid / kSlicesPerBlock;
    const int j = int(tid % kSlicesPerBlock);
    const bool active = blk_raw < n_blocks;  // whole 8-thread groups are active or not
    const size_t blk = active ? blk_raw : 0;
    const FilterPred<T, W> pred(refs ? refs[blk] : ref_scalar, lo, hi);  // SWAR_FIELD: p0 = splat(H - A), p1 = splat(H | B)
    const char* pk = packed + blk * (size_t(128) * W) + j * 16;
    constexpr int D = (FLB_PREFETCH < W) ? FLB_PREFETCH : W;
    Slice<T> w[W];
    seq_rows<D>([&](auto kc) {
        constexpr int k = decltype(kc)::value;
        w[k] = load_slice<T>(pk + k * 128);
    });
    constexpr R H = 0x80008000u;
    unsigned char* bm = bitmap + blk * 128 + j;
    uint32_t words[4] = {0u, 0u, 0u, 0u};  // the thread's 16 row bytes, for the popcount
    seq_rows<TB>([&](auto rc) {
        constexpr int row = decltype(rc)::value;
        constexpr int curr = (row * W) / TB;  // macros.rs:144
        constexpr bool first_of_word = (row == 0) || (((row - 1) * W) / TB != curr);
        if constexpr (first_of_word && curr > 0 && curr + D - 1 < W) w[curr + D - 1] = load_slice<T>(pk + (curr + D - 1) * 128);
        constexpr int nxt = (curr + 1 < W) ? curr + 1 : curr;  // only read when the field straddles (macros.rs:156)
        const Slice<T> v = extract_row<T, W, row>(w[curr], w[nxt]);
        uint32_t acc = 0;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const uint32_t u = (v.r[r] + pred.p0) & (pred.p1 - v.r[r]) & H;  // top bit of each lane: A <= v <= B
            acc = (r == 0) ? u : ((acc >> 2) | u);
        }
        uint32_t byte = ((acc >> 9) & 0x55u) | ((acc >> 24) & 0xAAu);
        byte = (byte ^ pred.invert) & 0xFFu;
        constexpr int off = fl_order(row / 8) * 2 + (row % 8) * 16;  // byte of index(row, 8j) in the block bitmap, minus j
        if (active) bm[off] = (unsigned char)byte;
        words[row / 4] |= byte << (8 * (row % 4));
    });
    if (counts != nullptr) {
        uint32_t cnt = uint32_t(__popc(words[0]) + __popc(words[1]) + __popc(words[2]) + __popc(words[3]));
#pragma unroll
        for (int d = 4; d >= 1; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
        if (active && j == 0) counts[blk] = cnt;
    }
}

// ---------------------------------------------------------------------------------------------------
// Delta scan: bit i of the block's bitmap = lo <= untranspose(undelta_pack(packed, base))[i] <= hi — the range scan
// over a delta-encoded (sorted ids, timestamps) column, answered in ORIGINAL value order without materialising the
// decoded block (src/delta.rs:48-63 + src/transpose.rs:18-22 + the caller-side loop of README.md:40-41).
// Decode and prefix-add exactly as unpack_warp_kernel<UOP_DELTA>; the predicate bits are kept lane-major so that
// every thread owns whole bytes of the original-order bitmap (fl_scan_bits.h, "delta scan").
// ---------------------------------------------------------------------------------------------------
// NB consecutive blocks per warp; the loads of block b + 1 (packed words and bases) are issued before block b is evaluated.
template <class T, int W, int NB>
__global__ void __launch_bounds__(kThreads)
delta_filter_warp_kernel(const char* __restrict__ packed, const char* __restrict__ base, unsigned char* __restrict__ bitmap,
                         uint32_t* __restrict__ counts, size_t n_blocks, T lo, T hi) {
    using R = typename Lay<T>::R;
    using WL = WarpLay<T>;
    constexpr int TB = Lay<T>::TB;
    constexpr int RPG = WL::RPG;
    constexpr int NR = Lay<T>::NR;
    static_assert(NB == 1 || NB % 2 == 0, "block loop unrolled by two");
    const size_t blk0 = ((size_t(blockIdx.x) * kThreads + threadIdx.x) >> 5) * NB;
    if (blk0 >= n_blocks) return;  // warp-uniform
    const int lane = threadIdx.x & 31;
    const int g = lane >> 3, j = lane & 7;
    const int q = WL::rank_of_group(g);
    __shared__ __align__(16) unsigned char scan_tile[kThreads / 32][128];
    unsigned char* tile = scan_tile[threadIdx.x >> 5];

    auto one_block = [&](size_t blk, bool more, const RunLoads<T, W>& cur, RunLoads<T, W>& nxt, const uint4& base_cur, uint4& base_nxt) {
    if (more) {
        nxt.issue(packed + (blk + 1) * (size_t(128) * W), q, j);
        base_nxt = ldg128_stream(base + (blk + 1) * 128 + j * 16);
    }
    Slice<T> carry = to_slice<T>(base_cur);  // prev = base[lane] (delta.rs:50)
    Slice<T> a[run_words<T, W>()];
    cur.assemble(q, a);
    Slice<T> v[RPG];
    warp_extract_rows<T, W>(a, v);

    // delta.rs:56-60: running wrapping sum along rows per lane (same as unpack_warp_kernel<UOP_DELTA>)
#pragma unroll
    for (int i = 1; i < RPG; ++i) v[i] = slice_add<T>(v[i], v[i - 1]);
#pragma unroll
    for (int qq = 0; qq < 3; ++qq) {  // totals of the runs that precede this one in row order
        const int src = WL::group_of_rank(qq) * 8 + j;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const R t = shfl_reg<R>(v[RPG - 1].r[r], src);
            if (qq < q) carry.r[r] = lane_add<T>(carry.r[r], t);
        }
    }
    // lo <= value <= hi  <=>  (value - lo) mod 2^T <= hi - lo.  The subtraction of lo rides on the carry (one lane-wise
    // subtract per register of the carry instead of one per value): v[i] becomes value - lo.  Bits LANE-major: bit k*RPG + i
    const Slice<T> cs = slice_splat<T>(lo), ss = slice_splat<T>(T(hi - lo));
    carry = slice_sub<T>(carry, cs);
#pragma unroll
    for (int i = 0; i < RPG; ++i) v[i] = slice_add<T>(v[i], carry);

    uint32_t x = 0;
    const R span = ss.r[0];
    if constexpr (sizeof(T) >= 4) {
        const R not_span = ~span;
#pragma unroll
        for (int r = NR - 1; r >= 0; --r)  // descending bit position: the last value shifted in lands at bit 0
#pragma unroll
            for (int i = RPG - 1; i >= 0; --i) shift_in_fail(x, v[i].r[r], not_span);
        x = ~x;
    } else {
        const R spanH = span | rep_value<T>(T(T(1) << (TB - 1)));
        seq_rows<RPG>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const uint32_t le = swar_leu_top<TB>(v[i].r[r], span, spanH);
                if constexpr (sizeof(T) == 2) x |= top_bits_lane_major_u16(le) << (8 * r + i);
                else x |= top_bits_lane_major_u8(le) << (8 * r + i);
            }
        });
    }
    if (hi < lo) x = 0;  // empty range

    // u8 / u16: rank == group, so the thread holding the same lanes at rank q^1 (q^2) is lane^8 (lane^16)
    if constexpr (sizeof(T) == 2) {
        x = merge_pair_bpt4(x, __shfl_xor_sync(0xffffffffu, x, 8), q);
    } else if constexpr (sizeof(T) == 1) {
        x = merge_pair_bpt2(x, __shfl_xor_sync(0xffffffffu, x, 8), q);
        x = merge_quad_bpt2(x, __shfl_xor_sync(0xffffffffu, x, 16), q);
    }
    scan_store_orig<TB>(tile, q, j, x);
    __syncwarp();
    const uint32_t word = reinterpret_cast<const uint32_t*>(tile)[lane];
    reinterpret_cast<uint32_t*>(bitmap + blk * 128)[lane] = word;
    if (counts != nullptr) {
        const uint32_t n = __reduce_add_sync(0xffffffffu, uint32_t(__popc(word)));
        if (lane == 0) counts[blk] = n;
    }
    if (NB > 1) __syncwarp();  // the tile is rewritten by the next block
    };

    const size_t blk_end = (n_blocks - blk0 < size_t(NB)) ? n_blocks : blk0 + NB;
    RunLoads<T, W> ld0, ld1;
    uint4 b0, b1;
    ld0.issue(packed + blk0 * (size_t(128) * W), q, j);
    b0 = ldg128_stream(base + blk0 * 128 + j * 16);
    if constexpr (NB == 1) {
        one_block(blk0, false, ld0, ld1, b0, b1);
    } else {
#pragma unroll 1
        for (size_t blk = blk0; blk < blk_end; blk += 2) {
            one_block(blk, blk + 1 < blk_end, ld0, ld1, b0, b1);
            if (blk + 1 >= blk_end) break;
            one_block(blk + 1, blk + 2 < blk_end, ld1, ld0, b1, b0);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// select (dense compaction of the values a bitmap selects): out[offsets[b] + k] = the k-th selected value of block b.
// One warp = one block at a time, the thread that decoded a value compacts it.  History of the store side and of the
// instruction count (profiles/ncu_s2_r01.md, profiles/ncu_r02_kernels.md, profiles/ncu_r02_select.md):
//   round 1   one predicated 1-element st.global per value straight from the decode registers: partial-sector writes cost
//             3.0x the algorithmic L2 write traffic.
//   round 2a  selected values are first compacted into a warp-private shared-memory staging buffer (predicated STS at the
//             value's rank, from the per-word popcount scan), laid out with the SAME 16-byte phase as the destination
//             out + offsets[b], and drained with full 16-byte coalesced STG.128; every output byte is written once, in
//             sector-sized pieces (write traffic 1.0x).  524 issued instructions per u32 block, issue slots 75 % busy.
//   round 2b  a lane-per-bitmap-word variant (index-order round trip through shared memory) — faster only for u8; deleted.
//   round 2c  (this kernel) per-row and per-block overheads removed:
//     * the (bitmap word, prefix) table holds the SHARED BYTE ADDRESS of the word's first output slot, so a row's first
//       store address is one multiply-add on the popcount of the lower bits (was: add, add, multiply-add);
//     * the slice's bitmap bits are rotated to positions 1 .. BPT so that one R2P moves them all into predicates;
//     * no "nothing selected in my slice of this row" branch — at any selectivity worth a dense output it was never
//       taken by all 32 lanes, so it only cost the test and the reconvergence pair;
//     * the SWAR types store the low byte / halfword of the shifted register (st.shared.u8 / .u16 truncate): no mask;
//     * the drain no longer diverges: full 16-byte vectors in one loop, the (at most two) partial edge vectors of the
//       block's output run by one predicated element store per lane;
//     * direct 128-bit loads instead of the TMA bulk load (no mbarrier round trip; as in the filter kernels);
//     * NB consecutive blocks per warp with all global loads of block b + 1 issued before block b is processed: ncu on
//       the one-block-per-warp version put 29 % of all warp samples on the first use of the bitmap word and the packed
//       words (two dependent DRAM round trips per warp lifetime).
//   u32 W = 8 at 25 % selectivity: 615 -> 468 us per 2^20 blocks (0.55 -> 0.75 of the measured HBM peak); 323 issued
//   instructions per block.
//   dynamic shared memory: kThreads/32 warps x select_stage_bytes<T>()
// ---------------------------------------------------------------------------------------------------
template <class T>
__host__ __device__ constexpr int select_stage_bytes() { return 1024 * int(sizeof(T)) + 16; }

// Stores the low sizeof(T) bytes of `x` at shared address `sa` (st.shared.u8 / .u16 truncate the wider source register, so the
// SWAR types pass their register shifted down to the lane without masking it: one instruction less per value).
template <class T>
__device__ __forceinline__ void sts_low(uint32_t sa, typename Lay<T>::R x) {
    if constexpr (sizeof(T) == 1) asm volatile("st.shared.u8 [%0], %1;" ::"r"(sa), "r"(x) : "memory");
    else if constexpr (sizeof(T) == 2) asm volatile("st.shared.u16 [%0], %1;" ::"r"(sa), "r"(x) : "memory");
    else if constexpr (sizeof(T) == 4) asm volatile("st.shared.b32 [%0], %1;" ::"r"(sa), "r"(x) : "memory");
    else asm volatile("st.shared.b64 [%0], %1;" ::"r"(sa), "l"((unsigned long long)x) : "memory");
}
// lane k of a 16-byte slice in the low bits of a register (the bits above it are other lanes)
template <class T>
__device__ __forceinline__ typename Lay<T>::R slice_lane_low(const Slice<T>& s, int k) {
    constexpr int LPR = Lay<T>::LPR;
    return s.r[k / LPR] >> (Lay<T>::TB * (k % LPR));
}

// everything a block's compaction needs from global memory, loaded one block ahead of its use
template <class T, int W>
struct SelectLoads {
    uint32_t mword;  // this lane's word of the block bitmap
    uint64_t obase;  // offsets[blk]
    T ref;
    RunLoads<T, W> run;  // this thread's slices of the word-rows of its run
};
template <class T, int W>
__device__ __forceinline__ void select_issue_loads(SelectLoads<T, W>& ld, size_t blk, int lane, int q, int j,
                                                   const char* __restrict__ packed, const unsigned char* __restrict__ bitmap,
                                                   const uint64_t* __restrict__ offsets, const T* __restrict__ refs, T ref_scalar) {
    ld.mword = reinterpret_cast<const uint32_t*>(bitmap + blk * 128)[lane];
    ld.obase = offsets[blk];
    ld.ref = refs ? refs[blk] : ref_scalar;
    ld.run.issue(packed + blk * (size_t(128) * W), q, j);
}

template <class T, int W, int NB>
__global__ void __launch_bounds__(kThreads)
select_warp_kernel(const char* __restrict__ packed, const unsigned char* __restrict__ bitmap,
                    const uint64_t* __restrict__ offsets, T* __restrict__ out, size_t n_blocks,
                    const T* __restrict__ refs, T ref_scalar) {
    using WL = WarpLay<T>;
    constexpr int TB = Lay<T>::TB;
    constexpr int RPG = WL::RPG;
    constexpr int BPT = 128 / TB;
    constexpr int S = int(sizeof(T));
    constexpr int EPV = 16 / S;  // elements per 16-byte vector
    static_assert(NB == 1 || NB % 2 == 0, "the block loop is unrolled by two (ping-pong of the prefetch registers)");
    const size_t blk0 = ((size_t(blockIdx.x) * kThreads + threadIdx.x) >> 5) * NB;
    if (blk0 >= n_blocks) return;  // warp-uniform
    const int lane = threadIdx.x & 31;
    const int g = lane >> 3, j = lane & 7;
    const int q = WL::rank_of_group(g);
    __shared__ uint2 sel_tile[kThreads / 32][32];  // (bitmap word, shared byte address of the word's first output slot)
    uint2* tile = sel_tile[threadIdx.x >> 5];
    extern __shared__ __align__(16) unsigned char select_stage_smem[];
    unsigned char* stage = select_stage_smem + (threadIdx.x >> 5) * select_stage_bytes<T>();  // 16-byte aligned
    const uint32_t stage_sa = smem_addr(stage);
    // Original index of this thread's first lane in local row i: index(q*RPG + i, j*BPT) (macros.rs:20-24).  Rows of one
    // 8-row band share FL_ORDER[r/8], so inside a band bit0 = c0 + (r%8)*128: the bit position inside the 32-bit bitmap
    // word (sh) and the mask of the lower bits are per-thread constants, the word index advances by 4 per row.
    constexpr int BANDS = RPG > 8 ? RPG / 8 : 1;
    const uint2* trow[BANDS];
    uint32_t sh[BANDS], lowmask[BANDS];
#pragma unroll
    for (int bnd = 0; bnd < BANDS; ++bnd) {
        const int c0 = select_band_origin<TB>(q, bnd, j);  // fl_scan_bits.h (checked on the CPU by tests/cpp/test_scan_bits.cpp)
        trow[bnd] = tile + (c0 >> 5);
        const uint32_t s0 = uint32_t(c0 & 31);
        sh[bnd] = select_rotation(s0);  // rotate right by s0 - 1: the slice's bits land at positions 1 .. BPT (see below)
        lowmask[bnd] = select_low_mask(s0);
    }

    // One block: `cur` was loaded during the previous block (or before the loop); the loads of block blk + 1 are issued
    // into `nxt` BEFORE this block is processed, so a warp waits for DRAM once per NB blocks, not once (or twice: bitmap,
    // then packed) per block.  ncu, one block per warp: 29 % of all warp samples sat on the first use of the bitmap word
    // and of the packed words (profiles/ncu_r02_select.md).
    auto one_block = [&](size_t blk, bool more, const SelectLoads<T, W>& cur, SelectLoads<T, W>& nxt) {
        if (more) select_issue_loads<T, W>(nxt, blk + 1, lane, q, j, packed, bitmap, offsets, refs, ref_scalar);
        const uint32_t mword = cur.mword;
        // exclusive prefix of the per-word popcounts: rank of the first bit of word `lane` among the block's set bits
        const uint32_t cnt = uint32_t(__popc(mword));
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        // Nothing selected: skip the decode.  The exit also depends on the packed words (`dep`): ptxas sinks loads below an
        // exit on whose path they are dead — with one block per warp that puts the bitmap and the packed round trips in
        // series.  Taking the long path with total == 0 is harmless (every store is predicated on a set bit, the drain is
        // empty), so the extra condition may be anything the compiler cannot fold.
        if (total == 0 && __all_sync(0xffffffffu, cur.run.any_word() != 0x5bd1e995u)) return;  // the vote keeps the branch uniform  // the vote keeps the branch warp-uniform
        T* o = out + cur.obase;
        const uint32_t mis = uint32_t((reinterpret_cast<uintptr_t>(o) & 15u) / sizeof(T));  // phase of the run inside a 16-byte vector
        // (the previous block's table look-ups ended before its pre-drain __syncwarp, and its drain reads end before any lane
        // passes the __syncwarp below, which is ahead of the first staging store of this block)
        tile[lane] = make_uint2(mword, stage_sa + (mis + incl - cnt) * uint32_t(S));
        __syncwarp();

        Slice<T> a[run_words<T, W>()];
        cur.run.assemble(q, a);
        Slice<T> v[RPG];
        warp_extract_rows<T, W>(a, v);
#pragma unroll
        for (int i = 0; i < RPG; ++i) {
            const uint2 e = trow[i / 8][4 * (i % 8)];  // word (c0 + (i%8)*128) / 32: one base, immediate offsets
            // bit k of the slice at position k + 1: the compiler moves register bits 1.. into predicates with one R2P, but
            // spends two extra instructions on bit 0 (P0 is its scratch predicate)
            const uint32_t bits = __funnelshift_r(e.x, e.x, sh[i / 8]);
            uint32_t sa = e.y + uint32_t(__popc(e.x & lowmask[i / 8])) * uint32_t(S);  // slot of the slice's first selected value
#pragma unroll
            for (int k = 0; k < BPT; ++k) {
                if (bits & (2u << k)) {  // predicated STS + predicated address bump
                    sts_low<T>(sa, slice_lane_low<T>(v[i], k));
                    sa += uint32_t(S);
                }
            }
        }
        __syncwarp();
        // drain: stage and o - mis are both 16-byte aligned; vector x covers run elements [x*EPV - mis, ...); the FoR
        // reference is added here, on the selected values only (ffor.rs:47)
        const T ref = cur.ref;
        const Slice<T> rs = slice_splat<T>(ref);
        const T* sT = reinterpret_cast<const T*>(stage);
        T* gbase = o - mis;
        const uint32_t end = mis + total;
        const uint32_t vfirst = mis != 0 ? 1u : 0u;  // first vector lying entirely inside the run
        const uint32_t vlast = end / EPV;            // one past the last such vector
        unsigned char* gb = reinterpret_cast<unsigned char*>(gbase);
#pragma unroll 1
        for (uint32_t off = (vfirst + lane) * 16u; off < vlast * 16u; off += 512u) {
            const Slice<T> val = slice_add<T>(to_slice<T>(*reinterpret_cast<const uint4*>(stage + off)), rs);
            stg128_stream(gb + off, from_slice<T>(val));
        }
        if (lane < EPV) {
            const uint32_t eh = uint32_t(lane);  // head: run elements inside vector 0 when the run starts mid-vector
            if (mis != 0 && eh >= mis && eh < end) gbase[eh] = T(sT[eh] + ref);
            const uint32_t et = vlast * EPV + uint32_t(lane);  // tail: the elements after the last full vector
            if (vlast >= vfirst && et < end) gbase[et] = T(sT[et] + ref);
        }
    };

    SelectLoads<T, W> ld0, ld1;
    select_issue_loads<T, W>(ld0, blk0, lane, q, j, packed, bitmap, offsets, refs, ref_scalar);
    if constexpr (NB == 1) {
        one_block(blk0, false, ld0, ld1);
    } else {
        const size_t blk_end = (n_blocks - blk0 < size_t(NB)) ? n_blocks : blk0 + NB;
#pragma unroll 1
        for (size_t blk = blk0; blk < blk_end; blk += 2) {
            one_block(blk, blk + 1 < blk_end, ld0, ld1);
            if (blk + 1 >= blk_end) break;
            one_block(blk + 1, blk + 2 < blk_end, ld1, ld0);
        }
    }
}

}  // namespace flb
