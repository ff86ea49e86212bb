// fl_kernels.cuh — batched sm_100a kernels of the FastLanes hot path.
//
// Two thread mappings over the same per-thread primitives (fl_device.cuh):
//
//  * WARP-BLOCK (shipped): one warp = one 1024-value block.  The four 8-thread groups split the T rows into runs of
//    T/4 consecutive rows; a thread owns the 16-byte column slice j of its group's rows.  Every warp-wide access is
//    4 x 128 B, contiguous for u32/u64 (512 B), and a block is read/written by a handful of back-to-back
//    instructions of one warp: measured 6.5-7.2 TB/s on every op (profiles/opbench_r01.txt).  Cross-group
//    dependencies (the delta prefix chain, packed words straddling two runs) go through warp shuffles; the
//    original-order variants (fused untranspose / transpose, standalone transposes) stage one block per warp in an
//    XOR-swizzled shared-memory tile guarded by __syncwarp only.
//        unpack_warp_kernel, pack_warp_kernel, delta_warp_kernel, transpose_warp_kernel
//
//  * ROW-SLICE (first correct path; kept for the A/B measurement in tools/kbench.cu, and used for u8 fused delta at
//    W < 8 where it is faster): 8 threads = one block, a warp = 4 consecutive blocks; no shared memory, no shuffles:
//    the serial per-lane chain of the reference (T rows, src/macros.rs:139-170; the prefix-add of
//    src/delta.rs:48-63) is a per-thread register chain.  5.8-6.2 TB/s: each warp instruction touches four
//    different blocks and a block takes T instructions to complete, which the DRAM controller likes less.
//        unpack_kernel, pack_kernel, delta_kernel
//
// Every kernel is HBM-bound integer shift/mask/add; algorithmic bytes per block are in DESIGN.md §5.
#pragma once
#include <type_traits>
#include <utility>
#include <array>

#include "fl_device.cuh"

namespace flb {

// UOP_DELTA_ORIG: fused undelta_pack + untranspose (output in ORIGINAL value order)  — SURVEY.md §8(f) rank 1
// POP_ORIG_DELTA: fused transpose + delta + pack   (input  in ORIGINAL value order)
enum UnpackOp : int { UOP_PLAIN = 0, UOP_FOR = 1, UOP_DELTA = 2, UOP_DELTA_ORIG = 3 };
// POP_FOR_AUTO:   for_pack with reference = the block's own minimum, computed in the same pass
enum PackOp : int { POP_PLAIN = 0, POP_FOR = 1, POP_ORIG_DELTA = 2, POP_FOR_AUTO = 3 };

#ifndef FLB_THREADS
#define FLB_THREADS 256
#endif
#ifndef FLB_U64_DELTA_OCC_W
#define FLB_U64_DELTA_OCC_W 12  // widths below this use the 3-CTAs-per-SM register cap in the u64 fused-delta kernel
#endif
constexpr int kThreads = FLB_THREADS;  // 256 threads = 32 blocks of 1024 values per CTA
constexpr int kSlicesPerBlock = 8;  // 8 x 16 B = one 128-byte row

// Software prefetch distance in word-rows (unpack) / rows (pack, delta): loads are issued this many
// iterations ahead of use.
#ifndef FLB_PREFETCH
#define FLB_PREFETCH 8
#endif

// seq_t! (src/lib.rs:41-47): compile-time unroll of `row in 0..N`, ROW available as a constant.
template <int... R, class F>
__device__ __forceinline__ void seq_impl(std::integer_sequence<int, R...>, F&& f) {
    (f(std::integral_constant<int, R>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void seq_rows(F&& f) {
    seq_impl(std::make_integer_sequence<int, N>{}, static_cast<F&&>(f));
}

// ---------------------------------------------------------------------------------------------------
// unpack family:  a9 BitPacking::unpack (src/bitpacking.rs:98-107), a19 FoR::unfor_pack
// (src/ffor.rs:38-50), a17 Delta::undelta_pack (src/delta.rs:48-63).  OP selects the closure the
// reference splices into unpack!.
//   packed : n_blocks x (128*W bytes)        out  : n_blocks x (128*T bytes)
//   refs   : per-block reference (UOP_FOR; nullptr -> ref_scalar)
//   base   : n_blocks x 128 bytes (UOP_DELTA; LANES elements per block)
// ---------------------------------------------------------------------------------------------------
template <class T, int W, int OP>
__device__ __forceinline__ void unpack_slice(const char* __restrict__ pk, char* __restrict__ o, Slice<T> extra) {
    constexpr int TB = Lay<T>::TB;

    auto emit = [&](int row_off, Slice<T> v) {
        if constexpr (OP == UOP_FOR) v = slice_add<T>(v, extra);  // ffor.rs:47
        if constexpr (OP == UOP_DELTA) {                           // delta.rs:58-60
            extra = slice_add<T>(extra, v);
            v = extra;
        }
        store_slice<T>(o + row_off, v);
    };

    if constexpr (W == 0) {
        // macros.rs:118-125: zeros, still visiting every row in order (the delta closure needs that)
        seq_rows<TB>([&](auto rc) { emit(row_byte_offset<T>(decltype(rc)::value), slice_zero<T>()); });
    } else {
        constexpr int D = (FLB_PREFETCH < W) ? FLB_PREFETCH : W;
        Slice<T> w[W];
        seq_rows<D>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            w[k] = load_slice<T>(pk + k * 128);
        });
        seq_rows<TB>([&](auto rc) {
            constexpr int row = decltype(rc)::value;
            constexpr int curr = (row * W) / TB;  // macros.rs:144
            constexpr bool first_of_word = (row == 0) || (((row - 1) * W) / TB != curr);
            // entering word-row `curr`: issue the load D word-rows ahead
            if constexpr (first_of_word && curr > 0 && curr + D - 1 < W) {
                w[curr + D - 1] = load_slice<T>(pk + (curr + D - 1) * 128);
            }
            if constexpr (W == TB) {
                emit(row_byte_offset<T>(row), w[row]);  // macros.rs:126-132
            } else {
                // `nxt` is only read when the field straddles a word boundary (then curr+1 < W, macros.rs:156)
                constexpr int nxt = (curr + 1 < W) ? curr + 1 : curr;
                emit(row_byte_offset<T>(row), extract_row<T, W, row>(w[curr], w[nxt]));
            }
        });
    }
}

template <class T, int W, int OP>
__global__ void __launch_bounds__(kThreads)
unpack_kernel(const char* __restrict__ packed, char* __restrict__ out, size_t n_blocks,
              const T* __restrict__ refs, T ref_scalar, const char* __restrict__ base) {
    const size_t tid = size_t(blockIdx.x) * kThreads + threadIdx.x;
    const size_t blk = tid / kSlicesPerBlock;
    const int j = int(tid % kSlicesPerBlock);
    if (blk >= n_blocks) return;
    Slice<T> extra = slice_zero<T>();
    if constexpr (OP == UOP_FOR) extra = slice_splat<T>(refs ? refs[blk] : ref_scalar);
    if constexpr (OP == UOP_DELTA) extra = load_slice<T>(base + blk * 128 + j * 16);  // prev = base[lane]
    unpack_slice<T, W, OP>(packed + blk * (size_t(128) * W) + j * 16,
                           out + blk * (size_t(128) * Lay<T>::TB) + j * 16, extra);
}

// ---------------------------------------------------------------------------------------------------
// unpack family, "warp-block" layout (the fast path; profiles/access_pattern_r01.txt shows why):
// ONE WARP = ONE BLOCK.  The four 8-thread groups of the warp split the T rows of the block into four runs
// of RPG = T/4 consecutive rows; thread (g, j) owns the 16-byte column slice j of its group's rows.  Each
// warp-wide store then writes 4 x 128 B that are contiguous (u32/u64: 512 B; u16: 2 x 256 B) and the
// whole block is written by T/4 back-to-back store instructions of one warp, which is what the HBM
// controller wants (row-slice layout: ~5.9 TB/s, this layout: ~6.7 TB/s on the same bytes).
//
// A group's rows occupy bits [q*RPG*W, (q+1)*RPG*W) of every lane stream (q = rank of the group in row
// order).  The group loads the NA(+1) word-rows covering that range and pre-shifts them by the run-time
// offset sh0 = (q*RPG*W) mod T with one funnel shift per word; after that every shift, mask and register
// index is a compile-time constant again (same extract_row<> as the row-slice kernel, ROW = local row).
// The fused delta prefix-add (src/delta.rs:48-63) becomes a per-thread scan over RPG rows plus a 4-group
// exclusive scan of the run totals through warp shuffles.
// ---------------------------------------------------------------------------------------------------
template <class T>
struct WarpLay {
    static constexpr int TB = Lay<T>::TB;
    static constexpr int RPG = TB / 4;  // rows per group
    // rank (position in row order) of group g.  u32/u64: {0,2,1,3} makes the 4 groups' rows adjacent in memory.
    __device__ static __forceinline__ int rank_of_group(int g) {
        if constexpr (sizeof(T) >= 4) return ((g << 1) & 2) | (g >> 1);  // swaps 1 and 2 (g < 4) without the branchy select chain
        else return g;
    }
    __device__ static __forceinline__ int group_of_rank(int q) { return rank_of_group(q); }  // involution / identity
};

// lane-wise funnel shift right by a RUN-TIME amount sh (0 <= sh < T): low T-sh bits from lo>>sh, rest from hi
template <class T>
__device__ __forceinline__ typename Lay<T>::R lane_funnel_rt(typename Lay<T>::R lo, typename Lay<T>::R hi, unsigned sh,
                                                             typename Lay<T>::R mlow) {
    using R = typename Lay<T>::R;
    if constexpr (sizeof(T) == 4) {
        return __funnelshift_r(lo, hi, sh);
    } else if constexpr (sizeof(T) == 8) {
        return (lo >> sh) | ((hi << 1) << (63u - sh));
    } else {
        constexpr unsigned TBu = Lay<T>::TB;
        return ((lo >> sh) & mlow) | ((hi << (TBu - sh)) & R(~mlow));
    }
}

// Order in which a thread visits its RPG local rows for global loads/stores: address-sequential.
// u32 (RPG 8) and u8/u16 are already sequential in i; u64 (RPG 16) alternates the two 8-row halves: 0,8,1,9,...
template <int RPG>
__host__ __device__ constexpr int warp_visit_row(int ii) { return (RPG == 16) ? ((ii & 1) * 8 + (ii >> 1)) : ii; }

// byte offset of local row i of the run of rank q (global row r = q*RPG + i), see row_byte_offset()
// LINEAR = the row order of the original cwida/FastLanes bit-packing (row r = values r*LANES .. r*LANES+LANES-1, i.e.
// byte offset r*128), which this crate reorders (README.md:49-56, src/macros.rs:1-9).
template <class T, int I, bool LINEAR = false>
__device__ __forceinline__ int warp_row_offset(int q) {
    constexpr int RPG = WarpLay<T>::RPG;
    if constexpr (LINEAR) {
        return (q * RPG + I) * 128;
    } else if constexpr (RPG >= 8) {
        const int oidx = q * (RPG / 8) + I / 8;  // r/8 ; r%8 = I%8
        return (fl_order_rt(oidx) * 16 + (I % 8) * 128) * int(sizeof(T));
    } else {
        const int r = q * RPG + I;
        return (fl_order_rt(r >> 3) * 16 + (r & 7) * 128) * int(sizeof(T));
    }
}

// Original-order placement (src/transpose.rs:29-36 composed with src/macros.rs:20-24): lane l of the transposed
// vector walks T CONSECUTIVE originals starting at start(l) = 64*(l%16) + 8*FL_ORDER[l/16], so
// the RPG rows a thread holds for one lane are RPG consecutive originals.  Byte offset of that run inside the
// block, for SWAR register r of slice j in the run of rank q (u32/u64 only: one lane per register):
template <class T>
__device__ __forceinline__ int orig_run_byte_offset(int q, int j, int k) {
    const int l = (16 / int(sizeof(T))) * j + k;  // lane index: thread j owns lanes [16/S * j, 16/S * (j+1))
    return (64 * (l & 15) + 8 * fl_order_rt(l >> 4) + q * WarpLay<T>::RPG) * int(sizeof(T));
}
// chunk m (16 bytes = EPC consecutive rows) of lane-register r out of the row-major register tile v[row].r[r]
template <class T, int RPG>
__device__ __forceinline__ uint4 gather_rows_chunk(const Slice<T> (&v)[RPG], int r, int m) {
    if constexpr (sizeof(T) == 4) {
        return make_uint4(v[4 * m].r[r], v[4 * m + 1].r[r], v[4 * m + 2].r[r], v[4 * m + 3].r[r]);
    } else {
        const uint64_t a = v[2 * m].r[r], b = v[2 * m + 1].r[r];
        return make_uint4(uint32_t(a), uint32_t(a >> 32), uint32_t(b), uint32_t(b >> 32));
    }
}
template <class T, int RPG>
__device__ __forceinline__ void scatter_rows_chunk(Slice<T> (&v)[RPG], int r, int m, uint4 c) {
    if constexpr (sizeof(T) == 4) {
        v[4 * m].r[r] = c.x; v[4 * m + 1].r[r] = c.y; v[4 * m + 2].r[r] = c.z; v[4 * m + 3].r[r] = c.w;
    } else {
        v[2 * m].r[r] = (uint64_t(c.y) << 32) | c.x;
        v[2 * m + 1].r[r] = (uint64_t(c.w) << 32) | c.z;
    }
}

// The fused original-order ops stage one block per warp in shared memory (warp-private tile, __syncwarp only):
// the register tile is scattered/gathered at its ORIGINAL byte offset A with 16-byte accesses, the global side
// is a linear 512-bytes-per-instruction copy.  XOR swizzle of the 16-byte bank group (A bits 4..6) with the
// address bits that vary across a quarter-warp in the scatter (u32: lane bits at A[7], A[10..11]; u64: A[10..12])
// makes BOTH sides bank-conflict free.
template <class T>
__device__ __forceinline__ int orig_tile_swizzle(int A) {
    if constexpr (sizeof(T) == 4) return A ^ ((((A >> 7) & 1) | (((A >> 10) & 3) << 1)) << 4);
    else if constexpr (sizeof(T) == 8) return A ^ (((A >> 10) & 7) << 4);
    else if constexpr (sizeof(T) == 2) return A ^ (((A >> 10) & 1) << 4);  // see orig_tile_scatter (u16)
    else return A;  // u8 scatters 2-byte pieces: no bank-group structure to fix
}

// register tile v[row].r[..]  ->  warp-private shared tile in ORIGINAL order (this thread's lanes, its run of rows)
template <class T, int RPG>
__device__ __forceinline__ void orig_tile_scatter(unsigned char* tile, const Slice<T> (&v)[RPG], int q, int j) {
    constexpr int NR = Lay<T>::NR;
    if constexpr (sizeof(T) >= 4) {
        constexpr int EPC = 16 / int(sizeof(T));
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const int A0 = orig_run_byte_offset<T>(q, j, r);
#pragma unroll
            for (int m = 0; m < RPG / EPC; ++m)
                *reinterpret_cast<uint4*>(tile + orig_tile_swizzle<T>(A0 + m * 16)) = gather_rows_chunk<T, RPG>(v, r, m);
        }
    } else if constexpr (sizeof(T) == 2) {
        // RPG = 4 rows x 8 lanes: per lane 4 consecutive u16 = 8 bytes at A = 128*(8(j&1) + k) + 16*FL_ORDER[j>>1] + 8q.
        // A 64-bit shared access is served per HALF-warp (two row groups q): its 16 lanes hit 4*FL_ORDER[j>>1] + 2q (+1),
        // 16 distinct 8-byte bank pairs, but j&1 (address bit 10) maps two lanes onto each pair: a 2-way conflict on
        // every STS.64 / LDS.64 (ncu: 16.8 M extra wavefronts per 2^20 blocks, profiles/ncu_r02_kernels.md).  Bit 4 (the
        // 16-byte bank group, = bit 2 of q's contribution, constant inside a half-warp) is XORed with bit 10.
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t sel = (k & 1) ? 0x7632u : 0x5410u;
            uint2 p;
            p.x = __byte_perm(v[0].r[k >> 1], v[1].r[k >> 1], sel);
            p.y = __byte_perm(v[2].r[k >> 1], v[3].r[k >> 1], sel);
            *reinterpret_cast<uint2*>(tile + orig_tile_swizzle<T>(orig_run_byte_offset<T>(q, j, k))) = p;
        }
    } else {
        // RPG = 2 rows x 16 lanes: per lane 2 consecutive u8 = 2 bytes
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const uint32_t b = k & 3;
            const uint32_t p = __byte_perm(v[0].r[k >> 2], v[1].r[k >> 2], ((4u + b) << 4) | b);
            *reinterpret_cast<uint16_t*>(tile + orig_run_byte_offset<T>(q, j, k)) = uint16_t(p);
        }
    }
}
// A 16-byte shared load the compiler may not narrow.  With u64 at W <= 32 only the low word of every value is live after
// the delta, and the compiler turns each LDS.128 of the gather below into two LDS.32 (offsets 0 and 8).  A 32-bit access
// is served for the whole warp at once, and the swizzle — built for 16-byte accesses, which are served per quarter-warp —
// leaves the four row groups on the same banks: 4 wavefronts per LDS.32 instead of 1, 96 excess wavefronts per block
// (ncu source page, profiles/ncu_r02_u64_orig.md).  The full 16-byte load costs 4 wavefronts for both words.
__device__ __forceinline__ uint4 lds128_full(const void* p) {
    uint4 v;
    asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_addr(p)) : "memory");
    return v;
}

// inverse: shared tile in ORIGINAL order -> register tile
template <class T, int RPG>
__device__ __forceinline__ void orig_tile_gather(const unsigned char* tile, Slice<T> (&v)[RPG], int q, int j) {
    constexpr int NR = Lay<T>::NR;
    if constexpr (sizeof(T) >= 4) {
        constexpr int EPC = 16 / int(sizeof(T));
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const int A0 = orig_run_byte_offset<T>(q, j, r);
#pragma unroll
            for (int m = 0; m < RPG / EPC; ++m) {
                const unsigned char* src = tile + orig_tile_swizzle<T>(A0 + m * 16);
                if constexpr (sizeof(T) == 8) scatter_rows_chunk<T, RPG>(v, r, m, lds128_full(src));
                else scatter_rows_chunk<T, RPG>(v, r, m, *reinterpret_cast<const uint4*>(src));
            }
        }
    } else if constexpr (sizeof(T) == 2) {
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
            const uint2 a = *reinterpret_cast<const uint2*>(tile + orig_tile_swizzle<T>(orig_run_byte_offset<T>(q, j, 2 * rr)));
            const uint2 b = *reinterpret_cast<const uint2*>(tile + orig_tile_swizzle<T>(orig_run_byte_offset<T>(q, j, 2 * rr + 1)));
            v[0].r[rr] = __byte_perm(a.x, b.x, 0x5410u);
            v[1].r[rr] = __byte_perm(a.x, b.x, 0x7632u);
            v[2].r[rr] = __byte_perm(a.y, b.y, 0x5410u);
            v[3].r[rr] = __byte_perm(a.y, b.y, 0x7632u);
        }
    } else {
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
            uint32_t p[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) p[x] = *reinterpret_cast<const uint16_t*>(tile + orig_run_byte_offset<T>(q, j, 4 * rr + x));
            const uint32_t t01 = __byte_perm(p[0], p[1], 0x5140u);  // (p0.row0, p1.row0, p0.row1, p1.row1)
            const uint32_t t23 = __byte_perm(p[2], p[3], 0x5140u);
            v[0].r[rr] = __byte_perm(t01, t23, 0x5410u);
            v[1].r[rr] = __byte_perm(t01, t23, 0x7632u);
        }
    }
}

template <class R>
__device__ __forceinline__ R shfl_reg(R v, int src) {
    if constexpr (sizeof(R) == 8) return R(__shfl_sync(0xffffffffu, (unsigned long long)v, src));
    else return R(__shfl_sync(0xffffffffu, v, src));
}

template <class R>
__device__ __forceinline__ R shfl_xor_reg(R v, int mask) {
    if constexpr (sizeof(R) == 8) return R(__shfl_xor_sync(0xffffffffu, (unsigned long long)v, mask));
    else return R(__shfl_xor_sync(0xffffffffu, v, mask));
}

// lane-wise funnel shift LEFT by a run-time amount sh (0 <= sh < T): (hi << sh) | (lo >> (T - sh))
template <class T>
__device__ __forceinline__ typename Lay<T>::R lane_funnel_left_rt(typename Lay<T>::R lo, typename Lay<T>::R hi, unsigned sh,
                                                                  typename Lay<T>::R mlow) {
    using R = typename Lay<T>::R;
    if constexpr (sizeof(T) == 4) {
        return __funnelshift_l(lo, hi, sh);
    } else if constexpr (sizeof(T) == 8) {
        return (hi << sh) | ((lo >> 1) >> (63u - sh));
    } else {
        constexpr unsigned TBu = Lay<T>::TB;
        return ((hi << sh) & R(~mlow)) | ((lo >> (TBu - sh)) & mlow);  // mlow = low `sh` bits of every lane
    }
}

// warp_load_run / warp_extract_rows / warp_decode_tile: the decode core shared by the unpack family and the fused
// scan kernels (this thread: group rank q, 16-byte slice j of block `blk_packed`).
//   warp_load_run     -> a[m]: the run_words<T,W>() word-rows holding the bits of the group's RPG rows, ALIGNED: local
//                        row i occupies bits [(i*W) % T, ...) of a[(i*W) / T], every position a compile-time constant
//   warp_extract_rows -> v[i] = slice j of row q*RPG + i (right-aligned W-bit values)
// TMA = true: the block's 128*W packed bytes arrive with ONE cp.async.bulk (TMA 1-D bulk copy) per warp into a
// warp-private shared buffer, completion on a per-warp mbarrier; the word-rows are then read with LDS.128.  Every
// packed byte crosses L2 exactly once (the direct path re-reads the word-rows that two groups share when W % 4 != 0).
// Callers issue their independent global loads (delta bases) BEFORE this call: the mbarrier wait is a compiler
// memory barrier, anything after it would be serialised behind the TMA round trip.
// RESERVE: static shared memory (bytes) the calling kernel declares for itself (counts against the 48 KiB limit).
template <class T, int W>
__host__ __device__ constexpr int run_words() {
    constexpr int TB = Lay<T>::TB, RPG = TB / 4;
    return W == 0 ? 1 : (W == TB ? RPG : (RPG * W + TB - 1) / TB);
}

// Assembles the aligned run a[] of group rank q from a word-row loader (k -> this thread's 16-byte slice of word-row k).
template <class T, int W, class Loader>
__device__ __forceinline__ void warp_run_from(Loader&& load_word_row, int q, Slice<T> (&a)[run_words<T, W>()]) {
    using R = typename Lay<T>::R;
    constexpr int TB = Lay<T>::TB;
    constexpr int RPG = WarpLay<T>::RPG;
    constexpr int NR = Lay<T>::NR;
    if constexpr (W == 0) {
        a[0] = slice_zero<T>();
    } else if constexpr (W == TB) {
        // macros.rs:126-132: row r is word-row r
        seq_rows<RPG>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            a[i] = load_word_row(unsigned(q * RPG + i));
        });
    } else {
        constexpr bool ALIGNED = (W % 4) == 0;              // then q*RPG*W is a multiple of T for every q
        constexpr int NA = (RPG * W + TB - 1) / TB;         // word-rows holding the run once aligned
        constexpr int NL = ALIGNED ? NA : NA + 1;           // word-rows to load
        const unsigned bit0 = unsigned(q) * (RPG * W);
        const unsigned k0 = bit0 / TB, sh0 = bit0 % TB;
        Slice<T> w[NL];
        seq_rows<NL>([&](auto mc) {
            constexpr int m = decltype(mc)::value;
            const unsigned k = (k0 + m < unsigned(W)) ? k0 + m : unsigned(W - 1);  // clamp: never read past the block
            w[m] = load_word_row(k);
        });
        if constexpr (ALIGNED) {
#pragma unroll
            for (int m = 0; m < NA; ++m) a[m] = w[m];
        } else {
            R mlow = 0;
            if constexpr (sizeof(T) == 2) { const uint32_t mm = 0xFFFFu >> sh0; mlow = mm | (mm << 16); }
            if constexpr (sizeof(T) == 1) { mlow = (0xFFu >> sh0) * 0x01010101u; }
            seq_rows<NA>([&](auto mc) {
                constexpr int m = decltype(mc)::value;
#pragma unroll
                for (int r = 0; r < NR; ++r) a[m].r[r] = lane_funnel_rt<T>(w[m].r[r], w[m + 1].r[r], sh0, mlow);
            });
        }
    }
}

template <class T, int W, bool TMA, int RESERVE = 0>
__device__ __forceinline__ void warp_load_run(const char* __restrict__ blk_packed, int lane, int q, int j,
                                              Slice<T> (&a)[run_words<T, W>()]) {
    const char* pk = blk_packed + j * 16;

    // word-row k of this block, this thread's 16-byte slice
    constexpr bool USE_TMA = TMA && W > 0 && (kThreads / 32) * 128 * W + 128 + RESERVE <= 48 * 1024;  // static shared-memory limit (u64: W <= 47)
    __shared__ __align__(128) unsigned char tma_buf[USE_TMA ? kThreads / 32 : 1][USE_TMA ? 128 * W : 16];
    __shared__ __align__(8) unsigned long long tma_bar[USE_TMA ? kThreads / 32 : 1];
    const unsigned char* sp = nullptr;
    if constexpr (USE_TMA) {
        const int wi = threadIdx.x >> 5;
        const unsigned bar = smem_addr(&tma_bar[wi]);
        if (lane == 0) mbar_init(bar, 1);
        __syncwarp();
        if (lane == 0) {
            mbar_arrive_expect_tx(bar, 128 * W);
            tma_bulk_load(smem_addr(&tma_buf[wi][0]), blk_packed, 128 * W, bar);
        }
        mbar_wait_parity(bar, 0);
        sp = &tma_buf[wi][0] + j * 16;
    }
    auto load_word_row = [&](unsigned k) -> Slice<T> {
        if constexpr (USE_TMA) return to_slice<T>(*reinterpret_cast<const uint4*>(sp + k * 128));
        else return load_slice<T>(pk + k * 128);
    };

    warp_run_from<T, W>(load_word_row, q, a);
}

template <class T, int W>
__device__ __forceinline__ void warp_extract_rows(const Slice<T> (&a)[run_words<T, W>()], Slice<T> (&v)[WarpLay<T>::RPG]) {
    constexpr int TB = Lay<T>::TB;
    constexpr int RPG = WarpLay<T>::RPG;
    if constexpr (W == 0) {
#pragma unroll
        for (int i = 0; i < RPG; ++i) v[i] = slice_zero<T>();
    } else if constexpr (W == TB) {
#pragma unroll
        for (int i = 0; i < RPG; ++i) v[i] = a[i];
    } else {
        constexpr int NA = run_words<T, W>();
        seq_rows<RPG>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            constexpr int idx = (i * W) / TB;
            constexpr int nxt = (idx + 1 < NA) ? idx + 1 : idx;  // only read when the field straddles
            v[i] = extract_row<T, W, i>(a[idx], a[nxt]);
        });
    }
}

template <class T, int W, bool TMA, int RESERVE = 0>
__device__ __forceinline__ void warp_decode_tile(const char* __restrict__ blk_packed, int lane, int q, int j,
                                                 Slice<T> (&v)[WarpLay<T>::RPG]) {
    Slice<T> a[run_words<T, W>()];
    warp_load_run<T, W, TMA, RESERVE>(blk_packed, lane, q, j, a);
    warp_extract_rows<T, W>(a, v);
}

// u64 fused delta: 16 rows x 2 x 64-bit per thread want 88 registers = 2 CTAs per SM; the dependent shuffle scan needs
// more resident warps than that at small W, so the compiler is held to 3 CTAs per SM (80 registers) there.
template <class T, int OP, int W>
constexpr int unpack_min_ctas() { return (sizeof(T) == 8 && OP == UOP_DELTA && W < FLB_U64_DELTA_OCC_W) ? 3 : 1; }

template <class T, int W, int OP, bool TMA = false, bool LINEAR = false, int MINB = unpack_min_ctas<T, OP, W>()>
__global__ void __launch_bounds__(kThreads, MINB)
unpack_warp_kernel(const char* __restrict__ packed, char* __restrict__ out, size_t n_blocks,
                   const T* __restrict__ refs, T ref_scalar, const char* __restrict__ base) {
    using R = typename Lay<T>::R;
    using WL = WarpLay<T>;
    constexpr int TB = Lay<T>::TB;
    constexpr int RPG = WL::RPG;
    constexpr int NR = Lay<T>::NR;
    const size_t blk = (size_t(blockIdx.x) * kThreads + threadIdx.x) >> 5;
    if (blk >= n_blocks) return;  // warp-uniform
    const int lane = threadIdx.x & 31;
    const int g = lane >> 3, j = lane & 7;
    static_assert(!LINEAR || OP == UOP_PLAIN || OP == UOP_FOR, "cwida row order: bit-packing and FoR only");
    const int q = LINEAR ? g : WL::rank_of_group(g);  // linear rows: groups in row order
    char* o = out + blk * (size_t(128) * TB) + j * 16;

    // prev = base[lane] (delta.rs:50): issued before the decode's TMA wait
    Slice<T> carry = slice_zero<T>();
    if constexpr (OP == UOP_DELTA || OP == UOP_DELTA_ORIG) carry = load_slice<T>(base + blk * 128 + j * 16);
    Slice<T> v[RPG];
    warp_decode_tile<T, W, TMA>(packed + blk * (size_t(128) * W), lane, q, j, v);

    if constexpr (OP == UOP_FOR) {
        const Slice<T> ref = slice_splat<T>(refs ? refs[blk] : ref_scalar);  // ffor.rs:47
#pragma unroll
        for (int i = 0; i < RPG; ++i) v[i] = slice_add<T>(v[i], ref);
    }
    if constexpr (OP == UOP_DELTA || OP == UOP_DELTA_ORIG) {
        // delta.rs:56-60: running wrapping sum along rows per lane, seeded with base[lane]
#pragma unroll
        for (int i = 1; i < RPG; ++i) v[i] = slice_add<T>(v[i], v[i - 1]);
#pragma unroll
        for (int qq = 0; qq < 3; ++qq) {  // totals of the runs that precede this one in row order
            const int src = WL::group_of_rank(qq) * 8 + j;
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                R t;
                if constexpr (sizeof(R) == 8) t = __shfl_sync(0xffffffffu, (unsigned long long)v[RPG - 1].r[r], src);
                else t = __shfl_sync(0xffffffffu, v[RPG - 1].r[r], src);
                if (qq < q) carry.r[r] = lane_add<T>(carry.r[r], t);
            }
        }
#pragma unroll
        for (int i = 0; i < RPG; ++i) v[i] = slice_add<T>(v[i], carry);
    }
    if constexpr (OP == UOP_DELTA_ORIG) {
        // untranspose fused into the store (transpose.rs:18-22): per lane, the RPG rows are RPG consecutive
        // originals -> RPG*sizeof(T)/16 contiguous 16-byte chunks; the row->chunk regrouping is register renaming.
        extern __shared__ __align__(128) unsigned char orig_tile_smem[];
        unsigned char* tile = orig_tile_smem + (threadIdx.x >> 5) * (128 * TB);
        orig_tile_scatter<T, RPG>(tile, v, q, j);
        __syncwarp();
        char* ob = out + blk * (size_t(128) * TB);
#pragma unroll
        for (int i = 0; i < (128 * TB) / 512; ++i) {
            const int A = i * 512 + lane * 16;
            stg128_stream(ob + A, *reinterpret_cast<const uint4*>(tile + orig_tile_swizzle<T>(A)));
        }
        return;
    }
    seq_rows<RPG>([&](auto ic) {
        // u64 (RPG = 16): visit local rows as 0,8,1,9,... so that consecutive warp stores are address-sequential
        constexpr int i = warp_visit_row<RPG>(decltype(ic)::value);
        // global row r = q*RPG + i: offset (FL_ORDER[r/8]*16 + (r%8)*128)*sizeof(T)  (macros.rs:20-24)
        int off;
        if constexpr (LINEAR) {
            off = (q * RPG + i) * 128;  // row r at byte r*128
        } else if constexpr (RPG >= 8) {
            // r/8 = q*(RPG/8) + i/8 ; r%8 = i%8
            const int oidx = q * (RPG / 8) + i / 8;
            off = (fl_order_rt(oidx) * 16 + (i % 8) * 128) * int(sizeof(T));
        } else {
            const int r = q * RPG + i;
            off = (fl_order_rt(r >> 3) * 16 + (r & 7) * 128) * int(sizeof(T));
        }
        store_slice<T>(o + off, v[i]);
    });
}

// ---------------------------------------------------------------------------------------------------
// pack family, warp-block layout: a7 BitPacking::pack (src/bitpacking.rs:65-74), a18 FoR::for_pack
// (src/ffor.rs:24-36).  Same thread mapping as unpack_warp_kernel.  Each group packs its T/4 rows into an
// aligned run (compile-time shifts), funnel-shifts the run left by sh0 into stream position, and word-rows
// that straddle two (or, for W < 4, up to four) groups are OR-merged through warp shuffles: the
// higher-rank group owns a shared word.
// ---------------------------------------------------------------------------------------------------
// TMA = true (plain / FoR ops): the 128*T-byte unpacked block arrives with one cp.async.bulk per warp.
template <class T, int W, int OP, bool TMA = false, bool LINEAR = false>
__global__ void __launch_bounds__(kThreads)
pack_warp_kernel(const char* __restrict__ in, char* __restrict__ packed, size_t n_blocks,
                 const T* __restrict__ refs, T ref_scalar, const char* __restrict__ base,
                 T* __restrict__ refs_out = nullptr, T* __restrict__ spans_out = nullptr) {
    using R = typename Lay<T>::R;
    using WL = WarpLay<T>;
    constexpr int TB = Lay<T>::TB;
    constexpr int RPG = WL::RPG;
    constexpr int NR = Lay<T>::NR;
    const size_t blk = (size_t(blockIdx.x) * kThreads + threadIdx.x) >> 5;
    if (blk >= n_blocks) return;  // warp-uniform
    if constexpr (W == 0 && OP != POP_FOR_AUTO) return;  // macros.rs:52 (FOR_AUTO still reports the statistics)
    const int lane = threadIdx.x & 31;
    const int g = lane >> 3, j = lane & 7;
    static_assert(!LINEAR || OP == POP_PLAIN || OP == POP_FOR, "cwida row order: bit-packing and FoR only");
    const int q = LINEAR ? g : WL::rank_of_group(g);
    const char* ip = in + blk * (size_t(128) * TB) + j * 16;
    char* pk = packed + blk * (size_t(128) * W) + j * 16;

    Slice<T> src[RPG];
    if constexpr (OP == POP_ORIG_DELTA) {
        // transpose fused into the load (transpose.rs:11-15), then delta along rows (delta.rs:24-33)
        extern __shared__ __align__(128) unsigned char orig_tile_smem[];
        unsigned char* tile = orig_tile_smem + (threadIdx.x >> 5) * (128 * TB);
        const char* ib = in + blk * (size_t(128) * TB);
#pragma unroll
        for (int i = 0; i < (128 * TB) / 512; ++i) {
            const int A = i * 512 + lane * 16;
            *reinterpret_cast<uint4*>(tile + orig_tile_swizzle<T>(A)) = ldg128_stream(ib + A);
        }
        __syncwarp();
        orig_tile_gather<T, RPG>(tile, src, q, j);
        const Slice<T> b0 = load_slice<T>(base + blk * 128 + j * 16);  // base[lane]
        const int srcl = WL::group_of_rank(q > 0 ? q - 1 : 0) * 8 + j;
        Slice<T> first_prev;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const R t = shfl_reg<R>(src[RPG - 1].r[r], srcl);
            first_prev.r[r] = (q == 0) ? b0.r[r] : t;
        }
#pragma unroll
        for (int i = RPG - 1; i >= 1; --i) src[i] = slice_sub<T>(src[i], src[i - 1]);
        src[0] = slice_sub<T>(src[0], first_prev);
    } else {
        if constexpr (TMA) {
            // dynamic shared memory: [8 warps x 128*T bytes | 8 mbarriers]  (u64 needs 64 KiB: opt-in attribute)
            extern __shared__ __align__(128) unsigned char orig_tile_smem[];
            const int wi = threadIdx.x >> 5;
            unsigned char* tma_in = orig_tile_smem + wi * (128 * TB);
            const unsigned bar = smem_addr(orig_tile_smem + (kThreads / 32) * (128 * TB) + wi * 8);
            if (lane == 0) mbar_init(bar, 1);
            __syncwarp();
            if (lane == 0) {
                mbar_arrive_expect_tx(bar, 128 * TB);
                tma_bulk_load(smem_addr(tma_in), in + blk * (size_t(128) * TB), 128 * TB, bar);
            }
            mbar_wait_parity(bar, 0);
            const unsigned char* sp = tma_in + j * 16;
            seq_rows<RPG>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                // u64: whole 16-byte loads even where only the low words are live (W <= 32) — narrowed to LDS.32 they
                // conflict 4-way across the row groups (see lds128_full)
                if constexpr (sizeof(T) == 8) src[i] = to_slice<T>(lds128_full(sp + warp_row_offset<T, i, LINEAR>(q)));
                else src[i] = to_slice<T>(*reinterpret_cast<const uint4*>(sp + warp_row_offset<T, i, LINEAR>(q)));
            });
        } else {
            seq_rows<RPG>([&](auto ic) {
                constexpr int i = warp_visit_row<RPG>(decltype(ic)::value);
                src[i] = load_slice<T>(ip + warp_row_offset<T, i, LINEAR>(q));
            });
        }
    }
    if constexpr (OP == POP_FOR) {
        const Slice<T> ref = slice_splat<T>(refs ? refs[blk] : ref_scalar);  // ffor.rs:33
#pragma unroll
        for (int i = 0; i < RPG; ++i) src[i] = slice_sub<T>(src[i], ref);
    }
    if constexpr (OP == POP_FOR_AUTO) {
        // The statistics pass a caller runs before FoR::for_pack (which takes `reference` as a given, ffor.rs:5-10),
        // fused: the warp already holds the whole block, so min / max are a SWAR reduction over the register tile plus
        // a 5-step butterfly; reference = min, spans_out = max - min (the caller checks it fits W bits).
        using M = MinMax<typename std::conditional<sizeof(T) == 1, uint16_t, T>::type>;  // u8 reduces as 16x2 (below)
        R lo, hi;
        if constexpr (sizeof(T) == 1) {
            uint32_t mn_e = 0x00FF00FFu, mn_o = 0x00FF00FFu, mx_e = 0, mx_o = 0;
#pragma unroll
            for (int i = 0; i < RPG; ++i)
#pragma unroll
                for (int r = 0; r < NR; ++r) u8_minmax_acc(src[i].r[r], mn_e, mn_o, mx_e, mx_o);
            lo = __vminu2(mn_e, mn_o); hi = __vmaxu2(mx_e, mx_o);  // two 16-bit lanes holding byte values
        } else {
            lo = src[0].r[0]; hi = lo;
#pragma unroll
            for (int i = 0; i < RPG; ++i)
#pragma unroll
                for (int r = 0; r < NR; ++r) { lo = M::mn(lo, src[i].r[r]); hi = M::mx(hi, src[i].r[r]); }
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            lo = M::mn(lo, shfl_xor_reg<R>(lo, d));
            hi = M::mx(hi, shfl_xor_reg<R>(hi, d));
        }
        T mn, mx;
        if constexpr (sizeof(T) == 1) { mn = T(min(lo & 0xFFFFu, lo >> 16)); mx = T(max(hi & 0xFFFFu, hi >> 16)); }
        else { mn = swar_reduce_min<T>(lo); mx = swar_reduce_max<T>(hi); }
        if (lane == 0) {
            refs_out[blk] = mn;
            if (spans_out != nullptr) spans_out[blk] = T(mx - mn);
        }
        const Slice<T> ref = slice_splat<T>(mn);
#pragma unroll
        for (int i = 0; i < RPG; ++i) src[i] = slice_sub<T>(src[i], ref);  // ffor.rs:33
    }
    if constexpr (W == 0) {
        return;  // macros.rs:52: nothing to store
    } else if constexpr (W == TB) {
        // macros.rs:54-59: verbatim, word-row = row
        seq_rows<RPG>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            store_slice<T>(pk + (q * RPG + i) * 128, src[i]);
        });
    } else {
        constexpr int NA = (RPG * W + TB - 1) / TB;
        constexpr R MW = rep_mask<T>(W);
        Slice<T> a[NA + 1];
#pragma unroll
        for (int m = 0; m <= NA; ++m) a[m] = slice_zero<T>();
        seq_rows<RPG>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            constexpr int idx = (i * W) / TB;
            constexpr int sh = (i * W) % TB;
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                const R v = src[i].r[r] & MW;  // macros.rs:73
                if constexpr (sh + W <= TB) {
                    a[idx].r[r] |= (sh == 0) ? v : R(v << sh);  // fits the lane: cannot leak (macros.rs:79)
                } else {
                    a[idx].r[r] |= lane_shl<T, sh>(v);
                    a[idx + 1].r[r] |= lane_shr_keep<T, TB - sh, W - (TB - sh)>(v);  // macros.rs:92
                }
            }
        });
        constexpr bool ALIGNED = (W % 4) == 0;
        const unsigned bit0 = unsigned(q) * (RPG * W);
        const unsigned k0 = bit0 / TB;
        if constexpr (ALIGNED) {
            seq_rows<NA>([&](auto mc) {
                constexpr int m = decltype(mc)::value;
                store_slice<T>(pk + (k0 + m) * 128, a[m]);
            });
        } else {
            const unsigned sh0 = bit0 % TB;
            R mlow = 0;
            if constexpr (sizeof(T) == 2) { const uint32_t mm = (1u << sh0) - 1u; mlow = mm | (mm << 16); }
            if constexpr (sizeof(T) == 1) { mlow = ((1u << sh0) - 1u) * 0x01010101u; }
            // s[m] = stream word k0+m restricted to this group's bits, m = 0..NA  (a[NA] == 0 by construction)
            Slice<T> s[NA + 1];
            seq_rows<NA + 1>([&](auto mc) {
                constexpr int m = decltype(mc)::value;
#pragma unroll
                for (int r = 0; r < NR; ++r) {
                    const R lo = (m == 0) ? R(0) : a[m > 0 ? m - 1 : 0].r[r];
                    s[m].r[r] = lane_funnel_left_rt<T>(lo, a[m].r[r], sh0, mlow);
                }
            });
            // d = number of stream words this group owns = index of the word shared with / owned by the next rank
            const bool full = (sh0 + unsigned(RPG * W)) / TB == unsigned(NA);  // d == NA, else d == NA-1
#pragma unroll
            for (int step = 1; step <= 3; ++step) {
                const int srcl = (LINEAR ? step - 1 : WL::group_of_rank(step - 1)) * 8 + j;
#pragma unroll
                for (int r = 0; r < NR; ++r) {
                    const R mine = full ? s[NA].r[r] : s[NA - 1].r[r];  // s[d]; includes earlier merges when NA == 1
                    const R c = shfl_reg<R>(mine, srcl);
                    if (q == step && sh0 != 0) s[0].r[r] |= c;
                }
            }
            seq_rows<NA>([&](auto mc) {
                constexpr int m = decltype(mc)::value;
                if constexpr (m < NA - 1) store_slice<T>(pk + (k0 + m) * 128, s[m]);
                else if (full) store_slice<T>(pk + (k0 + m) * 128, s[m]);
            });
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// a15 Delta::delta / a16 Delta::undelta (src/delta.rs:24-45), warp-block layout.
// ---------------------------------------------------------------------------------------------------
template <class T, bool UNDO, bool TMA = false>
__global__ void __launch_bounds__(kThreads)
delta_warp_kernel(const char* __restrict__ in, const char* __restrict__ base, char* __restrict__ out, size_t n_blocks) {
    using R = typename Lay<T>::R;
    using WL = WarpLay<T>;
    constexpr int TB = Lay<T>::TB;
    constexpr int RPG = WL::RPG;
    constexpr int NR = Lay<T>::NR;
    const size_t blk = (size_t(blockIdx.x) * kThreads + threadIdx.x) >> 5;
    if (blk >= n_blocks) return;
    const int lane = threadIdx.x & 31;
    const int g = lane >> 3, j = lane & 7;
    const int q = WL::rank_of_group(g);
    const char* ip = in + blk * (size_t(128) * TB) + j * 16;
    char* op = out + blk * (size_t(128) * TB) + j * 16;
    Slice<T> v[RPG];
    Slice<T> carry = load_slice<T>(base + blk * 128 + j * 16);  // delta.rs:26 / :38: prev = base[lane]
    if constexpr (TMA) {
        // one cp.async.bulk of the whole 128*T-byte block per warp (dynamic shared memory: 8 buffers + 8 mbarriers)
        extern __shared__ __align__(128) unsigned char orig_tile_smem[];
        const int wi = threadIdx.x >> 5;
        unsigned char* buf = orig_tile_smem + wi * (128 * TB);
        const unsigned bar = smem_addr(orig_tile_smem + (kThreads / 32) * (128 * TB) + wi * 8);
        if (lane == 0) mbar_init(bar, 1);
        __syncwarp();
        if (lane == 0) {
            mbar_arrive_expect_tx(bar, 128 * TB);
            tma_bulk_load(smem_addr(buf), in + blk * (size_t(128) * TB), 128 * TB, bar);
        }
        mbar_wait_parity(bar, 0);
        seq_rows<RPG>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            v[i] = to_slice<T>(*reinterpret_cast<const uint4*>(buf + j * 16 + warp_row_offset<T, i>(q)));
        });
    } else {
        seq_rows<RPG>([&](auto ic) {
            constexpr int i = warp_visit_row<RPG>(decltype(ic)::value);
            v[i] = load_slice<T>(ip + warp_row_offset<T, i>(q));
        });
    }
    if constexpr (UNDO) {
#pragma unroll
        for (int i = 1; i < RPG; ++i) v[i] = slice_add<T>(v[i], v[i - 1]);  // delta.rs:40-42 within the run
#pragma unroll
        for (int qq = 0; qq < 3; ++qq) {
            const int srcl = WL::group_of_rank(qq) * 8 + j;
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                const R t = shfl_reg<R>(v[RPG - 1].r[r], srcl);
                if (qq < q) carry.r[r] = lane_add<T>(carry.r[r], t);
            }
        }
#pragma unroll
        for (int i = 0; i < RPG; ++i) v[i] = slice_add<T>(v[i], carry);
    } else {
        // delta.rs:28-30: out = in - prev; prev of a run's first row is the previous run's last row
        const int srcl = WL::group_of_rank(q > 0 ? q - 1 : 0) * 8 + j;
        Slice<T> first_prev;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const R t = shfl_reg<R>(v[RPG - 1].r[r], srcl);
            first_prev.r[r] = (q == 0) ? carry.r[r] : t;
        }
#pragma unroll
        for (int i = RPG - 1; i >= 1; --i) v[i] = slice_sub<T>(v[i], v[i - 1]);
        v[0] = slice_sub<T>(v[0], first_prev);
    }
    seq_rows<RPG>([&](auto ic) {
        constexpr int i = warp_visit_row<RPG>(decltype(ic)::value);
        store_slice<T>(op + warp_row_offset<T, i>(q), v[i]);
    });
}

// ---------------------------------------------------------------------------------------------------
// a14 Transpose::transpose / untranspose (src/transpose.rs:11-22), warp-block layout: the same machinery as the
// fused original-order chains without the codec in between.  One warp = one block; the ORIGINAL-order side is a
// linear 512-bytes-per-instruction copy through the warp-private swizzled tile, the TRANSPOSED-order side is the
// row-major register tile (rows of the transposed vector are exactly the unpacked rows, macros.rs:20-24).
// ---------------------------------------------------------------------------------------------------
template <class T, bool UNDO>
__global__ void __launch_bounds__(kThreads)
transpose_warp_kernel(const char* __restrict__ in, char* __restrict__ out, size_t n_blocks) {
    using WL = WarpLay<T>;
    constexpr int TB = Lay<T>::TB;
    constexpr int RPG = WL::RPG;
    const size_t blk = (size_t(blockIdx.x) * kThreads + threadIdx.x) >> 5;
    if (blk >= n_blocks) return;
    const int lane = threadIdx.x & 31;
    const int g = lane >> 3, j = lane & 7;
    const int q = WL::rank_of_group(g);
    extern __shared__ __align__(128) unsigned char orig_tile_smem[];
    unsigned char* tile = orig_tile_smem + (threadIdx.x >> 5) * (128 * TB);
    const char* ib = in + blk * (size_t(128) * TB);
    char* ob = out + blk * (size_t(128) * TB);
    Slice<T> v[RPG];
    if constexpr (!UNDO) {
        // transposed[i] = original[t(i)]: original -> tile (linear) -> register rows -> transposed rows
#pragma unroll
        for (int i = 0; i < (128 * TB) / 512; ++i) {
            const int A = i * 512 + lane * 16;
            *reinterpret_cast<uint4*>(tile + orig_tile_swizzle<T>(A)) = ldg128_stream(ib + A);
        }
        __syncwarp();
        orig_tile_gather<T, RPG>(tile, v, q, j);
        seq_rows<RPG>([&](auto ic) {
            constexpr int i = warp_visit_row<RPG>(decltype(ic)::value);
            store_slice<T>(ob + j * 16 + warp_row_offset<T, i>(q), v[i]);
        });
    } else {
        seq_rows<RPG>([&](auto ic) {
            constexpr int i = warp_visit_row<RPG>(decltype(ic)::value);
            v[i] = load_slice<T>(ib + j * 16 + warp_row_offset<T, i>(q));
        });
        orig_tile_scatter<T, RPG>(tile, v, q, j);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < (128 * TB) / 512; ++i) {
            const int A = i * 512 + lane * 16;
            stg128_stream(ob + A, *reinterpret_cast<const uint4*>(tile + orig_tile_swizzle<T>(A)));
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// pack family:  a7 BitPacking::pack (src/bitpacking.rs:65-74), a18 FoR::for_pack (src/ffor.rs:24-36).
//   in : n_blocks x (128*T bytes)        packed : n_blocks x (128*W bytes)
// ---------------------------------------------------------------------------------------------------
template <class T, int W, int OP>
__device__ __forceinline__ void pack_slice(const char* __restrict__ in, char* __restrict__ pk, Slice<T> ref) {
    constexpr int TB = Lay<T>::TB;
    using R = typename Lay<T>::R;
    if constexpr (W == 0) {
        // macros.rs:52 — the packed array is zero bytes
    } else {
        constexpr int D = (FLB_PREFETCH < TB) ? FLB_PREFETCH : TB;
        Slice<T> src[TB];
        seq_rows<D>([&](auto rc) {
            constexpr int r = decltype(rc)::value;
            src[r] = load_slice<T>(in + row_byte_offset<T>(r));
        });
        Slice<T> tmp = slice_zero<T>();
        seq_rows<TB>([&](auto rc) {
            constexpr int row = decltype(rc)::value;
            if constexpr (row + D < TB) src[row + D] = load_slice<T>(in + row_byte_offset<T>(row + D));
            Slice<T> s = src[row];
            if constexpr (OP == POP_FOR) s = slice_sub<T>(s, ref);  // ffor.rs:33
            if constexpr (W == TB) {
                store_slice<T>(pk + row * 128, s);  // macros.rs:54-59 (no mask)
            } else {
                constexpr int shift = (row * W) % TB;
                constexpr int curr = (row * W) / TB;        // macros.rs:84
                constexpr int next = ((row + 1) * W) / TB;  // macros.rs:85
                constexpr R MW = rep_mask<T>(W);
#pragma unroll
                for (int i = 0; i < Lay<T>::NR; ++i) {
                    const R v = s.r[i] & MW;  // macros.rs:73
                    if constexpr (shift == 0) tmp.r[i] = v;  // macros.rs:76-77 (row 0, or a fresh word)
                    else if constexpr (shift + W <= TB) tmp.r[i] |= v << shift;  // cannot leak: fits the lane
                    else tmp.r[i] |= lane_shl<T, shift>(v);                      // macros.rs:79
                    s.r[i] = v;
                }
                if constexpr (next > curr) {  // macros.rs:88
                    store_slice<T>(pk + curr * 128, tmp);  // macros.rs:89
                    constexpr int rem = ((row + 1) * W) % TB;  // macros.rs:90
#pragma unroll
                    for (int i = 0; i < Lay<T>::NR; ++i) tmp.r[i] = lane_shr_keep<T, W - rem, rem>(s.r[i]);  // :92
                }
            }
        });
    }
}

template <class T, int W, int OP>
__global__ void __launch_bounds__(kThreads)
pack_kernel(const char* __restrict__ in, char* __restrict__ packed, size_t n_blocks,
            const T* __restrict__ refs, T ref_scalar) {
    const size_t tid = size_t(blockIdx.x) * kThreads + threadIdx.x;
    const size_t blk = tid / kSlicesPerBlock;
    const int j = int(tid % kSlicesPerBlock);
    if (blk >= n_blocks) return;
    Slice<T> ref = slice_zero<T>();
    if constexpr (OP == POP_FOR) ref = slice_splat<T>(refs ? refs[blk] : ref_scalar);
    pack_slice<T, W, OP>(in + blk * (size_t(128) * Lay<T>::TB) + j * 16,
                         packed + blk * (size_t(128) * W) + j * 16, ref);
}

// ---------------------------------------------------------------------------------------------------
// for_pack with reference = the block's own minimum (POP_FOR_AUTO), ROW-SLICE mapping for the small element types.
// In the warp-block kernel a u8 block (1 KiB) gives a thread two rows: the statistics are a 5-step, 2-register
// butterfly per 1 KiB, and the cross-group merge of the packed words comes on top (4.2-4.8 TB/s at W < 8,
// profiles/opbench_r01.txt).  With 8 threads per block a thread holds ALL rows of its 16-byte column slice: min / max are
// a register reduction plus a 3-step butterfly inside the 8-thread group, and packing is the shuffle-free register
// chain of pack_slice (src/macros.rs:62-94).  refs_out / spans_out as in pack_warp_kernel.
// ---------------------------------------------------------------------------------------------------
template <class T, int W>
__global__ void __launch_bounds__(kThreads)
for_pack_auto_slice_kernel(const char* __restrict__ in, char* __restrict__ packed, size_t n_blocks,
                           T* __restrict__ refs_out, T* __restrict__ spans_out) {
    using R = typename Lay<T>::R;
    constexpr int TB = Lay<T>::TB;
    constexpr int NR = Lay<T>::NR;
    static_assert(sizeof(T) <= 2, "row-slice statistics kernel: u8 / u16");
    const size_t tid = size_t(blockIdx.x) * kThreads + threadIdx.x;
    const size_t blk_raw = tid / kSlicesPerBlock;
    const int j = int(tid % kSlicesPerBlock);
    const bool active = blk_raw < n_blocks;  // whole 8-thread groups are active or not; the shuffles stay inside a group
    const size_t blk = active ? blk_raw : 0;
    const char* ip = in + blk * (size_t(128) * TB) + j * 16;
    Slice<T> src[TB];
    seq_rows<TB>([&](auto rc) {
        constexpr int r = decltype(rc)::value;
        src[r] = load_slice<T>(ip + row_byte_offset<T>(r));
    });
    // statistics: both types reduce as 16x2 vectors (u8: even / odd bytes, see u8_minmax_acc)
    uint32_t lo, hi;
    if constexpr (sizeof(T) == 1) {
        uint32_t mn_e = 0x00FF00FFu, mn_o = 0x00FF00FFu, mx_e = 0, mx_o = 0;
#pragma unroll
        for (int r = 0; r < TB; ++r)
#pragma unroll
            for (int i = 0; i < NR; ++i) u8_minmax_acc(src[r].r[i], mn_e, mn_o, mx_e, mx_o);
        lo = __vminu2(mn_e, mn_o); hi = __vmaxu2(mx_e, mx_o);
    } else {
        lo = src[0].r[0]; hi = lo;
#pragma unroll
        for (int r = 0; r < TB; ++r)
#pragma unroll
            for (int i = 0; i < NR; ++i) { lo = __vminu2(lo, src[r].r[i]); hi = __vmaxu2(hi, src[r].r[i]); }
    }
#pragma unroll
    for (int d = 4; d >= 1; d >>= 1) {
        lo = __vminu2(lo, __shfl_xor_sync(0xffffffffu, lo, d));
        hi = __vmaxu2(hi, __shfl_xor_sync(0xffffffffu, hi, d));
    }
    const T mn = T(min(lo & 0xFFFFu, lo >> 16)), mx = T(max(hi & 0xFFFFu, hi >> 16));
    if (active && j == 0) {
        refs_out[blk] = mn;
        if (spans_out != nullptr) spans_out[blk] = T(mx - mn);
    }
    if constexpr (W == 0) return;  // macros.rs:52: nothing to store
    if (!active) return;
    const Slice<T> ref = slice_splat<T>(mn);
    char* pk = packed + blk * (size_t(128) * W) + j * 16;
    Slice<T> tmp = slice_zero<T>();
    seq_rows<TB>([&](auto rc) {
        constexpr int row = decltype(rc)::value;
        Slice<T> s = slice_sub<T>(src[row], ref);  // ffor.rs:33
        if constexpr (W == TB) {
            store_slice<T>(pk + row * 128, s);  // macros.rs:54-59
        } else {
            constexpr int shift = (row * W) % TB;
            constexpr int curr = (row * W) / TB;        // macros.rs:84
            constexpr int next = ((row + 1) * W) / TB;  // macros.rs:85
            constexpr R MW = rep_mask<T>(W);
#pragma unroll
            for (int i = 0; i < NR; ++i) {
                const R v = s.r[i] & MW;  // macros.rs:73
                if constexpr (shift == 0) tmp.r[i] = v;
                else if constexpr (shift + W <= TB) tmp.r[i] |= v << shift;
                else tmp.r[i] |= lane_shl<T, shift>(v);  // macros.rs:79
                s.r[i] = v;
            }
            if constexpr (next > curr) {  // macros.rs:88-92
                store_slice<T>(pk + curr * 128, tmp);
                constexpr int rem = ((row + 1) * W) % TB;
#pragma unroll
                for (int i = 0; i < NR; ++i) tmp.r[i] = lane_shr_keep<T, W - rem, rem>(s.r[i]);
            }
        }
    });
}

// ---------------------------------------------------------------------------------------------------
// a15 Delta::delta (src/delta.rs:24-33) and a16 Delta::undelta (src/delta.rs:36-45): per lane, along
// rows in iterate! order (src/macros.rs:12-31).  in/out: n_blocks x (128*T bytes); base: n_blocks x 128 B.
// ---------------------------------------------------------------------------------------------------
template <class T, bool UNDO>
__global__ void __launch_bounds__(kThreads)
delta_kernel(const char* __restrict__ in, const char* __restrict__ base, char* __restrict__ out, size_t n_blocks) {
    constexpr int TB = Lay<T>::TB;
    const size_t tid = size_t(blockIdx.x) * kThreads + threadIdx.x;
    const size_t blk = tid / kSlicesPerBlock;
    const int j = int(tid % kSlicesPerBlock);
    if (blk >= n_blocks) return;
    const char* ip = in + blk * (size_t(128) * TB) + j * 16;
    char* op = out + blk * (size_t(128) * TB) + j * 16;
    Slice<T> prev = load_slice<T>(base + blk * 128 + j * 16);  // delta.rs:26 / :38
    constexpr int D = (FLB_PREFETCH < TB) ? FLB_PREFETCH : TB;
    Slice<T> src[TB];
    seq_rows<D>([&](auto rc) {
        constexpr int r = decltype(rc)::value;
        src[r] = load_slice<T>(ip + row_byte_offset<T>(r));
    });
    seq_rows<TB>([&](auto rc) {
        constexpr int row = decltype(rc)::value;
        if constexpr (row + D < TB) src[row + D] = load_slice<T>(ip + row_byte_offset<T>(row + D));
        if constexpr (UNDO) {
            prev = slice_add<T>(src[row], prev);  // delta.rs:40-42
            store_slice<T>(op + row_byte_offset<T>(row), prev);
        } else {
            store_slice<T>(op + row_byte_offset<T>(row), slice_sub<T>(src[row], prev));  // delta.rs:28-30
            prev = src[row];
        }
    });
}

// ---------------------------------------------------------------------------------------------------
// u8 fused original-order chains, ROW-SLICE layout (SURVEY.md §8f rank 1, u8 only).
// A u8 block is 1 KiB: in the warp-block layout a thread holds only 2 rows, the 4-group shuffle scan and the
// 16 STS.U16 + drain of the shared tile dominate (4.6-5.5 TB/s at W < 8).  With 8 threads per block a thread
// holds ALL 8 rows of its 16 lanes, so (i) the delta chain is a register chain (no shuffles) and (ii) for every
// lane the 8 rows are the 8 CONSECUTIVE originals start(l) .. start(l)+7, start(l) = 64*(l%16) + 8*FL_ORDER[l/16]
// (src/transpose.rs:29-36 composed with src/macros.rs:20-24): one 8-byte global access per lane, assembled by
// 4x4 byte transposes (PRMT).  Thread j owns lanes 16j+k: address 64*k + 8*FL_ORDER[j]; for a fixed k the 8 threads
// of a block cover 64 contiguous bytes (two full sectors), a warp instruction four such segments.  No shared memory.
// ---------------------------------------------------------------------------------------------------
// o[b] = (a0.byte b, a1.byte b, a2.byte b, a3.byte b): 4x4 byte transpose, an involution
__device__ __forceinline__ void byte_transpose4(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t (&o)[4]) {
    const uint32_t t0 = __byte_perm(a0, a1, 0x5140u), t1 = __byte_perm(a2, a3, 0x5140u);
    const uint32_t t2 = __byte_perm(a0, a1, 0x7362u), t3 = __byte_perm(a2, a3, 0x7362u);
    o[0] = __byte_perm(t0, t1, 0x5410u);
    o[1] = __byte_perm(t0, t1, 0x7632u);
    o[2] = __byte_perm(t2, t3, 0x5410u);
    o[3] = __byte_perm(t2, t3, 0x7632u);
}
__device__ __forceinline__ void stg64_stream(void* p, uint32_t x, uint32_t y) {
    asm volatile("st.global.cs.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ uint2 ldg64_stream(const void* p) {
    uint2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}

// untranspose(undelta_pack::<W>(packed, base)) for u8  (src/delta.rs:48-63 then src/transpose.rs:18-22)
template <int W>
__global__ void __launch_bounds__(kThreads)
undelta_orig_u8_slice_kernel(const char* __restrict__ packed, char* __restrict__ out, size_t n_blocks,
                             const char* __restrict__ base) {
    using T = uint8_t;
    const size_t tid = size_t(blockIdx.x) * kThreads + threadIdx.x;
    const size_t blk = tid / kSlicesPerBlock;
    const int j = int(tid % kSlicesPerBlock);
    if (blk >= n_blocks) return;
    Slice<T> prev = load_slice<T>(base + blk * 128 + j * 16);  // delta.rs:50
    const char* pk = packed + blk * (size_t(128) * W) + j * 16;
    Slice<T> w[W > 0 ? W : 1];
    seq_rows<W>([&](auto kc) {
        constexpr int k = decltype(kc)::value;
        w[k] = load_slice<T>(pk + k * 128);
    });
    Slice<T> v[8];
    seq_rows<8>([&](auto rc) {
        constexpr int row = decltype(rc)::value;
        Slice<T> x;
        if constexpr (W == 0) x = slice_zero<T>();       // macros.rs:118-125
        else if constexpr (W == 8) x = w[row];            // macros.rs:126-132
        else {
            constexpr int curr = (row * W) / 8;
            constexpr int nxt = (curr + 1 < W) ? curr + 1 : curr;
            x = extract_row<T, W, row>(w[curr], w[nxt]);
        }
        prev = slice_add<T>(prev, x);  // delta.rs:58-60
        v[row] = prev;
    });
    char* ob = out + blk * 1024 + 8 * fl_order_rt(j);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        uint32_t lo[4], hi[4];
        byte_transpose4(v[0].r[r], v[1].r[r], v[2].r[r], v[3].r[r], lo);  // lane 4r+b: rows 0..3
        byte_transpose4(v[4].r[r], v[5].r[r], v[6].r[r], v[7].r[r], hi);  // lane 4r+b: rows 4..7
#pragma unroll
        for (int b = 0; b < 4; ++b) stg64_stream(ob + 64 * (4 * r + b), lo[b], hi[b]);
    }
}

// pack::<W>(delta(transpose(in), base)) for u8  (src/transpose.rs:11-15, src/delta.rs:24-33, src/macros.rs:35-97)
template <int W>
__global__ void __launch_bounds__(kThreads)
orig_delta_pack_u8_slice_kernel(const char* __restrict__ in, char* __restrict__ packed, size_t n_blocks,
                                const char* __restrict__ base) {
    using T = uint8_t;
    using R = uint32_t;
    if constexpr (W == 0) return;  // macros.rs:52
    const size_t tid = size_t(blockIdx.x) * kThreads + threadIdx.x;
    const size_t blk = tid / kSlicesPerBlock;
    const int j = int(tid % kSlicesPerBlock);
    if (blk >= n_blocks) return;
    Slice<T> prev = load_slice<T>(base + blk * 128 + j * 16);  // delta.rs:26
    const char* ib = in + blk * 1024 + 8 * fl_order_rt(j);
    uint2 c[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) c[k] = ldg64_stream(ib + 64 * k);
    Slice<T> v[8];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        uint32_t lo[4], hi[4];
        byte_transpose4(c[4 * r].x, c[4 * r + 1].x, c[4 * r + 2].x, c[4 * r + 3].x, lo);  // rows 0..3 of lanes 4r..4r+3
        byte_transpose4(c[4 * r].y, c[4 * r + 1].y, c[4 * r + 2].y, c[4 * r + 3].y, hi);  // rows 4..7
#pragma unroll
        for (int b = 0; b < 4; ++b) { v[b].r[r] = lo[b]; v[4 + b].r[r] = hi[b]; }
    }
    char* pk = packed + blk * (size_t(128) * W) + j * 16;
    Slice<T> tmp = slice_zero<T>();
    seq_rows<8>([&](auto rc) {
        constexpr int row = decltype(rc)::value;
        Slice<T> s = slice_sub<T>(v[row], prev);  // delta.rs:28-30
        prev = v[row];
        if constexpr (W == 8) {
            store_slice<T>(pk + row * 128, s);  // macros.rs:54-59
        } else {
            constexpr int shift = (row * W) % 8;
            constexpr int curr = (row * W) / 8;
            constexpr int next = ((row + 1) * W) / 8;
            constexpr R MW = rep_mask<T>(W);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const R x = s.r[i] & MW;  // macros.rs:73
                if constexpr (shift == 0) tmp.r[i] = x;
                else if constexpr (shift + W <= 8) tmp.r[i] |= x << shift;
                else tmp.r[i] |= lane_shl<T, shift>(x);  // macros.rs:79
                s.r[i] = x;
            }
            if constexpr (next > curr) {  // macros.rs:88-92
                store_slice<T>(pk + curr * 128, tmp);
                constexpr int rem = ((row + 1) * W) % 8;
#pragma unroll
                for (int i = 0; i < 4; ++i) tmp.r[i] = lane_shr_keep<T, W - rem, rem>(s.r[i]);
            }
        }
    });
}

// ---------------------------------------------------------------------------------------------------
// u16 fused ENCODE chain, ROW-SLICE layout, W = 1 (round 2).  Same idea as the u8 kernels above: thread j of an 8-thread
// group owns lanes 8j..8j+7 for ALL 16 rows, so the delta chain is a register chain and, for every lane l, the 16 rows
// are the 16 CONSECUTIVE originals start(l) .. start(l)+15 (src/transpose.rs:29-36 composed with src/macros.rs:20-24) =
// 32 contiguous bytes at 2*start(l) = 128*(8*(j&1) + k) + 16*FL_ORDER[j>>1] (l = 8j + k): two 16-byte loads per lane,
// re-assembled into row-major SWAR registers by PRMT, one SWAR register (two lanes) at a time.  No shared memory, no
// shuffles.  Measured (profiles/opbench_u16_orig_slice_r02.txt): transpose_delta_pack u16 W=1 6.06 -> 7.09 TB/s, but
// slower than the warp-block kernel from W = 2 on (the packed stores of 8 threads are 16-byte pieces of different
// rows), so it is used at W = 1 only.  The mirror-image DECODE kernel (two 16-byte stores per lane at a 32-byte stride
// across the group) reached only 2.4 TB/s — half-sector writes — and was dropped; the decode keeps the warp-block
// kernel with the extra-occupancy instantiation (FLB_ORIG_OCC).
// ---------------------------------------------------------------------------------------------------
// pack::<W>(delta(transpose(in), base)) for u16 at small W  (src/transpose.rs:11-15, src/delta.rs:24-33, src/macros.rs:35-97)
template <int W>
__global__ void __launch_bounds__(kThreads)
orig_delta_pack_u16_slice_kernel(const char* __restrict__ in, char* __restrict__ packed, size_t n_blocks,
                                 const char* __restrict__ base) {
    using T = uint16_t;
    using R = uint32_t;
    constexpr int TB = 16;
    if constexpr (W == 0) return;  // macros.rs:52
    const size_t tid = size_t(blockIdx.x) * kThreads + threadIdx.x;
    const size_t blk = tid / kSlicesPerBlock;
    const int j = int(tid % kSlicesPerBlock);
    if (blk >= n_blocks) return;
    const Slice<T> b0 = load_slice<T>(base + blk * 128 + j * 16);  // delta.rs:26
    const char* ib = in + blk * 2048 + 1024 * (j & 1) + 16 * fl_order_rt(j >> 1);
    Slice<T> outw[W > 0 ? W : 1];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const uint4 a0 = ldg128_stream(ib + 128 * (2 * r)), a1 = ldg128_stream(ib + 128 * (2 * r) + 16);          // lane 2r
        const uint4 c0 = ldg128_stream(ib + 128 * (2 * r + 1)), c1 = ldg128_stream(ib + 128 * (2 * r + 1) + 16);  // lane 2r+1
        const uint32_t la[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const uint32_t lc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
        R v[TB];
#pragma unroll
        for (int i = 0; i < 8; ++i) {  // word i of a lane = rows 2i, 2i+1
            v[2 * i] = __byte_perm(la[i], lc[i], 0x5410u);
            v[2 * i + 1] = __byte_perm(la[i], lc[i], 0x7632u);
        }
        R prev = b0.r[r];
        R tmp = 0;
        seq_rows<TB>([&](auto rc) {
            constexpr int row = decltype(rc)::value;
            R s = lane_sub<T>(v[row], prev);  // delta.rs:28-30
            prev = v[row];
            if constexpr (W == TB) {
                outw[row].r[r] = s;  // macros.rs:54-59
            } else {
                constexpr int shift = (row * W) % TB;
                constexpr int curr = (row * W) / TB;
                constexpr int next = ((row + 1) * W) / TB;
                constexpr R MW = rep_mask<T>(W);
                const R x = s & MW;  // macros.rs:73
                if constexpr (shift == 0) tmp = x;
                else if constexpr (shift + W <= TB) tmp |= x << shift;
                else tmp |= lane_shl<T, shift>(x);  // macros.rs:79
                if constexpr (next > curr) {  // macros.rs:88-92
                    outw[curr].r[r] = tmp;
                    constexpr int rem = ((row + 1) * W) % TB;
                    tmp = lane_shr_keep<T, W - rem, rem>(x);
                }
            }
        });
    }
    char* pk = packed + blk * (size_t(128) * W) + j * 16;
    seq_rows<W>([&](auto kc) {
        constexpr int k = decltype(kc)::value;
        store_slice<T>(pk + k * 128, outw[k]);
    });
}

}  // namespace flb
