// fl_codec_inst.cu — instantiates the width-templated kernels for ONE element type and ONE part
// (compile with -DFLB_TBITS=8|16|32|64 -DFLB_PART=0|1|2|3: 0 = unpack family, 1 = pack family, 2 = delta and
// transpose, 3 = fused scan kernels)
// and builds the runtime-width dispatch tables: the `match width { W => ::<W>() }` of
// src/bitpacking.rs:82-95, :115-128 in the reference.
#include <cstdlib>
#include <cstring>

#include "fl_internal.h"
#include "fl_kernels.cuh"

#ifndef FLB_U8_FILTER_DEFAULT_SLICE
#define FLB_U8_FILTER_DEFAULT_SLICE 1  // measured: 550-733 us vs 706-872 (profiles/opbench_scan_r01.txt)
#endif
#ifndef FLB_U8_ORIG_DEFAULT_SLICE
#define FLB_U8_ORIG_DEFAULT_SLICE 1  // measured: 6.5-7.1 TB/s vs 4.6-6.4 (profiles/opbench_u8orig_r01.txt)
#endif
#ifndef FLB_U16_ORIG_SLICE_W
#define FLB_U16_ORIG_SLICE_W 1  // widths up to this use the row-slice kernel for the u16 fused ENCODE chain (measured: faster at W = 1 only)
#endif
#ifndef FLB_U16_ORIG_DEFAULT_SLICE
#define FLB_U16_ORIG_DEFAULT_SLICE 1
#endif
#ifndef FLB_U8_DELTA_W8_DEFAULT_SLICE
#define FLB_U8_DELTA_W8_DEFAULT_SLICE 1  // measured: 1375 vs 1408 us per 2^22 blocks (profiles/opbench_u8_delta_w8_r02.txt)
#endif
#ifndef FLB_ORIG_OCC_W
#define FLB_ORIG_OCC_W 4  // widths up to this have the extra-occupancy instantiation of the fused original-order decode
#endif
#ifndef FLB_U16_DELTA_OCC_DEFAULT
#define FLB_U16_DELTA_OCC_DEFAULT 1  // measured: W=1 787 -> 734 us per 2^21 blocks (profiles/opbench_u16_delta_occ_r02.txt)
#endif
#ifndef FLB_ORIG_OCC_DEFAULT
#define FLB_ORIG_OCC_DEFAULT 2
#endif
#ifndef FLB_U16_FILTER_DEFAULT_SLICE
#define FLB_U16_FILTER_DEFAULT_SLICE 1
#endif
#ifndef FLB_AUTO_SLICE_DEFAULT
#define FLB_AUTO_SLICE_DEFAULT 1
#endif
#if FLB_PART == 3
#include "fl_scan.cuh"
#endif

namespace flb {

#if FLB_TBITS == 8
using elem_t = uint8_t;
#elif FLB_TBITS == 16
using elem_t = uint16_t;
#elif FLB_TBITS == 32
using elem_t = uint32_t;
#elif FLB_TBITS == 64
using elem_t = uint64_t;
#else
#error "FLB_TBITS must be 8/16/32/64"
#endif

using launch_fn = cudaError_t (*)(const LaunchArgs&);

[[maybe_unused]] static inline unsigned grid_for(size_t n_blocks) {
    return unsigned((n_blocks * kSlicesPerBlock + kThreads - 1) / kThreads);
}

// u8 fused original-order chains: FLB_U8_ORIG=warp|slice selects the warp-block (shared tile) or the row-slice
// (8-byte global accesses, no shared memory) kernel for A/B measurement; the default is the measured best.
[[maybe_unused]] static inline bool u8_orig_slice() {
    static const bool v = [] {
        const char* e = std::getenv("FLB_U8_ORIG");
        if (e && std::strcmp(e, "warp") == 0) return false;
        if (e && std::strcmp(e, "slice") == 0) return true;
        return FLB_U8_ORIG_DEFAULT_SLICE != 0;
    }();
    return v;
}

// u16 fused encode chain at W = 1: FLB_U16_ORIG=warp|slice (A/B; default = measured best)
[[maybe_unused]] static inline bool u16_orig_slice() {
    static const bool v = [] {
        const char* e = std::getenv("FLB_U16_ORIG");
        if (e && std::strcmp(e, "warp") == 0) return false;
        if (e && std::strcmp(e, "slice") == 0) return true;
        return FLB_U16_ORIG_DEFAULT_SLICE != 0;
    }();
    return v;
}

#if FLB_PART == 0
template <class T, int W, int OP>
static cudaError_t do_unpack(const LaunchArgs& a) {
    if constexpr (sizeof(T) == 1 && OP == UOP_DELTA && W == 8) {
        // u8 at W = 8 (verbatim rows): FLB_U8_DELTA_W8=slice|warp, A/B; default = measured best
        static const bool slice = [] {
            const char* e = std::getenv("FLB_U8_DELTA_W8");
            if (e && std::strcmp(e, "slice") == 0) return true;
            if (e && std::strcmp(e, "warp") == 0) return false;
            return FLB_U8_DELTA_W8_DEFAULT_SLICE != 0;
        }();
        if (slice) {
            unpack_kernel<T, W, OP><<<grid_for(a.n_blocks), kThreads, 0, a.stream>>>(
                static_cast<const char*>(a.in), static_cast<char*>(a.out), a.n_blocks, static_cast<const T*>(a.refs),
                T(a.ref_scalar), static_cast<const char*>(a.base));
            return cudaGetLastError();
        }
    }
    if constexpr (sizeof(T) == 1 && OP == UOP_DELTA && W < 8) {
        // u8 fused delta at W < 8: a 1 KiB block gives a warp too little work to amortise the 4-group shuffle
        // scan (ALU-bound, 5.1-6.1 TB/s); the row-slice kernel keeps the whole 8-row chain in one thread
        // (6.3-6.5 TB/s measured, profiles/kbench_r01_u8_u16_delta.txt).
        unpack_kernel<T, W, OP><<<grid_for(a.n_blocks), kThreads, 0, a.stream>>>(
            static_cast<const char*>(a.in), static_cast<char*>(a.out), a.n_blocks, static_cast<const T*>(a.refs),
            T(a.ref_scalar), static_cast<const char*>(a.base));
        return cudaGetLastError();
    }
    if constexpr (sizeof(T) == 1 && OP == UOP_DELTA_ORIG) {
        if (u8_orig_slice()) {
            undelta_orig_u8_slice_kernel<W><<<grid_for(a.n_blocks), kThreads, 0, a.stream>>>(
                static_cast<const char*>(a.in), static_cast<char*>(a.out), a.n_blocks, static_cast<const char*>(a.base));
            return cudaGetLastError();
        }
    }
    // warp-block layout: one warp per 1024-value block (see fl_kernels.cuh)
    const unsigned grid = unsigned((a.n_blocks * 32 + kThreads - 1) / kThreads);
    // TMA bulk load of the packed block (one cp.async.bulk per warp): measured +0.5..8% on u32 (largest at W >= 25,
    // profiles/kbench_r01_tma_u32.txt).  u8 blocks (<= 1 KiB packed) are too small to amortise the mbarrier round
    // trip (measured slower): direct loads.  The original-order variant also owns a dynamic shared tile; there the
    // extra static buffer costs occupancy and the bulk copy only wins where measured: u32 at W >= 8, u64 at W % 4 != 0.
    constexpr bool kTma = (OP != UOP_DELTA_ORIG) ? (sizeof(T) >= 2)
                                                 : ((sizeof(T) == 4 && W >= 8) || (sizeof(T) == 8 && W % 4 != 0));
    size_t smem = 0;
    if constexpr (OP == UOP_DELTA_ORIG) {  // one block staged per warp
        smem = size_t(kThreads / 32) * 128 * Lay<T>::TB;
        static SmemOptIn opt_in;
        if (const cudaError_t attr = opt_in.ensure(unpack_warp_kernel<T, W, OP, kTma>, smem); attr != cudaSuccess) return attr;
    }
    if constexpr (OP == UOP_DELTA_ORIG && sizeof(T) >= 2 && W <= FLB_ORIG_OCC_W) {
        // Small W: the kernel is latency-bound (ncu u32 W=1: 72 registers -> 3 CTAs per SM, 31 % of the warps resident,
        // DRAM 71 %, profiles/ncu_r02_kernels.md).  A second instantiation whose launch bound holds the compiler to more
        // resident CTAs (u32: 4 = 64 registers; u64: 3 = the shared-memory limit; u16: 6 = 40 registers) is launched
        // instead.  Measured (profiles/opbench_orig_occ_r02.txt): u32 W=1 811 -> 702 us, u64 W=1 779 -> 710 us,
        // u16 W=1 827 -> 760 us per 4 GiB of output.
        // FLB_ORIG_OCC=0|1|2: off / u32 + u64 / also u16 (A/B; default = measured best).
        static const int occ = [] {
            const char* e = std::getenv("FLB_ORIG_OCC");
            return (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : FLB_ORIG_OCC_DEFAULT;
        }();
        if ((occ >= 1 && sizeof(T) >= 4) || occ == 2) {
            constexpr int kMinB = sizeof(T) == 2 ? 6 : (sizeof(T) == 4 ? 4 : 3);
            static SmemOptIn opt_in2;
            if (const cudaError_t attr = opt_in2.ensure(unpack_warp_kernel<T, W, OP, kTma, false, kMinB>, smem); attr != cudaSuccess) return attr;
            unpack_warp_kernel<T, W, OP, kTma, false, kMinB><<<grid, kThreads, smem, a.stream>>>(
                static_cast<const char*>(a.in), static_cast<char*>(a.out), a.n_blocks, static_cast<const T*>(a.refs),
                T(a.ref_scalar), static_cast<const char*>(a.base));
            return cudaGetLastError();
        }
    }
    if constexpr (OP == UOP_DELTA && sizeof(T) == 2 && W <= FLB_ORIG_OCC_W) {
        // u16 fused delta at small W (ncu W=1: 48 registers, 51 % of the warps resident, issue 57 %, DRAM 73 %): the same
        // occupancy lever, 6 CTAs per SM = 40 registers.  FLB_U16_DELTA_OCC=0|1 (A/B; default = measured best).
        static const bool occ = [] {
            const char* e = std::getenv("FLB_U16_DELTA_OCC");
            return e ? e[0] == '1' : FLB_U16_DELTA_OCC_DEFAULT != 0;
        }();
        if (occ) {
            unpack_warp_kernel<T, W, OP, kTma, false, 6><<<grid, kThreads, smem, a.stream>>>(
                static_cast<const char*>(a.in), static_cast<char*>(a.out), a.n_blocks, static_cast<const T*>(a.refs),
                T(a.ref_scalar), static_cast<const char*>(a.base));
            return cudaGetLastError();
        }
    }
    unpack_warp_kernel<T, W, OP, kTma><<<grid, kThreads, smem, a.stream>>>(
        static_cast<const char*>(a.in), static_cast<char*>(a.out), a.n_blocks, static_cast<const T*>(a.refs),
        T(a.ref_scalar), static_cast<const char*>(a.base));
    return cudaGetLastError();
}
template <class T, int OP, int... W>
static constexpr std::array<launch_fn, sizeof...(W)> unpack_table(std::integer_sequence<int, W...>) {
    return {{&do_unpack<T, W, OP>...}};
}
// cwida row order (linear rows; bit-packing and FoR only): same kernel, LINEAR = true, same load-path choice
template <class T, int W, int OP>
static cudaError_t do_unpack_linear(const LaunchArgs& a) {
    const unsigned grid = unsigned((a.n_blocks * 32 + kThreads - 1) / kThreads);
    unpack_warp_kernel<T, W, OP, (sizeof(T) >= 2), true><<<grid, kThreads, 0, a.stream>>>(
        static_cast<const char*>(a.in), static_cast<char*>(a.out), a.n_blocks, static_cast<const T*>(a.refs),
        T(a.ref_scalar), nullptr);
    return cudaGetLastError();
}
template <class T, int OP, int... W>
static constexpr std::array<launch_fn, sizeof...(W)> unpack_linear_table(std::integer_sequence<int, W...>) {
    return {{&do_unpack_linear<T, W, OP>...}};
}
// fused undelta_pack + untranspose (all four types)
template <class T>
static cudaError_t unpack_delta_orig(const LaunchArgs& a) {
    static constexpr auto tab = unpack_table<T, UOP_DELTA_ORIG>(std::make_integer_sequence<int, Lay<T>::TB + 1>{});
    return tab[a.width](a);
}
template <>
cudaError_t launch_unpack<elem_t>(int op, const LaunchArgs& a) {
    using seq = std::make_integer_sequence<int, Lay<elem_t>::TB + 1>;
    static constexpr auto plain = unpack_table<elem_t, UOP_PLAIN>(seq{});
    static constexpr auto ffor = unpack_table<elem_t, UOP_FOR>(seq{});
    static constexpr auto delta = unpack_table<elem_t, UOP_DELTA>(seq{});
    if (op == kUnpackDeltaOrig) return unpack_delta_orig<elem_t>(a);
    if (op == kUnpackPlainLinear || op == kUnpackForLinear) {
        static constexpr auto lplain = unpack_linear_table<elem_t, UOP_PLAIN>(seq{});
        static constexpr auto lfor = unpack_linear_table<elem_t, UOP_FOR>(seq{});
        return (op == kUnpackPlainLinear ? lplain : lfor)[a.width](a);
    }
    switch (op) {
        case kUnpackPlain: return plain[a.width](a);
        case kUnpackFor: return ffor[a.width](a);
        case kUnpackDelta: return delta[a.width](a);
        default: return cudaErrorNotSupported;
    }
}
#elif FLB_PART == 1
template <class T, int W, int OP>
static cudaError_t do_pack(const LaunchArgs& a) {
    if constexpr (sizeof(T) == 2 && OP == POP_ORIG_DELTA && W <= FLB_U16_ORIG_SLICE_W) {
        if (u16_orig_slice()) {
            orig_delta_pack_u16_slice_kernel<W><<<grid_for(a.n_blocks), kThreads, 0, a.stream>>>(
                static_cast<const char*>(a.in), static_cast<char*>(a.out), a.n_blocks, static_cast<const char*>(a.base));
            return cudaGetLastError();
        }
    }
    if constexpr (sizeof(T) == 1 && OP == POP_ORIG_DELTA) {
        if (u8_orig_slice()) {
            orig_delta_pack_u8_slice_kernel<W><<<grid_for(a.n_blocks), kThreads, 0, a.stream>>>(
                static_cast<const char*>(a.in), static_cast<char*>(a.out), a.n_blocks, static_cast<const char*>(a.base));
            return cudaGetLastError();
        }
    }
    if constexpr (sizeof(T) == 1 && (OP == POP_PLAIN || OP == POP_FOR) && W > 0 && W < 8) {
        // u8 at small W: FLB_U8_PACK=slice|warp selects the row-slice kernel (8 threads per block, no cross-group merge
        // shuffles) or the warp-block kernel below, for A/B measurement; the default is the measured best per width.
        static const int mode = [] {
            const char* e = std::getenv("FLB_U8_PACK");
            if (e && std::strcmp(e, "slice") == 0) return 1;
            if (e && std::strcmp(e, "warp") == 0) return 0;
            return -1;
        }();
        // measured (profiles/opbench_u8pack_r01.txt): for_pack is faster row-sliced at every W < 8 (5.8-6.5 -> 6.6-6.8 TB/s),
        // plain pack only at W = 1 (6.2 -> 6.5)
        if (mode == 1 || (mode < 0 && (OP == POP_FOR || W < 2))) {
            pack_kernel<T, W, OP><<<grid_for(a.n_blocks), kThreads, 0, a.stream>>>(
                static_cast<const char*>(a.in), static_cast<char*>(a.out), a.n_blocks, static_cast<const T*>(a.refs), T(a.ref_scalar));
            return cudaGetLastError();
        }
    }
    if constexpr (sizeof(T) <= 2 && OP == POP_FOR_AUTO) {
        // fused statistics + FoR for u8 / u16: FLB_AUTO_SLICE=0|1|2 selects the warp-block kernel (0), the row-slice kernel
        // where it measured faster (1: u8 below W = 8, u16 at W <= 2) or for u8 and u16 at every width (2); A/B
        static const int mode = [] {
            const char* e = std::getenv("FLB_AUTO_SLICE");
            return (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : FLB_AUTO_SLICE_DEFAULT;
        }();
        // measured (profiles/opbench_auto_slice_r02.txt): u8 W < 8: 4.15-5.11 -> 6.56-6.78 TB/s; u16 W = 1: 5.76 -> 6.88,
        // W = 4: 7.19 -> 6.84 (worse), W >= 9: equal within 2 %
        if ((mode == 1 && ((sizeof(T) == 1 && W < 8) || (sizeof(T) == 2 && W <= 2))) || mode == 2) {
            for_pack_auto_slice_kernel<T, W><<<grid_for(a.n_blocks), kThreads, 0, a.stream>>>(
                static_cast<const char*>(a.in), static_cast<char*>(a.out), a.n_blocks, static_cast<T*>(a.refs_out),
                static_cast<T*>(a.spans_out));
            return cudaGetLastError();
        }
    }
    const unsigned grid = unsigned((a.n_blocks * 32 + kThreads - 1) / kThreads);  // warp-block layout
    // Every variant stages one block per warp in dynamic shared memory: the original-order op as its swizzled tile,
    // the plain / FoR ops as the landing buffer of the TMA bulk load (+3..7% measured, profiles/kbench_r01_tma_pack_u32.txt).
    // (u8: a 1 KiB block does not amortise the mbarrier round trip — measured slower — so it keeps direct loads.)
    constexpr bool kTma = (OP != POP_ORIG_DELTA) && sizeof(T) >= 2;
    const size_t smem = (kTma || OP == POP_ORIG_DELTA) ? size_t(kThreads / 32) * 128 * Lay<T>::TB + (kTma ? (kThreads / 32) * 8 : 0) : 0;
    static SmemOptIn opt_in;
    if (const cudaError_t attr = opt_in.ensure(pack_warp_kernel<T, W, OP, kTma>, smem); attr != cudaSuccess) return attr;
    pack_warp_kernel<T, W, OP, kTma><<<grid, kThreads, smem, a.stream>>>(
        static_cast<const char*>(a.in), static_cast<char*>(a.out), a.n_blocks, static_cast<const T*>(a.refs),
        T(a.ref_scalar), static_cast<const char*>(a.base), static_cast<T*>(a.refs_out), static_cast<T*>(a.spans_out));
    return cudaGetLastError();
}
template <class T, int OP, int... W>
static constexpr std::array<launch_fn, sizeof...(W)> pack_table(std::integer_sequence<int, W...>) {
    return {{&do_pack<T, W, OP>...}};
}
// cwida row order (linear rows)
template <class T, int W, int OP>
static cudaError_t do_pack_linear(const LaunchArgs& a) {
    const unsigned grid = unsigned((a.n_blocks * 32 + kThreads - 1) / kThreads);
    constexpr bool kTma = sizeof(T) >= 2;
    const size_t smem = kTma ? size_t(kThreads / 32) * (128 * Lay<T>::TB + 8) : 0;
    static SmemOptIn opt_in;
    if (const cudaError_t attr = opt_in.ensure(pack_warp_kernel<T, W, OP, kTma, true>, smem); attr != cudaSuccess) return attr;
    pack_warp_kernel<T, W, OP, kTma, true><<<grid, kThreads, smem, a.stream>>>(
        static_cast<const char*>(a.in), static_cast<char*>(a.out), a.n_blocks, static_cast<const T*>(a.refs),
        T(a.ref_scalar), nullptr, nullptr, nullptr);
    return cudaGetLastError();
}
template <class T, int OP, int... W>
static constexpr std::array<launch_fn, sizeof...(W)> pack_linear_table(std::integer_sequence<int, W...>) {
    return {{&do_pack_linear<T, W, OP>...}};
}
// fused transpose + delta + pack (all four types)
template <class T>
static cudaError_t pack_orig_delta(const LaunchArgs& a) {
    static constexpr auto tab = pack_table<T, POP_ORIG_DELTA>(std::make_integer_sequence<int, Lay<T>::TB + 1>{});
    return tab[a.width](a);
}
// for_pack with the block minimum as reference, statistics fused into the pack pass
template <class T>
static cudaError_t pack_for_auto(const LaunchArgs& a) {
    static constexpr auto tab = pack_table<T, POP_FOR_AUTO>(std::make_integer_sequence<int, Lay<T>::TB + 1>{});
    return tab[a.width](a);
}
template <>
cudaError_t launch_pack<elem_t>(int op, const LaunchArgs& a) {
    using seq = std::make_integer_sequence<int, Lay<elem_t>::TB + 1>;
    static constexpr auto plain = pack_table<elem_t, POP_PLAIN>(seq{});
    static constexpr auto ffor = pack_table<elem_t, POP_FOR>(seq{});
    if (op == kPackOrigDelta) return pack_orig_delta<elem_t>(a);
    if (op == kPackForAuto) return pack_for_auto<elem_t>(a);
    if (op == kPackPlainLinear || op == kPackForLinear) {
        static constexpr auto lplain = pack_linear_table<elem_t, POP_PLAIN>(seq{});
        static constexpr auto lfor = pack_linear_table<elem_t, POP_FOR>(seq{});
        return (op == kPackPlainLinear ? lplain : lfor)[a.width](a);
    }
    if (op == kPackPlain) return plain[a.width](a);
    if (op == kPackFor) return ffor[a.width](a);
    return cudaErrorNotSupported;
}
#elif FLB_PART == 3
// fused decode + predicate kernels (fl_scan.cuh)
template <class T, int W>
static cudaError_t do_filter(const LaunchArgs& a) {
    if constexpr (sizeof(T) == 1) {
        // u8: FLB_U8_FILTER=warp|slice selects the warp-block kernel below or the row-slice kernel (A/B; default = measured best)
        static const bool slice = [] {
            const char* e = std::getenv("FLB_U8_FILTER");
            if (e && std::strcmp(e, "warp") == 0) return false;
            if (e && std::strcmp(e, "slice") == 0) return true;
            return FLB_U8_FILTER_DEFAULT_SLICE != 0;
        }();
        if (slice) {
            filter_u8_slice_kernel<W><<<grid_for(a.n_blocks), kThreads, 0, a.stream>>>(
                static_cast<const char*>(a.in), static_cast<unsigned char*>(a.out), a.counts, a.n_blocks,
                static_cast<const uint8_t*>(a.refs), uint8_t(a.ref_scalar), uint8_t(a.flo), uint8_t(a.fhi));
            return cudaGetLastError();
        }
    }
    if constexpr (sizeof(T) == 2 && W > 0 && W < 16) {
        // u16: FLB_U16_FILTER=warp|slice (A/B; default = measured best)
        static const bool slice = [] {
            const char* e = std::getenv("FLB_U16_FILTER");
            if (e && std::strcmp(e, "warp") == 0) return false;
            if (e && std::strcmp(e, "slice") == 0) return true;
            return FLB_U16_FILTER_DEFAULT_SLICE != 0;
        }();
        if (slice) {
            filter_u16_slice_kernel<W><<<grid_for(a.n_blocks), kThreads, 0, a.stream>>>(
                static_cast<const char*>(a.in), static_cast<unsigned char*>(a.out), a.counts, a.n_blocks,
                static_cast<const uint16_t*>(a.refs), uint16_t(a.ref_scalar), uint16_t(a.flo), uint16_t(a.fhi));
            return cudaGetLastError();
        }
    }
    // blocks per warp: 4 amortises the per-warp set-up while the filter is issue-bound (u32 W=8: 285 -> 208 us); only the
    // verbatim width W = T of u16/u32/u64 measured (slightly) faster with one block per warp; u64 W=61 loses 25 % with one
    // (profiles/opbench_scan_r01.txt)
    constexpr int kNB = (sizeof(T) > 1 && W == Lay<T>::TB) ? 1 : 4;
    const size_t warps = (a.n_blocks + kNB - 1) / kNB;
    const unsigned grid = unsigned((warps * 32 + kThreads - 1) / kThreads);
    // Direct 128-bit loads, not the TMA bulk load: the filter reads little per block and is ALU-bound below W ~ 3T/4,
    // where the mbarrier round trip costs 10-15 % (u32 W=8: 300 vs 346 us; equal at W >= 29 —
    // profiles/opbench_filter_tma_r01.txt).
    // One-block-ahead loads (fl_scan.cuh, RunLoads) with 8 blocks per warp where they measured faster: u64 W = 16 242 -> 195 us,
    // u32 W = 1 / 8 203 -> 189 / 211 -> 205 us per 4 GiB of values; u64 W = 33 / 61 lose 20-40 % to the second register set,
    // everything else is within 1 % (profiles/opbench_filter_pipe_r02.txt).
    constexpr bool kPipe = kNB > 1 && ((sizeof(T) == 8 && W <= 16) || (sizeof(T) == 4 && W <= 8));
    if constexpr (kPipe) {
        const size_t warps8 = (a.n_blocks + 7) / 8;
        filter_warp_kernel<T, W, false, 8, true><<<unsigned((warps8 * 32 + kThreads - 1) / kThreads), kThreads, 0, a.stream>>>(
            static_cast<const char*>(a.in), static_cast<unsigned char*>(a.out), a.counts, a.n_blocks,
            static_cast<const T*>(a.refs), T(a.ref_scalar), T(a.flo), T(a.fhi));
        return cudaGetLastError();
    }
    filter_warp_kernel<T, W, false, kNB><<<grid, kThreads, 0, a.stream>>>(
        static_cast<const char*>(a.in), static_cast<unsigned char*>(a.out), a.counts, a.n_blocks,
        static_cast<const T*>(a.refs), T(a.ref_scalar), T(a.flo), T(a.fhi));
    return cudaGetLastError();
}
// Blocks per warp of the select kernel, from a full-width sweep of NB = 1 / 4 / 8 on one box at 25 % selectivity
// (profiles/select_sweep_r02.txt).  NB > 1 hides the load latency behind the previous block but holds a second set of
// packed words in registers: it pays where the run is short (small or 4-aligned W) and for u64, whose 8 KiB staging buffer
// per warp caps the resident warps anyway; it loses where the unaligned run already needs 40+ registers (u8, u16, u32 W > 16).
template <class T, int W>
constexpr int select_nb() {
    if constexpr (sizeof(T) == 1) return W == 8 ? 8 : 1;
    else if constexpr (sizeof(T) == 2) return (W % 4 == 0 && W < 16) ? 8 : 1;
    else if constexpr (sizeof(T) == 4) {
        constexpr uint64_t nb8 = (1ull << 0) | (1ull << 2) | (1ull << 3) | (1ull << 4) | (0x1FFull << 8) /* 8..16 */ | (1ull << 20) |
                                 (7ull << 25) /* 25..27 */ | (1ull << 32);
        return ((nb8 >> W) & 1) ? 8 : 1;
    } else return 8;
}
template <class T, int W>
static cudaError_t do_select(const LaunchArgs& a) {
    constexpr int NB = select_nb<T, W>();
    // per-warp staging buffer of the compacted values (fl_scan.cuh); u64: 64 KiB per CTA
    const size_t smem = size_t(kThreads / 32) * select_stage_bytes<T>();
    static SmemOptIn opt_in;
    if (const cudaError_t attr = opt_in.ensure(select_warp_kernel<T, W, NB>, smem); attr != cudaSuccess) return attr;
    const size_t warps = (a.n_blocks + NB - 1) / NB;
    select_warp_kernel<T, W, NB><<<unsigned((warps * 32 + kThreads - 1) / kThreads), kThreads, smem, a.stream>>>(
        static_cast<const char*>(a.in), static_cast<const unsigned char*>(a.bitmap), a.offsets, static_cast<T*>(a.out),
        a.n_blocks, static_cast<const T*>(a.refs), T(a.ref_scalar));
    return cudaGetLastError();
}
template <class T, int W>
static cudaError_t do_delta_filter(const LaunchArgs& a) {
    // Blocks per warp, with the loads of block b + 1 issued before block b is evaluated (fl_scan.cuh).  8 where it measured
    // faster (u8: 1.89 -> 1.58 ms at W = 1, 2.09 -> 1.66 at W = 8; u16 W <= 4: 876 -> 756 us; u32 W <= 17: 471 -> 380 us at
    // W = 1), 1 where the second register set costs more than the hidden latency (u16 W >= 9, u32 W >= 29, u64 W = 16);
    // profiles/opbench_filter_pipe_r02.txt.
    constexpr int kNB = sizeof(T) == 1 ? 8 : (sizeof(T) == 2 ? (W <= 4 ? 8 : 1) : (sizeof(T) == 4 ? (W <= 17 ? 8 : 1) : 1));
    const unsigned grid = unsigned((((a.n_blocks + kNB - 1) / kNB) * 32 + kThreads - 1) / kThreads);
    delta_filter_warp_kernel<T, W, kNB><<<grid, kThreads, 0, a.stream>>>(
        static_cast<const char*>(a.in), static_cast<const char*>(a.base), static_cast<unsigned char*>(a.out), a.counts,
        a.n_blocks, T(a.flo), T(a.fhi));
    return cudaGetLastError();
}
template <class T, int... W>
static constexpr std::array<launch_fn, sizeof...(W)> delta_filter_table(std::integer_sequence<int, W...>) {
    return {{&do_delta_filter<T, W>...}};
}
template <>
cudaError_t launch_delta_filter<elem_t>(const LaunchArgs& a) {
    static constexpr auto tab = delta_filter_table<elem_t>(std::make_integer_sequence<int, Lay<elem_t>::TB + 1>{});
    return tab[a.width](a);
}
template <class T, int... W>
static constexpr std::array<launch_fn, sizeof...(W)> filter_table(std::integer_sequence<int, W...>) {
    return {{&do_filter<T, W>...}};
}
template <class T, int... W>
static constexpr std::array<launch_fn, sizeof...(W)> select_table(std::integer_sequence<int, W...>) {
    return {{&do_select<T, W>...}};
}
template <>
cudaError_t launch_filter<elem_t>(const LaunchArgs& a) {
    static constexpr auto tab = filter_table<elem_t>(std::make_integer_sequence<int, Lay<elem_t>::TB + 1>{});
    return tab[a.width](a);
}
template <>
cudaError_t launch_select<elem_t>(const LaunchArgs& a) {
    static constexpr auto tab = select_table<elem_t>(std::make_integer_sequence<int, Lay<elem_t>::TB + 1>{});
    return tab[a.width](a);
}
#else
template <>
cudaError_t launch_delta<elem_t>(bool undo, const LaunchArgs& a) {
    const char* in = static_cast<const char*>(a.in);
    const char* base = static_cast<const char*>(a.base);
    char* out = static_cast<char*>(a.out);
    const unsigned grid = unsigned((a.n_blocks * 32 + kThreads - 1) / kThreads);  // warp-block layout
    constexpr bool kTma = sizeof(elem_t) >= 2;  // TMA bulk load of the block (u8: too small to pay, see do_pack)
    const size_t smem = kTma ? size_t(kThreads / 32) * (128 * Lay<elem_t>::TB + 8) : 0;
    if (undo) {
        static SmemOptIn opt_in;
        if (const cudaError_t attr = opt_in.ensure(delta_warp_kernel<elem_t, true, kTma>, smem); attr != cudaSuccess) return attr;
        delta_warp_kernel<elem_t, true, kTma><<<grid, kThreads, smem, a.stream>>>(in, base, out, a.n_blocks);
    } else {
        static SmemOptIn opt_in;
        if (const cudaError_t attr = opt_in.ensure(delta_warp_kernel<elem_t, false, kTma>, smem); attr != cudaSuccess) return attr;
        delta_warp_kernel<elem_t, false, kTma><<<grid, kThreads, smem, a.stream>>>(in, base, out, a.n_blocks);
    }
    return cudaGetLastError();
}

template <class T, bool UNDO>
static cudaError_t do_transpose_warp(const LaunchArgs& a) {
    const unsigned grid = unsigned((a.n_blocks * 32 + kThreads - 1) / kThreads);
    const size_t smem = size_t(kThreads / 32) * 128 * Lay<T>::TB;
    static SmemOptIn opt_in;
    if (const cudaError_t attr = opt_in.ensure(transpose_warp_kernel<T, UNDO>, smem); attr != cudaSuccess) return attr;
    transpose_warp_kernel<T, UNDO><<<grid, kThreads, smem, a.stream>>>(static_cast<const char*>(a.in), static_cast<char*>(a.out), a.n_blocks);
    return cudaGetLastError();
}
template <>
cudaError_t launch_transpose_warp<elem_t>(bool undo, const LaunchArgs& a) {
    return undo ? do_transpose_warp<elem_t, true>(a) : do_transpose_warp<elem_t, false>(a);
}
#endif

}  // namespace flb
