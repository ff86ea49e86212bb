// fl_kernels.cuh — batched sm_100a kernels of the FastLanes hot path (row-slice layout, see fl_device.cuh).
//
// One thread = one 16-byte column slice of one block; 8 threads = one block; a warp = 4 consecutive
// blocks.  No shared memory, no shuffles, no divergence: the serial per-lane chain of the reference
// (T rows, src/macros.rs:139-170; the T-long prefix-add of src/delta.rs:48-63) is a per-thread register
// chain, and the 8 x (warps) threads supply the memory-level parallelism.  Every kernel is HBM-bound;
// algorithmic bytes per block are in DESIGN.md.
#pragma once
#include <type_traits>
#include <utility>
#include <array>

#include "fl_device.cuh"

namespace flb {

enum UnpackOp : int { UOP_PLAIN = 0, UOP_FOR = 1, UOP_DELTA = 2 };
enum PackOp : int { POP_PLAIN = 0, POP_FOR = 1 };

#ifndef FLB_THREADS
#define FLB_THREADS 256
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

}  // namespace flb
