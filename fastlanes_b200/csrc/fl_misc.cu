// fl_misc.cu — transpose / untranspose (a14) and batched unpack_single (a11/a12) kernels, all types.
#include <cstdlib>

#include "fl_device.cuh"
#include "fl_internal.h"

namespace flb {

// ---------------------------------------------------------------------------------------------------
// Transpose::transpose / untranspose (src/transpose.rs:11-22).
//   transposed[i] = original[t(i)],  t(i) = (i%16)*64 + FL_ORDER[(i/16)%8]*8 + i/128   (:29-36)
// Writing i = a + 16 b + 128 c and o = t(i) = 64 a + 8 FL_ORDER[b] + c the permutation is the axis
// reversal [a:16][f:8][c:8] -> [c:8][b:8][a:16] with b = FL_ORDER[f] (FL_ORDER is an involution,
// src/lib.rs:53-59).
// ---------------------------------------------------------------------------------------------------
constexpr int kTrThreads = 256;

__device__ __forceinline__ int transpose_index(int i) {
    return (i % 16) * 64 + fl_order((i / 16) % 8) * 8 + i / 128;  // transpose.rs:31-35
}

// ---------------------------------------------------------------------------------------------------
// CTA-tile variant: 16-byte-chunk tiles, zero shared-memory bank conflicts, no per-element traffic.
// (Selected with FLB_TRANSPOSE=tile; the default is transpose_warp_kernel in fl_kernels.cuh, which measured faster.)
// A thread owns the square tile  {EPC consecutive a} x {the EPC elements of one original chunk}  (EPC = 16/sizeof(T)):
// it moves EPC original chunks <-> EPC transposed chunks and the EPC x EPC element transpose between them is a
// pure register renaming.  The original-order side lives in shared memory with an XOR swizzle on the chunk
// position (so both the linear fill/drain and the tile accesses are conflict-free); the transposed-order side is
// accessed directly in global memory, a warp covering 512 contiguous bytes per instruction.
//   u32: tile (ag,f,ch): chunks (a=4ag+e, f, c=4ch..4ch+3)  <->  chunks' (c=4ch+e', b=FL[f], a=4ag..4ag+3)
//   u64: tile (ap,f,cq): chunks (a=2ap+e, f, c=2cq..2cq+1)  <->  chunks' (c=2cq+e', b=FL[f], a=2ap..2ap+1)
// ---------------------------------------------------------------------------------------------------
template <class T> struct TileCfg;
template <> struct TileCfg<uint32_t> {
    static constexpr int EPC = 4, BLOCKS = 8, TILES = 64, CHUNKS = 256, PAD = 0;
    // tile id -> (lane-major so that a warp's transposed chunks are contiguous): lane = ag*8 + f, warp = ch
    __device__ static void decode(int tile, int& arow0, int& f, int& inner) { inner = tile >> 5; arow0 = ((tile >> 3) & 3) * 4; f = tile & 7; }
    // swizzled position (16-byte units) of original chunk (a, f, inner) inside a block's 4 KiB tile
    __device__ static int smem_chunk(int a, int f, int inner) { return 16 * a + ((2 * f + inner) ^ (f >> 2)); }
    // swizzled position of LINEAR original chunk index x (x = 16a + 2f + ch)
    __device__ static int smem_linear(int x) { const int v = x & 15; return (x & ~15) | (v ^ (v >> 3)); }
    // transposed chunk index for element c of the tile
    __device__ static int tr_chunk(int arow0, int f, int c) { return (arow0 >> 2) + 4 * fl_order(f) + 32 * c; }
};
template <> struct TileCfg<uint64_t> {
    static constexpr int EPC = 2, BLOCKS = 4, TILES = 256, CHUNKS = 512, PAD = 0;
    // lane = bq*8 + ap, warp = (h, cq):  b = 4h + bq, f = FL[b]
    __device__ static void decode(int tile, int& arow0, int& f, int& inner) {
        const int ap = tile & 7, bq = (tile >> 3) & 3, h = (tile >> 5) & 1;
        inner = tile >> 6; arow0 = 2 * ap; f = fl_order_rt(4 * h + bq);
    }
    __device__ static int smem_chunk(int a, int f, int inner) { return 32 * a + ((4 * f + inner) ^ ((a >> 1) & 7)); }
    __device__ static int smem_linear(int x) { const int a = x >> 5; return (x & ~31) | ((x & 31) ^ ((a >> 1) & 7)); }
    __device__ static int tr_chunk(int arow0, int f, int c) { return (arow0 >> 1) + 8 * fl_order_rt(f) + 64 * c; }
};

//   u16: tile (ah,f):    chunks (a=8ah+e, f, c=0..7)             <->  chunks' (c, b=FL[f], a=8ah..8ah+7)   (8x8 halfwords)
//   u8 : tile (fp):      chunks (a=e, f=2fp..2fp+1, c=0..7)       <->  chunks' (c, b=FL[f], a=0..15)        (16x16 bytes)
// (for u8/u16 the element transpose is no longer free: the compiler emits PRMT byte permutes, ~1 per output byte/2.)
template <> struct TileCfg<uint16_t> {
    static constexpr int EPC = 8, BLOCKS = 16, TILES = 16, CHUNKS = 128, PAD = 0;
    __device__ static void decode(int tile, int& arow0, int& f, int& inner) { inner = 0; arow0 = (tile >> 3) * 8; f = tile & 7; }
    __device__ static int smem_chunk(int a, int f, int) { return 8 * a + f; }  // quarter-warp = 8 consecutive f: conflict-free
    __device__ static int smem_linear(int x) { return x; }
    __device__ static int tr_chunk(int arow0, int f, int c) { return (arow0 >> 3) + 2 * fl_order_rt(f) + 16 * c; }
};
template <> struct TileCfg<uint8_t> {
    // PAD: 4 chunks of skew per block so that the two blocks sharing a quarter-warp hit different bank groups
    static constexpr int EPC = 16, BLOCKS = 32, TILES = 4, CHUNKS = 64, PAD = 4;
    __device__ static void decode(int tile, int& arow0, int& f, int& inner) { inner = 0; arow0 = 0; f = tile; }  // f := fp
    __device__ static int smem_chunk(int a, int fp, int) { return 4 * a + fp; }
    __device__ static int smem_linear(int x) { return x; }
    // element x of the original chunk is (f = 2fp + x/8, c = x%8); transposed chunk index = b + 8c
    __device__ static int tr_chunk(int, int fp, int x) { return fl_order_rt(2 * fp + (x >> 3)) + 8 * (x & 7); }
};

template <class T, bool UNDO>
__global__ void __launch_bounds__(kTrThreads)
transpose_tile_kernel(const T* __restrict__ in, T* __restrict__ out, size_t n_blocks) {
    using C = TileCfg<T>;
    constexpr int EPC = C::EPC;
    __shared__ __align__(16) uint4 tile[C::BLOCKS][C::CHUNKS + C::PAD];
    const size_t blk0 = size_t(blockIdx.x) * C::BLOCKS;
    const int nb = int(min(size_t(C::BLOCKS), n_blocks - blk0));

    if constexpr (!UNDO) {
        // fill: original order, linear & coalesced, swizzled shared position
        for (int x = threadIdx.x; x < nb * C::CHUNKS; x += kTrThreads) {
            const int b = x / C::CHUNKS, q = x % C::CHUNKS;
            tile[b][C::smem_linear(q)] = ldg128_stream(reinterpret_cast<const char*>(in + (blk0 + b) * 1024) + q * 16);
        }
        __syncthreads();
    }
    for (int t = threadIdx.x; t < nb * C::TILES; t += kTrThreads) {
        const int b = t / C::TILES;
        int arow0, f, inner;
        C::decode(t % C::TILES, arow0, f, inner);
        alignas(16) T m[EPC][EPC];  // m[e][k]: original chunk (a = arow0+e), element k  (c = inner*EPC + k)
        if constexpr (!UNDO) {
#pragma unroll
            for (int e = 0; e < EPC; ++e) *reinterpret_cast<uint4*>(m[e]) = tile[b][C::smem_chunk(arow0 + e, f, inner)];
            char* o = reinterpret_cast<char*>(out + (blk0 + b) * 1024);
#pragma unroll
            for (int k = 0; k < EPC; ++k) {
                alignas(16) T v[EPC];
#pragma unroll
                for (int e = 0; e < EPC; ++e) v[e] = m[e][k];
                stg128_stream(o + C::tr_chunk(arow0, f, inner * EPC + k) * 16, *reinterpret_cast<uint4*>(v));
            }
        } else {
            const char* ip = reinterpret_cast<const char*>(in + (blk0 + b) * 1024);
#pragma unroll
            for (int k = 0; k < EPC; ++k) {
                alignas(16) T v[EPC];
                *reinterpret_cast<uint4*>(v) = ldg128_stream(ip + C::tr_chunk(arow0, f, inner * EPC + k) * 16);
#pragma unroll
                for (int e = 0; e < EPC; ++e) m[e][k] = v[e];
            }
#pragma unroll
            for (int e = 0; e < EPC; ++e) tile[b][C::smem_chunk(arow0 + e, f, inner)] = *reinterpret_cast<uint4*>(m[e]);
        }
    }
    if constexpr (UNDO) {
        __syncthreads();
        for (int x = threadIdx.x; x < nb * C::CHUNKS; x += kTrThreads) {
            const int b = x / C::CHUNKS, q = x % C::CHUNKS;
            stg128_stream(reinterpret_cast<char*>(out + (blk0 + b) * 1024) + q * 16, tile[b][C::smem_linear(q)]);
        }
    }
}

template <class T>
cudaError_t launch_transpose(bool undo, const LaunchArgs& a) {
    const T* in = static_cast<const T*>(a.in);
    T* out = static_cast<T*>(a.out);
    using TC = TileCfg<T>;
    const unsigned grid = unsigned((a.n_blocks + TC::BLOCKS - 1) / TC::BLOCKS);
    if (undo) transpose_tile_kernel<T, true><<<grid, kTrThreads, 0, a.stream>>>(in, out, a.n_blocks);
    else transpose_tile_kernel<T, false><<<grid, kTrThreads, 0, a.stream>>>(in, out, a.n_blocks);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// Per-block min / max: the statistics a caller needs before FoR::for_pack
// (src/ffor.rs:24-36 takes `reference` and W as givens): reference = min, W = bits(max - min).
// One warp = one block, linear 512-bytes-per-instruction reads, SWAR lane-wise min/max (u16: VIMNMX.U16x2),
// then a butterfly over the warp.  Read-only: 128*T bytes per block.
// ---------------------------------------------------------------------------------------------------
// G threads cooperate on one block (u8: 8, u16: 16, u32/u64: 32) so that every thread reduces >= 8 chunks
// in registers before the log2(G)-step butterfly; a warp covers 32/G consecutive blocks.
template <class T, int G>
__global__ void __launch_bounds__(256)
block_minmax_kernel(const char* __restrict__ in, T* __restrict__ mins, T* __restrict__ maxs, size_t n_blocks) {
    using R = typename Lay<T>::R;
    using M = MinMax<T>;
    constexpr int TB = Lay<T>::TB;
    constexpr int NR = Lay<T>::NR;
    constexpr int CHUNKS = (128 * TB) / 16;  // 16-byte chunks per block
    const size_t tid = size_t(blockIdx.x) * 256 + threadIdx.x;
    const size_t blk = tid / G;
    const int t = int(tid % G);
    const bool active = blk < n_blocks;  // whole groups are active or not; shuffles below stay inside a group
    const char* ib = in + (active ? blk : 0) * (size_t(128) * TB) + t * 16;
    R l, h;
    if constexpr (sizeof(T) == 1) {
        // u8: 16x2 split accumulators (u8_minmax_acc); the SWAR 8x4 intrinsics made this kernel ALU-bound at 5.1 TB/s
        uint32_t mn_e = 0x00FF00FFu, mn_o = 0x00FF00FFu, mx_e = 0, mx_o = 0;
#pragma unroll
        for (int i = 0; i < CHUNKS / G; ++i) {
            const Slice<T> v = load_slice<T>(ib + i * (G * 16));
#pragma unroll
            for (int r = 0; r < NR; ++r) u8_minmax_acc(v.r[r], mn_e, mn_o, mx_e, mx_o);
        }
        l = __vminu2(mn_e, mn_o);  // two 16-bit lanes, values <= 255
        h = __vmaxu2(mx_e, mx_o);
    } else {
        Slice<T> lo = load_slice<T>(ib), hi = lo;
#pragma unroll
        for (int i = 1; i < CHUNKS / G; ++i) {
            const Slice<T> v = load_slice<T>(ib + i * (G * 16));
#pragma unroll
            for (int r = 0; r < NR; ++r) { lo.r[r] = M::mn(lo.r[r], v.r[r]); hi.r[r] = M::mx(hi.r[r], v.r[r]); }
        }
        l = lo.r[0]; h = hi.r[0];
#pragma unroll
        for (int r = 1; r < NR; ++r) { l = M::mn(l, lo.r[r]); h = M::mx(h, hi.r[r]); }
    }
#pragma unroll
    for (int d = G / 2; d >= 1; d >>= 1) {
        R ol, oh;
        if constexpr (sizeof(R) == 8) {
            ol = __shfl_xor_sync(0xffffffffu, (unsigned long long)l, d); oh = __shfl_xor_sync(0xffffffffu, (unsigned long long)h, d);
        } else {
            ol = __shfl_xor_sync(0xffffffffu, l, d); oh = __shfl_xor_sync(0xffffffffu, h, d);
        }
        if constexpr (sizeof(T) == 1) { l = __vminu2(l, ol); h = __vmaxu2(h, oh); }
        else { l = M::mn(l, ol); h = M::mx(h, oh); }
    }
    if (active && t == 0) {
        if constexpr (sizeof(T) == 1) {  // l / h hold two 16-bit lanes
            mins[blk] = T(min(l & 0xFFFFu, l >> 16)); maxs[blk] = T(max(h & 0xFFFFu, h >> 16));
        } else {
            mins[blk] = swar_reduce_min<T>(l); maxs[blk] = swar_reduce_max<T>(h);
        }
    }
}

template <class T>
cudaError_t launch_block_minmax(size_t n_blocks, const T* in, T* mins, T* maxs, cudaStream_t stream) {
    if (n_blocks == 0) return cudaSuccess;
    if constexpr (sizeof(T) == 1) {
        // FLB_MINMAX_G8=8|16|32: threads per u8 block (A/B measurement; default = measured best)
        static const int g8 = [] {
            const char* e = std::getenv("FLB_MINMAX_G8");
            const int v = e ? std::atoi(e) : 8;
            return (v == 16 || v == 32) ? v : 8;
        }();
        const unsigned grid = unsigned((n_blocks * g8 + 255) / 256);
        const char* p = reinterpret_cast<const char*>(in);
        if (g8 == 16) block_minmax_kernel<T, 16><<<grid, 256, 0, stream>>>(p, mins, maxs, n_blocks);
        else if (g8 == 32) block_minmax_kernel<T, 32><<<grid, 256, 0, stream>>>(p, mins, maxs, n_blocks);
        else block_minmax_kernel<T, 8><<<grid, 256, 0, stream>>>(p, mins, maxs, n_blocks);
    } else {
        constexpr int G = (sizeof(T) == 2) ? 16 : 32;
        const unsigned grid = unsigned((n_blocks * G + 255) / 256);
        block_minmax_kernel<T, G><<<grid, 256, 0, stream>>>(reinterpret_cast<const char*>(in), mins, maxs, n_blocks);
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// Batched BitPacking::unpack_single (src/bitpacking.rs:132-179): one thread per query, runtime width.
// (lane,row) follow lanes_by_index / rows_by_index (:207-232) in closed form instead of the 1 KiB tables.
// ---------------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(256)
gather_kernel(unsigned W, size_t n_blocks, const T* __restrict__ packed, const uint64_t* __restrict__ gidx, size_t n,
              T* __restrict__ out, int* __restrict__ oob_flag) {
    constexpr unsigned TB = Lay<T>::TB;
    constexpr unsigned L = Lay<T>::L;
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t g = gidx[i];
    const uint64_t blk = g >> 10;
    if (blk >= n_blocks) {  // the reference's assert!(index < 1024) generalised to the batch (:152)
        out[i] = 0;
        if (oob_flag) *oob_flag = 1;
        return;
    }
    if (W == 0) {  // :137-140
        out[i] = 0;
        return;
    }
    const unsigned index = unsigned(g & 1023);
    const unsigned lane = index % L;                       // :210
    const unsigned s = index / 128;                        // :225
    const unsigned f = (index - s * 128 - lane) / 16;      // :226
    const unsigned row = unsigned(fl_order(int(f))) * 8 + s;  // :227-229
    const T* p = packed + blk * (size_t(1024) * W / TB);
    if (W == TB) {  // :159-162
        out[i] = __ldg(p + L * row + lane);
        return;
    }
    const T mask = T((T(1) << W) - 1);          // :164
    const unsigned start_bit = row * W;          // :165
    const unsigned start_word = start_bit / TB;  // :166
    const unsigned lo_shift = start_bit % TB;    // :167
    const unsigned remaining = TB - lo_shift;    // :168
    const T lo = T(__ldg(p + L * start_word + lane) >> lo_shift);  // :170
    if (remaining >= W) {
        out[i] = T(lo & mask);  // :171-173
    } else {
        const T hi = T(__ldg(p + L * (start_word + 1) + lane) << remaining);  // :176
        out[i] = T(T(lo | hi) & mask);                                        // :177
    }
}

template <class T>
cudaError_t launch_gather(unsigned width, size_t n_blocks, const T* packed, const uint64_t* global_index, size_t n,
                          T* out, int* oob_flag, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const unsigned grid = unsigned((n + 255) / 256);
    gather_kernel<T><<<grid, 256, 0, stream>>>(width, n_blocks, packed, global_index, n, out, oob_flag);
    return cudaGetLastError();
}

#define FLB_INST(T)                                                                                          \
    template cudaError_t launch_transpose<T>(bool, const LaunchArgs&);                                       \
    template cudaError_t launch_gather<T>(unsigned, size_t, const T*, const uint64_t*, size_t, T*, int*,     \
                                          cudaStream_t);                                                    \
    template cudaError_t launch_block_minmax<T>(size_t, const T*, T*, T*, cudaStream_t);
FLB_INST(uint8_t)
FLB_INST(uint16_t)
FLB_INST(uint32_t)
FLB_INST(uint64_t)

}  // namespace flb
