// fl_internal.h — internal launch interface between the per-type kernel translation units and the C ABI.
#pragma once
#include <atomic>
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace flb {

// Opt-in to more than 48 KiB of dynamic shared memory.  The attribute belongs to the (function, device) pair, so it is
// cached per device ordinal — one bit per device in a per-call-site mask — and a failure is never cached.
struct SmemOptIn {
    std::atomic<uint64_t> done[4] = {};  // 256 device ordinals
    template <class K>
    cudaError_t ensure(K kernel, size_t smem) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        const bool cacheable = dev >= 0 && dev < 256;
        const uint64_t bit = uint64_t(1) << (dev & 63);
        if (cacheable && (done[dev >> 6].load(std::memory_order_acquire) & bit)) return cudaSuccess;
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e == cudaSuccess && cacheable) done[dev >> 6].fetch_or(bit, std::memory_order_release);
        return e;
    }
};

struct LaunchArgs {
    const void* in = nullptr;    // unpacked input (pack/delta/transpose) or packed input (unpack family)
    void* out = nullptr;
    const void* base = nullptr;  // n_blocks x LANES (delta family)
    const void* refs = nullptr;  // per-block references (FoR family) or nullptr
    uint64_t ref_scalar = 0;     // used when refs == nullptr
    size_t n_blocks = 0;
    unsigned width = 0;
    cudaStream_t stream = nullptr;
    // fused scan kernels (fl_scan.cuh)
    const void* bitmap = nullptr;       // select: n_blocks x 128 bytes (filter writes its bitmap to `out`)
    const uint64_t* offsets = nullptr;  // select: exclusive prefix of the per-block selected counts
    uint32_t* counts = nullptr;         // filter: optional per-block popcount
    uint64_t flo = 0, fhi = 0;          // filter: inclusive value range
    void* refs_out = nullptr;           // for_pack_auto: per-block reference (= block minimum) written by the kernel
    void* spans_out = nullptr;          // for_pack_auto: optional per-block max - min
};

// op codes of launch_unpack / launch_pack (match UnpackOp / PackOp in fl_kernels.cuh)
enum : int { kUnpackPlain = 0, kUnpackFor = 1, kUnpackDelta = 2, kUnpackDeltaOrig = 3, kUnpackPlainLinear = 4, kUnpackForLinear = 5 };
enum : int { kPackPlain = 0, kPackFor = 1, kPackOrigDelta = 2, kPackForAuto = 3, kPackPlainLinear = 4, kPackForLinear = 5 };

// Defined once per element type in fl_codec_inst.cu (compiled with -DFLB_TBITS=8/16/32/64).
template <class T> cudaError_t launch_unpack(int op, const LaunchArgs& a);
template <class T> cudaError_t launch_pack(int op, const LaunchArgs& a);
template <class T> cudaError_t launch_delta(bool undo, const LaunchArgs& a);
template <class T> cudaError_t launch_transpose_warp(bool undo, const LaunchArgs& a);
template <class T> cudaError_t launch_filter(const LaunchArgs& a);  // in = packed, out = bitmap
template <class T> cudaError_t launch_select(const LaunchArgs& a);  // in = packed, out = dense values
template <class T> cudaError_t launch_delta_filter(const LaunchArgs& a);  // in = packed, base, out = bitmap

// Defined for all types in fl_misc.cu.
template <class T> cudaError_t launch_transpose(bool undo, const LaunchArgs& a);
template <class T>
cudaError_t launch_gather(unsigned width, size_t n_blocks, const T* packed, const uint64_t* global_index, size_t n,
                          T* out, int* oob_flag, cudaStream_t stream);

template <class T>
cudaError_t launch_block_minmax(size_t n_blocks, const T* in, T* mins, T* maxs, cudaStream_t stream);

}  // namespace flb
