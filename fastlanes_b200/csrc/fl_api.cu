// fl_api.cu — the extern "C" boundary declared in include/fastlanes_b200.h.
//
// Device family: argument checks + one kernel launch on the caller's stream.
// Host family:   chunked H2D -> kernel -> D2H pipeline over internal streams (per-device context).
// There is no CPU compute path anywhere in this library: every value is produced by a CUDA kernel.
#include <sched.h>

#include <atomic>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/fastlanes_b200.h"
#include "fl_device.cuh"
#include "fl_internal.h"

namespace {

using flb::LaunchArgs;

thread_local std::string g_err = "";

fl_status fail(fl_status s, const char* what) {
    g_err = what;
    return s;
}
fl_status cuda_fail(cudaError_t e, const char* where) {
    g_err = std::string(where) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    return FL_ERR_CUDA;
}
#define FL_CUDA(call)                                         \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// One launch covers at most 2^31 blocks: every kernel maps >= 1 block per warp and 8 warps per CTA, so the grid stays
// below 2^28 CTAs (gridDim.x limit 2^31 - 1).  2^31 u8 blocks are already 2 TiB unpacked — beyond any single GPU.
constexpr size_t kMaxBlocksPerLaunch = size_t(1) << 31;

enum class Op { Pack, Unpack, ForPack, UnforPack, Delta, Undelta, UndeltaPack, Transpose, Untranspose,
                UndeltaPackUntranspose, TransposeDeltaPack,
                PackLinear, UnpackLinear, ForPackLinear, UnforPackLinear };  // cwida row order (linear rows)

inline bool op_has_width(Op op) {
    return op == Op::Pack || op == Op::Unpack || op == Op::ForPack || op == Op::UnforPack || op == Op::UndeltaPack ||
           op == Op::UndeltaPackUntranspose || op == Op::TransposeDeltaPack || op == Op::PackLinear ||
           op == Op::UnpackLinear || op == Op::ForPackLinear || op == Op::UnforPackLinear;
}
inline bool op_input_packed(Op op) {
    return op == Op::Unpack || op == Op::UnforPack || op == Op::UndeltaPack || op == Op::UndeltaPackUntranspose ||
           op == Op::UnpackLinear || op == Op::UnforPackLinear;
}
inline bool op_output_packed(Op op) {
    return op == Op::Pack || op == Op::ForPack || op == Op::TransposeDeltaPack || op == Op::PackLinear || op == Op::ForPackLinear;
}
inline bool op_has_base(Op op) {
    return op == Op::Delta || op == Op::Undelta || op == Op::UndeltaPack || op == Op::UndeltaPackUntranspose ||
           op == Op::TransposeDeltaPack;
}

// bytes per block on each side
inline size_t in_block_bytes(Op op, unsigned tbits, unsigned width) {
    return op_input_packed(op) ? size_t(128) * width : size_t(128) * tbits;
}
inline size_t out_block_bytes(Op op, unsigned tbits, unsigned width) {
    return op_output_packed(op) ? size_t(128) * width : size_t(128) * tbits;
}

// Two transpose implementations are built: the CTA-tile kernel (fl_misc.cu) and the warp-block kernel
// (fl_kernels.cuh).  FLB_TRANSPOSE=tile|warp selects one for A/B measurement; the default is the measured best.
inline int transpose_variant() {
    static const int v = [] {
        const char* e = std::getenv("FLB_TRANSPOSE");
        if (e && std::strcmp(e, "tile") == 0) return 0;
        return 1;
    }();
    return v;
}

template <class T>
fl_status device_op(Op op, unsigned width, size_t n_blocks, const void* in, void* out, const void* base,
                    const void* refs, uint64_t ref_scalar, cudaStream_t stream) {
    constexpr unsigned TB = sizeof(T) * 8;
    if (op_has_width(op) && width > TB) return fail(FL_ERR_WIDTH, "width exceeds the bit size of the element type");
    if (!op_has_width(op)) width = 0;
    if (n_blocks == 0) return FL_OK;
    if (n_blocks > kMaxBlocksPerLaunch) return fail(FL_ERR_LEN, "n_blocks too large for one launch (limit 2^31)");
    const bool in_used = in_block_bytes(op, TB, width) != 0;
    const bool out_used = out_block_bytes(op, TB, width) != 0;
    if ((in_used && !in) || (out_used && !out)) return fail(FL_ERR_NULL, "null data pointer");
    if (op_has_base(op) && !base) return fail(FL_ERR_NULL, "null base pointer");
    if ((in_used && !aligned16(in)) || (out_used && !aligned16(out)) || (op_has_base(op) && !aligned16(base)))
        return fail(FL_ERR_ALIGN, "device pointers must be 16-byte aligned");
    if (!out_used) return FL_OK;  // pack at width 0 writes nothing (src/macros.rs:52)

    LaunchArgs a;
    a.in = in; a.out = out; a.base = base; a.refs = refs; a.ref_scalar = ref_scalar;
    a.n_blocks = n_blocks; a.width = width; a.stream = stream;
    cudaError_t e = cudaSuccess;
    switch (op) {
        case Op::Pack: e = flb::launch_pack<T>(flb::kPackPlain, a); break;
        case Op::ForPack: e = flb::launch_pack<T>(flb::kPackFor, a); break;
        case Op::Unpack: e = flb::launch_unpack<T>(flb::kUnpackPlain, a); break;
        case Op::UnforPack: e = flb::launch_unpack<T>(flb::kUnpackFor, a); break;
        case Op::UndeltaPack: e = flb::launch_unpack<T>(flb::kUnpackDelta, a); break;
        case Op::UndeltaPackUntranspose: e = flb::launch_unpack<T>(flb::kUnpackDeltaOrig, a); break;
        case Op::TransposeDeltaPack: e = flb::launch_pack<T>(flb::kPackOrigDelta, a); break;
        case Op::PackLinear: e = flb::launch_pack<T>(flb::kPackPlainLinear, a); break;
        case Op::ForPackLinear: e = flb::launch_pack<T>(flb::kPackForLinear, a); break;
        case Op::UnpackLinear: e = flb::launch_unpack<T>(flb::kUnpackPlainLinear, a); break;
        case Op::UnforPackLinear: e = flb::launch_unpack<T>(flb::kUnpackForLinear, a); break;
        case Op::Delta: e = flb::launch_delta<T>(false, a); break;
        case Op::Undelta: e = flb::launch_delta<T>(true, a); break;
        case Op::Transpose: e = transpose_variant() ? flb::launch_transpose_warp<T>(false, a) : flb::launch_transpose<T>(false, a); break;
        case Op::Untranspose: e = transpose_variant() ? flb::launch_transpose_warp<T>(true, a) : flb::launch_transpose<T>(true, a); break;
    }
    if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
    return FL_OK;
}

// ---- NUMA placement of page-locked host memory ---------------------------------------------------
// node of the PCIe device behind CUDA device `dev` (sysfs), or -1 when unknown
int device_numa_node(int dev) {
    char id[32] = {0};
    if (cudaDeviceGetPCIBusId(id, int(sizeof(id)), dev) != cudaSuccess) { (void)cudaGetLastError(); return -1; }
    for (char* c = id; *c; ++c) *c = char(std::tolower(static_cast<unsigned char>(*c)));
    const std::string path = std::string("/sys/bus/pci/devices/") + id + "/numa_node";
    FILE* f = std::fopen(path.c_str(), "r");
    if (!f) return -1;
    int node = -1;
    if (std::fscanf(f, "%d", &node) != 1) node = -1;
    std::fclose(f);
    return node;
}
// CPUs of a NUMA node from /sys/devices/system/node/node<N>/cpulist ("0-31,64-95")
bool node_cpuset(int node, cpu_set_t* set) {
    if (node < 0) return false;
    const std::string path = "/sys/devices/system/node/node" + std::to_string(node) + "/cpulist";
    FILE* f = std::fopen(path.c_str(), "r");
    if (!f) return false;
    char buf[4096] = {0};
    const bool ok = std::fgets(buf, sizeof(buf), f) != nullptr;
    std::fclose(f);
    if (!ok) return false;
    CPU_ZERO(set);
    int n = 0;
    for (const char* p = buf; *p && *p != '\n';) {
        char* end = nullptr;
        const long a = std::strtol(p, &end, 10);
        if (end == p) return false;
        long b = a;
        p = end;
        if (*p == '-') { b = std::strtol(p + 1, &end, 10); if (end == p + 1) return false; p = end; }
        for (long c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET(int(c), set); ++n; }
        if (*p == ',') ++p;
    }
    return n > 0;
}

// ---- host path ----------------------------------------------------------------------------------
struct Slot {
    cudaStream_t stream = nullptr;
    void* d_in = nullptr;
    void* d_out = nullptr;
    void* d_base = nullptr;
    size_t in_cap = 0, out_cap = 0, base_cap = 0;
};

struct HostCtx {
    int device = -1;
    std::vector<Slot> slots;
    std::mutex mu;
};

std::mutex g_ctx_mu;
std::vector<HostCtx*> g_ctxs;
std::atomic<size_t> g_chunk_blocks{16384};  // fl_host_configure may race with fl_host_* calls on other threads
std::atomic<int> g_n_streams{3};

fl_status get_ctx(HostCtx** out) {
    int dev = -1;
    FL_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    for (HostCtx* c : g_ctxs)
        if (c->device == dev) { *out = c; return FL_OK; }
    HostCtx* c = new HostCtx;
    c->device = dev;
    g_ctxs.push_back(c);
    *out = c;
    return FL_OK;
}

fl_status ensure(void** p, size_t* cap, size_t need) {
    if (*cap >= need) return FL_OK;
    if (*p) FL_CUDA(cudaFree(*p));
    *p = nullptr; *cap = 0;
    FL_CUDA(cudaMalloc(p, need));
    *cap = need;
    return FL_OK;
}

template <class T>
fl_status host_op(Op op, unsigned width, size_t n_blocks, const void* in, void* out, const void* base,
                  uint64_t ref_scalar) {
    constexpr unsigned TB = sizeof(T) * 8;
    if (op_has_width(op) && width > TB) return fail(FL_ERR_WIDTH, "width exceeds the bit size of the element type");
    if (!op_has_width(op)) width = 0;
    if (n_blocks == 0) return FL_OK;
    const size_t ib = in_block_bytes(op, TB, width), ob = out_block_bytes(op, TB, width);
    if ((ib && !in) || (ob && !out)) return fail(FL_ERR_NULL, "null data pointer");
    if (op_has_base(op) && !base) return fail(FL_ERR_NULL, "null base pointer");
    if (ob == 0) return FL_OK;

    HostCtx* ctx = nullptr;
    if (fl_status s = get_ctx(&ctx)) return s;
    std::lock_guard<std::mutex> lk(ctx->mu);
    const size_t chunk_cfg = g_chunk_blocks.load();
    const size_t chunk = chunk_cfg ? chunk_cfg : 16384;
    const size_t n_chunks = (n_blocks + chunk - 1) / chunk;
    const int streams_cfg = g_n_streams.load();
    const size_t n_slots = size_t(streams_cfg > 0 ? streams_cfg : 3);
    if (ctx->slots.size() < n_slots) ctx->slots.resize(n_slots);
    const size_t use_slots = n_chunks < n_slots ? n_chunks : n_slots;
    const size_t cb = n_blocks < chunk ? n_blocks : chunk;
    for (size_t s = 0; s < use_slots; ++s) {
        Slot& sl = ctx->slots[s];
        if (!sl.stream) FL_CUDA(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
        if (ib) if (fl_status st = ensure(&sl.d_in, &sl.in_cap, cb * ib)) return st;
        if (fl_status st = ensure(&sl.d_out, &sl.out_cap, cb * ob)) return st;
        if (op_has_base(op)) if (fl_status st = ensure(&sl.d_base, &sl.base_cap, cb * 128)) return st;
    }
    fl_status result = FL_OK;
    for (size_t c = 0; c < n_chunks && result == FL_OK; ++c) {
        Slot& sl = ctx->slots[c % use_slots];
        const size_t b0 = c * chunk;
        const size_t nb = (n_blocks - b0) < chunk ? (n_blocks - b0) : chunk;
        // stream order serialises reuse of this slot's device buffers with the previous chunk's D2H.
        // Errors break out to the drain loop below: no copy touching the caller's buffers may outlive the call.
        cudaError_t e = cudaSuccess;
        if (ib) e = cudaMemcpyAsync(sl.d_in, static_cast<const char*>(in) + b0 * ib, nb * ib, cudaMemcpyHostToDevice, sl.stream);
        if (e == cudaSuccess && op_has_base(op))
            e = cudaMemcpyAsync(sl.d_base, static_cast<const char*>(base) + b0 * 128, nb * 128, cudaMemcpyHostToDevice, sl.stream);
        if (e != cudaSuccess) { result = cuda_fail(e, "cudaMemcpyAsync H2D"); break; }
        result = device_op<T>(op, width, nb, sl.d_in, sl.d_out, sl.d_base, nullptr, ref_scalar, sl.stream);
        if (result != FL_OK) break;
        e = cudaMemcpyAsync(static_cast<char*>(out) + b0 * ob, sl.d_out, nb * ob, cudaMemcpyDeviceToHost, sl.stream);
        if (e != cudaSuccess) { result = cuda_fail(e, "cudaMemcpyAsync D2H"); break; }
    }
    for (size_t s = 0; s < use_slots; ++s) {
        cudaError_t e = cudaStreamSynchronize(ctx->slots[s].stream);
        if (e != cudaSuccess && result == FL_OK) result = cuda_fail(e, "cudaStreamSynchronize");
    }
    return result;
}

template <class T>
fl_status device_minmax(size_t n_blocks, const T* in, T* mins, T* maxs, cudaStream_t stream) {
    if (n_blocks == 0) return FL_OK;
    if (n_blocks > kMaxBlocksPerLaunch) return fail(FL_ERR_LEN, "n_blocks too large for one launch (limit 2^31)");
    if (!in || !mins || !maxs) return fail(FL_ERR_NULL, "null pointer");
    if (!aligned16(in)) return fail(FL_ERR_ALIGN, "device pointers must be 16-byte aligned");
    cudaError_t e = flb::launch_block_minmax<T>(n_blocks, in, mins, maxs, stream);
    if (e != cudaSuccess) return cuda_fail(e, "minmax launch");
    return FL_OK;
}

template <class T>
fl_status host_minmax(size_t n_blocks, const T* in, T* mins, T* maxs) {
    if (n_blocks == 0) return FL_OK;
    if (!in || !mins || !maxs) return fail(FL_ERR_NULL, "null pointer");
    HostCtx* ctx = nullptr;
    if (fl_status s = get_ctx(&ctx)) return s;
    std::lock_guard<std::mutex> lk(ctx->mu);
    const size_t chunk_cfg = g_chunk_blocks.load();
    const size_t chunk = chunk_cfg ? chunk_cfg : 16384;
    const int streams_cfg = g_n_streams.load();
    const size_t n_slots = size_t(streams_cfg > 0 ? streams_cfg : 3);
    if (ctx->slots.size() < n_slots) ctx->slots.resize(n_slots);
    const size_t n_chunks = (n_blocks + chunk - 1) / chunk;
    const size_t use_slots = n_chunks < n_slots ? n_chunks : n_slots;
    const size_t cb = n_blocks < chunk ? n_blocks : chunk;
    const size_t half = (cb * sizeof(T) + 15) & ~size_t(15);
    for (size_t s = 0; s < use_slots; ++s) {
        Slot& sl = ctx->slots[s];
        if (!sl.stream) FL_CUDA(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
        if (fl_status st = ensure(&sl.d_in, &sl.in_cap, cb * 1024 * sizeof(T))) return st;
        if (fl_status st = ensure(&sl.d_out, &sl.out_cap, 2 * half)) return st;
    }
    fl_status result = FL_OK;
    for (size_t c = 0; c < n_chunks && result == FL_OK; ++c) {
        Slot& sl = ctx->slots[c % use_slots];
        const size_t b0 = c * chunk;
        const size_t nb = (n_blocks - b0) < chunk ? (n_blocks - b0) : chunk;
        T* d_min = static_cast<T*>(sl.d_out);
        T* d_max = reinterpret_cast<T*>(static_cast<char*>(sl.d_out) + half);
        cudaError_t e = cudaMemcpyAsync(sl.d_in, in + b0 * 1024, nb * 1024 * sizeof(T), cudaMemcpyHostToDevice, sl.stream);
        if (e != cudaSuccess) { result = cuda_fail(e, "cudaMemcpyAsync H2D"); break; }
        result = device_minmax<T>(nb, static_cast<const T*>(sl.d_in), d_min, d_max, sl.stream);
        if (result != FL_OK) break;
        e = cudaMemcpyAsync(mins + b0, d_min, nb * sizeof(T), cudaMemcpyDeviceToHost, sl.stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(maxs + b0, d_max, nb * sizeof(T), cudaMemcpyDeviceToHost, sl.stream);
        if (e != cudaSuccess) { result = cuda_fail(e, "cudaMemcpyAsync D2H"); break; }
    }
    for (size_t s = 0; s < use_slots; ++s) {
        cudaError_t e = cudaStreamSynchronize(ctx->slots[s].stream);
        if (e != cudaSuccess && result == FL_OK) result = cuda_fail(e, "cudaStreamSynchronize");
    }
    return result;
}

// for_pack with reference = block minimum, statistics fused into the pack pass
template <class T>
fl_status device_for_pack_auto(unsigned width, size_t n_blocks, const T* in, T* refs_out, T* spans_out, T* packed,
                               cudaStream_t stream) {
    if (width > sizeof(T) * 8) return fail(FL_ERR_WIDTH, "width exceeds the bit size of the element type");
    if (n_blocks == 0) return FL_OK;
    if (n_blocks > kMaxBlocksPerLaunch) return fail(FL_ERR_LEN, "n_blocks too large for one launch (limit 2^31)");
    if (!in || !refs_out || (width && !packed)) return fail(FL_ERR_NULL, "null pointer");
    if (!aligned16(in) || (width && !aligned16(packed))) return fail(FL_ERR_ALIGN, "device pointers must be 16-byte aligned");
    LaunchArgs a;
    a.in = in; a.out = packed; a.refs_out = refs_out; a.spans_out = spans_out;
    a.n_blocks = n_blocks; a.width = width; a.stream = stream;
    const cudaError_t e = flb::launch_pack<T>(flb::kPackForAuto, a);
    if (e != cudaSuccess) return cuda_fail(e, "for_pack_auto launch");
    return FL_OK;
}

// ---- fused scan (fl_scan.cuh) -------------------------------------------------------------------
template <class T>
fl_status device_filter(unsigned width, size_t n_blocks, const T* packed, const T* refs, T reference, T lo, T hi,
                        uint8_t* bitmap, uint32_t* counts, cudaStream_t stream) {
    if (width > sizeof(T) * 8) return fail(FL_ERR_WIDTH, "width exceeds the bit size of the element type");
    if (n_blocks == 0) return FL_OK;
    if (n_blocks > kMaxBlocksPerLaunch) return fail(FL_ERR_LEN, "n_blocks too large for one launch (limit 2^31)");
    if (!bitmap || (width && !packed)) return fail(FL_ERR_NULL, "null pointer");
    if ((width && !aligned16(packed)) || !aligned16(bitmap)) return fail(FL_ERR_ALIGN, "device pointers must be 16-byte aligned");
    LaunchArgs a;
    a.in = packed; a.out = bitmap; a.counts = counts; a.refs = refs; a.ref_scalar = reference;
    a.flo = lo; a.fhi = hi; a.n_blocks = n_blocks; a.width = width; a.stream = stream;
    const cudaError_t e = flb::launch_filter<T>(a);
    if (e != cudaSuccess) return cuda_fail(e, "filter launch");
    return FL_OK;
}

template <class T>
fl_status device_select(unsigned width, size_t n_blocks, const T* packed, const T* refs, T reference,
                        const uint8_t* bitmap, const uint64_t* offsets, T* out, cudaStream_t stream) {
    if (width > sizeof(T) * 8) return fail(FL_ERR_WIDTH, "width exceeds the bit size of the element type");
    if (n_blocks == 0) return FL_OK;
    if (n_blocks > kMaxBlocksPerLaunch) return fail(FL_ERR_LEN, "n_blocks too large for one launch (limit 2^31)");
    if (!bitmap || !offsets || !out || (width && !packed)) return fail(FL_ERR_NULL, "null pointer");
    if ((width && !aligned16(packed)) || !aligned16(bitmap)) return fail(FL_ERR_ALIGN, "device pointers must be 16-byte aligned");
    LaunchArgs a;
    a.in = packed; a.out = out; a.bitmap = bitmap; a.offsets = offsets; a.refs = refs; a.ref_scalar = reference;
    a.n_blocks = n_blocks; a.width = width; a.stream = stream;
    const cudaError_t e = flb::launch_select<T>(a);
    if (e != cudaSuccess) return cuda_fail(e, "select launch");
    return FL_OK;
}

template <class T>
fl_status device_delta_filter(unsigned width, size_t n_blocks, const T* packed, const T* base, T lo, T hi, uint8_t* bitmap,
                              uint32_t* counts, cudaStream_t stream) {
    if (width > sizeof(T) * 8) return fail(FL_ERR_WIDTH, "width exceeds the bit size of the element type");
    if (n_blocks == 0) return FL_OK;
    if (n_blocks > kMaxBlocksPerLaunch) return fail(FL_ERR_LEN, "n_blocks too large for one launch (limit 2^31)");
    if (!bitmap || !base || (width && !packed)) return fail(FL_ERR_NULL, "null pointer");
    if ((width && !aligned16(packed)) || !aligned16(bitmap) || !aligned16(base))
        return fail(FL_ERR_ALIGN, "device pointers must be 16-byte aligned");
    LaunchArgs a;
    a.in = packed; a.base = base; a.out = bitmap; a.counts = counts; a.flo = lo; a.fhi = hi;
    a.n_blocks = n_blocks; a.width = width; a.stream = stream;
    const cudaError_t e = flb::launch_delta_filter<T>(a);
    if (e != cudaSuccess) return cuda_fail(e, "delta filter launch");
    return FL_OK;
}

// host buffers: H2D of the packed chunk, filter kernel, D2H of 128 (+4) bytes per block — the decoded values never
// cross the PCIe link
// `base` != nullptr selects the delta scan (reference unused)
template <class T>
fl_status host_filter(unsigned width, size_t n_blocks, const T* packed, const T* base, T reference, T lo, T hi,
                      uint8_t* bitmap, uint32_t* counts) {
    if (width > sizeof(T) * 8) return fail(FL_ERR_WIDTH, "width exceeds the bit size of the element type");
    if (n_blocks == 0) return FL_OK;
    if (!bitmap || (width && !packed)) return fail(FL_ERR_NULL, "null pointer");
    HostCtx* ctx = nullptr;
    if (fl_status s = get_ctx(&ctx)) return s;
    std::lock_guard<std::mutex> lk(ctx->mu);
    const size_t chunk_cfg = g_chunk_blocks.load();
    const size_t chunk = chunk_cfg ? chunk_cfg : 16384;
    const int streams_cfg = g_n_streams.load();
    const size_t n_slots = size_t(streams_cfg > 0 ? streams_cfg : 3);
    if (ctx->slots.size() < n_slots) ctx->slots.resize(n_slots);
    const size_t n_chunks = (n_blocks + chunk - 1) / chunk;
    const size_t use_slots = n_chunks < n_slots ? n_chunks : n_slots;
    const size_t cb = n_blocks < chunk ? n_blocks : chunk;
    const size_t ib = size_t(128) * width;
    for (size_t s = 0; s < use_slots; ++s) {
        Slot& sl = ctx->slots[s];
        if (!sl.stream) FL_CUDA(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
        if (ib) if (fl_status st = ensure(&sl.d_in, &sl.in_cap, cb * ib)) return st;
        if (fl_status st = ensure(&sl.d_out, &sl.out_cap, cb * 128 + cb * sizeof(uint32_t))) return st;
        if (base) if (fl_status st = ensure(&sl.d_base, &sl.base_cap, cb * 128)) return st;
    }
    fl_status result = FL_OK;
    for (size_t c = 0; c < n_chunks && result == FL_OK; ++c) {
        Slot& sl = ctx->slots[c % use_slots];
        const size_t b0 = c * chunk;
        const size_t nb = (n_blocks - b0) < chunk ? (n_blocks - b0) : chunk;
        uint8_t* d_bitmap = static_cast<uint8_t*>(sl.d_out);
        uint32_t* d_counts = reinterpret_cast<uint32_t*>(d_bitmap + cb * 128);
        cudaError_t e = cudaSuccess;
        if (ib) e = cudaMemcpyAsync(sl.d_in, reinterpret_cast<const char*>(packed) + b0 * ib, nb * ib, cudaMemcpyHostToDevice, sl.stream);
        if (e == cudaSuccess && base)
            e = cudaMemcpyAsync(sl.d_base, reinterpret_cast<const char*>(base) + b0 * 128, nb * 128, cudaMemcpyHostToDevice, sl.stream);
        if (e != cudaSuccess) { result = cuda_fail(e, "cudaMemcpyAsync H2D"); break; }
        if (base)
            result = device_delta_filter<T>(width, nb, static_cast<const T*>(sl.d_in), static_cast<const T*>(sl.d_base), lo, hi,
                                            d_bitmap, counts ? d_counts : nullptr, sl.stream);
        else
            result = device_filter<T>(width, nb, static_cast<const T*>(sl.d_in), nullptr, reference, lo, hi, d_bitmap,
                                      counts ? d_counts : nullptr, sl.stream);
        if (result != FL_OK) break;
        e = cudaMemcpyAsync(bitmap + b0 * 128, d_bitmap, nb * 128, cudaMemcpyDeviceToHost, sl.stream);
        if (e == cudaSuccess && counts)
            e = cudaMemcpyAsync(counts + b0, d_counts, nb * sizeof(uint32_t), cudaMemcpyDeviceToHost, sl.stream);
        if (e != cudaSuccess) { result = cuda_fail(e, "cudaMemcpyAsync D2H"); break; }
    }
    for (size_t s = 0; s < use_slots; ++s) {  // always drain: no copy may outlive the call
        cudaError_t e = cudaStreamSynchronize(ctx->slots[s].stream);
        if (e != cudaSuccess && result == FL_OK) result = cuda_fail(e, "cudaStreamSynchronize");
    }
    return result;
}

template <class T>
fl_status device_gather(unsigned width, size_t n_blocks, const T* packed, const uint64_t* gidx, size_t n, T* out,
                        int* oob_flag, cudaStream_t stream) {
    if (width > sizeof(T) * 8) return fail(FL_ERR_WIDTH, "width exceeds the bit size of the element type");
    if (n == 0) return FL_OK;
    if (!gidx || !out || (width && n_blocks && !packed)) return fail(FL_ERR_NULL, "null pointer");
    cudaError_t e = flb::launch_gather<T>(width, n_blocks, packed, gidx, n, out, oob_flag, stream);
    if (e != cudaSuccess) return cuda_fail(e, "gather launch");
    return FL_OK;
}

template <class T>
fl_status host_gather(unsigned width, size_t n_blocks, const T* packed, const uint64_t* gidx, size_t n, T* out) {
    if (width > sizeof(T) * 8) return fail(FL_ERR_WIDTH, "width exceeds the bit size of the element type");
    if (n == 0) return FL_OK;
    if (!gidx || !out || (width && n_blocks && !packed)) return fail(FL_ERR_NULL, "null pointer");
    HostCtx* ctx = nullptr;
    if (fl_status s = get_ctx(&ctx)) return s;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (ctx->slots.empty()) ctx->slots.resize(1);
    Slot& sl = ctx->slots[0];
    if (!sl.stream) FL_CUDA(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
    const size_t pbytes = n_blocks * size_t(128) * width;
    // d_in: packed blocks; d_out: [indices | values | oob flag]
    if (pbytes) if (fl_status st = ensure(&sl.d_in, &sl.in_cap, pbytes)) return st;
    const size_t idx_bytes = n * sizeof(uint64_t), val_bytes = (n * sizeof(T) + 15) & ~size_t(15);
    if (fl_status st = ensure(&sl.d_out, &sl.out_cap, idx_bytes + val_bytes + 16)) return st;
    char* d = static_cast<char*>(sl.d_out);
    uint64_t* d_idx = reinterpret_cast<uint64_t*>(d);
    T* d_val = reinterpret_cast<T*>(d + idx_bytes);
    int* d_flag = reinterpret_cast<int*>(d + idx_bytes + val_bytes);
    int flag = 0;
    fl_status result = FL_OK;
    cudaError_t e = cudaSuccess;
    if (pbytes) e = cudaMemcpyAsync(sl.d_in, packed, pbytes, cudaMemcpyHostToDevice, sl.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_idx, gidx, idx_bytes, cudaMemcpyHostToDevice, sl.stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_flag, 0, sizeof(int), sl.stream);
    if (e != cudaSuccess) result = cuda_fail(e, "cudaMemcpyAsync H2D");
    if (result == FL_OK)
        result = device_gather<T>(width, n_blocks, static_cast<const T*>(sl.d_in), d_idx, n, d_val, d_flag, sl.stream);
    if (result == FL_OK) {
        e = cudaMemcpyAsync(out, d_val, n * sizeof(T), cudaMemcpyDeviceToHost, sl.stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, sl.stream);
        if (e != cudaSuccess) result = cuda_fail(e, "cudaMemcpyAsync D2H");
    }
    e = cudaStreamSynchronize(sl.stream);  // always drain: `flag` and the caller's buffers must not be written later
    if (e != cudaSuccess && result == FL_OK) result = cuda_fail(e, "cudaStreamSynchronize");
    if (result != FL_OK) return result;
    if (flag) return fail(FL_ERR_INDEX, "index out of range");  // src/bitpacking.rs:152
    return FL_OK;
}

}  // namespace

extern "C" {

const char* fl_version(void) { return "fastlanes_b200 0.1.0 (sm_100a; wire format spiraldb/fastlanes 0.1.8)"; }
const char* fl_last_error_string(void) { return g_err.c_str(); }
const char* fl_status_string(fl_status s) {
    switch (s) {
        case FL_OK: return "FL_OK";
        case FL_ERR_WIDTH: return "FL_ERR_WIDTH";
        case FL_ERR_LEN: return "FL_ERR_LEN";
        case FL_ERR_INDEX: return "FL_ERR_INDEX";
        case FL_ERR_ALIGN: return "FL_ERR_ALIGN";
        case FL_ERR_CUDA: return "FL_ERR_CUDA";
        case FL_ERR_NULL: return "FL_ERR_NULL";
        case FL_ERR_UNSUPPORTED: return "FL_ERR_UNSUPPORTED";
        default: return "FL_ERR_UNKNOWN";
    }
}
int fl_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    return n;
}
fl_status fl_init(int device) {
    FL_CUDA(cudaSetDevice(device));
    FL_CUDA(cudaFree(nullptr));  // force primary-context creation
    HostCtx* ctx = nullptr;
    if (fl_status s = get_ctx(&ctx)) return s;
    std::lock_guard<std::mutex> lk(ctx->mu);
    const int streams_cfg = g_n_streams.load();
    const size_t n_slots = size_t(streams_cfg > 0 ? streams_cfg : 3);
    if (ctx->slots.size() < n_slots) ctx->slots.resize(n_slots);
    for (Slot& sl : ctx->slots)
        if (!sl.stream) FL_CUDA(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
    return FL_OK;
}
fl_status fl_host_configure(size_t chunk_blocks, int n_streams) {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    g_chunk_blocks.store(chunk_blocks ? chunk_blocks : 16384);
    g_n_streams.store(n_streams > 0 ? (n_streams > 16 ? 16 : n_streams) : 3);
    return FL_OK;
}
int fl_device_numa_node(int device) { return device_numa_node(device); }
fl_status fl_host_alloc(void** p, size_t bytes) {
    if (!p) return fail(FL_ERR_NULL, "null pointer");
    // NUMA-local staging: the pages of a page-locked allocation are placed on the node of the allocating thread, so
    // run the allocation on a CPU of the GPU's own node (a D2H stream that crosses the socket interconnect loses
    // bandwidth, and with one rank per GPU every rank would otherwise land on whatever node it was started on).
    // FLB_NUMA=0 disables; any failure falls back to a plain allocation.
    int dev = -1;
    cpu_set_t local, saved;
    bool bound = false;
    const char* e = std::getenv("FLB_NUMA");
    if (!(e && e[0] == '0') && cudaGetDevice(&dev) == cudaSuccess && node_cpuset(device_numa_node(dev), &local) &&
        sched_getaffinity(0, sizeof(saved), &saved) == 0) {
        cpu_set_t want;
        CPU_AND(&want, &local, &saved);  // stay inside the CPUs this thread may use
        if (CPU_COUNT(&want) > 0 && sched_setaffinity(0, sizeof(want), &want) == 0) bound = true;
    }
    const cudaError_t err = cudaHostAlloc(p, bytes, cudaHostAllocDefault);
    if (bound) (void)sched_setaffinity(0, sizeof(saved), &saved);
    if (err != cudaSuccess) return cuda_fail(err, "cudaHostAlloc");
    return FL_OK;
}
fl_status fl_host_free(void* p) {
    if (p) FL_CUDA(cudaFreeHost(p));
    return FL_OK;
}
fl_status fl_host_register(void* p, size_t bytes) {
    if (!p) return fail(FL_ERR_NULL, "null pointer");
    FL_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterDefault));
    return FL_OK;
}
fl_status fl_host_unregister(void* p) {
    if (!p) return fail(FL_ERR_NULL, "null pointer");
    FL_CUDA(cudaHostUnregister(p));
    return FL_OK;
}
fl_status fl_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    int cur = -1;
    (void)cudaGetDevice(&cur);
    for (HostCtx* c : g_ctxs) {
        std::lock_guard<std::mutex> lk2(c->mu);
        if (cudaSetDevice(c->device) != cudaSuccess) continue;
        for (Slot& s : c->slots) {
            if (s.stream) { cudaStreamSynchronize(s.stream); cudaStreamDestroy(s.stream); }
            if (s.d_in) cudaFree(s.d_in);
            if (s.d_out) cudaFree(s.d_out);
            if (s.d_base) cudaFree(s.d_base);
            s = Slot{};
        }
        c->slots.clear();
    }
    if (cur >= 0) (void)cudaSetDevice(cur);
    (void)cudaGetLastError();
    return FL_OK;
}

#define FL_DEFINE_TYPE(SFX, T)                                                                                          \
    fl_status fl_pack_##SFX(unsigned width, size_t n, const T* in, T* packed, void* st) {                               \
        return device_op<T>(Op::Pack, width, n, in, packed, nullptr, nullptr, 0, (cudaStream_t)st);                     \
    }                                                                                                                   \
    fl_status fl_host_pack_##SFX(unsigned width, size_t n, const T* in, T* packed) {                                    \
        return host_op<T>(Op::Pack, width, n, in, packed, nullptr, 0);                                                  \
    }                                                                                                                   \
    fl_status fl_unpack_##SFX(unsigned width, size_t n, const T* packed, T* out, void* st) {                            \
        return device_op<T>(Op::Unpack, width, n, packed, out, nullptr, nullptr, 0, (cudaStream_t)st);                  \
    }                                                                                                                   \
    fl_status fl_host_unpack_##SFX(unsigned width, size_t n, const T* packed, T* out) {                                 \
        return host_op<T>(Op::Unpack, width, n, packed, out, nullptr, 0);                                               \
    }                                                                                                                   \
    fl_status fl_unpack_gather_##SFX(unsigned width, size_t n_blocks, const T* packed, const uint64_t* gi, size_t n,    \
                                     T* out, int* oob, void* st) {                                                      \
        return device_gather<T>(width, n_blocks, packed, gi, n, out, oob, (cudaStream_t)st);                            \
    }                                                                                                                   \
    fl_status fl_host_unpack_gather_##SFX(unsigned width, size_t n_blocks, const T* packed, const uint64_t* gi,         \
                                          size_t n, T* out) {                                                           \
        return host_gather<T>(width, n_blocks, packed, gi, n, out);                                                     \
    }                                                                                                                   \
    fl_status fl_host_unpack_single_##SFX(unsigned width, const T* packed, size_t index, T* value) {                    \
        if (index >= 1024) return fail(FL_ERR_INDEX, "index must be less than 1024"); /* bitpacking.rs:152 */           \
        const uint64_t gi = index;                                                                                      \
        return host_gather<T>(width, 1, packed, &gi, 1, value);                                                         \
    }                                                                                                                   \
    fl_status fl_for_pack_##SFX(unsigned width, size_t n, const T* in, T reference, T* packed, void* st) {              \
        return device_op<T>(Op::ForPack, width, n, in, packed, nullptr, nullptr, reference, (cudaStream_t)st);          \
    }                                                                                                                   \
    fl_status fl_for_pack_refs_##SFX(unsigned width, size_t n, const T* in, const T* refs, T* packed, void* st) {       \
        if (!refs) return fail(FL_ERR_NULL, "null refs pointer");                                                       \
        return device_op<T>(Op::ForPack, width, n, in, packed, nullptr, refs, 0, (cudaStream_t)st);                     \
    }                                                                                                                   \
    fl_status fl_host_for_pack_##SFX(unsigned width, size_t n, const T* in, T reference, T* packed) {                   \
        return host_op<T>(Op::ForPack, width, n, in, packed, nullptr, reference);                                       \
    }                                                                                                                   \
    fl_status fl_unfor_pack_##SFX(unsigned width, size_t n, const T* packed, T reference, T* out, void* st) {           \
        return device_op<T>(Op::UnforPack, width, n, packed, out, nullptr, nullptr, reference, (cudaStream_t)st);       \
    }                                                                                                                   \
    fl_status fl_unfor_pack_refs_##SFX(unsigned width, size_t n, const T* packed, const T* refs, T* out, void* st) {    \
        if (!refs) return fail(FL_ERR_NULL, "null refs pointer");                                                       \
        return device_op<T>(Op::UnforPack, width, n, packed, out, nullptr, refs, 0, (cudaStream_t)st);                  \
    }                                                                                                                   \
    fl_status fl_host_unfor_pack_##SFX(unsigned width, size_t n, const T* packed, T reference, T* out) {                \
        return host_op<T>(Op::UnforPack, width, n, packed, out, nullptr, reference);                                    \
    }                                                                                                                   \
    fl_status fl_delta_##SFX(size_t n, const T* in, const T* base, T* out, void* st) {                                  \
        return device_op<T>(Op::Delta, 0, n, in, out, base, nullptr, 0, (cudaStream_t)st);                              \
    }                                                                                                                   \
    fl_status fl_host_delta_##SFX(size_t n, const T* in, const T* base, T* out) {                                       \
        return host_op<T>(Op::Delta, 0, n, in, out, base, 0);                                                           \
    }                                                                                                                   \
    fl_status fl_undelta_##SFX(size_t n, const T* in, const T* base, T* out, void* st) {                                \
        return device_op<T>(Op::Undelta, 0, n, in, out, base, nullptr, 0, (cudaStream_t)st);                            \
    }                                                                                                                   \
    fl_status fl_host_undelta_##SFX(size_t n, const T* in, const T* base, T* out) {                                     \
        return host_op<T>(Op::Undelta, 0, n, in, out, base, 0);                                                         \
    }                                                                                                                   \
    fl_status fl_undelta_pack_##SFX(unsigned width, size_t n, const T* packed, const T* base, T* out, void* st) {       \
        return device_op<T>(Op::UndeltaPack, width, n, packed, out, base, nullptr, 0, (cudaStream_t)st);                \
    }                                                                                                                   \
    fl_status fl_host_undelta_pack_##SFX(unsigned width, size_t n, const T* packed, const T* base, T* out) {            \
        return host_op<T>(Op::UndeltaPack, width, n, packed, out, base, 0);                                             \
    }                                                                                                                   \
    fl_status fl_undelta_pack_untranspose_##SFX(unsigned width, size_t n, const T* packed, const T* base, T* out,       \
                                                void* st) {                                                            \
        return device_op<T>(Op::UndeltaPackUntranspose, width, n, packed, out, base, nullptr, 0, (cudaStream_t)st);     \
    }                                                                                                                   \
    fl_status fl_host_undelta_pack_untranspose_##SFX(unsigned width, size_t n, const T* packed, const T* base, T* out) { \
        return host_op<T>(Op::UndeltaPackUntranspose, width, n, packed, out, base, 0);                                  \
    }                                                                                                                   \
    fl_status fl_transpose_delta_pack_##SFX(unsigned width, size_t n, const T* in, const T* base, T* packed, void* st) { \
        return device_op<T>(Op::TransposeDeltaPack, width, n, in, packed, base, nullptr, 0, (cudaStream_t)st);          \
    }                                                                                                                   \
    fl_status fl_host_transpose_delta_pack_##SFX(unsigned width, size_t n, const T* in, const T* base, T* packed) {     \
        return host_op<T>(Op::TransposeDeltaPack, width, n, in, packed, base, 0);                                       \
    }                                                                                                                   \
    fl_status fl_block_minmax_##SFX(size_t n, const T* in, T* mins, T* maxs, void* st) {                                \
        return device_minmax<T>(n, in, mins, maxs, (cudaStream_t)st);                                                   \
    }                                                                                                                   \
    fl_status fl_host_block_minmax_##SFX(size_t n, const T* in, T* mins, T* maxs) {                                     \
        return host_minmax<T>(n, in, mins, maxs);                                                                       \
    }                                                                                                                   \
    fl_status fl_pack_cwida_##SFX(unsigned width, size_t n, const T* in, T* packed, void* st) {                         \
        return device_op<T>(Op::PackLinear, width, n, in, packed, nullptr, nullptr, 0, (cudaStream_t)st);               \
    }                                                                                                                   \
    fl_status fl_unpack_cwida_##SFX(unsigned width, size_t n, const T* packed, T* out, void* st) {                      \
        return device_op<T>(Op::UnpackLinear, width, n, packed, out, nullptr, nullptr, 0, (cudaStream_t)st);            \
    }                                                                                                                   \
    fl_status fl_for_pack_cwida_##SFX(unsigned width, size_t n, const T* in, T reference, T* packed, void* st) {        \
        return device_op<T>(Op::ForPackLinear, width, n, in, packed, nullptr, nullptr, reference, (cudaStream_t)st);    \
    }                                                                                                                   \
    fl_status fl_unfor_pack_cwida_##SFX(unsigned width, size_t n, const T* packed, T reference, T* out, void* st) {     \
        return device_op<T>(Op::UnforPackLinear, width, n, packed, out, nullptr, nullptr, reference, (cudaStream_t)st); \
    }                                                                                                                   \
    fl_status fl_for_pack_auto_##SFX(unsigned width, size_t n, const T* in, T* refs_out, T* spans_out, T* packed,       \
                                     void* st) {                                                                       \
        return device_for_pack_auto<T>(width, n, in, refs_out, spans_out, packed, (cudaStream_t)st);                    \
    }                                                                                                                   \
    fl_status fl_unpack_filter_##SFX(unsigned width, size_t n, const T* packed, const T* refs, T reference, T lo, T hi, \
                                     uint8_t* bitmap, uint32_t* counts, void* st) {                                    \
        return device_filter<T>(width, n, packed, refs, reference, lo, hi, bitmap, counts, (cudaStream_t)st);           \
    }                                                                                                                   \
    fl_status fl_host_unpack_filter_##SFX(unsigned width, size_t n, const T* packed, T reference, T lo, T hi,           \
                                          uint8_t* bitmap, uint32_t* counts) {                                          \
        return host_filter<T>(width, n, packed, nullptr, reference, lo, hi, bitmap, counts);                            \
    }                                                                                                                   \
    fl_status fl_undelta_pack_filter_##SFX(unsigned width, size_t n, const T* packed, const T* base, T lo, T hi,        \
                                           uint8_t* bitmap, uint32_t* counts, void* st) {                               \
        return device_delta_filter<T>(width, n, packed, base, lo, hi, bitmap, counts, (cudaStream_t)st);                \
    }                                                                                                                   \
    fl_status fl_host_undelta_pack_filter_##SFX(unsigned width, size_t n, const T* packed, const T* base, T lo, T hi,   \
                                                uint8_t* bitmap, uint32_t* counts) {                                    \
        if (!base) return fail(FL_ERR_NULL, "null base pointer");                                                       \
        return host_filter<T>(width, n, packed, base, 0, lo, hi, bitmap, counts);                                       \
    }                                                                                                                   \
    fl_status fl_unpack_select_##SFX(unsigned width, size_t n, const T* packed, const T* refs, T reference,             \
                                     const uint8_t* bitmap, const uint64_t* offsets, T* out, void* st) {                \
        return device_select<T>(width, n, packed, refs, reference, bitmap, offsets, out, (cudaStream_t)st);             \
    }                                                                                                                   \
    fl_status fl_transpose_##SFX(size_t n, const T* in, T* out, void* st) {                                             \
        return device_op<T>(Op::Transpose, 0, n, in, out, nullptr, nullptr, 0, (cudaStream_t)st);                       \
    }                                                                                                                   \
    fl_status fl_untranspose_##SFX(size_t n, const T* in, T* out, void* st) {                                           \
        return device_op<T>(Op::Untranspose, 0, n, in, out, nullptr, nullptr, 0, (cudaStream_t)st);                     \
    }                                                                                                                   \
    fl_status fl_host_transpose_##SFX(size_t n, const T* in, T* out) {                                                  \
        return host_op<T>(Op::Transpose, 0, n, in, out, nullptr, 0);                                                    \
    }                                                                                                                   \
    fl_status fl_host_untranspose_##SFX(size_t n, const T* in, T* out) {                                                \
        return host_op<T>(Op::Untranspose, 0, n, in, out, nullptr, 0);                                                  \
    }

FL_DEFINE_TYPE(u8, uint8_t)
FL_DEFINE_TYPE(u16, uint16_t)
FL_DEFINE_TYPE(u32, uint32_t)
FL_DEFINE_TYPE(u64, uint64_t)

}  // extern "C"
