// fl_api.cu — the extern "C" boundary declared in include/fastlanes_b200.h.
//
// Device family:  argument checks + one kernel launch on the caller's stream.
// Host family:    chunked H2D -> kernel -> D2H pipeline over internal streams (per-device lanes), a zero-copy
//                 low-latency path for the reference's single-block trait calls, NUMA-placed page-locked buffers.
// Context family: fl_ctx — one process, several devices, contiguous block shards, one worker thread per device.
// There is no CPU compute path anywhere in this library: every value is produced by a CUDA kernel.
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cctype>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
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

// =====================================================================================================
// Host family.  Everything below moves bytes and launches the kernels above; no value is computed on the CPU.
// =====================================================================================================

// ---- NUMA placement of page-locked host memory ---------------------------------------------------
// The host family is PCIe-bound, and a GPU's DMA engine reaches only the memory of its own socket at full speed: with
// eight GPUs on a two-socket box, buffers that all sit on one node push half of the traffic over the socket
// interconnect (round 1: 0.18 end-to-end scaling efficiency at 8 GPUs, SCALE_r01.json).  So pinned buffers are placed
// on the node of the device that will copy them.  Node lookup, first hit wins:
//   FLB_NUMA_MAP="n0,n1,..."   explicit node per device ordinal (escape hatch)
//   cudaDevAttrHostNumaId       the driver's answer
//   /sys/bus/pci/devices/<id>/numa_node
// Placement is by mbind(MPOL_PREFERRED) on an anonymous mapping that is then cudaHostRegister-ed (pages are faulted in
// under the policy while being pinned); a cpuset-restricted launcher (CPUs 0-31 only) cannot defeat it the way it
// defeats first-touch.  Any failure falls back to cudaHostAlloc.  FLB_NUMA=0 disables all of it.
inline long sys_mbind(void* p, size_t len, int mode, const unsigned long* mask, unsigned long maxnode, unsigned flags) {
    return syscall(SYS_mbind, p, len, mode, mask, maxnode, flags);
}
constexpr int kMpolPreferred = 1;

bool numa_enabled() {
    static const bool on = [] {
        const char* e = std::getenv("FLB_NUMA");
        return !(e && e[0] == '0');
    }();
    return on;
}
int n_memory_nodes() {
    static const int n = [] {
        int k = 0;
        for (; k < 64; ++k) {
            const std::string path = "/sys/devices/system/node/node" + std::to_string(k);
            if (access(path.c_str(), F_OK) != 0) break;
        }
        return k;
    }();
    return n;
}
int device_numa_node(int dev) {
    if (!numa_enabled()) return -1;
    if (const char* m = std::getenv("FLB_NUMA_MAP")) {
        int i = 0;
        for (const char* p = m; *p; ++i) {
            char* end = nullptr;
            const long v = std::strtol(p, &end, 10);
            if (end == p) break;
            if (i == dev) return int(v);
            p = (*end == ',') ? end + 1 : end;
        }
    }
    int node = -1;
    if (cudaDeviceGetAttribute(&node, cudaDevAttrHostNumaId, dev) == cudaSuccess && node >= 0) return node;
    (void)cudaGetLastError();
    char id[32] = {0};
    if (cudaDeviceGetPCIBusId(id, int(sizeof(id)), dev) != cudaSuccess) { (void)cudaGetLastError(); return -1; }
    for (char* c = id; *c; ++c) *c = char(std::tolower(static_cast<unsigned char>(*c)));
    const std::string path = std::string("/sys/bus/pci/devices/") + id + "/numa_node";
    FILE* f = std::fopen(path.c_str(), "r");
    if (!f) return -1;
    node = -1;
    if (std::fscanf(f, "%d", &node) != 1) node = -1;
    std::fclose(f);
    return node;
}

// registry of mmap + cudaHostRegister allocations (fl_host_free must undo exactly what fl_host_alloc did)
struct MappedAlloc { size_t bytes; };
std::mutex g_alloc_mu;
std::unordered_map<void*, MappedAlloc> g_allocs;

struct NodeRange { size_t begin, end; int node; };  // byte range [begin, end) -> preferred node (-1: leave default)

// Page-locked allocation whose byte ranges prefer the given nodes.  Returns false when placement is impossible here
// (no NUMA, mmap/mbind/register refused): the caller then uses cudaHostAlloc.
bool alloc_placed(void** out, size_t bytes, const std::vector<NodeRange>& ranges) {
    if (!numa_enabled() || n_memory_nodes() < 2 || bytes == 0) return false;
    bool any = false;
    for (const NodeRange& r : ranges) any = any || r.node >= 0;
    if (!any) return false;
    const size_t page = size_t(sysconf(_SC_PAGESIZE));
    const size_t len = (bytes + page - 1) / page * page;
    void* p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) return false;
    (void)madvise(p, len, MADV_HUGEPAGE);
    bool bound = false;
    for (const NodeRange& r : ranges) {
        if (r.node < 0 || r.node >= 64) continue;
        const size_t b = r.begin / page * page;                       // ranges are contiguous: a shared page goes to the
        const size_t e = std::min(len, (r.end + page - 1) / page * page);  // later range, which rebinds it
        if (e <= b) continue;
        const unsigned long mask = 1ul << r.node;
        if (sys_mbind(static_cast<char*>(p) + b, e - b, kMpolPreferred, &mask, 64, 0) == 0) bound = true;
    }
    if (!bound) { munmap(p, len); return false; }  // mbind refused (seccomp / no permission): nothing gained
    if (cudaHostRegister(p, len, cudaHostRegisterPortable | cudaHostRegisterMapped) != cudaSuccess) {
        (void)cudaGetLastError();
        munmap(p, len);
        return false;
    }
    {
        std::lock_guard<std::mutex> lk(g_alloc_mu);
        g_allocs[p] = MappedAlloc{len};
    }
    *out = p;
    return true;
}

// ---- per-device pipeline state ---------------------------------------------------------------------
// A LANE is one independent host pipeline: n_streams slots (stream + device staging) and one small page-locked buffer for
// the low-latency path.  A device owns up to kMaxLanes lanes; a host call takes a free one, so concurrent callers on one
// device do not serialise behind a device-wide mutex (round 1) unless all lanes are busy.
struct Slot {
    cudaStream_t stream = nullptr;
    void* d_in = nullptr;
    void* d_out = nullptr;
    void* d_base = nullptr;
    size_t in_cap = 0, out_cap = 0, base_cap = 0;
};
#ifndef FLB_DIRECT_MAX_DEFAULT
#define FLB_DIRECT_MAX_DEFAULT (size_t(512) << 20)
#endif
constexpr size_t kSmallBytes = size_t(256) << 10;  // low-latency path: in + base + out of the whole call fit in this
struct Lane {
    std::mutex mu;
    std::vector<Slot> slots;
    char* h_small = nullptr;  // page-locked, device-accessible (UVA): kSmallBytes
};
constexpr size_t kMaxLanes = 4;
struct HostCtx {
    int device = -1;
    std::mutex mu;  // guards `lanes`
    std::vector<std::unique_ptr<Lane>> lanes;
};

std::mutex g_ctx_mu;
std::vector<HostCtx*> g_ctxs;
std::atomic<size_t> g_chunk_blocks{16384};  // process-wide (fl_host_configure); may race with fl_host_* calls on other threads
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

// RAII: a locked lane of the current device
struct LaneLock {
    Lane* lane = nullptr;
    ~LaneLock() { if (lane) lane->mu.unlock(); }
    LaneLock() = default;
    LaneLock(const LaneLock&) = delete;
    LaneLock& operator=(const LaneLock&) = delete;
};
fl_status acquire_lane(LaneLock* out) {
    HostCtx* ctx = nullptr;
    if (fl_status s = get_ctx(&ctx)) return s;
    Lane* busy = nullptr;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        for (auto& l : ctx->lanes)
            if (l->mu.try_lock()) { out->lane = l.get(); return FL_OK; }
        if (ctx->lanes.size() < kMaxLanes) {
            ctx->lanes.emplace_back(new Lane);
            ctx->lanes.back()->mu.lock();
            out->lane = ctx->lanes.back().get();
            return FL_OK;
        }
        static std::atomic<unsigned> rr{0};
        busy = ctx->lanes[rr.fetch_add(1) % ctx->lanes.size()].get();
    }
    busy->mu.lock();  // all lanes busy: wait for one (outside the context mutex)
    out->lane = busy;
    return FL_OK;
}

fl_status ensure(void** p, size_t* cap, size_t need) {
    if (*cap >= need) return FL_OK;
    void* old = *p;
    *p = nullptr; *cap = 0;  // never leave a dangling pointer behind, whatever cudaFree says
    if (old) FL_CUDA(cudaFree(old));
    FL_CUDA(cudaMalloc(p, need));
    *cap = need;
    return FL_OK;
}
fl_status ensure_stream(Slot& sl) {
    if (!sl.stream) FL_CUDA(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
    return FL_OK;
}
fl_status ensure_small(Lane& lane) {
    if (!lane.h_small) FL_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&lane.h_small), kSmallBytes, cudaHostAllocDefault));
    return FL_OK;
}

struct PipeCfg { size_t chunk, n_slots; };
inline PipeCfg pipe_cfg() {
    const size_t c = g_chunk_blocks.load();
    const int s = g_n_streams.load();
    return PipeCfg{c ? c : 16384, size_t(s > 0 ? s : 3)};
}

// The chunked pipeline shared by every bulk host call: chunk c of `chunk` blocks runs on slot c % n_slots — H2D copies,
// kernel(s), D2H copies, all enqueued by `body(slot, first_block, n_blocks_in_chunk)` on the slot's stream, whose order
// also protects the slot's staging buffers from the next chunk that reuses them.  `prepare(slot, max_blocks_per_chunk)`
// sizes the staging buffers.  Every stream is drained before returning, error or not: no copy touching the caller's
// buffers may outlive the call.
template <class Prepare, class Body>
fl_status run_pipeline(Lane& lane, size_t n_blocks, Prepare&& prepare, Body&& body) {
    const PipeCfg cfg = pipe_cfg();
    if (lane.slots.size() < cfg.n_slots) lane.slots.resize(cfg.n_slots);
    const size_t n_chunks = (n_blocks + cfg.chunk - 1) / cfg.chunk;
    const size_t use_slots = std::min(n_chunks, cfg.n_slots);
    const size_t cb = std::min(n_blocks, cfg.chunk);
    for (size_t s = 0; s < use_slots; ++s) {
        if (fl_status st = ensure_stream(lane.slots[s])) return st;
        if (fl_status st = prepare(lane.slots[s], cb)) return st;
    }
    fl_status result = FL_OK;
    for (size_t c = 0; c < n_chunks && result == FL_OK; ++c) {
        const size_t b0 = c * cfg.chunk;
        result = body(lane.slots[c % use_slots], b0, std::min(cfg.chunk, n_blocks - b0));
    }
    for (size_t s = 0; s < use_slots; ++s) {
        const cudaError_t e = cudaStreamSynchronize(lane.slots[s].stream);
        if (e != cudaSuccess && result == FL_OK) result = cuda_fail(e, "cudaStreamSynchronize");
    }
    return result;
}

inline fl_status copy_async(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t st, const char* what) {
    if (bytes == 0) return FL_OK;
    const cudaError_t e = cudaMemcpyAsync(dst, src, bytes, kind, st);
    return e == cudaSuccess ? FL_OK : cuda_fail(e, what);
}

// FLB_SMALL=0 disables the low-latency path (A/B measurement, tools/refbench.py); 1 = bounce buffer only for pageable
// caller memory is not attempted: the buffer is always used (default).
inline bool small_path_enabled() {
    static const bool on = [] {
        const char* e = std::getenv("FLB_SMALL");
        return !(e && e[0] == '0');
    }();
    return on;
}

// FLB_DIRECT_MAX=<bytes>: largest call (input + bases + output) that takes the direct path below; 0 disables it.
inline size_t direct_max_bytes() {
    static const size_t v = [] {
        const char* e = std::getenv("FLB_DIRECT_MAX");
        return e ? size_t(std::strtoull(e, nullptr, 10)) : size_t(FLB_DIRECT_MAX_DEFAULT);
    }();
    return v;
}
// Device-visible address of a PAGE-LOCKED host pointer (cudaHostAlloc / cudaHostRegister / fl_host_alloc), or false for
// pageable memory.  ~0.3 us per call.
inline bool pinned_device_ptr(const void* host, const void** dev) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, host) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    if (at.type != cudaMemoryTypeHost || at.devicePointer == nullptr) return false;
    *dev = at.devicePointer;
    return true;
}

template <class T>
fl_status host_op(Op op, unsigned width, size_t n_blocks, const void* in, void* out, const void* base,
                  uint64_t ref_scalar) {
    constexpr unsigned TB = sizeof(T) * 8;
    if (op_has_width(op) && width > TB) return fail(FL_ERR_WIDTH, "width exceeds the bit size of the element type");
    if (!op_has_width(op)) width = 0;
    if (n_blocks == 0) return FL_OK;
    const size_t ib = in_block_bytes(op, TB, width), ob = out_block_bytes(op, TB, width);
    const size_t bb = op_has_base(op) ? 128 : 0;
    if ((ib && !in) || (ob && !out)) return fail(FL_ERR_NULL, "null data pointer");
    if (bb && !base) return fail(FL_ERR_NULL, "null base pointer");
    if (ob == 0) return FL_OK;

    LaneLock lk;
    if (fl_status s = acquire_lane(&lk)) return s;
    Lane& lane = *lk.lane;

    // ---- low-latency path: the reference's single-block trait call (BitPacking::unpack(&packed, &mut out)) ----------
    // Round 1 paid H2D copy + kernel + D2H copy + sync on three streams: 21.6 us for one u16 block against 4.3 us on the
    // CPU (profiles/refbench_r01.txt).  Here the caller's bytes are memcpy-ed into a persistent page-locked buffer that
    // the GPU addresses directly (UVA zero-copy): ONE kernel launch reads its input from and writes its output to host
    // memory over PCIe, one stream synchronise, memcpy out.  No copy-engine operations, no allocation.
    const size_t a256 = 255;
    const size_t in_sz = (n_blocks * ib + a256) & ~a256, base_sz = (n_blocks * bb + a256) & ~a256, out_sz = n_blocks * ob;
    if (small_path_enabled() && in_sz + base_sz + out_sz <= kSmallBytes) {
        if (lane.slots.empty()) lane.slots.resize(1);
        if (fl_status st = ensure_stream(lane.slots[0])) return st;
        if (fl_status st = ensure_small(lane)) return st;
        char* h_in = lane.h_small;
        char* h_base = h_in + in_sz;
        char* h_out = h_base + base_sz;
        if (ib) std::memcpy(h_in, in, n_blocks * ib);
        if (bb) std::memcpy(h_base, base, n_blocks * bb);
        fl_status result = device_op<T>(op, width, n_blocks, h_in, h_out, h_base, nullptr, ref_scalar, lane.slots[0].stream);
        const cudaError_t e = cudaStreamSynchronize(lane.slots[0].stream);
        if (e != cudaSuccess && result == FL_OK) result = cuda_fail(e, "cudaStreamSynchronize");
        if (result == FL_OK) std::memcpy(out, h_out, out_sz);
        return result;
    }

    // ---- direct path: mid-size calls on page-locked caller buffers ----------------------------------------------------
    // The reference's own throughput bench shape (benches/bitpacking.rs:67-98: 1024 blocks = 2 MiB) is one pipeline chunk:
    // H2D copy, kernel, D2H copy run back to back on one stream, three operations and two trips through HBM for a transfer
    // the link finishes in ~40 us.  When every buffer of the call is page-locked the kernel addresses the caller's memory
    // itself (UVA zero-copy, as the low-latency path does through its bounce buffer): ONE launch streams the packed words in
    // and the decoded values out over PCIe concurrently, one synchronise.  Measured (tools/midbench.cpp,
    // profiles/midbench_r02.txt): 1024 blocks u16 W=3 69.7 -> 55.2 us, 64 blocks u32 W=10 32.7 -> 18.7 us, 65536 blocks u32
    // (335 MB moved) 6.3-6.8 -> 5.6 ms.  SM-issued PCIe writes top out near 48 GB/s where the copy engines reach 52-55, so
    // above FLB_DIRECT_MAX (default 512 MiB per call) the chunked copy-engine pipeline, which needs many chunks in flight
    // to reach that ceiling, takes over.
    if (n_blocks * (ib + bb + ob) <= direct_max_bytes() && n_blocks <= kMaxBlocksPerLaunch) {
        const void *d_in = nullptr, *d_out = nullptr, *d_base = nullptr;
        const bool aligned = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(base)) & 15u) == 0;
        if (aligned && (!ib || pinned_device_ptr(in, &d_in)) && pinned_device_ptr(out, &d_out) && (!bb || pinned_device_ptr(base, &d_base))) {
            if (lane.slots.empty()) lane.slots.resize(1);
            if (fl_status st = ensure_stream(lane.slots[0])) return st;
            fl_status result = device_op<T>(op, width, n_blocks, d_in, const_cast<void*>(d_out), d_base, nullptr, ref_scalar, lane.slots[0].stream);
            const cudaError_t e = cudaStreamSynchronize(lane.slots[0].stream);
            if (e != cudaSuccess && result == FL_OK) result = cuda_fail(e, "cudaStreamSynchronize");
            return result;
        }
    }

    return run_pipeline(
        lane, n_blocks,
        [&](Slot& sl, size_t cb) -> fl_status {
            if (ib) if (fl_status st = ensure(&sl.d_in, &sl.in_cap, cb * ib)) return st;
            if (fl_status st = ensure(&sl.d_out, &sl.out_cap, cb * ob)) return st;
            if (bb) if (fl_status st = ensure(&sl.d_base, &sl.base_cap, cb * bb)) return st;
            return FL_OK;
        },
        [&](Slot& sl, size_t b0, size_t nb) -> fl_status {
            if (fl_status st = copy_async(sl.d_in, static_cast<const char*>(in) + b0 * ib, nb * ib, cudaMemcpyHostToDevice, sl.stream, "cudaMemcpyAsync H2D")) return st;
            if (bb) if (fl_status st = copy_async(sl.d_base, static_cast<const char*>(base) + b0 * bb, nb * bb, cudaMemcpyHostToDevice, sl.stream, "cudaMemcpyAsync H2D")) return st;
            if (fl_status st = device_op<T>(op, width, nb, sl.d_in, sl.d_out, sl.d_base, nullptr, ref_scalar, sl.stream)) return st;
            return copy_async(static_cast<char*>(out) + b0 * ob, sl.d_out, nb * ob, cudaMemcpyDeviceToHost, sl.stream, "cudaMemcpyAsync D2H");
        });
}

// The copies of the host pipeline WITHOUT the kernel: in_bytes_per_block go H2D, out_bytes_per_block come back D2H (from
// whatever the staging buffer holds), chunked and overlapped exactly like host_op.  This is the PCIe ceiling of a host
// call with those byte counts on this box — bench.py reports e2e as a fraction of it (e2e.link_ceiling).
fl_status host_copy_probe(size_t in_bytes_per_block, size_t out_bytes_per_block, size_t n_blocks, const void* in, void* out) {
    if (n_blocks == 0) return FL_OK;
    if ((in_bytes_per_block && !in) || (out_bytes_per_block && !out)) return fail(FL_ERR_NULL, "null data pointer");
    LaneLock lk;
    if (fl_status s = acquire_lane(&lk)) return s;
    const size_t ib = in_bytes_per_block, ob = out_bytes_per_block;
    return run_pipeline(
        *lk.lane, n_blocks,
        [&](Slot& sl, size_t cb) -> fl_status {
            if (ib) if (fl_status st = ensure(&sl.d_in, &sl.in_cap, cb * ib)) return st;
            if (ob) if (fl_status st = ensure(&sl.d_out, &sl.out_cap, cb * ob)) return st;
            return FL_OK;
        },
        [&](Slot& sl, size_t b0, size_t nb) -> fl_status {
            if (fl_status st = copy_async(sl.d_in, static_cast<const char*>(in) + b0 * ib, nb * ib, cudaMemcpyHostToDevice, sl.stream, "cudaMemcpyAsync H2D")) return st;
            return copy_async(static_cast<char*>(out) + b0 * ob, sl.d_out, nb * ob, cudaMemcpyDeviceToHost, sl.stream, "cudaMemcpyAsync D2H");
        });
}

template <class T>
fl_status host_minmax(size_t n_blocks, const T* in, T* mins, T* maxs) {
    if (n_blocks == 0) return FL_OK;
    if (!in || !mins || !maxs) return fail(FL_ERR_NULL, "null pointer");
    LaneLock lk;
    if (fl_status s = acquire_lane(&lk)) return s;
    const size_t cb_max = std::min(n_blocks, pipe_cfg().chunk);
    const size_t half = (cb_max * sizeof(T) + 15) & ~size_t(15);
    return run_pipeline(
        *lk.lane, n_blocks,
        [&](Slot& sl, size_t cb) -> fl_status {
            if (fl_status st = ensure(&sl.d_in, &sl.in_cap, cb * 1024 * sizeof(T))) return st;
            return ensure(&sl.d_out, &sl.out_cap, 2 * half);
        },
        [&](Slot& sl, size_t b0, size_t nb) -> fl_status {
            T* d_min = static_cast<T*>(sl.d_out);
            T* d_max = reinterpret_cast<T*>(static_cast<char*>(sl.d_out) + half);
            if (fl_status st = copy_async(sl.d_in, in + b0 * 1024, nb * 1024 * sizeof(T), cudaMemcpyHostToDevice, sl.stream, "cudaMemcpyAsync H2D")) return st;
            if (fl_status st = device_minmax<T>(nb, static_cast<const T*>(sl.d_in), d_min, d_max, sl.stream)) return st;
            if (fl_status st = copy_async(mins + b0, d_min, nb * sizeof(T), cudaMemcpyDeviceToHost, sl.stream, "cudaMemcpyAsync D2H")) return st;
            return copy_async(maxs + b0, d_max, nb * sizeof(T), cudaMemcpyDeviceToHost, sl.stream, "cudaMemcpyAsync D2H");
        });
}

// host buffers: H2D of the packed chunk, filter kernel, D2H of 128 (+4) bytes per block — the decoded values never
// cross the PCIe link.  `base` != nullptr selects the delta scan (reference unused).
template <class T>
fl_status host_filter(unsigned width, size_t n_blocks, const T* packed, const T* base, T reference, T lo, T hi,
                      uint8_t* bitmap, uint32_t* counts) {
    if (width > sizeof(T) * 8) return fail(FL_ERR_WIDTH, "width exceeds the bit size of the element type");
    if (n_blocks == 0) return FL_OK;
    if (!bitmap || (width && !packed)) return fail(FL_ERR_NULL, "null pointer");
    LaneLock lk;
    if (fl_status s = acquire_lane(&lk)) return s;
    const size_t ib = size_t(128) * width;
    // direct path (see host_op): every buffer page-locked and the call below FLB_DIRECT_MAX -> one launch on the caller's memory
    if (n_blocks * (ib + 132 + (base ? 128 : 0)) <= direct_max_bytes() && n_blocks <= kMaxBlocksPerLaunch) {
        const void *d_in = nullptr, *d_base = nullptr, *d_bm = nullptr, *d_cnt = nullptr;
        const bool aligned = ((reinterpret_cast<uintptr_t>(packed) | reinterpret_cast<uintptr_t>(base) | reinterpret_cast<uintptr_t>(bitmap)) & 15u) == 0 &&
                             (reinterpret_cast<uintptr_t>(counts) & 3u) == 0;
        if (aligned && (!ib || pinned_device_ptr(packed, &d_in)) && (!base || pinned_device_ptr(base, &d_base)) &&
            pinned_device_ptr(bitmap, &d_bm) && (!counts || pinned_device_ptr(counts, &d_cnt))) {
            Lane& lane = *lk.lane;
            if (lane.slots.empty()) lane.slots.resize(1);
            if (fl_status st = ensure_stream(lane.slots[0])) return st;
            cudaStream_t stream = lane.slots[0].stream;
            uint8_t* bm = static_cast<uint8_t*>(const_cast<void*>(d_bm));
            uint32_t* cn = static_cast<uint32_t*>(const_cast<void*>(d_cnt));
            fl_status result = base ? device_delta_filter<T>(width, n_blocks, static_cast<const T*>(d_in), static_cast<const T*>(d_base), lo, hi, bm, cn, stream)
                                    : device_filter<T>(width, n_blocks, static_cast<const T*>(d_in), nullptr, reference, lo, hi, bm, cn, stream);
            const cudaError_t e = cudaStreamSynchronize(stream);
            if (e != cudaSuccess && result == FL_OK) result = cuda_fail(e, "cudaStreamSynchronize");
            return result;
        }
    }
    const size_t cb_max = std::min(n_blocks, pipe_cfg().chunk);
    return run_pipeline(
        *lk.lane, n_blocks,
        [&](Slot& sl, size_t cb) -> fl_status {
            if (ib) if (fl_status st = ensure(&sl.d_in, &sl.in_cap, cb * ib)) return st;
            if (fl_status st = ensure(&sl.d_out, &sl.out_cap, cb * 128 + cb * sizeof(uint32_t))) return st;
            if (base) if (fl_status st = ensure(&sl.d_base, &sl.base_cap, cb * 128)) return st;
            return FL_OK;
        },
        [&](Slot& sl, size_t b0, size_t nb) -> fl_status {
            uint8_t* d_bitmap = static_cast<uint8_t*>(sl.d_out);
            uint32_t* d_counts = reinterpret_cast<uint32_t*>(d_bitmap + cb_max * 128);
            if (fl_status st = copy_async(sl.d_in, reinterpret_cast<const char*>(packed) + b0 * ib, nb * ib, cudaMemcpyHostToDevice, sl.stream, "cudaMemcpyAsync H2D")) return st;
            if (base) if (fl_status st = copy_async(sl.d_base, reinterpret_cast<const char*>(base) + b0 * 128, nb * 128, cudaMemcpyHostToDevice, sl.stream, "cudaMemcpyAsync H2D")) return st;
            fl_status st;
            if (base)
                st = device_delta_filter<T>(width, nb, static_cast<const T*>(sl.d_in), static_cast<const T*>(sl.d_base), lo, hi,
                                            d_bitmap, counts ? d_counts : nullptr, sl.stream);
            else
                st = device_filter<T>(width, nb, static_cast<const T*>(sl.d_in), nullptr, reference, lo, hi, d_bitmap,
                                      counts ? d_counts : nullptr, sl.stream);
            if (st != FL_OK) return st;
            if (fl_status s2 = copy_async(bitmap + b0 * 128, d_bitmap, nb * 128, cudaMemcpyDeviceToHost, sl.stream, "cudaMemcpyAsync D2H")) return s2;
            if (counts) return copy_async(counts + b0, d_counts, nb * sizeof(uint32_t), cudaMemcpyDeviceToHost, sl.stream, "cudaMemcpyAsync D2H");
            return FL_OK;
        });
}

// Batched unpack_single on host buffers.  Only the blocks the indices actually reference cross the link: the indices
// are rewritten against a compacted list of distinct blocks (host-side bookkeeping, no decode), so a single lookup
// moves one block instead of the whole column (round 1 copied all n_blocks).  Small requests take the
// zero-copy path (one launch, no copy-engine operations).
template <class T>
fl_status host_gather(unsigned width, size_t n_blocks, const T* packed, const uint64_t* gidx, size_t n, T* out) {
    if (width > sizeof(T) * 8) return fail(FL_ERR_WIDTH, "width exceeds the bit size of the element type");
    if (n == 0) return FL_OK;
    if (!gidx || !out || (width && n_blocks && !packed)) return fail(FL_ERR_NULL, "null pointer");
    const size_t pb = size_t(128) * width;
    const uint64_t limit = uint64_t(n_blocks) * 1024;
    for (size_t i = 0; i < n; ++i)
        if (gidx[i] >= limit) return fail(FL_ERR_INDEX, "index out of range");  // src/bitpacking.rs:152
    // distinct blocks in first-use order; compaction only pays while it moves fewer bytes than the whole column
    const bool compact = n < n_blocks;
    std::vector<uint64_t> blocks;
    std::vector<uint64_t> local;
    if (compact) {
        std::unordered_map<uint64_t, uint64_t> slot_of;
        slot_of.reserve(n * 2);
        local.resize(n);
        for (size_t i = 0; i < n; ++i) {
            const uint64_t b = gidx[i] >> 10;
            auto it = slot_of.find(b);
            if (it == slot_of.end()) { it = slot_of.emplace(b, uint64_t(blocks.size())).first; blocks.push_back(b); }
            local[i] = (it->second << 10) | (gidx[i] & 1023);
        }
    }
    const size_t nb_copy = compact ? blocks.size() : n_blocks;
    const uint64_t* idx_src = compact ? local.data() : gidx;

    LaneLock lk;
    if (fl_status s = acquire_lane(&lk)) return s;
    Lane& lane = *lk.lane;
    if (lane.slots.empty()) lane.slots.resize(1);
    Slot& sl = lane.slots[0];
    if (fl_status st = ensure_stream(sl)) return st;
    const size_t idx_bytes = n * sizeof(uint64_t), val_bytes = (n * sizeof(T) + 15) & ~size_t(15);
    const size_t blk_bytes = (nb_copy * pb + 255) & ~size_t(255);

    if (small_path_enabled() && blk_bytes + idx_bytes + val_bytes + 16 <= kSmallBytes) {
        // zero-copy: [blocks | indices | values | flag] in the lane's page-locked buffer, addressed by the kernel directly
        // (this is also the reference's unpack_single: one block, one index, one launch)
        if (fl_status st = ensure_small(lane)) return st;
        char* h = lane.h_small;
        if (!compact) {
            if (pb) std::memcpy(h, packed, n_blocks * pb);
        } else {
            for (size_t k = 0; k < blocks.size(); ++k)
                if (pb) std::memcpy(h + k * pb, reinterpret_cast<const char*>(packed) + blocks[k] * pb, pb);
        }
        uint64_t* h_idx = reinterpret_cast<uint64_t*>(h + blk_bytes);
        std::memcpy(h_idx, idx_src, idx_bytes);
        T* h_val = reinterpret_cast<T*>(h + blk_bytes + idx_bytes);
        int* h_flag = reinterpret_cast<int*>(h + blk_bytes + idx_bytes + val_bytes);
        *h_flag = 0;
        fl_status result = device_gather<T>(width, nb_copy, reinterpret_cast<const T*>(h), h_idx, n, h_val, h_flag, sl.stream);
        const cudaError_t e = cudaStreamSynchronize(sl.stream);
        if (e != cudaSuccess && result == FL_OK) result = cuda_fail(e, "cudaStreamSynchronize");
        if (result != FL_OK) return result;
        if (*h_flag) return fail(FL_ERR_INDEX, "index out of range");
        std::memcpy(out, h_val, n * sizeof(T));
        return FL_OK;
    }

    // d_in: packed blocks; d_out: [indices | values | oob flag]
    if (blk_bytes) if (fl_status st = ensure(&sl.d_in, &sl.in_cap, blk_bytes)) return st;
    if (fl_status st = ensure(&sl.d_out, &sl.out_cap, idx_bytes + val_bytes + 16)) return st;
    char* d = static_cast<char*>(sl.d_out);
    uint64_t* d_idx = reinterpret_cast<uint64_t*>(d);
    T* d_val = reinterpret_cast<T*>(d + idx_bytes);
    int* d_flag = reinterpret_cast<int*>(d + idx_bytes + val_bytes);
    int flag = 0;
    fl_status result = FL_OK;
    if (pb) {
        if (compact) {  // runs of consecutive distinct blocks go as one copy
            for (size_t k = 0; k < blocks.size() && result == FL_OK;) {
                size_t e = k + 1;
                while (e < blocks.size() && blocks[e] == blocks[e - 1] + 1) ++e;
                result = copy_async(static_cast<char*>(sl.d_in) + k * pb, reinterpret_cast<const char*>(packed) + blocks[k] * pb,
                                    (e - k) * pb, cudaMemcpyHostToDevice, sl.stream, "cudaMemcpyAsync H2D");
                k = e;
            }
        } else {
            result = copy_async(sl.d_in, packed, n_blocks * pb, cudaMemcpyHostToDevice, sl.stream, "cudaMemcpyAsync H2D");
        }
    }
    if (result == FL_OK) result = copy_async(d_idx, idx_src, idx_bytes, cudaMemcpyHostToDevice, sl.stream, "cudaMemcpyAsync H2D");
    if (result == FL_OK) {
        const cudaError_t e = cudaMemsetAsync(d_flag, 0, sizeof(int), sl.stream);
        if (e != cudaSuccess) result = cuda_fail(e, "cudaMemsetAsync");
    }
    if (result == FL_OK)
        result = device_gather<T>(width, nb_copy, static_cast<const T*>(sl.d_in), d_idx, n, d_val, d_flag, sl.stream);
    if (result == FL_OK) result = copy_async(out, d_val, n * sizeof(T), cudaMemcpyDeviceToHost, sl.stream, "cudaMemcpyAsync D2H");
    if (result == FL_OK) result = copy_async(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, sl.stream, "cudaMemcpyAsync D2H");
    const cudaError_t e = cudaStreamSynchronize(sl.stream);  // always drain: `flag`, `local` and the caller's buffers must not be touched later
    if (e != cudaSuccess && result == FL_OK) result = cuda_fail(e, "cudaStreamSynchronize");
    if (result != FL_OK) return result;
    if (flag) return fail(FL_ERR_INDEX, "index out of range");
    return FL_OK;
}

// =====================================================================================================
// Multi-device context (fl_ctx): one process, several GPUs, contiguous block shards — SURVEY.md §8(e) behind the C ABI.
// Blocks are independent (no cross-block state anywhere in the reference), so device i of G gets blocks
// [n*i/G, n*(i+1)/G) of every array of the call (packed, unpacked and base shard on the same block index) and runs the
// ordinary single-device host pipeline on them from its own persistent worker thread: G PCIe links, G copy pipelines,
// no exchange between devices, no collective.
// =====================================================================================================
struct Worker {
    int device = -1;
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<void()> task;
    bool has_task = false, stop = false;
    void loop() {
        (void)cudaSetDevice(device);
        for (;;) {
            std::function<void()> t;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return has_task || stop; });
                if (!has_task && stop) return;
                t = std::move(task);
                has_task = false;
            }
            t();
        }
    }
    void post(std::function<void()> t) {
        {
            std::lock_guard<std::mutex> lk(mu);
            task = std::move(t);
            has_task = true;
        }
        cv.notify_one();
    }
};

}  // namespace

struct fl_ctx {
    std::vector<int> devices;
    std::vector<int> nodes;  // NUMA node of each device (-1 unknown)
    std::vector<std::unique_ptr<Worker>> workers;
    std::mutex call_mu;  // one sharded call at a time per context (each worker holds one task)
};

namespace {

inline void shard_range(size_t n_blocks, size_t i, size_t g, size_t* b0, size_t* b1) {
    // same partition as fastlanes_b200.shard.block_shard: ranks differ by at most one block
    *b0 = size_t((unsigned __int128)n_blocks * i / g);
    *b1 = size_t((unsigned __int128)n_blocks * (i + 1) / g);
}

// Runs fn(shard_index, first_block, n_blocks_of_shard) on every device's worker thread (current device = that device)
// and returns the first failure, with its message copied into the caller's thread-local error string.
template <class Fn>
fl_status ctx_run(fl_ctx* ctx, size_t n_blocks, Fn&& fn) {
    if (!ctx) return fail(FL_ERR_NULL, "null context");
    const size_t g = ctx->devices.size();
    std::lock_guard<std::mutex> call_lk(ctx->call_mu);
    std::vector<fl_status> status(g, FL_OK);
    std::vector<std::string> message(g);
    std::mutex done_mu;
    std::condition_variable done_cv;
    size_t pending = 0;
    for (size_t i = 0; i < g; ++i) {
        size_t b0, b1;
        shard_range(n_blocks, i, g, &b0, &b1);
        if (b1 == b0) continue;
        {
            std::lock_guard<std::mutex> lk(done_mu);
            ++pending;
        }
        ctx->workers[i]->post([&, i, b0, b1] {
            status[i] = fn(i, b0, b1 - b0);
            if (status[i] != FL_OK) message[i] = g_err;  // the worker's thread-local message
            std::lock_guard<std::mutex> lk(done_mu);
            if (--pending == 0) done_cv.notify_one();
        });
    }
    {
        std::unique_lock<std::mutex> lk(done_mu);
        done_cv.wait(lk, [&] { return pending == 0; });
    }
    for (size_t i = 0; i < g; ++i)
        if (status[i] != FL_OK) {
            g_err = "device " + std::to_string(ctx->devices[i]) + ": " + message[i];
            return status[i];
        }
    return FL_OK;
}

template <class T>
fl_status ctx_host_op(fl_ctx* ctx, Op op, unsigned width, size_t n_blocks, const void* in, void* out, const void* base,
                      uint64_t ref_scalar) {
    constexpr unsigned TB = sizeof(T) * 8;
    if (op_has_width(op) && width > TB) return fail(FL_ERR_WIDTH, "width exceeds the bit size of the element type");
    const unsigned w = op_has_width(op) ? width : 0;
    const size_t ib = in_block_bytes(op, TB, w), ob = out_block_bytes(op, TB, w);
    return ctx_run(ctx, n_blocks, [=](size_t, size_t b0, size_t nb) {
        return host_op<T>(op, width, nb, in ? static_cast<const char*>(in) + b0 * ib : nullptr,
                          out ? static_cast<char*>(out) + b0 * ob : nullptr,
                          base ? static_cast<const char*>(base) + b0 * 128 : nullptr, ref_scalar);
    });
}
template <class T>
fl_status ctx_host_filter(fl_ctx* ctx, unsigned width, size_t n_blocks, const T* packed, const T* base, T reference, T lo,
                          T hi, uint8_t* bitmap, uint32_t* counts) {
    if (width > sizeof(T) * 8) return fail(FL_ERR_WIDTH, "width exceeds the bit size of the element type");
    const size_t pe = size_t(1024) * width / (sizeof(T) * 8), le = 1024 / (sizeof(T) * 8);
    return ctx_run(ctx, n_blocks, [=](size_t, size_t b0, size_t nb) {
        return host_filter<T>(width, nb, packed ? packed + b0 * pe : nullptr, base ? base + b0 * le : nullptr, reference, lo, hi,
                              bitmap ? bitmap + b0 * 128 : nullptr, counts ? counts + b0 : nullptr);
    });
}
template <class T>
fl_status ctx_host_minmax(fl_ctx* ctx, size_t n_blocks, const T* in, T* mins, T* maxs) {
    return ctx_run(ctx, n_blocks, [=](size_t, size_t b0, size_t nb) {
        return host_minmax<T>(nb, in ? in + b0 * 1024 : nullptr, mins ? mins + b0 : nullptr, maxs ? maxs + b0 : nullptr);
    });
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
    LaneLock lk;
    if (fl_status s = acquire_lane(&lk)) return s;
    Lane& lane = *lk.lane;
    const PipeCfg cfg = pipe_cfg();
    if (lane.slots.size() < cfg.n_slots) lane.slots.resize(cfg.n_slots);
    for (Slot& sl : lane.slots)
        if (fl_status st = ensure_stream(sl)) return st;
    return ensure_small(lane);
}
fl_status fl_host_configure(size_t chunk_blocks, int n_streams) {
    // a chunk is one kernel launch: keep it inside the launch limit (and the chunk arithmetic inside size_t)
    if (chunk_blocks > kMaxBlocksPerLaunch) chunk_blocks = kMaxBlocksPerLaunch;
    g_chunk_blocks.store(chunk_blocks ? chunk_blocks : 16384);
    g_n_streams.store(n_streams > 0 ? (n_streams > 16 ? 16 : n_streams) : 3);
    return FL_OK;
}
int fl_device_numa_node(int device) { return device_numa_node(device); }
fl_status fl_host_alloc(void** p, size_t bytes) {
    if (!p) return fail(FL_ERR_NULL, "null pointer");
    int dev = -1;
    if (cudaGetDevice(&dev) == cudaSuccess) {
        const int node = device_numa_node(dev);
        if (node >= 0 && alloc_placed(p, bytes, {NodeRange{0, bytes, node}})) return FL_OK;
    }
    (void)cudaGetLastError();
    const cudaError_t err = cudaHostAlloc(p, bytes, cudaHostAllocPortable | cudaHostAllocMapped);
    if (err != cudaSuccess) return cuda_fail(err, "cudaHostAlloc");
    return FL_OK;
}
fl_status fl_host_free(void* p) {
    if (!p) return FL_OK;
    size_t mapped_len = 0;
    {
        std::lock_guard<std::mutex> lk(g_alloc_mu);
        auto it = g_allocs.find(p);
        if (it != g_allocs.end()) { mapped_len = it->second.bytes; g_allocs.erase(it); }
    }
    if (mapped_len) {
        const cudaError_t e = cudaHostUnregister(p);
        munmap(p, mapped_len);
        if (e != cudaSuccess) return cuda_fail(e, "cudaHostUnregister");
        return FL_OK;
    }
    FL_CUDA(cudaFreeHost(p));
    return FL_OK;
}
int fl_host_buffer_node(const void* p) {
    // NUMA node of the page holding *p (move_pages query), or -1: lets a caller / the bench verify the placement
    void* page = const_cast<void*>(p);
    int status = -1;
    if (syscall(SYS_move_pages, 0, 1ul, &page, nullptr, &status, 0) != 0) return -1;
    return status;
}
fl_status fl_host_register(void* p, size_t bytes) {
    if (!p) return fail(FL_ERR_NULL, "null pointer");
    FL_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    return FL_OK;
}
fl_status fl_host_unregister(void* p) {
    if (!p) return fail(FL_ERR_NULL, "null pointer");
    FL_CUDA(cudaHostUnregister(p));
    return FL_OK;
}
fl_status fl_host_copy_probe(size_t in_bytes_per_block, size_t out_bytes_per_block, size_t n_blocks, const void* in, void* out) {
    return host_copy_probe(in_bytes_per_block, out_bytes_per_block, n_blocks, in, out);
}
fl_status fl_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    int cur = -1;
    (void)cudaGetDevice(&cur);
    for (HostCtx* c : g_ctxs) {
        std::lock_guard<std::mutex> lk2(c->mu);
        if (cudaSetDevice(c->device) != cudaSuccess) continue;
        for (auto& lane : c->lanes) {
            std::lock_guard<std::mutex> lk3(lane->mu);
            for (Slot& s : lane->slots) {
                if (s.stream) { cudaStreamSynchronize(s.stream); cudaStreamDestroy(s.stream); }
                if (s.d_in) cudaFree(s.d_in);
                if (s.d_out) cudaFree(s.d_out);
                if (s.d_base) cudaFree(s.d_base);
                s = Slot{};
            }
            lane->slots.clear();
            if (lane->h_small) { cudaFreeHost(lane->h_small); lane->h_small = nullptr; }
        }
    }
    if (cur >= 0) (void)cudaSetDevice(cur);
    (void)cudaGetLastError();
    return FL_OK;
}

/* ---- multi-device context ---------------------------------------------------------------------- */
fl_status fl_ctx_create(const int* devices, int n_devices, fl_ctx** out) {
    if (!out) return fail(FL_ERR_NULL, "null pointer");
    *out = nullptr;
    int visible = 0;
    FL_CUDA(cudaGetDeviceCount(&visible));
    std::vector<int> devs;
    if (!devices || n_devices <= 0) {
        for (int d = 0; d < visible; ++d) devs.push_back(d);  // all visible devices
    } else {
        for (int i = 0; i < n_devices; ++i) {
            if (devices[i] < 0 || devices[i] >= visible) return fail(FL_ERR_CUDA, "fl_ctx_create: no such device");
            devs.push_back(devices[i]);  // a device may be listed more than once: each entry is an independent shard worker
        }
    }
    if (devs.empty()) return fail(FL_ERR_CUDA, "fl_ctx_create: no CUDA device");
    int cur = -1;
    (void)cudaGetDevice(&cur);
    fl_ctx* ctx = new fl_ctx;
    ctx->devices = devs;
    for (int d : devs) {
        const cudaError_t e = cudaSetDevice(d);
        if (e == cudaSuccess) (void)cudaFree(nullptr);
        ctx->nodes.push_back(device_numa_node(d));
    }
    // peer access for fl_ctx_scatter_blocks / fl_ctx_gather_blocks (best effort: the copies also work without it)
    for (int a : devs)
        for (int b : devs) {
            int can = 0;
            if (a != b && cudaDeviceCanAccessPeer(&can, a, b) == cudaSuccess && can && cudaSetDevice(a) == cudaSuccess)
                (void)cudaDeviceEnablePeerAccess(b, 0);
        }
    (void)cudaGetLastError();
    if (cur >= 0) (void)cudaSetDevice(cur);
    for (int d : devs) {
        ctx->workers.emplace_back(new Worker);
        Worker* w = ctx->workers.back().get();
        w->device = d;
        w->th = std::thread([w] { w->loop(); });
    }
    *out = ctx;
    return FL_OK;
}
fl_status fl_ctx_destroy(fl_ctx* ctx) {
    if (!ctx) return FL_OK;
    for (auto& w : ctx->workers) {
        {
            std::lock_guard<std::mutex> lk(w->mu);
            w->stop = true;
        }
        w->cv.notify_one();
        if (w->th.joinable()) w->th.join();
    }
    delete ctx;
    return FL_OK;
}
int fl_ctx_device_count(const fl_ctx* ctx) { return ctx ? int(ctx->devices.size()) : 0; }
int fl_ctx_device(const fl_ctx* ctx, int i) {
    return (ctx && i >= 0 && size_t(i) < ctx->devices.size()) ? ctx->devices[size_t(i)] : -1;
}
fl_status fl_ctx_block_range(const fl_ctx* ctx, size_t n_blocks, int i, size_t* first, size_t* end) {
    if (!ctx || !first || !end) return fail(FL_ERR_NULL, "null pointer");
    if (i < 0 || size_t(i) >= ctx->devices.size()) return fail(FL_ERR_INDEX, "shard index out of range");
    shard_range(n_blocks, size_t(i), ctx->devices.size(), first, end);
    return FL_OK;
}
fl_status fl_ctx_host_alloc(fl_ctx* ctx, size_t n_blocks, size_t bytes_per_block, void** p) {
    if (!ctx || !p) return fail(FL_ERR_NULL, "null pointer");
    const size_t bytes = n_blocks * bytes_per_block;
    std::vector<NodeRange> ranges;
    for (size_t i = 0; i < ctx->devices.size(); ++i) {
        size_t b0, b1;
        shard_range(n_blocks, i, ctx->devices.size(), &b0, &b1);
        ranges.push_back(NodeRange{b0 * bytes_per_block, b1 * bytes_per_block, ctx->nodes[i]});
    }
    if (alloc_placed(p, bytes, ranges)) return FL_OK;
    const cudaError_t err = cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocPortable);
    if (err != cudaSuccess) return cuda_fail(err, "cudaHostAlloc");
    return FL_OK;
}
fl_status fl_ctx_host_copy_probe(fl_ctx* ctx, size_t in_bytes_per_block, size_t out_bytes_per_block, size_t n_blocks,
                                 const void* in, void* out) {
    return ctx_run(ctx, n_blocks, [=](size_t, size_t b0, size_t nb) {
        return host_copy_probe(in_bytes_per_block, out_bytes_per_block, nb,
                               in ? static_cast<const char*>(in) + b0 * in_bytes_per_block : nullptr,
                               out ? static_cast<char*>(out) + b0 * out_bytes_per_block : nullptr);
    });
}
/* The "trivial block shard / gather" of north_star inside one process: peer copies over NVLink (cudaMemcpyPeerAsync),
 * outside any decode.  `whole` lives on context device `root`; shards[i] on context device i holds block_range(i). */
static fl_status ctx_peer_move(fl_ctx* ctx, size_t bytes_per_block, size_t n_blocks, void* whole, int root, void* const* shards,
                               bool scatter) {
    if (!ctx || !shards || (n_blocks && !whole)) return fail(FL_ERR_NULL, "null pointer");
    if (root < 0 || size_t(root) >= ctx->devices.size()) return fail(FL_ERR_INDEX, "root index out of range");
    const int root_dev = ctx->devices[size_t(root)];
    return ctx_run(ctx, n_blocks, [=](size_t i, size_t b0, size_t nb) -> fl_status {
        if (!shards[i]) return fail(FL_ERR_NULL, "null shard pointer");
        char* w = static_cast<char*>(whole) + b0 * bytes_per_block;
        const int dev = ctx->devices[i];
        const cudaError_t e = scatter ? cudaMemcpyPeer(shards[i], dev, w, root_dev, nb * bytes_per_block)
                                      : cudaMemcpyPeer(w, root_dev, shards[i], dev, nb * bytes_per_block);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyPeer");
        return FL_OK;
    });
}
fl_status fl_ctx_scatter_blocks(fl_ctx* ctx, size_t bytes_per_block, size_t n_blocks, const void* src, int root,
                                void* const* shards) {
    return ctx_peer_move(ctx, bytes_per_block, n_blocks, const_cast<void*>(src), root, shards, true);
}
fl_status fl_ctx_gather_blocks(fl_ctx* ctx, size_t bytes_per_block, size_t n_blocks, const void* const* shards, int root,
                               void* dst) {
    return ctx_peer_move(ctx, bytes_per_block, n_blocks, dst, root, const_cast<void* const*>(shards), false);
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
    }                                                                                                                   \
    /* context family: the same host calls, block-sharded over the context's devices */                                 \
    fl_status fl_ctx_host_pack_##SFX(fl_ctx* c, unsigned width, size_t n, const T* in, T* packed) {                     \
        return ctx_host_op<T>(c, Op::Pack, width, n, in, packed, nullptr, 0);                                           \
    }                                                                                                                   \
    fl_status fl_ctx_host_unpack_##SFX(fl_ctx* c, unsigned width, size_t n, const T* packed, T* out) {                  \
        return ctx_host_op<T>(c, Op::Unpack, width, n, packed, out, nullptr, 0);                                        \
    }                                                                                                                   \
    fl_status fl_ctx_host_for_pack_##SFX(fl_ctx* c, unsigned width, size_t n, const T* in, T reference, T* packed) {    \
        return ctx_host_op<T>(c, Op::ForPack, width, n, in, packed, nullptr, reference);                                \
    }                                                                                                                   \
    fl_status fl_ctx_host_unfor_pack_##SFX(fl_ctx* c, unsigned width, size_t n, const T* packed, T reference, T* out) { \
        return ctx_host_op<T>(c, Op::UnforPack, width, n, packed, out, nullptr, reference);                             \
    }                                                                                                                   \
    fl_status fl_ctx_host_delta_##SFX(fl_ctx* c, size_t n, const T* in, const T* base, T* out) {                        \
        return ctx_host_op<T>(c, Op::Delta, 0, n, in, out, base, 0);                                                    \
    }                                                                                                                   \
    fl_status fl_ctx_host_undelta_##SFX(fl_ctx* c, size_t n, const T* in, const T* base, T* out) {                      \
        return ctx_host_op<T>(c, Op::Undelta, 0, n, in, out, base, 0);                                                  \
    }                                                                                                                   \
    fl_status fl_ctx_host_undelta_pack_##SFX(fl_ctx* c, unsigned width, size_t n, const T* packed, const T* base, T* out) { \
        return ctx_host_op<T>(c, Op::UndeltaPack, width, n, packed, out, base, 0);                                      \
    }                                                                                                                   \
    fl_status fl_ctx_host_undelta_pack_untranspose_##SFX(fl_ctx* c, unsigned width, size_t n, const T* packed,          \
                                                         const T* base, T* out) {                                       \
        return ctx_host_op<T>(c, Op::UndeltaPackUntranspose, width, n, packed, out, base, 0);                           \
    }                                                                                                                   \
    fl_status fl_ctx_host_transpose_delta_pack_##SFX(fl_ctx* c, unsigned width, size_t n, const T* in, const T* base,   \
                                                     T* packed) {                                                       \
        return ctx_host_op<T>(c, Op::TransposeDeltaPack, width, n, in, packed, base, 0);                                \
    }                                                                                                                   \
    fl_status fl_ctx_host_transpose_##SFX(fl_ctx* c, size_t n, const T* in, T* out) {                                   \
        return ctx_host_op<T>(c, Op::Transpose, 0, n, in, out, nullptr, 0);                                             \
    }                                                                                                                   \
    fl_status fl_ctx_host_untranspose_##SFX(fl_ctx* c, size_t n, const T* in, T* out) {                                 \
        return ctx_host_op<T>(c, Op::Untranspose, 0, n, in, out, nullptr, 0);                                           \
    }                                                                                                                   \
    fl_status fl_ctx_host_block_minmax_##SFX(fl_ctx* c, size_t n, const T* in, T* mins, T* maxs) {                      \
        return ctx_host_minmax<T>(c, n, in, mins, maxs);                                                                \
    }                                                                                                                   \
    fl_status fl_ctx_host_unpack_filter_##SFX(fl_ctx* c, unsigned width, size_t n, const T* packed, T reference, T lo,  \
                                              T hi, uint8_t* bitmap, uint32_t* counts) {                                \
        return ctx_host_filter<T>(c, width, n, packed, nullptr, reference, lo, hi, bitmap, counts);                     \
    }                                                                                                                   \
    fl_status fl_ctx_host_undelta_pack_filter_##SFX(fl_ctx* c, unsigned width, size_t n, const T* packed,               \
                                                    const T* base, T lo, T hi, uint8_t* bitmap, uint32_t* counts) {     \
        if (!base) return fail(FL_ERR_NULL, "null base pointer");                                                       \
        return ctx_host_filter<T>(c, width, n, packed, base, 0, lo, hi, bitmap, counts);                                \
    }

FL_DEFINE_TYPE(u8, uint8_t)
FL_DEFINE_TYPE(u16, uint16_t)
FL_DEFINE_TYPE(u32, uint32_t)
FL_DEFINE_TYPE(u64, uint64_t)

}  // extern "C"
