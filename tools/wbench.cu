// tools/wbench.cu — HBM store/load pattern micro-benchmark (development tool).
// Question it answers: which store pattern reaches cudaMemset-class write bandwidth on B200, so the unpack
// kernels (write-dominated: 4096 B out per 128*W B in) can be shaped accordingly.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1);} } while (0)

__device__ __forceinline__ void st_cs(void* p, uint4 v) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_plain(void* p, uint4 v) {
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_nc(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// P1: linear — consecutive threads write consecutive 16 B; each thread writes ITER chunks strided by the CTA span.
template <int ITER, bool CS>
__global__ void __launch_bounds__(256) w_linear(uint4* out, size_t n16) {
    size_t base = size_t(blockIdx.x) * (256 * ITER) + threadIdx.x;
    uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3, 4);
#pragma unroll
    for (int i = 0; i < ITER; ++i) {
        size_t idx = base + size_t(i) * 256;
        if (idx < n16) { if (CS) st_cs(out + idx, v); else st_plain(out + idx, v); }
    }
}
// P2: unpack-like row-slice: 8 threads per 4 KB block, 32 rows of 128 B, thread writes its 16 B of every row.
__global__ void __launch_bounds__(256) w_rowslice(char* out, size_t n_blocks) {
    size_t tid = size_t(blockIdx.x) * 256 + threadIdx.x;
    size_t blk = tid >> 3; int j = tid & 7;
    if (blk >= n_blocks) return;
    char* o = out + blk * 4096 + j * 16;
    uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3, 4);
#pragma unroll
    for (int r = 0; r < 32; ++r) st_cs(o + r * 128, v);
}
// P3: warp per 4 KB block: each warp store instruction writes 512 contiguous bytes, 8 instructions per block.
__global__ void __launch_bounds__(256) w_warpblock(char* out, size_t n_blocks) {
    size_t warp = (size_t(blockIdx.x) * 256 + threadIdx.x) >> 5; int lane = threadIdx.x & 31;
    if (warp >= n_blocks) return;
    char* o = out + warp * 4096 + lane * 16;
    uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3, 4);
#pragma unroll
    for (int r = 0; r < 8; ++r) st_cs(o + r * 512, v);
}
// P3b: warp per 4 blocks (16 KB): 32 instructions x 512 contiguous bytes
__global__ void __launch_bounds__(256) w_warp4(char* out, size_t n_blocks) {
    size_t warp = (size_t(blockIdx.x) * 256 + threadIdx.x) >> 5; int lane = threadIdx.x & 31;
    if (warp * 4 >= n_blocks) return;
    char* o = out + warp * 16384 + lane * 16;
    uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3, 4);
#pragma unroll
    for (int r = 0; r < 32; ++r) st_cs(o + r * 512, v);
}

// P4: TMA bulk store from shared memory: CTA tile of TILE bytes, NBUF buffers, persistent grid.
template <int TILE, int NBUF>
__global__ void __launch_bounds__(256) w_bulk(char* out, size_t n_tiles) {
    extern __shared__ __align__(128) char smem[];
    uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3, 4);
    int buf = 0;
    for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        char* s = smem + buf * TILE;
        // buffer reuse: wait until the bulk store issued NBUF iterations ago has finished READING smem
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NBUF - 1) : "memory");
        __syncthreads();
        for (int i = threadIdx.x * 16; i < TILE; i += 256 * 16) *reinterpret_cast<uint4*>(s + i) = v;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned sa = (unsigned)__cvta_generic_to_shared(s);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + t * TILE), "r"(sa), "n"(TILE) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        buf = (buf + 1) % NBUF;
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// C1: linear copy, consecutive threads consecutive 16 B, ITER in flight per thread
template <int ITER>
__global__ void __launch_bounds__(256) c_linear(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n16) {
    size_t base = size_t(blockIdx.x) * (256 * ITER) + threadIdx.x;
    uint4 v[ITER];
#pragma unroll
    for (int i = 0; i < ITER; ++i) { size_t idx = base + size_t(i) * 256; if (idx < n16) v[i] = ld_nc(in + idx); }
#pragma unroll
    for (int i = 0; i < ITER; ++i) { size_t idx = base + size_t(i) * 256; if (idx < n16) st_cs(out + idx, v[i]); }
}
// R1: linear read
template <int ITER>
__global__ void __launch_bounds__(256) r_linear(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n16) {
    size_t base = size_t(blockIdx.x) * (256 * ITER) + threadIdx.x;
    uint4 acc = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int i = 0; i < ITER; ++i) { size_t idx = base + size_t(i) * 256; if (idx < n16) { uint4 v = ld_nc(in + idx); acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w; } }
    if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345u) out[0] = acc;
}
// M1: unpack-like mix with W/32 read fraction: row-slice writes + linear-ish reads of W rows (pattern only, no math)
template <int W>
__global__ void __launch_bounds__(256) m_rowslice(const char* __restrict__ in, char* __restrict__ out, size_t n_blocks) {
    size_t tid = size_t(blockIdx.x) * 256 + threadIdx.x;
    size_t blk = tid >> 3; int j = tid & 7;
    if (blk >= n_blocks) return;
    const char* p = in + blk * (128 * W) + j * 16;
    char* o = out + blk * 4096 + j * 16;
    uint4 acc = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int k = 0; k < W; ++k) { uint4 v = ld_nc(p + k * 128); acc.x ^= v.x; acc.y += v.y; acc.z ^= v.z; acc.w += v.w; }
#pragma unroll
    for (int r = 0; r < 32; ++r) { acc.x += r; st_cs(o + r * 128, acc); }
}

struct T { cudaStream_t s; cudaEvent_t e0, e1; };
template <class F> static float run(T& t, F&& f, int iters = 10) {
    for (int i = 0; i < 3; ++i) f();
    CK(cudaGetLastError());
    std::vector<float> ts;
    for (int i = 0; i < iters; ++i) {
        CK(cudaEventRecord(t.e0, t.s)); f(); CK(cudaEventRecord(t.e1, t.s)); CK(cudaEventSynchronize(t.e1));
        float ms; CK(cudaEventElapsedTime(&ms, t.e0, t.e1)); ts.push_back(ms);
    }
    CK(cudaGetLastError());
    std::sort(ts.begin(), ts.end());
    return ts[ts.size() / 2];
}
static void rep(const char* name, double bytes, float ms) { printf("%-34s %9.1f us  %8.1f GB/s\n", name, ms * 1e3, bytes / (ms * 1e-3) / 1e9); fflush(stdout); }

int main() {
    const size_t bytes = size_t(4) << 30;
    const size_t n16 = bytes / 16, n_blocks = bytes / 4096;
    char *a, *b;
    CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes));
    CK(cudaMemset(a, 1, bytes)); CK(cudaMemset(b, 2, bytes));
    T t; CK(cudaStreamCreate(&t.s)); CK(cudaEventCreate(&t.e0)); CK(cudaEventCreate(&t.e1));
    float ms;
    ms = run(t, [&] { CK(cudaMemsetAsync(b, 0, bytes, t.s)); }); rep("cudaMemset (write)", bytes, ms);
    ms = run(t, [&] { CK(cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice, t.s)); }); rep("cudaMemcpy D2D (r+w)", 2.0 * bytes, ms);
    ms = run(t, [&] { w_linear<1, true><<<unsigned(n16 / 256), 256, 0, t.s>>>((uint4*)b, n16); }); rep("w_linear<1,cs>", bytes, ms);
    ms = run(t, [&] { w_linear<4, true><<<unsigned(n16 / 1024), 256, 0, t.s>>>((uint4*)b, n16); }); rep("w_linear<4,cs>", bytes, ms);
    ms = run(t, [&] { w_linear<8, true><<<unsigned(n16 / 2048), 256, 0, t.s>>>((uint4*)b, n16); }); rep("w_linear<8,cs>", bytes, ms);
    ms = run(t, [&] { w_linear<8, false><<<unsigned(n16 / 2048), 256, 0, t.s>>>((uint4*)b, n16); }); rep("w_linear<8,plain>", bytes, ms);
    ms = run(t, [&] { w_linear<32, true><<<unsigned(n16 / 8192), 256, 0, t.s>>>((uint4*)b, n16); }); rep("w_linear<32,cs>", bytes, ms);
    ms = run(t, [&] { w_rowslice<<<unsigned(n_blocks * 8 / 256), 256, 0, t.s>>>(b, n_blocks); }); rep("w_rowslice (unpack pattern)", bytes, ms);
    ms = run(t, [&] { w_warpblock<<<unsigned(n_blocks * 32 / 256), 256, 0, t.s>>>(b, n_blocks); }); rep("w_warpblock (512B/instr)", bytes, ms);
    ms = run(t, [&] { w_warp4<<<unsigned(n_blocks * 8 / 256), 256, 0, t.s>>>(b, n_blocks); }); rep("w_warp4 (512B/instr x32)", bytes, ms);
    {
        constexpr int TILE = 16384, NBUF = 2;
        CK(cudaFuncSetAttribute(w_bulk<TILE, NBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE * NBUF));
        for (int mult : {1, 2, 4, 6}) {
            ms = run(t, [&] { w_bulk<TILE, NBUF><<<148 * mult, 256, TILE * NBUF, t.s>>>(b, bytes / TILE); });
            char nm[64]; snprintf(nm, 64, "w_bulk<16K,2> grid=148x%d", mult); rep(nm, bytes, ms);
        }
    }
    {
        constexpr int TILE = 32768, NBUF = 2;
        CK(cudaFuncSetAttribute(w_bulk<TILE, NBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE * NBUF));
        for (int mult : {1, 2, 3}) {
            ms = run(t, [&] { w_bulk<TILE, NBUF><<<148 * mult, 256, TILE * NBUF, t.s>>>(b, bytes / TILE); });
            char nm[64]; snprintf(nm, 64, "w_bulk<32K,2> grid=148x%d", mult); rep(nm, bytes, ms);
        }
    }
    {
        constexpr int TILE = 4096, NBUF = 4;
        CK(cudaFuncSetAttribute(w_bulk<TILE, NBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE * NBUF));
        for (int mult : {4, 8}) {
            ms = run(t, [&] { w_bulk<TILE, NBUF><<<148 * mult, 256, TILE * NBUF, t.s>>>(b, bytes / TILE); });
            char nm[64]; snprintf(nm, 64, "w_bulk<4K,4> grid=148x%d", mult); rep(nm, bytes, ms);
        }
    }
    ms = run(t, [&] { r_linear<8><<<unsigned(n16 / 2048), 256, 0, t.s>>>((const uint4*)a, (uint4*)b, n16); }); rep("r_linear<8> (read)", bytes, ms);
    ms = run(t, [&] { c_linear<4><<<unsigned(n16 / 1024), 256, 0, t.s>>>((const uint4*)a, (uint4*)b, n16); }); rep("c_linear<4> (r+w)", 2.0 * bytes, ms);
    ms = run(t, [&] { c_linear<8><<<unsigned(n16 / 2048), 256, 0, t.s>>>((const uint4*)a, (uint4*)b, n16); }); rep("c_linear<8> (r+w)", 2.0 * bytes, ms);
    ms = run(t, [&] { m_rowslice<1><<<unsigned(n_blocks * 8 / 256), 256, 0, t.s>>>(a, b, n_blocks); }); rep("m_rowslice<W=1> (4224 B/blk)", 4224.0 * n_blocks, ms);
    ms = run(t, [&] { m_rowslice<8><<<unsigned(n_blocks * 8 / 256), 256, 0, t.s>>>(a, b, n_blocks); }); rep("m_rowslice<W=8> (5120 B/blk)", 5120.0 * n_blocks, ms);
    ms = run(t, [&] { m_rowslice<16><<<unsigned(n_blocks * 8 / 256), 256, 0, t.s>>>(a, b, n_blocks); }); rep("m_rowslice<W=16> (6144 B/blk)", 6144.0 * n_blocks, ms);
    ms = run(t, [&] { m_rowslice<32><<<unsigned(n_blocks * 8 / 256), 256, 0, t.s>>>(a, b, n_blocks); }); rep("m_rowslice<W=32> (8192 B/blk)", 8192.0 * n_blocks, ms);
    return 0;
}
