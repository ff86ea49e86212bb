// tools/wbench2.cu — which side of the unpack access pattern costs bandwidth?  (development tool)
// Pattern-only kernels (no bit math): RD in {rowslice, warpblock, linear} x WR in {rowslice, warpblock}.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1);} } while (0)
__device__ __forceinline__ void st_cs(void* p, uint4 v) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_nc(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void mix(uint4& a, uint4 v) { a.x ^= v.x; a.y += v.y; a.z ^= v.z; a.w += v.w; }

// A: rowslice read + rowslice write (the shipped layout): 8 threads per block
template <int W>
__global__ void __launch_bounds__(256) p_rs_rs(const char* __restrict__ in, char* __restrict__ out, size_t nb) {
    size_t tid = size_t(blockIdx.x) * 256 + threadIdx.x; size_t blk = tid >> 3; int j = tid & 7;
    if (blk >= nb) return;
    const char* p = in + blk * (128 * W) + j * 16; char* o = out + blk * 4096 + j * 16;
    uint4 acc = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int k = 0; k < W; ++k) mix(acc, ld_nc(p + k * 128));
#pragma unroll
    for (int r = 0; r < 32; ++r) { acc.x += r; st_cs(o + r * 128, acc); }
}
// B: warp per block: reads 512 contiguous B per instruction (W/4 instr, last partial), writes 8 x 512 contiguous
template <int W>
__global__ void __launch_bounds__(256) p_wb_wb(const char* __restrict__ in, char* __restrict__ out, size_t nb) {
    size_t warp = (size_t(blockIdx.x) * 256 + threadIdx.x) >> 5; int lane = threadIdx.x & 31;
    if (warp >= nb) return;
    const char* p = in + warp * (128 * W); char* o = out + warp * 4096 + lane * 16;
    uint4 acc = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int k = 0; k < (W * 128 + 511) / 512; ++k) { int off = k * 512 + lane * 16; if (off < W * 128) mix(acc, ld_nc(p + off)); }
#pragma unroll
    for (int r = 0; r < 8; ++r) { acc.x += r; st_cs(o + r * 512, acc); }
}
// C: rowslice read + warp-contiguous write via exchange-free trick (pattern only): 8 threads/block read, but each warp
//    (4 blocks) writes 512 contiguous bytes per instruction covering block after block (32 instr x 512 B = 16 KB)
template <int W>
__global__ void __launch_bounds__(256) p_rs_wb(const char* __restrict__ in, char* __restrict__ out, size_t nb) {
    size_t tid = size_t(blockIdx.x) * 256 + threadIdx.x; size_t blk = tid >> 3; int j = tid & 7;
    if (blk >= nb) return;
    const char* p = in + blk * (128 * W) + j * 16;
    size_t warp = tid >> 5; int lane = threadIdx.x & 31;
    char* o = out + warp * 16384 + lane * 16;
    uint4 acc = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int k = 0; k < W; ++k) mix(acc, ld_nc(p + k * 128));
#pragma unroll
    for (int r = 0; r < 32; ++r) { acc.x += r; st_cs(o + r * 512, acc); }
}
// D: warp-contiguous read + rowslice write
template <int W>
__global__ void __launch_bounds__(256) p_wb_rs(const char* __restrict__ in, char* __restrict__ out, size_t nb) {
    size_t tid = size_t(blockIdx.x) * 256 + threadIdx.x; size_t blk = tid >> 3; int j = tid & 7;
    if (blk >= nb) return;
    size_t warp = tid >> 5; int lane = threadIdx.x & 31;
    const char* p = in + warp * (512 * W) + lane * 16;   // 4 blocks' packed data, contiguous
    char* o = out + blk * 4096 + j * 16;
    uint4 acc = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int k = 0; k < W; ++k) mix(acc, ld_nc(p + k * 512));
#pragma unroll
    for (int r = 0; r < 32; ++r) { acc.x += r; st_cs(o + r * 128, acc); }
}
// E: rowslice both, but 2 blocks per thread group (16 threads... no: each thread handles 2 consecutive blocks sequentially)
template <int W>
__global__ void __launch_bounds__(256) p_rs_rs2(const char* __restrict__ in, char* __restrict__ out, size_t nb) {
    size_t tid = size_t(blockIdx.x) * 256 + threadIdx.x; size_t g = tid >> 3; int j = tid & 7;
    // warp covers 8 consecutive blocks: groups 0..3 take blocks 4q..4q+3 then 4q+4..4q+7? keep contiguity per warp: blk = (warp*8) + (grp) and +4
    size_t warp = tid >> 5; int grp = (threadIdx.x >> 3) & 3;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        size_t blk = warp * 8 + h * 4 + grp;
        if (blk >= nb) return;
        const char* p = in + blk * (128 * W) + j * 16; char* o = out + blk * 4096 + j * 16;
        uint4 acc = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int k = 0; k < W; ++k) mix(acc, ld_nc(p + k * 128));
#pragma unroll
        for (int r = 0; r < 32; ++r) { acc.x += r; st_cs(o + r * 128, acc); }
    }
    (void)g;
}
// F: persistent grid-stride rowslice (grid = 148*k CTAs)
template <int W>
__global__ void __launch_bounds__(256) p_rs_rs_persist(const char* __restrict__ in, char* __restrict__ out, size_t nb) {
    const int j = threadIdx.x & 7;
    for (size_t blk = (size_t(blockIdx.x) * 256 + threadIdx.x) >> 3; blk < nb; blk += size_t(gridDim.x) * 32) {
        const char* p = in + blk * (128 * W) + j * 16; char* o = out + blk * 4096 + j * 16;
        uint4 acc = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int k = 0; k < W; ++k) mix(acc, ld_nc(p + k * 128));
#pragma unroll
        for (int r = 0; r < 32; ++r) { acc.x += r; st_cs(o + r * 128, acc); }
    }
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
static void rep(const char* name, int W, double bytes, float ms) { printf("%-22s W=%-2d %9.1f us  %8.1f GB/s\n", name, W, ms * 1e3, bytes / (ms * 1e-3) / 1e9); fflush(stdout); }

template <int W> static void sweep(T& t, const char* a, char* b, size_t nb) {
    const double bytes = double(128 * (W + 32)) * nb;
    float ms;
    ms = run(t, [&] { p_rs_rs<W><<<unsigned(nb * 8 / 256), 256, 0, t.s>>>(a, b, nb); }); rep("A rs->rs", W, bytes, ms);
    ms = run(t, [&] { p_wb_wb<W><<<unsigned(nb * 32 / 256), 256, 0, t.s>>>(a, b, nb); }); rep("B warpblk->warpblk", W, bytes, ms);
    ms = run(t, [&] { p_rs_wb<W><<<unsigned(nb * 8 / 256), 256, 0, t.s>>>(a, b, nb); }); rep("C rs->warp512", W, bytes, ms);
    ms = run(t, [&] { p_wb_rs<W><<<unsigned(nb * 8 / 256), 256, 0, t.s>>>(a, b, nb); }); rep("D warp512->rs", W, bytes, ms);
    ms = run(t, [&] { p_rs_rs2<W><<<unsigned(nb * 4 / 256), 256, 0, t.s>>>(a, b, nb); }); rep("E rs->rs 2blk/thr", W, bytes, ms);
    for (int mult : {4, 8}) {
        ms = run(t, [&] { p_rs_rs_persist<W><<<148 * mult, 256, 0, t.s>>>(a, b, nb); });
        char nm[32]; snprintf(nm, 32, "F rs persist x%d", mult); rep(nm, W, bytes, ms);
    }
}

int main() {
    const size_t bytes = size_t(4) << 30; const size_t nb = bytes / 4096;
    char *a, *b; CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes));
    CK(cudaMemset(a, 1, bytes)); CK(cudaMemset(b, 2, bytes));
    T t; CK(cudaStreamCreate(&t.s)); CK(cudaEventCreate(&t.e0)); CK(cudaEventCreate(&t.e1));
    sweep<1>(t, a, b, nb); sweep<4>(t, a, b, nb); sweep<8>(t, a, b, nb); sweep<16>(t, a, b, nb); sweep<32>(t, a, b, nb);
    return 0;
}
