// tools/wbench3.cu — does TMA (cp.async.bulk) help the PACKED-BLOCK LOAD of the warp-block unpack pattern?  (development tool)
// north_star suggests "TMA for the packed-block bulk copy".  Pattern-only kernels (no bit math), u32 shape:
//   B0  warp-block, direct LDG.128 (the shipped pattern)
//   T1  warp-block, one-shot: lane 0 issues ONE cp.async.bulk global->shared of the block's 128*W bytes on a per-warp
//       mbarrier, the warp waits, reads shared (LDS.128) and stores 8 x 512 B
//   T2  persistent warps, 2-stage ring: the bulk load of block n+1 is in flight while block n is consumed
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
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile("{\n .reg .pred p;\n WAIT:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE;\n bra WAIT;\n DONE:\n}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int W>
__global__ void __launch_bounds__(256) p_direct(const char* __restrict__ in, char* __restrict__ out, size_t nb) {
    size_t warp = (size_t(blockIdx.x) * 256 + threadIdx.x) >> 5; int lane = threadIdx.x & 31;
    if (warp >= nb) return;
    const char* p = in + warp * (128 * W); char* o = out + warp * 4096 + lane * 16;
    uint4 acc = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int k = 0; k < (W * 128 + 511) / 512; ++k) { int off = k * 512 + lane * 16; if (off < W * 128) mix(acc, ld_nc(p + off)); }
#pragma unroll
    for (int r = 0; r < 8; ++r) { acc.x += r; st_cs(o + r * 512, acc); }
}

template <int W>
__global__ void __launch_bounds__(256) p_tma_oneshot(const char* __restrict__ in, char* __restrict__ out, size_t nb) {
    __shared__ __align__(128) unsigned char buf[8][128 * W];
    __shared__ __align__(8) unsigned long long bars[8];
    const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
    size_t warp = (size_t(blockIdx.x) * 256 + threadIdx.x) >> 5;
    if (warp >= nb) return;
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&bars[wi]);
    const unsigned dst = (unsigned)__cvta_generic_to_shared(&buf[wi][0]);
    if (lane == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    if (lane == 0) { mbar_expect_tx(bar, 128 * W); bulk_g2s(dst, in + warp * (128 * W), 128 * W, bar); }
    mbar_wait(bar, 0);
    uint4 acc = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int k = 0; k < (W * 128 + 511) / 512; ++k) { int off = k * 512 + lane * 16; if (off < W * 128) mix(acc, *reinterpret_cast<const uint4*>(&buf[wi][off])); }
    char* o = out + warp * 4096 + lane * 16;
#pragma unroll
    for (int r = 0; r < 8; ++r) { acc.x += r; st_cs(o + r * 512, acc); }
}

template <int W>
__global__ void __launch_bounds__(256) p_tma_ring(const char* __restrict__ in, char* __restrict__ out, size_t nb) {
    __shared__ __align__(128) unsigned char buf[8][2][128 * W];
    __shared__ __align__(8) unsigned long long bars[8][2];
    const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t nwarps = size_t(gridDim.x) * 8;
    size_t blk = size_t(blockIdx.x) * 8 + wi;
    unsigned bar[2] = {(unsigned)__cvta_generic_to_shared(&bars[wi][0]), (unsigned)__cvta_generic_to_shared(&bars[wi][1])};
    unsigned dst[2] = {(unsigned)__cvta_generic_to_shared(&buf[wi][0][0]), (unsigned)__cvta_generic_to_shared(&buf[wi][1][0])};
    if (lane == 0) { mbar_init(bar[0], 1); mbar_init(bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    if (blk < nb && lane == 0) { mbar_expect_tx(bar[0], 128 * W); bulk_g2s(dst[0], in + blk * (128 * W), 128 * W, bar[0]); }
    unsigned phase[2] = {0, 0};
    int s = 0;
    for (; blk < nb; blk += nwarps) {
        const size_t nxt = blk + nwarps;
        if (nxt < nb && lane == 0) { mbar_expect_tx(bar[s ^ 1], 128 * W); bulk_g2s(dst[s ^ 1], in + nxt * (128 * W), 128 * W, bar[s ^ 1]); }
        mbar_wait(bar[s], phase[s]); phase[s] ^= 1;
        uint4 acc = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int k = 0; k < (W * 128 + 511) / 512; ++k) { int off = k * 512 + lane * 16; if (off < W * 128) mix(acc, *reinterpret_cast<const uint4*>(&buf[wi][s][off])); }
        __syncwarp();  // all lanes done reading buf[s] before it is refilled two iterations later
        char* o = out + blk * 4096 + lane * 16;
#pragma unroll
        for (int r = 0; r < 8; ++r) { acc.x += r; st_cs(o + r * 512, acc); }
        s ^= 1;
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
static void rep(const char* name, int W, double bytes, float ms) { printf("%-28s W=%-2d %9.1f us  %8.1f GB/s\n", name, W, ms * 1e3, bytes / (ms * 1e-3) / 1e9); fflush(stdout); }

template <int W> static void sweep(T& t, const char* a, char* b, size_t nb) {
    const double bytes = double(128 * (W + 32)) * nb;
    float ms;
    ms = run(t, [&] { p_direct<W><<<unsigned(nb / 8), 256, 0, t.s>>>(a, b, nb); }); rep("B0 direct LDG.128", W, bytes, ms);
    ms = run(t, [&] { p_tma_oneshot<W><<<unsigned(nb / 8), 256, 0, t.s>>>(a, b, nb); }); rep("T1 TMA bulk load, one-shot", W, bytes, ms);
    for (int mult : {4, 8}) {
        ms = run(t, [&] { p_tma_ring<W><<<148 * mult, 256, 0, t.s>>>(a, b, nb); });
        char nm[48]; snprintf(nm, 48, "T2 TMA ring x%d CTAs/SM", mult); rep(nm, W, bytes, ms);
    }
}

int main() {
    const size_t bytes = size_t(4) << 30; const size_t nb = bytes / 4096;
    char *a, *b; CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes));
    CK(cudaMemset(a, 1, bytes)); CK(cudaMemset(b, 2, bytes));
    T t; CK(cudaStreamCreate(&t.s)); CK(cudaEventCreate(&t.e0)); CK(cudaEventCreate(&t.e1));
    sweep<4>(t, a, b, nb); sweep<8>(t, a, b, nb); sweep<16>(t, a, b, nb); sweep<20>(t, a, b, nb);
    return 0;
}
