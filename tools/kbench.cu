// tools/kbench.cu — kernel micro-benchmark harness (development tool, not part of the product library).
// Times the codec kernels with CUDA events on device-resident synthetic data and prints one line per
// (op, type, width): microseconds, algorithmic GB/s, Gint/s.  Compile-time variants (prefetch distance,
// cache hints, CTA size) are selected with -DFLB_PREFETCH / -DFLB_LD_MODE / -DFLB_ST_MODE / -DFLB_THREADS.
//
//   kbench [tbits=32] [op=unpack|pack|undelta_pack|unfor_pack|for_pack|delta|undelta|copy] [log2_blocks=20] [iters=10] [w_lo] [w_hi]
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "fl_kernels.cuh"

using namespace flb;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1);} } while (0)

__global__ void fill_kernel(uint4* p, size_t n, uint64_t seed) {
    size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    for (; i < n; i += stride) {
        uint64_t x = seed + i * 0x9E3779B97F4A7C15ull;
        x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
        uint64_t y = x * 0xD6E8FEB86659FD93ull; y ^= y >> 32;
        p[i] = make_uint4(uint32_t(x), uint32_t(x >> 32), uint32_t(y), uint32_t(y >> 32));
    }
}

__global__ void __launch_bounds__(256) copy_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n) {
    size_t i = (size_t(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n) {
        uint4 a = ldg128_stream(in + i), b = ldg128_stream(in + i + 1), c = ldg128_stream(in + i + 2), d = ldg128_stream(in + i + 3);
        stg128_stream(out + i, a); stg128_stream(out + i + 1, b); stg128_stream(out + i + 2, c); stg128_stream(out + i + 3, d);
    }
}
__global__ void __launch_bounds__(256) write_kernel(uint4* __restrict__ out, size_t n) {
    size_t i = (size_t(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n) {
        uint4 a = make_uint4(1, 2, 3, 4);
        stg128_stream(out + i, a); stg128_stream(out + i + 1, a); stg128_stream(out + i + 2, a); stg128_stream(out + i + 3, a);
    }
}
__global__ void __launch_bounds__(256) read_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n) {
    size_t i = (size_t(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n) {
        uint4 a = ldg128_stream(in + i), b = ldg128_stream(in + i + 1), c = ldg128_stream(in + i + 2), d = ldg128_stream(in + i + 3);
        if ((a.x ^ b.y ^ c.z ^ d.w) == 0x12345678u && a.y == 42) out[0] = a;  // practically never
    }
}

__global__ void checksum_kernel(const uint4* p, size_t n, unsigned long long* out) {
    unsigned long long acc = 0;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        uint4 v = p[i];
        acc += (unsigned long long)v.x * 3 + (unsigned long long)v.y * 5 + (unsigned long long)v.z * 7 + (unsigned long long)v.w * 11 + (i & 0xffff);
    }
    atomicAdd(out, acc);
}
static unsigned long long checksum(const void* p, size_t bytes, cudaStream_t s) {
    unsigned long long* d; unsigned long long h = 0;
    CK(cudaMalloc(&d, 8)); CK(cudaMemsetAsync(d, 0, 8, s));
    checksum_kernel<<<148 * 4, 256, 0, s>>>(static_cast<const uint4*>(p), bytes / 16, d);
    CK(cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, s)); CK(cudaStreamSynchronize(s)); CK(cudaFree(d));
    return h;
}

struct Ctx {
    char* in; char* out; char* base; size_t n_blocks; int iters; cudaStream_t s; cudaEvent_t e0, e1;
};

template <class F>
static float time_ms(const Ctx& c, F&& launch) {
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaGetLastError());
    std::vector<float> ts;
    for (int i = 0; i < c.iters; ++i) {
        CK(cudaEventRecord(c.e0, c.s));
        launch();
        CK(cudaEventRecord(c.e1, c.s));
        CK(cudaEventSynchronize(c.e1));
        float ms; CK(cudaEventElapsedTime(&ms, c.e0, c.e1));
        ts.push_back(ms);
    }
    CK(cudaGetLastError());
    std::sort(ts.begin(), ts.end());
    return ts[ts.size() / 2];
}

static void report(const char* op, int tb, int w, size_t n_blocks, size_t bytes_per_block, float ms) {
    const double gbs = double(n_blocks) * bytes_per_block / (ms * 1e-3) / 1e9;
    const double gints = double(n_blocks) * 1024 / (ms * 1e-3) / 1e9;
    printf("%-13s u%-2d W=%-2d blocks=%zu  %9.1f us  %8.1f GB/s  %8.1f Gint/s\n", op, tb, w, n_blocks, ms * 1e3, gbs, gints);
    fflush(stdout);
}

template <class T, int W>
static void bench_width(const Ctx& c, const std::string& op) {
    constexpr int TB = Lay<T>::TB;
    const unsigned grid = unsigned((c.n_blocks * kSlicesPerBlock + kThreads - 1) / kThreads);
    if (op == "unpack") {
        float ms = time_ms(c, [&] { unpack_kernel<T, W, UOP_PLAIN><<<grid, kThreads, 0, c.s>>>(c.in, c.out, c.n_blocks, nullptr, T(0), nullptr); });
        report("unpack", TB, W, c.n_blocks, 128 * (W + TB), ms);
    } else if (op == "unpackB" || op == "unpackT" || op == "unfor_packB" || op == "undelta_packB") {
        const unsigned gridB = unsigned((c.n_blocks * 32 + kThreads - 1) / kThreads);
        if (op == "unpackT") {
            float ms = time_ms(c, [&] { unpack_warp_kernel<T, W, UOP_PLAIN, true><<<gridB, kThreads, 0, c.s>>>(c.in, c.out, c.n_blocks, nullptr, T(0), nullptr); });
            report("unpackT", TB, W, c.n_blocks, 128 * (W + TB), ms);
            const unsigned long long ct = checksum(c.out, c.n_blocks * 128 * size_t(TB), c.s);
            unpack_warp_kernel<T, W, UOP_PLAIN><<<gridB, kThreads, 0, c.s>>>(c.in, c.out, c.n_blocks, nullptr, T(0), nullptr);
            const unsigned long long cb = checksum(c.out, c.n_blocks * 128 * size_t(TB), c.s);
            if (ct != cb) printf("  !! unpackT checksum MISMATCH at W=%d\n", W);
        } else if (op == "unpackB") {
            float ms = time_ms(c, [&] { unpack_warp_kernel<T, W, UOP_PLAIN><<<gridB, kThreads, 0, c.s>>>(c.in, c.out, c.n_blocks, nullptr, T(0), nullptr); });
            report("unpackB", TB, W, c.n_blocks, 128 * (W + TB), ms);
        } else if (op == "unfor_packB") {
            float ms = time_ms(c, [&] { unpack_warp_kernel<T, W, UOP_FOR><<<gridB, kThreads, 0, c.s>>>(c.in, c.out, c.n_blocks, nullptr, T(12345), nullptr); });
            report("unfor_packB", TB, W, c.n_blocks, 128 * (W + TB), ms);
        } else {
            float ms = time_ms(c, [&] { unpack_warp_kernel<T, W, UOP_DELTA><<<gridB, kThreads, 0, c.s>>>(c.in, c.out, c.n_blocks, nullptr, T(0), c.base); });
            report("undelta_packB", TB, W, c.n_blocks, 128 * (W + TB + 1), ms);
        }
    } else if (op == "unfor_pack") {
        float ms = time_ms(c, [&] { unpack_kernel<T, W, UOP_FOR><<<grid, kThreads, 0, c.s>>>(c.in, c.out, c.n_blocks, nullptr, T(12345), nullptr); });
        report("unfor_pack", TB, W, c.n_blocks, 128 * (W + TB), ms);
    } else if (op == "undelta_pack") {
        float ms = time_ms(c, [&] { unpack_kernel<T, W, UOP_DELTA><<<grid, kThreads, 0, c.s>>>(c.in, c.out, c.n_blocks, nullptr, T(0), c.base); });
        report("undelta_pack", TB, W, c.n_blocks, 128 * (W + TB + 1), ms);
    } else if (op == "packB" || op == "packT" || op == "for_packB") {
        const unsigned gridB = unsigned((c.n_blocks * 32 + kThreads - 1) / kThreads);
        if (op == "packT") {
            float ms = time_ms(c, [&] { pack_warp_kernel<T, W, POP_PLAIN, true><<<gridB, kThreads, (kThreads / 32) * (128 * TB + 8), c.s>>>(c.out, c.in, c.n_blocks, nullptr, T(0), nullptr); });
            report("packT", TB, W, c.n_blocks, 128 * (W + TB), ms);
            const unsigned long long ct = checksum(c.in, c.n_blocks * 128 * size_t(W), c.s);
            pack_warp_kernel<T, W, POP_PLAIN><<<gridB, kThreads, 0, c.s>>>(c.out, c.in, c.n_blocks, nullptr, T(0), nullptr);
            const unsigned long long cb = checksum(c.in, c.n_blocks * 128 * size_t(W), c.s);
            if (ct != cb) printf("  !! packT checksum MISMATCH at W=%d\n", W);
        } else if (op == "packB") {
            float ms = time_ms(c, [&] { pack_warp_kernel<T, W, POP_PLAIN><<<gridB, kThreads, 0, c.s>>>(c.out, c.in, c.n_blocks, nullptr, T(0), nullptr); });
            report("packB", TB, W, c.n_blocks, 128 * (W + TB), ms);
        } else {
            float ms = time_ms(c, [&] { pack_warp_kernel<T, W, POP_FOR><<<gridB, kThreads, 0, c.s>>>(c.out, c.in, c.n_blocks, nullptr, T(12345), nullptr); });
            report("for_packB", TB, W, c.n_blocks, 128 * (W + TB), ms);
        }
    } else if (op == "pack") {
        // input = the "out" buffer (unpacked side), output = the "in" buffer (packed side)
        float ms = time_ms(c, [&] { pack_kernel<T, W, POP_PLAIN><<<grid, kThreads, 0, c.s>>>(c.out, c.in, c.n_blocks, nullptr, T(0)); });
        report("pack", TB, W, c.n_blocks, 128 * (W + TB), ms);
    } else if (op == "for_pack") {
        float ms = time_ms(c, [&] { pack_kernel<T, W, POP_FOR><<<grid, kThreads, 0, c.s>>>(c.out, c.in, c.n_blocks, nullptr, T(12345)); });
        report("for_pack", TB, W, c.n_blocks, 128 * (W + TB), ms);
    }
}

template <class T, int... W>
static void bench_all(const Ctx& c, const std::string& op, int lo, int hi, std::integer_sequence<int, W...>) {
    ((W >= lo && W <= hi ? bench_width<T, W>(c, op) : void()), ...);
}

template <class T>
static void run_type(Ctx& c, const std::string& op, int lo, int hi) {
    constexpr int TB = Lay<T>::TB;
    const unsigned grid = unsigned((c.n_blocks * kSlicesPerBlock + kThreads - 1) / kThreads);
    if (op == "deltaB" || op == "undeltaB") {
        const unsigned gridB = unsigned((c.n_blocks * 32 + kThreads - 1) / kThreads);
        char* tmp = c.in;
        float ms = (op == "deltaB")
            ? time_ms(c, [&] { delta_warp_kernel<T, false><<<gridB, kThreads, 0, c.s>>>(c.out, c.base, tmp, c.n_blocks); })
            : time_ms(c, [&] { delta_warp_kernel<T, true><<<gridB, kThreads, 0, c.s>>>(c.out, c.base, tmp, c.n_blocks); });
        report(op.c_str(), TB, 0, c.n_blocks, 128 * (2 * TB + 1), ms);
        return;
    }
    if (op == "delta" || op == "undelta") {
        char* tmp = c.in;  // both sides are unpacked-size: in buffer is sized for W = TB
        float ms = (op == "delta")
            ? time_ms(c, [&] { delta_kernel<T, false><<<grid, kThreads, 0, c.s>>>(c.out, c.base, tmp, c.n_blocks); })
            : time_ms(c, [&] { delta_kernel<T, true><<<grid, kThreads, 0, c.s>>>(c.out, c.base, tmp, c.n_blocks); });
        report(op.c_str(), TB, 0, c.n_blocks, 128 * (2 * TB + 1), ms);
        return;
    }
    bench_all<T>(c, op, lo, hi, std::make_integer_sequence<int, TB + 1>{});
}

int main(int argc, char** argv) {
    const int tb = argc > 1 ? atoi(argv[1]) : 32;
    const std::string op = argc > 2 ? argv[2] : "unpack";
    const int lg = argc > 3 ? atoi(argv[3]) : 20;
    Ctx c{};
    c.iters = argc > 4 ? atoi(argv[4]) : 10;
    const int lo = argc > 5 ? atoi(argv[5]) : 1;
    const int hi = argc > 6 ? atoi(argv[6]) : tb;
    c.n_blocks = size_t(1) << lg;
    const size_t unpacked = c.n_blocks * 128 * size_t(tb);
    CK(cudaMalloc(&c.in, unpacked));   // packed side, sized for W = T
    CK(cudaMalloc(&c.out, unpacked));  // unpacked side
    CK(cudaMalloc(&c.base, c.n_blocks * 128));
    CK(cudaStreamCreate(&c.s));
    CK(cudaEventCreate(&c.e0)); CK(cudaEventCreate(&c.e1));
    fill_kernel<<<148 * 8, 256, 0, c.s>>>(reinterpret_cast<uint4*>(c.in), unpacked / 16, 42);
    fill_kernel<<<148 * 8, 256, 0, c.s>>>(reinterpret_cast<uint4*>(c.out), unpacked / 16, 43);
    fill_kernel<<<148 * 8, 256, 0, c.s>>>(reinterpret_cast<uint4*>(c.base), c.n_blocks * 8, 44);
    CK(cudaStreamSynchronize(c.s));
    printf("# kbench tbits=%d op=%s blocks=2^%d iters=%d threads=%d prefetch=%d ld_mode=%d st_mode=%d\n", tb, op.c_str(), lg,
           c.iters, kThreads, FLB_PREFETCH, FLB_LD_MODE, FLB_ST_MODE);
    if (op == "copy") {
        const size_t n16 = unpacked / 16;
        const unsigned grid = unsigned((n16 / 4 + 255) / 256);
        float ms = time_ms(c, [&] { copy_kernel<<<grid, 256, 0, c.s>>>((const uint4*)c.in, (uint4*)c.out, n16); });
        printf("copy_kernel   %zu B   %9.1f us  %8.1f GB/s (read+write)\n", unpacked, ms * 1e3, 2.0 * unpacked / (ms * 1e-3) / 1e9);
        ms = time_ms(c, [&] { CK(cudaMemcpyAsync(c.out, c.in, unpacked, cudaMemcpyDeviceToDevice, c.s)); });
        printf("cudaMemcpyD2D %zu B   %9.1f us  %8.1f GB/s (read+write)\n", unpacked, ms * 1e3, 2.0 * unpacked / (ms * 1e-3) / 1e9);
        ms = time_ms(c, [&] { write_kernel<<<grid, 256, 0, c.s>>>((uint4*)c.out, n16); });
        printf("write_kernel  %zu B   %9.1f us  %8.1f GB/s (write only)\n", unpacked, ms * 1e3, 1.0 * unpacked / (ms * 1e-3) / 1e9);
        ms = time_ms(c, [&] { read_kernel<<<grid, 256, 0, c.s>>>((const uint4*)c.in, (uint4*)c.out, n16); });
        printf("read_kernel   %zu B   %9.1f us  %8.1f GB/s (read only)\n", unpacked, ms * 1e3, 1.0 * unpacked / (ms * 1e-3) / 1e9);
        ms = time_ms(c, [&] { CK(cudaMemsetAsync(c.out, 0, unpacked, c.s)); });
        printf("cudaMemset    %zu B   %9.1f us  %8.1f GB/s (write only)\n", unpacked, ms * 1e3, 1.0 * unpacked / (ms * 1e-3) / 1e9);
        return 0;
    }
    switch (tb) {
#ifndef KB_ONLY_U32
        case 8: run_type<uint8_t>(c, op, lo, hi); break;
        case 16: run_type<uint16_t>(c, op, lo, hi); break;
        case 64: run_type<uint64_t>(c, op, lo, hi); break;
#endif
        case 32: run_type<uint32_t>(c, op, lo, hi); break;
        default: fprintf(stderr, "bad tbits\n"); return 2;
    }
    return 0;
}
