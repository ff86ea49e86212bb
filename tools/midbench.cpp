// midbench.cpp — mid-size host calls through the C ABI from compiled code: the reference's own throughput bench shape
// (benches/bitpacking.rs:67-98: u16 W=3, 1024 blocks) and its neighbours, with pageable and with page-locked buffers.
//   build: make build/midbench      run: build/midbench [iters]     (FLB_DIRECT_MAX=0 disables the direct path: A/B)
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "fastlanes_b200.h"

template <class F>
static double median_us(F&& f, int iters) {
    for (int i = 0; i < 5; ++i) f();
    std::vector<double> t(iters);
    for (int i = 0; i < iters; ++i) {
        auto a = std::chrono::steady_clock::now();
        f();
        t[i] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - a).count();
    }
    std::sort(t.begin(), t.end());
    return t[iters / 2];
}

template <class T>
static T* alloc(bool pinned, size_t n) {
    void* p = nullptr;
    if (pinned) { if (fl_host_alloc(&p, n * sizeof(T)) != FL_OK) { std::printf("fl_host_alloc: %s\n", fl_last_error_string()); std::exit(1); } }
    else p = std::aligned_alloc(64, (n * sizeof(T) + 63) / 64 * 64);
    std::memset(p, 0, n * sizeof(T));
    return static_cast<T*>(p);
}
template <class T>
static void release(bool pinned, T* p) { if (pinned) fl_host_free(p); else std::free(p); }

int main(int argc, char** argv) {
    const int iters = argc > 1 ? std::atoi(argv[1]) : 200;
    if (fl_init(0) != FL_OK) { std::printf("fl_init failed: %s\n", fl_last_error_string()); return 1; }
    const char* dm = std::getenv("FLB_DIRECT_MAX");
    std::printf("FLB_DIRECT_MAX=%s\n", dm ? dm : "(default)");
    for (int pinned = 0; pinned <= 1; ++pinned) {
        for (size_t n : {64, 256, 1024, 4096, 16384, 65536}) {
            const int it = n >= 16384 ? std::max(20, iters / 10) : iters;
            uint16_t* v16 = alloc<uint16_t>(pinned, n * 1024); uint16_t* p16 = alloc<uint16_t>(pinned, n * 192); uint16_t* u16 = alloc<uint16_t>(pinned, n * 1024);
            for (size_t i = 0; i < n * 1024; ++i) v16[i] = uint16_t(i % 8);  // benches/bitpacking.rs:70
            const double tp = median_us([&] { fl_host_pack_u16(3, n, v16, p16); }, it);
            const double tu = median_us([&] { fl_host_unpack_u16(3, n, p16, u16); }, it);
            if (std::memcmp(v16, u16, n * 2048) != 0) { std::printf("MISMATCH u16 n=%zu\n", n); return 1; }
            const double bytes = double(n) * 2048;
            std::printf("%-9s %6zu blocks u16 W=3 : pack %9.1f us (%6.2f GB/s)  unpack %9.1f us (%6.2f GB/s)   [unpacked bytes / time]\n",
                        pinned ? "pinned" : "pageable", n, tp, bytes / tp / 1e3, tu, bytes / tu / 1e3);
            release(pinned, v16); release(pinned, p16); release(pinned, u16);
            uint32_t* v32 = alloc<uint32_t>(pinned, n * 1024); uint32_t* p32 = alloc<uint32_t>(pinned, n * 320); uint32_t* u32 = alloc<uint32_t>(pinned, n * 1024);
            for (size_t i = 0; i < n * 1024; ++i) v32[i] = (uint32_t(i) * 2654435761u) >> 22;
            fl_host_pack_u32(10, n, v32, p32);
            const double t32 = median_us([&] { fl_host_unpack_u32(10, n, p32, u32); }, it);
            if (std::memcmp(v32, u32, n * 4096) != 0) { std::printf("MISMATCH u32 n=%zu\n", n); return 1; }
            std::printf("%-9s %6zu blocks u32 W=10: unpack %9.1f us (%6.2f GB/s, %6.2f Gint/s)\n", pinned ? "pinned" : "pageable", n, t32,
                        double(n) * 4096 / t32 / 1e3, double(n) * 1024 / t32 / 1e3);
            release(pinned, v32); release(pinned, p32); release(pinned, u32);
        }
    }
    fl_shutdown();
    return 0;
}
