// latbench.cpp — latency of the drop-in single-block trait call through the C ABI, from compiled host code (what the
// Rust shim's `impl BitPacking for u16` delivers), next to nothing else: no Python in the timed path.
//   build: make build/latbench      run: build/latbench [iters]     (FLB_SMALL=0 selects the round-1 copy path)
// <= 10 us per single-block fl_host_unpack_u16 call (was 21.6 us; the CPU loop takes 4.3 us).
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "fastlanes_b200.h"

template <class F>
static double median_us(F&& f, int iters) {
    for (int i = 0; i < 50; ++i) f();
    std::vector<double> t(iters);
    for (int i = 0; i < iters; ++i) {
        auto a = std::chrono::steady_clock::now();
        f();
        t[i] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - a).count();
    }
    std::sort(t.begin(), t.end());
    return t[iters / 2];
}

int main(int argc, char** argv) {
    const int iters = argc > 1 ? std::atoi(argv[1]) : 2000;
    if (fl_init(0) != FL_OK) { std::printf("fl_init failed: %s\n", fl_last_error_string()); return 1; }
    const char* mode = std::getenv("FLB_SMALL");
    std::printf("FLB_SMALL=%s\n", mode ? mode : "(default: zero-copy low-latency path)");
    for (int n : {1, 4, 16}) {
        std::vector<uint16_t> v16(size_t(n) * 1024, 3), p16(size_t(n) * 192), u16(size_t(n) * 1024);
        std::vector<uint32_t> v32(size_t(n) * 1024, 77), p32(size_t(n) * 320), u32(size_t(n) * 1024), b32(size_t(n) * 32, 5);
        std::vector<uint64_t> v64(size_t(n) * 1024, 9), p64(size_t(n) * 16 * 33), u64(size_t(n) * 1024);
        std::printf("%2d block(s): fl_host_pack_u16 W=3        %8.2f us\n", n, median_us([&] { fl_host_pack_u16(3, n, v16.data(), p16.data()); }, iters));
        std::printf("%2d block(s): fl_host_unpack_u16 W=3      %8.2f us\n", n, median_us([&] { fl_host_unpack_u16(3, n, p16.data(), u16.data()); }, iters));
        std::printf("%2d block(s): fl_host_pack_u32 W=10       %8.2f us\n", n, median_us([&] { fl_host_pack_u32(10, n, v32.data(), p32.data()); }, iters));
        std::printf("%2d block(s): fl_host_unpack_u32 W=10     %8.2f us\n", n, median_us([&] { fl_host_unpack_u32(10, n, p32.data(), u32.data()); }, iters));
        std::printf("%2d block(s): fl_host_undelta_pack_u32    %8.2f us\n", n, median_us([&] { fl_host_undelta_pack_u32(10, n, p32.data(), b32.data(), u32.data()); }, iters));
        std::printf("%2d block(s): fl_host_unpack_u64 W=33     %8.2f us\n", n, median_us([&] { fl_host_unpack_u64(33, n, p64.data(), u64.data()); }, iters));
        for (size_t i = 0; i < u16.size(); ++i)
            if (u16[i] != 3) { std::printf("MISMATCH u16 at %zu\n", i); return 1; }
    }
    uint16_t one = 0;
    std::vector<uint16_t> v(1024, 5), p(192);
    fl_host_pack_u16(3, 1, v.data(), p.data());
    std::printf("fl_host_unpack_single_u16 W=3            %8.2f us\n", median_us([&] { fl_host_unpack_single_u16(3, p.data(), 517, &one); }, iters));
    if (one != 5) { std::printf("MISMATCH single\n"); return 1; }
    fl_shutdown();
    return 0;
}
