// CPU check of the bit arithmetic of the fused scan kernels (fastlanes_b200/csrc/fl_scan_bits.h): emulates the 32
// threads of a warp in lockstep (phase by phase, shuffles = array reads) and compares the assembled 128-byte block
// bitmap with a brute-force one built from index(row, lane) = FL_ORDER[row/8]*16 + (row%8)*128 + lane
// (/root/reference src/macros.rs:20-24).  No CUDA needed: g++ -std=c++17 -Ifastlanes_b200/csrc.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "fl_scan_bits.h"

static const int FL_ORDER[8] = {0, 4, 2, 6, 1, 5, 3, 7};
static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint64_t rnd() {
    uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

template <int TBITS>
static int run(int trials) {
    constexpr int L = 1024 / TBITS, RPG = TBITS / 4, BPT = 128 / TBITS;
    int bad = 0;
    for (int t = 0; t < trials; ++t) {
        std::vector<unsigned char> pred(TBITS * L);
        const int density = t % 5;  // 0: random, 1: sparse, 2: dense, 3: all, 4: none
        for (auto& p : pred) {
            const uint64_t r = rnd();
            p = density == 0 ? (r & 1) : density == 1 ? ((r & 31) == 0) : density == 2 ? ((r & 31) != 0) : density == 3;
        }
        unsigned char expect[128] = {0}, tile[128];
        std::memset(tile, 0xAA, sizeof tile);  // every byte must be overwritten
        for (int r = 0; r < TBITS; ++r)
            for (int l = 0; l < L; ++l)
                if (pred[r * L + l]) {
                    const int idx = FL_ORDER[r / 8] * 16 + (r % 8) * 128 + l;
                    expect[idx >> 3] |= (unsigned char)(1u << (idx & 7));
                }
        uint32_t X[32], Y[32], Z[32];
        int Q[32];
        for (int th = 0; th < 32; ++th) {
            const int g = th >> 3, j = th & 7;
            const int q = (TBITS >= 32) ? (g == 1 ? 2 : (g == 2 ? 1 : g)) : g;  // WarpLay<T>::rank_of_group
            Q[th] = q;
            uint32_t x = 0;
            for (int i = 0; i < RPG; ++i)
                for (int k = 0; k < BPT; ++k)
                    if (pred[(q * RPG + i) * L + j * BPT + k]) x |= 1u << (i * BPT + k);
            X[th] = x;
        }
        for (int th = 0; th < 32; ++th) {
            const int j = th & 7;
            if (TBITS == 32) Z[th] = flb::merge_pair_bpt4(X[th], X[th ^ 1], j);
            else if (TBITS == 64) Y[th] = flb::merge_pair_bpt2(X[th], X[th ^ 1], j);
            else Z[th] = X[th];
        }
        if (TBITS == 64)
            for (int th = 0; th < 32; ++th) Z[th] = flb::merge_quad_bpt2(Y[th], Y[th ^ 2], th & 7);
        for (int th = 0; th < 32; ++th) flb::scan_store<TBITS>(tile, Q[th], th & 7, Z[th]);
        if (std::memcmp(tile, expect, 128) != 0) {
            if (!bad) std::printf("u%d trial %d: bitmap mismatch\n", TBITS, t);
            ++bad;
        }
    }
    return bad;
}

int main() {
    int bad = 0;
    // mask -> bits helpers
    for (int m = 0; m < 16; ++m) {
        uint32_t w = 0;
        for (int k = 0; k < 4; ++k) if (m >> k & 1) w |= 0xFFu << (8 * k);
        if (flb::mask_bytes_to_bits(w) != (uint32_t)m) { std::printf("mask_bytes_to_bits(%08x)\n", w); ++bad; }
    }
    for (int m = 0; m < 4; ++m) {
        uint32_t w = 0;
        for (int k = 0; k < 2; ++k) if (m >> k & 1) w |= 0xFFFFu << (16 * k);
        if (flb::mask_halves_to_bits(w) != (uint32_t)m) { std::printf("mask_halves_to_bits(%08x)\n", w); ++bad; }
    }
    bad += run<8>(200);
    bad += run<16>(200);
    bad += run<32>(200);
    bad += run<64>(200);
    std::printf(bad ? "FAILED (%d)\n" : "scan bits ok\n", bad);
    return bad ? 1 : 0;
}
