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

// delta scan: predicate given in ORIGINAL order; the warp holds it in transposed order (value (row, lane) of the
// unpacked vector is original[t(index(row, lane))], src/transpose.rs:29-36), lane-major bits per thread.
static int transpose_index(int i) { return (i % 16) * 64 + FL_ORDER[(i / 16) % 8] * 8 + i / 128; }

template <int TBITS>
static int run_orig(int trials) {
    constexpr int RPG = TBITS / 4, BPT = 128 / TBITS;
    int bad = 0;
    for (int t = 0; t < trials; ++t) {
        unsigned char pred[1024];
        const int density = t % 5;
        for (auto& p : pred) {
            const uint64_t r = rnd();
            p = density == 0 ? (r & 1) : density == 1 ? ((r & 31) == 0) : density == 2 ? ((r & 31) != 0) : density == 3;
        }
        unsigned char expect[128] = {0}, tile[128];
        std::memset(tile, 0xAA, sizeof tile);
        for (int i = 0; i < 1024; ++i) if (pred[i]) expect[i >> 3] |= (unsigned char)(1u << (i & 7));
        uint32_t X[32], Y[32], Z[32];
        int Q[32];
        for (int th = 0; th < 32; ++th) {
            const int g = th >> 3, j = th & 7;
            const int q = (TBITS >= 32) ? (g == 1 ? 2 : (g == 2 ? 1 : g)) : g;
            Q[th] = q;
            uint32_t x = 0;
            for (int k = 0; k < BPT; ++k)
                for (int i = 0; i < RPG; ++i) {
                    const int row = q * RPG + i, lane = j * BPT + k;
                    const int idx = FL_ORDER[row / 8] * 16 + (row % 8) * 128 + lane;  // position in the unpacked vector
                    if (pred[transpose_index(idx)]) x |= 1u << (k * RPG + i);
                }
            X[th] = x;
        }
        // u8 / u16: rank == group, so rank q^1 is thread th^8 and rank q^2 is thread th^16
        for (int th = 0; th < 32; ++th) {
            if (TBITS == 16) Z[th] = flb::merge_pair_bpt4(X[th], X[th ^ 8], Q[th]);
            else if (TBITS == 8) Y[th] = flb::merge_pair_bpt2(X[th], X[th ^ 8], Q[th]);
            else Z[th] = X[th];
        }
        if (TBITS == 8)
            for (int th = 0; th < 32; ++th) Z[th] = flb::merge_quad_bpt2(Y[th], Y[th ^ 16], Q[th]);
        for (int th = 0; th < 32; ++th) flb::scan_store_orig<TBITS>(tile, Q[th], th & 7, Z[th]);
        if (std::memcmp(tile, expect, 128) != 0) {
            if (!bad) std::printf("u%d trial %d: original-order bitmap mismatch\n", TBITS, t);
            ++bad;
        }
    }
    return bad;
}

// select: the compaction of select_warp_kernel on an emulated warp.  Values are their own original index, so the staged
// output must be the ascending list of the selected indices.  Uses the kernel's helpers for the slice position, the
// rotation that puts bit k of the slice at register position k + 1, and the rank inside the word.
template <int TBITS>
static int run_select(int trials) {
    const int L = 1024 / TBITS, RPG = TBITS / 4, BPT = 128 / TBITS, BANDS = RPG > 8 ? RPG / 8 : 1;
    int bad = 0;
    for (int t = 0; t < trials; ++t) {
        uint32_t bitmap[32];
        const int density = t % 5;  // 0: empty, 1: sparse, 2: quarter, 3: half, 4: full
        for (int w = 0; w < 32; ++w) {
            const uint32_t a = (uint32_t)rnd(), b = (uint32_t)rnd(), c = (uint32_t)rnd();
            bitmap[w] = density == 0 ? 0u : density == 1 ? (a & b & c & (uint32_t)rnd()) : density == 2 ? (a & b) : density == 3 ? a : ~0u;
        }
        uint32_t prefix[32], total = 0;
        for (int w = 0; w < 32; ++w) { prefix[w] = total; total += (uint32_t)__builtin_popcount(bitmap[w]); }
        std::vector<int> stage(1024 + 32, -1);
        for (int th = 0; th < 32; ++th) {
            const int g = th >> 3, j = th & 7;
            const int q = (TBITS >= 32) ? (((g << 1) & 2) | (g >> 1)) : g;  // WarpLay<T>::rank_of_group
            for (int i = 0; i < RPG; ++i) {
                const int band = BANDS > 1 ? i / 8 : 0;
                const int c0 = flb::select_band_origin<TBITS>(q, band, j);
                const int word = (c0 >> 5) + 4 * (i % 8);
                const uint32_t s0 = uint32_t(c0 & 31);
                if (int(s0) + BPT > 32 || word > 31) { if (!bad) std::printf("u%d: slice leaves its bitmap word\n", TBITS); ++bad; continue; }
                const int r = q * RPG + i;
                const int first = FL_ORDER[r / 8] * 16 + (r % 8) * 128 + j * BPT;  // index(r, j*BPT), macros.rs:20-24
                if (first != word * 32 + int(s0)) { if (!bad) std::printf("u%d: slice position mismatch\n", TBITS); ++bad; continue; }
                const uint32_t bits = flb::select_rotate(bitmap[word], flb::select_rotation(s0));
                uint32_t rank = prefix[word] + (uint32_t)__builtin_popcount(bitmap[word] & flb::select_low_mask(s0));
                for (int k = 0; k < BPT; ++k)
                    if (bits & (2u << k)) stage[rank++] = first + k;
            }
        }
        uint32_t n = 0;
        for (int idx = 0; idx < 1024; ++idx)
            if (bitmap[idx >> 5] >> (idx & 31) & 1u) { if (stage[n] != idx) { if (!bad) std::printf("u%d trial %d: selected value %u out of order\n", TBITS, t, n); ++bad; break; } ++n; }
        if (n == total && stage[total] != -1) { if (!bad) std::printf("u%d trial %d: wrote past the selected count\n", TBITS, t); ++bad; }
        (void)L;
    }
    return bad;
}

int main() {
    int bad = 0;
    bad += run_select<8>(100);
    bad += run_select<16>(100);
    bad += run_select<32>(100);
    bad += run_select<64>(100);
    // SWAR lane-wise x <= y and the top-bit compression: u8 exhaustive over (x, y) in every lane position with
    // random neighbours; u16 edge values x random
    for (int x = 0; x < 256; ++x)
        for (int y = 0; y < 256; ++y) {
            const uint32_t nx = (uint32_t)rnd(), ny = (uint32_t)rnd();
            for (int k = 0; k < 4; ++k) {
                const uint32_t X = (nx & ~(0xFFu << (8 * k))) | ((uint32_t)x << (8 * k));
                const uint32_t Y = (ny & ~(0xFFu << (8 * k))) | ((uint32_t)y << (8 * k));
                const uint32_t bits = flb::top_bits_u8(flb::swar_leu_top<8>(X, Y, Y | 0x80808080u));
                uint32_t want = 0;
                for (int l = 0; l < 4; ++l) if (((X >> (8 * l)) & 0xFF) <= ((Y >> (8 * l)) & 0xFF)) want |= 1u << l;
                if (bits != want) { if (bad < 5) std::printf("u8 leu X=%08x Y=%08x got %x want %x\n", X, Y, bits, want); ++bad; }
            }
        }
    {
        const uint32_t edge[] = {0, 1, 2, 0x7FFE, 0x7FFF, 0x8000, 0x8001, 0xFFFE, 0xFFFF, 0x1234, 0xABCD};
        for (int t = 0; t < 200000; ++t) {
            uint32_t X = (uint32_t)rnd(), Y = (uint32_t)rnd();
            if (t % 3 == 0) X = (X & 0xFFFF0000u) | edge[rnd() % 11];
            if (t % 5 == 0) Y = (Y & 0x0000FFFFu) | (edge[rnd() % 11] << 16);
            if (t % 7 == 0) Y = X;
            if (t % 11 == 0) X = (X & 0xFFFFu) | (Y & 0xFFFF0000u);
            const uint32_t bits = flb::top_bits_u16(flb::swar_leu_top<16>(X, Y, Y | 0x80008000u));
            const uint32_t want = uint32_t((X & 0xFFFF) <= (Y & 0xFFFF)) | (uint32_t((X >> 16) <= (Y >> 16)) << 1);
            if (bits != want) { if (bad < 5) std::printf("u16 leu X=%08x Y=%08x got %x want %x\n", X, Y, bits, want); ++bad; }
        }
    }
    // lane-major top-bit placement
    for (int m = 0; m < 16; ++m) {
        uint32_t le = (uint32_t)rnd() & 0x7F7F7F7Fu, want = 0;
        for (int k = 0; k < 4; ++k) if (m >> k & 1) { le |= 0x80u << (8 * k); want |= 1u << (2 * k); }
        if (flb::top_bits_lane_major_u8(le) != want) { std::printf("top_bits_lane_major_u8(%08x)\n", le); ++bad; }
    }
    for (int m = 0; m < 4; ++m) {
        uint32_t le = (uint32_t)rnd() & 0x7FFF7FFFu, want = 0;
        for (int k = 0; k < 2; ++k) if (m >> k & 1) { le |= 0x8000u << (16 * k); want |= 1u << (4 * k); }
        if (flb::top_bits_lane_major_u16(le) != want) { std::printf("top_bits_lane_major_u16(%08x)\n", le); ++bad; }
    }
    bad += run_orig<8>(200);
    bad += run_orig<16>(200);
    bad += run_orig<32>(200);
    bad += run_orig<64>(200);
    bad += run<8>(200);
    bad += run<16>(200);
    bad += run<32>(200);
    bad += run<64>(200);
    std::printf(bad ? "FAILED (%d)\n" : "scan bits ok\n", bad);
    return bad ? 1 : 0;
}
