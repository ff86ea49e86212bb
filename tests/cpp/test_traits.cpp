// tests/cpp/test_traits.cpp — the reference's unit tests restated against the C++ trait mirror
// (include/fastlanes_b200.hpp -> C ABI -> sm_100a kernels).  Needs a GPU; run by tests/test_gpu_cpp_traits.py.
//   test_round_trip_*   src/bitpacking.rs:273-315      test_unchecked_pack  :249-256
//   test_unpack_single  src/bitpacking.rs:259-271      test_delta           src/delta.rs:81-107
//   test_ffor           src/ffor.rs:67-88              README example       README.md:14-47
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <utility>
#include <vector>

#include "fastlanes_b200.hpp"

using namespace fastlanes;

static int g_fail = 0;
#define EXPECT(c) do { if (!(c)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); ++g_fail; } } while (0)

template <class T, std::size_t W>
void try_round_trip() {
    std::array<T, 1024> values{};
    for (std::size_t i = 0; i < 1024; ++i) values[i] = T(i % (std::size_t(1) << (W % FastLanes<T>::T_BITS)));
    Packed<T, W> packed{};
    BitPacking<T>::template pack<W>(values, packed);
    std::array<T, 1024> unpacked{};
    unpacked.fill(T(0xEE));
    BitPacking<T>::template unpack<W>(packed, unpacked);
    EXPECT(unpacked == values);
    for (std::size_t i = 0; i < 1024; i += 61) {  // every index is covered by the Python host-path test
        EXPECT(BitPacking<T>::template unpack_single<W>(packed, i) == values[i]);
        EXPECT(BitPacking<T>::unchecked_unpack_single(W, packed.data(), i) == values[i]);
    }
}
template <class T, std::size_t... W>
void round_trips(std::index_sequence<W...>) { (try_round_trip<T, W>(), ...); }

int main() {
    if (fl_device_count() < 1) { std::printf("no CUDA device\n"); return 2; }
    round_trips<uint8_t>(std::make_index_sequence<9>{});
    round_trips<uint16_t>(std::make_index_sequence<17>{});
    round_trips<uint32_t>(std::make_index_sequence<33>{});
    round_trips<uint64_t>(std::make_index_sequence<65>{});
    {   // test_unchecked_pack
        std::array<uint32_t, 1024> input{}, output{};
        for (std::size_t i = 0; i < 1024; ++i) input[i] = uint32_t(i);
        std::array<uint32_t, 320> packed{};
        BitPacking<uint32_t>::unchecked_pack(10, input.data(), 1024, packed.data(), 320);
        BitPacking<uint32_t>::unchecked_unpack(10, packed.data(), 320, output.data(), 1024);
        EXPECT(input == output);
    }
    {   // test_delta
        constexpr std::size_t W = 15;
        std::array<uint16_t, 1024> values{}, transposed{}, deltas{}, unpacked{}, undelta{};
        for (std::size_t i = 0; i < 1024; ++i) values[i] = uint16_t(i / 8);
        Transpose<uint16_t>::transpose(values, transposed);
        Delta<uint16_t>::Base base{};
        Delta<uint16_t>::delta(transposed, base, deltas);
        Packed<uint16_t, W> packed{};
        BitPacking<uint16_t>::pack<W>(deltas, packed);
        Delta<uint16_t>::undelta_pack<W>(packed, base, unpacked);
        EXPECT(transposed == unpacked);
        BitPacking<uint16_t>::unpack<W>(packed, unpacked);
        Delta<uint16_t>::undelta(unpacked, base, undelta);
        EXPECT(transposed == undelta);
        for (std::size_t i = 0; i < 1024; ++i) EXPECT(transposed[i] == values[transpose(i)]);
    }
    {   // fused chains == compositions (u32): encode from original order, decode back to original order
        constexpr std::size_t W = 12;
        std::array<uint32_t, 1024> values{}, transposed{}, deltas{}, decoded{};
        for (std::size_t i = 0; i < 1024; ++i) values[i] = uint32_t(i * 3 + 5);
        Delta<uint32_t>::Base base{};
        Packed<uint32_t, W> fused{}, chained{};
        Delta<uint32_t>::transpose_delta_pack<W>(values, base, fused);
        Transpose<uint32_t>::transpose(values, transposed);
        Delta<uint32_t>::delta(transposed, base, deltas);
        BitPacking<uint32_t>::pack<W>(deltas, chained);
        EXPECT(fused == chained);
        Delta<uint32_t>::undelta_pack_untranspose<W>(fused, base, decoded);
        EXPECT(decoded == values);
    }
    {   // test_ffor
        constexpr std::size_t W = 15;
        std::array<uint16_t, 1024> values{}, unpacked{};
        for (std::size_t i = 0; i < 1024; ++i) values[i] = uint16_t(i % (1 << W));
        Packed<uint16_t, W> packed{};
        FoR<uint16_t>::for_pack<W>(values, 10, packed);
        BitPacking<uint16_t>::unpack<W>(packed, unpacked);
        for (std::size_t i = 0; i < 1024; ++i) EXPECT(uint16_t((values[i] - 10) & ((1 << W) - 1)) == unpacked[i]);
    }
    {   // fused scan == unfor_pack + the caller-side predicate loop (README.md:40-41), u16 W=15 ref=10 (test_ffor's data)
        constexpr std::size_t W = 15;
        std::array<uint16_t, 1024> values{}, decoded{};
        for (std::size_t i = 0; i < 1024; ++i) values[i] = uint16_t(10 + i * 7 % (1 << W));
        Packed<uint16_t, W> packed{};
        FoR<uint16_t>::for_pack<W>(values, 10, packed);
        FoR<uint16_t>::unfor_pack<W>(packed, 10, decoded);
        EXPECT(decoded == values);
        Scan<uint16_t>::Bitmap bitmap{};
        const uint32_t count = Scan<uint16_t>::filter_range<W>(packed, 10, 100, 2000, bitmap);
        uint32_t want = 0;
        for (std::size_t i = 0; i < 1024; ++i) {
            const bool sel = decoded[i] >= 100 && decoded[i] <= 2000;
            want += sel;
            EXPECT(bool((bitmap[i / 8] >> (i % 8)) & 1) == sel);
        }
        EXPECT(count == want);
    }
    {   // delta scan == untranspose(undelta_pack) + predicate loop, on test_delta's data (src/delta.rs:81-107)
        constexpr std::size_t W = 15;
        std::array<uint16_t, 1024> values{}, transposed{}, deltas{};
        for (std::size_t i = 0; i < 1024; ++i) values[i] = uint16_t(i / 8);
        Transpose<uint16_t>::transpose(values, transposed);
        Delta<uint16_t>::Base base{};
        Delta<uint16_t>::delta(transposed, base, deltas);
        Packed<uint16_t, W> packed{};
        BitPacking<uint16_t>::pack<W>(deltas, packed);
        Scan<uint16_t>::Bitmap bitmap{};
        const uint32_t count = Scan<uint16_t>::filter_range_delta<W>(packed, base, 17, 99, bitmap);
        uint32_t want = 0;
        for (std::size_t i = 0; i < 1024; ++i) {
            const bool sel = values[i] >= 17 && values[i] <= 99;
            want += sel;
            EXPECT(bool((bitmap[i / 8] >> (i % 8)) & 1) == sel);
        }
        EXPECT(count == want);
    }
    {   // index >= 1024 panics (src/bitpacking.rs:152)
        Packed<uint16_t, 3> packed{};
        bool threw = false;
        try { (void)BitPacking<uint16_t>::unpack_single<3>(packed, 1024); } catch (const Panic& p) { threw = p.status == FL_ERR_INDEX; }
        EXPECT(threw);
    }
    {   // context family: a whole column sharded over the devices (device 0 listed twice on a 1-GPU box: two shard
        // workers), against the single-block trait calls on every block; round trip through pack
        constexpr std::size_t W = 11, N = 37;
        std::vector<int> devs = fl_device_count() >= 2 ? std::vector<int>{} : std::vector<int>{0, 0};
        Context ctx(devs);
        EXPECT(ctx.device_count() >= 2);
        EXPECT(ctx.block_range(N, 0).first == 0 && ctx.block_range(N, ctx.device_count() - 1).second == N);
        std::vector<uint32_t> values(N * 1024), packed(N * 32 * W), decoded(N * 1024), base(N * 32, 7u), dd(N * 1024);
        for (std::size_t i = 0; i < values.size(); ++i) values[i] = uint32_t((i * 2654435761u) >> 7) & ((1u << W) - 1);
        ctx.unchecked_pack<uint32_t>(W, values, packed);
        ctx.unchecked_unpack<uint32_t>(W, packed, decoded);
        EXPECT(decoded == values);
        ctx.undelta_pack<uint32_t>(W, packed, base, dd);
        for (std::size_t b = 0; b < N; b += 9) {
            std::array<uint32_t, 1024> one{}, ud{};
            Packed<uint32_t, W> pk{};
            Delta<uint32_t>::Base bs{};
            bs.fill(7u);
            std::copy(values.begin() + b * 1024, values.begin() + (b + 1) * 1024, one.begin());
            BitPacking<uint32_t>::pack<W>(one, pk);
            EXPECT(std::equal(pk.begin(), pk.end(), packed.begin() + b * 32 * W));
            Delta<uint32_t>::undelta_pack<W>(pk, bs, ud);
            EXPECT(std::equal(ud.begin(), ud.end(), dd.begin() + b * 1024));
        }
        bool threw = false;
        try { std::vector<uint32_t> bad(5); ctx.unchecked_unpack<uint32_t>(W, bad, decoded); } catch (const Panic& p) { threw = p.status == FL_ERR_LEN; }
        EXPECT(threw);
    }
    std::printf(g_fail ? "FAILED (%d)\n" : "ALL OK\n", g_fail);
    return g_fail ? 1 : 0;
}
