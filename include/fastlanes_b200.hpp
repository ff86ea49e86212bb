// fastlanes_b200.hpp — C++ host-side mirror of the spiraldb/fastlanes trait surface over the C ABI.
//
// The reference is a compiled (Rust) library whose toolchain is absent from this image, so the host
// side above the C-ABI is C++ (bindings/rust/ carries the Rust shim as source).  Names, argument order
// and error behaviour follow the reference traits:
//     BitPacking (src/bitpacking.rs:16-59)  FoR (src/ffor.rs:4-18)
//     Delta      (src/delta.rs:6-17)        Transpose (src/transpose.rs:4-7)
// with the const-generic W as a template parameter and [T; N] as std::array<T, N>.  Where the reference
// panics, fastlanes::Panic is thrown (width > T cannot be expressed: it is a compile error here too,
// via static_assert — the reference's BitPackWidth<W>: SupportedBitPackWidth<T> bound, :8-13).
#pragma once
#include <array>
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "fastlanes_b200.h"

namespace fastlanes {

inline constexpr std::array<std::size_t, 8> FL_ORDER = {0, 4, 2, 6, 1, 5, 3, 7};  // src/lib.rs:22

struct Panic : std::runtime_error {
    fl_status status;
    Panic(fl_status s, const std::string& what) : std::runtime_error(what), status(s) {}
};

namespace detail {
inline void check(fl_status s, const char* what) {
    if (s != FL_OK) throw Panic(s, std::string(what) + ": " + fl_status_string(s) + ": " + fl_last_error_string());
}
template <class T> struct Abi;
#define FLB_ABI(T, SFX)                                                                                         \
    template <> struct Abi<T> {                                                                                  \
        static fl_status pack(unsigned w, size_t n, const T* i, T* o) { return fl_host_pack_##SFX(w, n, i, o); } \
        static fl_status unpack(unsigned w, size_t n, const T* i, T* o) { return fl_host_unpack_##SFX(w, n, i, o); } \
        static fl_status single(unsigned w, const T* p, size_t idx, T* v) { return fl_host_unpack_single_##SFX(w, p, idx, v); } \
        static fl_status for_pack(unsigned w, size_t n, const T* i, T r, T* o) { return fl_host_for_pack_##SFX(w, n, i, r, o); } \
        static fl_status unfor_pack(unsigned w, size_t n, const T* i, T r, T* o) { return fl_host_unfor_pack_##SFX(w, n, i, r, o); } \
        static fl_status delta(size_t n, const T* i, const T* b, T* o) { return fl_host_delta_##SFX(n, i, b, o); } \
        static fl_status undelta(size_t n, const T* i, const T* b, T* o) { return fl_host_undelta_##SFX(n, i, b, o); } \
        static fl_status undelta_pack(unsigned w, size_t n, const T* i, const T* b, T* o) { return fl_host_undelta_pack_##SFX(w, n, i, b, o); } \
        static fl_status undelta_pack_untranspose(unsigned w, size_t n, const T* i, const T* b, T* o) { return fl_host_undelta_pack_untranspose_##SFX(w, n, i, b, o); } \
        static fl_status transpose_delta_pack(unsigned w, size_t n, const T* i, const T* b, T* o) { return fl_host_transpose_delta_pack_##SFX(w, n, i, b, o); } \
        static fl_status filter(unsigned w, size_t n, const T* i, T r, T lo, T hi, uint8_t* bm, uint32_t* c) { return fl_host_unpack_filter_##SFX(w, n, i, r, lo, hi, bm, c); } \
        static fl_status delta_filter(unsigned w, size_t n, const T* i, const T* b, T lo, T hi, uint8_t* bm, uint32_t* c) { return fl_host_undelta_pack_filter_##SFX(w, n, i, b, lo, hi, bm, c); } \
        static fl_status transpose(size_t n, const T* i, T* o) { return fl_host_transpose_##SFX(n, i, o); }     \
        static fl_status untranspose(size_t n, const T* i, T* o) { return fl_host_untranspose_##SFX(n, i, o); } \
        static fl_status ctx_pack(fl_ctx* c, unsigned w, size_t n, const T* i, T* o) { return fl_ctx_host_pack_##SFX(c, w, n, i, o); } \
        static fl_status ctx_unpack(fl_ctx* c, unsigned w, size_t n, const T* i, T* o) { return fl_ctx_host_unpack_##SFX(c, w, n, i, o); } \
        static fl_status ctx_for_pack(fl_ctx* c, unsigned w, size_t n, const T* i, T r, T* o) { return fl_ctx_host_for_pack_##SFX(c, w, n, i, r, o); } \
        static fl_status ctx_unfor_pack(fl_ctx* c, unsigned w, size_t n, const T* i, T r, T* o) { return fl_ctx_host_unfor_pack_##SFX(c, w, n, i, r, o); } \
        static fl_status ctx_undelta_pack(fl_ctx* c, unsigned w, size_t n, const T* i, const T* b, T* o) { return fl_ctx_host_undelta_pack_##SFX(c, w, n, i, b, o); } \
        static fl_status ctx_filter(fl_ctx* c, unsigned w, size_t n, const T* i, T r, T lo, T hi, uint8_t* bm, uint32_t* cn) { return fl_ctx_host_unpack_filter_##SFX(c, w, n, i, r, lo, hi, bm, cn); } \
    };
FLB_ABI(uint8_t, u8)
FLB_ABI(uint16_t, u16)
FLB_ABI(uint32_t, u32)
FLB_ABI(uint64_t, u64)
#undef FLB_ABI
}  // namespace detail

// trait FastLanes (src/lib.rs:24-27)
template <class T>
struct FastLanes {
    static constexpr std::size_t T_BITS = sizeof(T) * 8;
    static constexpr std::size_t LANES = 1024 / T_BITS;
};

template <class T, std::size_t W>
using Packed = std::array<T, 1024 * W / FastLanes<T>::T_BITS>;  // [Self; 1024 * W / Self::T]

// trait BitPacking (src/bitpacking.rs:16-59)
template <class T>
struct BitPacking {
    template <std::size_t W>
    static void pack(const std::array<T, 1024>& input, Packed<T, W>& output) {
        static_assert(W <= FastLanes<T>::T_BITS, "BitPackWidth<W>: SupportedBitPackWidth<T>");
        detail::check(detail::Abi<T>::pack(W, 1, input.data(), output.data()), "pack");
    }
    static void unchecked_pack(std::size_t width, const T* input, std::size_t in_len, T* output, std::size_t out_len) {
        if (in_len != 1024 || out_len != 128 * width / sizeof(T)) throw Panic(FL_ERR_LEN, "Output buffer must be of size 1024 * W / T");
        detail::check(detail::Abi<T>::pack(unsigned(width), 1, input, output), "unchecked_pack");
    }
    template <std::size_t W>
    static void unpack(const Packed<T, W>& input, std::array<T, 1024>& output) {
        static_assert(W <= FastLanes<T>::T_BITS, "BitPackWidth<W>: SupportedBitPackWidth<T>");
        detail::check(detail::Abi<T>::unpack(W, 1, input.data(), output.data()), "unpack");
    }
    static void unchecked_unpack(std::size_t width, const T* input, std::size_t in_len, T* output, std::size_t out_len) {
        if (out_len != 1024 || in_len != 128 * width / sizeof(T)) throw Panic(FL_ERR_LEN, "Input buffer must be of size 1024 * W / T");
        detail::check(detail::Abi<T>::unpack(unsigned(width), 1, input, output), "unchecked_unpack");
    }
    template <std::size_t W>
    static T unpack_single(const Packed<T, W>& packed, std::size_t index) {
        T v{};
        detail::check(detail::Abi<T>::single(W, packed.data(), index, &v), "unpack_single");
        return v;
    }
    static T unchecked_unpack_single(std::size_t width, const T* packed, std::size_t index) {
        T v{};
        detail::check(detail::Abi<T>::single(unsigned(width), packed, index, &v), "unchecked_unpack_single");
        return v;
    }
};

// trait FoR (src/ffor.rs:4-18)
template <class T>
struct FoR {
    template <std::size_t W>
    static void for_pack(const std::array<T, 1024>& input, T reference, Packed<T, W>& output) {
        detail::check(detail::Abi<T>::for_pack(W, 1, input.data(), reference, output.data()), "for_pack");
    }
    template <std::size_t W>
    static void unfor_pack(const Packed<T, W>& input, T reference, std::array<T, 1024>& output) {
        detail::check(detail::Abi<T>::unfor_pack(W, 1, input.data(), reference, output.data()), "unfor_pack");
    }
};

// trait Delta (src/delta.rs:6-17)
template <class T>
struct Delta {
    using Base = std::array<T, FastLanes<T>::LANES>;
    static void delta(const std::array<T, 1024>& input, const Base& base, std::array<T, 1024>& output) {
        detail::check(detail::Abi<T>::delta(1, input.data(), base.data(), output.data()), "delta");
    }
    static void undelta(const std::array<T, 1024>& input, const Base& base, std::array<T, 1024>& output) {
        detail::check(detail::Abi<T>::undelta(1, input.data(), base.data(), output.data()), "undelta");
    }
    template <std::size_t W>
    static void undelta_pack(const Packed<T, W>& input, const Base& base, std::array<T, 1024>& output) {
        detail::check(detail::Abi<T>::undelta_pack(W, 1, input.data(), base.data(), output.data()), "undelta_pack");
    }
    // Fused chains (compositions of the reference methods, one GPU pass; u32/u64):
    //   untranspose(undelta_pack::<W>(input, base))  — src/delta.rs:99 + src/transpose.rs:18-22
    template <std::size_t W>
    static void undelta_pack_untranspose(const Packed<T, W>& input, const Base& base, std::array<T, 1024>& output) {
        detail::check(detail::Abi<T>::undelta_pack_untranspose(W, 1, input.data(), base.data(), output.data()), "undelta_pack_untranspose");
    }
    //   pack::<W>(delta(transpose(input), base))      — src/delta.rs:88-95
    template <std::size_t W>
    static void transpose_delta_pack(const std::array<T, 1024>& input, const Base& base, Packed<T, W>& output) {
        detail::check(detail::Abi<T>::transpose_delta_pack(W, 1, input.data(), base.data(), output.data()), "transpose_delta_pack");
    }
};

// Fused scan (not a trait of the reference: FoR::unfor_pack + the caller-side loop of README.md:40-41 in one GPU
// pass).  Returns the number of selected values; bit i of `bitmap` = lo <= unfor_pack::<W>(input, reference)[i] <= hi.
template <class T>
struct Scan {
    using Bitmap = std::array<uint8_t, 128>;
    template <std::size_t W>
    static uint32_t filter_range(const Packed<T, W>& input, T reference, T lo, T hi, Bitmap& bitmap) {
        static_assert(W <= FastLanes<T>::T_BITS, "BitPackWidth<W>: SupportedBitPackWidth<T>");
        uint32_t count = 0;
        detail::check(detail::Abi<T>::filter(W, 1, input.data(), reference, lo, hi, bitmap.data(), &count), "filter_range");
        return count;
    }
    // bit i = lo <= untranspose(undelta_pack::<W>(input, base))[i] <= hi  (src/delta.rs:48-63, src/transpose.rs:18-22)
    template <std::size_t W>
    static uint32_t filter_range_delta(const Packed<T, W>& input, const typename Delta<T>::Base& base, T lo, T hi, Bitmap& bitmap) {
        static_assert(W <= FastLanes<T>::T_BITS, "BitPackWidth<W>: SupportedBitPackWidth<T>");
        uint32_t count = 0;
        detail::check(detail::Abi<T>::delta_filter(W, 1, input.data(), base.data(), lo, hi, bitmap.data(), &count), "filter_range_delta");
        return count;
    }
};

// trait Transpose + const fn transpose (src/transpose.rs:4-7, :29-36)
template <class T>
struct Transpose {
    static void transpose(const std::array<T, 1024>& input, std::array<T, 1024>& output) {
        detail::check(detail::Abi<T>::transpose(1, input.data(), output.data()), "transpose");
    }
    static void untranspose(const std::array<T, 1024>& input, std::array<T, 1024>& output) {
        detail::check(detail::Abi<T>::untranspose(1, input.data(), output.data()), "untranspose");
    }
};
constexpr std::size_t transpose(std::size_t idx) { return (idx % 16) * 64 + FL_ORDER[(idx / 16) % 8] * 8 + idx / 128; }

// Multi-GPU in one process (fl_ctx, include/fastlanes_b200.h): batched forms of the runtime-width family
// (src/bitpacking.rs:109-129) over whole columns, contiguous block shards on the context's devices.
class Context {
  public:
    explicit Context(const std::vector<int>& devices = {}) {
        detail::check(fl_ctx_create(devices.empty() ? nullptr : devices.data(), int(devices.size()), &ctx_), "fl_ctx_create");
    }
    ~Context() { fl_ctx_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    int device_count() const { return fl_ctx_device_count(ctx_); }
    std::pair<std::size_t, std::size_t> block_range(std::size_t n_blocks, int i) const {
        std::size_t a = 0, b = 0;
        detail::check(fl_ctx_block_range(ctx_, n_blocks, i, &a, &b), "fl_ctx_block_range");
        return {a, b};
    }
    // slices hold whole blocks: unpacked n*1024 elements, packed n*1024*W/T (the reference's debug_asserts, :78-80,:111-113)
    template <class T>
    void unchecked_pack(std::size_t width, const std::vector<T>& input, std::vector<T>& output) {
        const std::size_t n = blocks_of<T>(input, output, width);
        detail::check(detail::Abi<T>::ctx_pack(ctx_, unsigned(width), n, input.data(), output.data()), "ctx pack");
    }
    template <class T>
    void unchecked_unpack(std::size_t width, const std::vector<T>& input, std::vector<T>& output) {
        const std::size_t n = blocks_of<T>(output, input, width);
        detail::check(detail::Abi<T>::ctx_unpack(ctx_, unsigned(width), n, input.data(), output.data()), "ctx unpack");
    }
    template <class T>
    void for_pack(std::size_t width, const std::vector<T>& input, T reference, std::vector<T>& output) {
        const std::size_t n = blocks_of<T>(input, output, width);
        detail::check(detail::Abi<T>::ctx_for_pack(ctx_, unsigned(width), n, input.data(), reference, output.data()), "ctx for_pack");
    }
    template <class T>
    void unfor_pack(std::size_t width, const std::vector<T>& input, T reference, std::vector<T>& output) {
        const std::size_t n = blocks_of<T>(output, input, width);
        detail::check(detail::Abi<T>::ctx_unfor_pack(ctx_, unsigned(width), n, input.data(), reference, output.data()), "ctx unfor_pack");
    }
    template <class T>
    void undelta_pack(std::size_t width, const std::vector<T>& input, const std::vector<T>& base, std::vector<T>& output) {
        const std::size_t n = blocks_of<T>(output, input, width);
        if (base.size() != n * (1024 / (sizeof(T) * 8))) throw Panic(FL_ERR_LEN, "Base buffer must hold LANES elements per block");
        detail::check(detail::Abi<T>::ctx_undelta_pack(ctx_, unsigned(width), n, input.data(), base.data(), output.data()), "ctx undelta_pack");
    }
    fl_ctx* raw() { return ctx_; }

  private:
    template <class T>
    static std::size_t blocks_of(const std::vector<T>& unpacked, const std::vector<T>& packed, std::size_t width) {
        if (unpacked.size() % 1024) throw Panic(FL_ERR_LEN, "unpacked buffer must be a whole number of 1024-element blocks");
        const std::size_t n = unpacked.size() / 1024;
        if (width <= sizeof(T) * 8 && packed.size() != n * 1024 * width / (sizeof(T) * 8))
            throw Panic(FL_ERR_LEN, "packed buffer must be of size n * 1024 * W / T");
        return n;
    }
    fl_ctx* ctx_ = nullptr;
};

}  // namespace fastlanes
