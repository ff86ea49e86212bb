// fl_device.cuh — layout constants and per-thread SWAR primitives of the B200 FastLanes codec.
//
// Wire format (spiraldb/fastlanes v0.1.8; citations relative to /root/reference):
//   * a block is 1024 values of T bits, LANES = 1024/T (src/lib.rs:24-27);
//   * value i lives in lane i%LANES at row FL_ORDER[..]*8+s (src/bitpacking.rs:207-232), i.e. row r of
//     lane l is value index(r,l) = FL_ORDER[r/8]*16 + (r%8)*128 + l (src/macros.rs:20-24);
//   * lane l's bit-stream is the concatenation of packed[LANES*k + l], k = 0..W-1, LSB first, and row r
//     occupies bits [r*W, r*W+W) (src/macros.rs:35-97, 101-173).
//
// B200 mapping ("row-slice" layout).  For EVERY T a packed word-row k is 128 contiguous bytes and an
// unpacked row r is 128 contiguous bytes (LANES*T/8 = 128).  A thread owns one 16-byte column slice of
// every row of one block: 8 threads cover a block, a warp covers 4 blocks, every global access is a
// 128-bit LDG/STG and every quarter-warp touches exactly one full 128-byte line.  The 16 bytes hold
// 16/8/4/2 lanes of u8/u16/u32/u64, processed SWAR in 32-bit (u8/u16/u32) or 64-bit (u64) registers, so
// all shifts, masks and word indices are compile-time immediates (template<W>, rows fully unrolled)
// exactly as the reference's seq_t! unrolling makes them (src/lib.rs:41-47).
#pragma once
#include <cstddef>
#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>

namespace flb {

// src/lib.rs:22
__host__ __device__ constexpr int fl_order(int i) {
    constexpr int o[8] = {0, 4, 2, 6, 1, 5, 3, 7};
    return o[i];
}

// run-time FL_ORDER lookup without a memory table: nibble i of 0x73516240 is FL_ORDER[i]
__device__ __forceinline__ int fl_order_rt(int i) { return int((0x73516240u >> (4 * i)) & 7u); }

template <class T>
struct Lay {
    static constexpr int TB = int(sizeof(T)) * 8;  // src/lib.rs:25
    static constexpr int L = 1024 / TB;            // src/lib.rs:26
    using R = typename std::conditional<sizeof(T) == 8, uint64_t, uint32_t>::type;  // SWAR register
    static constexpr int RB = int(sizeof(R)) * 8;
    static constexpr int NR = 16 / int(sizeof(R));       // registers per 16-byte slice (4 or 2)
    static constexpr int LPR = int(sizeof(R) / sizeof(T));  // lanes per register (4,2,1,1)
};

// Byte offset of unpacked row `row` inside a block: index(row, 0) * sizeof(T)  (src/macros.rs:20-24).
template <class T>
__host__ __device__ constexpr int row_byte_offset(int row) {
    return (fl_order(row / 8) * 16 + (row % 8) * 128) * int(sizeof(T));
}

// Low-`n`-bit mask of a T-bit lane, replicated into every lane of the SWAR register.
template <class T>
__host__ __device__ constexpr typename Lay<T>::R rep_mask(int n) {
    using R = typename Lay<T>::R;
    constexpr int TB = Lay<T>::TB;
    const uint64_t lane = (n >= 64) ? ~uint64_t(0) : ((uint64_t(1) << n) - 1);
    uint64_t r = 0;
    for (int i = 0; i < Lay<T>::LPR; ++i) r |= lane << (i * TB);
    return R(r);
}

// Broadcast one T value into every lane of the SWAR register.
template <class T>
__host__ __device__ constexpr typename Lay<T>::R rep_value(T v) {
    using R = typename Lay<T>::R;
    uint64_t r = 0;
    for (int i = 0; i < Lay<T>::LPR; ++i) r |= uint64_t(v) << (i * Lay<T>::TB);
    return R(r);
}

// 16-byte slice of a row.
template <class T>
struct Slice {
    typename Lay<T>::R r[Lay<T>::NR];
};

// ---- 128-bit global accesses.  Data is touched exactly once: bypass L1, mark streaming. ----------
// FLB_LD_MODE / FLB_ST_MODE select the cache hints (tools/kbench.cu sweeps them; defaults = measured best).
#ifndef FLB_LD_MODE
#define FLB_LD_MODE 0
#endif
#ifndef FLB_ST_MODE
#define FLB_ST_MODE 0
#endif
__device__ __forceinline__ uint4 ldg128_stream(const void* p) {
    uint4 v;
#if FLB_LD_MODE == 0
    asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];"
#elif FLB_LD_MODE == 1
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];"
#elif FLB_LD_MODE == 2
    asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];"
#else
    asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4];"
#endif
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void stg128_stream(void* p, uint4 v) {
#if FLB_ST_MODE == 0
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};"
#elif FLB_ST_MODE == 1
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};"
#elif FLB_ST_MODE == 2
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
#else
    asm volatile("st.global.wt.v4.u32 [%0], {%1,%2,%3,%4};"
#endif
                 ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}

// ---- TMA 1-D bulk copy global -> shared::cta completing on an mbarrier (SASS: UBLKCP + SYNCS) ----------------
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n .reg .pred p;\n FLB_WAIT:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra FLB_DONE;\n bra FLB_WAIT;\n FLB_DONE:\n}"
        ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_load(unsigned dst_smem, const void* src_global, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src_global), "r"(bytes), "r"(bar) : "memory");
}

template <class T>
__device__ __forceinline__ Slice<T> to_slice(uint4 v) {
    Slice<T> s;
    if constexpr (sizeof(T) == 8) {
        s.r[0] = (uint64_t(v.y) << 32) | v.x;
        s.r[1] = (uint64_t(v.w) << 32) | v.z;
    } else {
        s.r[0] = v.x; s.r[1] = v.y; s.r[2] = v.z; s.r[3] = v.w;
    }
    return s;
}
template <class T>
__device__ __forceinline__ uint4 from_slice(const Slice<T>& s) {
    if constexpr (sizeof(T) == 8) {
        return make_uint4(uint32_t(s.r[0]), uint32_t(s.r[0] >> 32), uint32_t(s.r[1]), uint32_t(s.r[1] >> 32));
    } else {
        return make_uint4(s.r[0], s.r[1], s.r[2], s.r[3]);
    }
}
template <class T>
__device__ __forceinline__ Slice<T> load_slice(const void* p) { return to_slice<T>(ldg128_stream(p)); }
template <class T>
__device__ __forceinline__ void store_slice(void* p, const Slice<T>& s) { stg128_stream(p, from_slice<T>(s)); }

// ---- lane-wise wrapping add / sub (the reference's wrapping_add / wrapping_sub, delta.rs, ffor.rs) ----
template <class T>
__device__ __forceinline__ typename Lay<T>::R lane_add(typename Lay<T>::R a, typename Lay<T>::R b) {
    using R = typename Lay<T>::R;
    if constexpr (Lay<T>::LPR == 1) {
        return a + b;
    } else if constexpr (Lay<T>::LPR == 2) {
        return __vadd2(a, b);  // sm_100a: one VIADD.16x2
    } else {
        constexpr R H = rep_value<T>(T(T(1) << (Lay<T>::TB - 1)));
        return ((a & ~H) + (b & ~H)) ^ ((a ^ b) & H);
    }
}
template <class T>
__device__ __forceinline__ typename Lay<T>::R lane_sub(typename Lay<T>::R a, typename Lay<T>::R b) {
    using R = typename Lay<T>::R;
    if constexpr (Lay<T>::LPR == 1) {
        return a - b;
    } else if constexpr (Lay<T>::LPR == 2) {
        return __vsub2(a, b);  // sm_100a: one VIADD.16x2 with a negated operand
    } else {
        constexpr R H = rep_value<T>(T(T(1) << (Lay<T>::TB - 1)));
        return ((a | H) - (b & ~H)) ^ ((a ^ ~b) & H);
    }
}
template <class T>
__device__ __forceinline__ Slice<T> slice_add(const Slice<T>& a, const Slice<T>& b) {
    Slice<T> o;
#pragma unroll
    for (int i = 0; i < Lay<T>::NR; ++i) o.r[i] = lane_add<T>(a.r[i], b.r[i]);
    return o;
}
template <class T>
__device__ __forceinline__ Slice<T> slice_sub(const Slice<T>& a, const Slice<T>& b) {
    Slice<T> o;
#pragma unroll
    for (int i = 0; i < Lay<T>::NR; ++i) o.r[i] = lane_sub<T>(a.r[i], b.r[i]);
    return o;
}
template <class T>
__device__ __forceinline__ Slice<T> slice_splat(T v) {
    Slice<T> o;
    typename Lay<T>::R x = typename Lay<T>::R(v);
    if constexpr (Lay<T>::LPR == 2) x |= x << 16;
    if constexpr (Lay<T>::LPR == 4) { x |= x << 8; x |= x << 16; }
#pragma unroll
    for (int i = 0; i < Lay<T>::NR; ++i) o.r[i] = x;
    return o;
}
template <class T>
__device__ __forceinline__ Slice<T> slice_zero() {
    Slice<T> o;
#pragma unroll
    for (int i = 0; i < Lay<T>::NR; ++i) o.r[i] = 0;
    return o;
}

// ---- lane-wise unsigned min / max on the SWAR register (u16: one VIMNMX.U16x2) --------------------------------
template <class T> struct MinMax;
template <> struct MinMax<uint8_t> {
    __device__ static uint32_t mn(uint32_t a, uint32_t b) { return __vminu4(a, b); }
    __device__ static uint32_t mx(uint32_t a, uint32_t b) { return __vmaxu4(a, b); }
};
template <> struct MinMax<uint16_t> {
    __device__ static uint32_t mn(uint32_t a, uint32_t b) { return __vminu2(a, b); }
    __device__ static uint32_t mx(uint32_t a, uint32_t b) { return __vmaxu2(a, b); }
};
template <> struct MinMax<uint32_t> {
    __device__ static uint32_t mn(uint32_t a, uint32_t b) { return min(a, b); }
    __device__ static uint32_t mx(uint32_t a, uint32_t b) { return max(a, b); }
};
template <> struct MinMax<uint64_t> {
    __device__ static uint64_t mn(uint64_t a, uint64_t b) { return min(a, b); }
    __device__ static uint64_t mx(uint64_t a, uint64_t b) { return max(a, b); }
};
// u8 has no 8x4 min/max instruction (__vminu4 / __vmaxu4 expand to ~8 instructions each).  Split a register into its
// even and odd bytes as two 16x2 vectors (2 PRMT) and use the native VIMNMX.U16x2: 6 instructions per register for
// min AND max.  Accumulators start at mn = 0x00FF00FF, mx = 0.
__device__ __forceinline__ void u8_minmax_acc(uint32_t x, uint32_t& mn_e, uint32_t& mn_o, uint32_t& mx_e, uint32_t& mx_o) {
    const uint32_t e = __byte_perm(x, 0u, 0x4240u), o = __byte_perm(x, 0u, 0x4341u);  // (b0, 0, b2, 0), (b1, 0, b3, 0)
    mn_e = __vminu2(mn_e, e); mn_o = __vminu2(mn_o, o);
    mx_e = __vmaxu2(mx_e, e); mx_o = __vmaxu2(mx_o, o);
}

// reduce the SWAR lanes of one register to a single T
template <class T>
__device__ __forceinline__ T swar_reduce_min(typename Lay<T>::R l) {
    using M = MinMax<T>;
    if constexpr (sizeof(T) == 1) { uint32_t a = M::mn(l, l >> 16); a = M::mn(a, a >> 8); return T(a & 0xFF); }
    else if constexpr (sizeof(T) == 2) return T(M::mn(l, l >> 16) & 0xFFFF);
    else return T(l);
}
template <class T>
__device__ __forceinline__ T swar_reduce_max(typename Lay<T>::R h) {
    using M = MinMax<T>;
    if constexpr (sizeof(T) == 1) { uint32_t b = M::mx(h, h >> 16); b = M::mx(b, b >> 8); return T(b & 0xFF); }
    else if constexpr (sizeof(T) == 2) return T(M::mx(h, h >> 16) & 0xFFFF);
    else return T(h);
}

// ---- unpack: bits [ROW*W, ROW*W+W) of every lane's stream  (src/macros.rs:139-170) -------------------
// `cur` is word-row (ROW*W)/T, `nxt` word-row +1 (only read when the field straddles a word boundary).
template <class T, int W, int ROW>
__device__ __forceinline__ typename Lay<T>::R extract_field(typename Lay<T>::R cur, typename Lay<T>::R nxt) {
    using R = typename Lay<T>::R;
    constexpr int TB = Lay<T>::TB;
    constexpr int shift = (ROW * W) % TB;  // macros.rs:147
    constexpr R MW = rep_mask<T>(W);
    if constexpr (shift + W <= TB) {
        // macros.rs:164 (and :154 with remaining_bits == 0)
        if constexpr (Lay<T>::LPR == 1 && shift + W == TB) return cur >> shift;
        else if constexpr (shift == 0) return cur & MW;
        else if constexpr (sizeof(T) == 4 && W == 8) return __byte_perm(cur, 0u, 0x4440u + shift / 8);  // one PRMT, not SHF + LOP3
        else return (cur >> shift) & MW;
    } else {
        // macros.rs:149-161: low current_bits from cur, the remaining bits from the next word-row
        constexpr int cb = TB - shift;  // current_bits
        if constexpr (sizeof(T) == 4) {
            return __funnelshift_r(cur, nxt, shift) & MW;
        } else if constexpr (sizeof(T) == 8) {
            return ((cur >> shift) | (nxt << cb)) & MW;
        } else {
            constexpr R MC = rep_mask<T>(cb);
            return ((cur >> shift) & MC) | ((nxt << cb) & (MW ^ MC));
        }
    }
}

template <class T, int W, int ROW>
__device__ __forceinline__ Slice<T> extract_row(const Slice<T>& cur, const Slice<T>& nxt) {
    Slice<T> o;
#pragma unroll
    for (int i = 0; i < Lay<T>::NR; ++i) o.r[i] = extract_field<T, W, ROW>(cur.r[i], nxt.r[i]);
    return o;
}

// ---- pack: lane-wise shifts that must not leak into the neighbouring SWAR lane ----------------------
// (src << shift) within each T-bit lane  (macros.rs:79)
template <class T, int SH>
__device__ __forceinline__ typename Lay<T>::R lane_shl(typename Lay<T>::R x) {
    using R = typename Lay<T>::R;
    if constexpr (SH == 0) return x;
    else if constexpr (Lay<T>::LPR == 1) return x << SH;
    else return (x << SH) & R(~rep_mask<T>(SH));
}
// (src >> SH) within each lane, where src has at most W significant bits  (macros.rs:92)
template <class T, int SH, int KEEP>
__device__ __forceinline__ typename Lay<T>::R lane_shr_keep(typename Lay<T>::R x) {
    using R = typename Lay<T>::R;
    if constexpr (KEEP == 0) return R(0);
    else if constexpr (Lay<T>::LPR == 1) return x >> SH;
    else return (x >> SH) & rep_mask<T>(KEEP);
}

}  // namespace flb
