// fl_scan_bits.h — pure bit arithmetic of the fused scan kernels (fl_scan.cuh), kept free of CUDA intrinsics so
// that tests/cpp/test_scan_bits.cpp can run the exact same functions on the CPU against a brute-force bitmap.
//
// Setting.  In the warp-block layout (fl_kernels.cuh) thread (q, j) of a warp holds, for each of its RPG = T/4
// rows i (global row r = q*RPG + i), the 16-byte slice j of the row = BPT = 128/T consecutive lanes.  A predicate
// over those values gives the thread exactly RPG*BPT = 32 bits X, row i at bits [i*BPT, (i+1)*BPT).  The block's
// selection bitmap is indexed by the ORIGINAL value index (bit index(r, lane) = FL_ORDER[r/8]*16 + (r%8)*128 + lane,
// src/macros.rs:20-24), so row r owns the aligned run of L = 1024/T bits starting at row_bit_offset(r), and thread j
// owns BPT of them.  For BPT >= 8 (u8, u16) a thread's bits are whole bytes; for BPT = 4 / 2 (u32, u64) two / four
// neighbouring threads exchange X through warp shuffles and each assembles 4 whole bytes (4 rows x 8 lanes).
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define FLB_HD __host__ __device__ __forceinline__
#else
#define FLB_HD inline
#endif

namespace flb {

// 4 nibbles (16 bits) -> the low nibbles of 4 bytes
FLB_HD uint32_t spread4(uint32_t a) {
    a = (a | (a << 8)) & 0x00FF00FFu;
    a = (a | (a << 4)) & 0x0F0F0F0Fu;
    return a;
}
// 8 two-bit fields (16 bits) -> the low two bits of 8 nibbles
FLB_HD uint32_t spread2(uint32_t a) {
    a = spread4(a);
    a = (a | (a << 2)) & 0x33333333u;
    return a;
}

// FL_ORDER[i] (src/lib.rs:22) without a table: nibble i of 0x73516240
FLB_HD int scan_fl_order(int i) { return int((0x73516240u >> (4 * i)) & 7u); }

// byte offset inside the 128-byte block bitmap of the L-bit run owned by row r: index(r, 0) / 8
FLB_HD int row_bitmap_byte(int r) { return scan_fl_order(r >> 3) * 2 + (r & 7) * 16; }

// u32 (BPT = 4, RPG = 8): `mine` = this thread's X, `other` = X of thread j^1.  Returns 4 bytes: byte ii = lanes
// 8*(j>>1) .. +7 of local row 4*(j&1) + ii.
FLB_HD uint32_t merge_pair_bpt4(uint32_t mine, uint32_t other, int j) {
    const int h = j & 1;
    const uint32_t lo = h ? other : mine, hi = h ? mine : other;
    return spread4((lo >> (16 * h)) & 0xFFFFu) | (spread4((hi >> (16 * h)) & 0xFFFFu) << 4);
}
// u64 (BPT = 2, RPG = 16), step 1: returns 8 nibbles: nibble ii = lanes 4*(j>>1) .. +3 of local row 8*(j&1) + ii.
FLB_HD uint32_t merge_pair_bpt2(uint32_t mine, uint32_t other, int j) {
    const int h = j & 1;
    const uint32_t lo = h ? other : mine, hi = h ? mine : other;
    return spread2((lo >> (16 * h)) & 0xFFFFu) | (spread2((hi >> (16 * h)) & 0xFFFFu) << 2);
}
// u64 step 2: `mine` / `other` = step-1 results of threads j and j^2.  Returns 4 bytes: byte ii = lanes
// 8*(j>>2) .. +7 of local row 8*(j&1) + 4*((j>>1)&1) + ii.
FLB_HD uint32_t merge_quad_bpt2(uint32_t mine, uint32_t other, int j) {
    const int h = (j >> 1) & 1;
    const uint32_t lo = h ? other : mine, hi = h ? mine : other;
    return spread4((lo >> (16 * h)) & 0xFFFFu) | (spread4((hi >> (16 * h)) & 0xFFFFu) << 4);
}

// Lane-wise unsigned "x <= y" on the SWAR lanes of a 32-bit register (TBITS = 8: 4 lanes, 16: 2 lanes); the
// result is the TOP bit of every lane.  yH = y | H (H = top bit of every lane).  d = (y|H) - (x&~H) never borrows
// across lanes and its top bit says low(y) >= low(x); the top bits of x and y decide otherwise (one LOP3 on the GPU).
template <int TBITS>
FLB_HD uint32_t swar_leu_top(uint32_t x, uint32_t y, uint32_t yH) {
    const uint32_t H = (TBITS == 8) ? 0x80808080u : 0x80008000u;
    const uint32_t d = yH - (x & ~H);
    return ((~x & y) | (~(x ^ y) & d)) & H;
}
// top bits of the lanes -> packed predicate bits (bit k = lane k)
FLB_HD uint32_t top_bits_u8(uint32_t le) { return ((((le >> 7) & 0x01010101u) * 0x00204081u) >> 21) & 15u; }
FLB_HD uint32_t top_bits_u16(uint32_t le) {
    const uint32_t m = le >> 15;  // bit 0, bit 16
    return (m | (m >> 15)) & 3u;
}

// Stores a thread's assembled word into the warp's 128-byte bitmap tile.  `z` is, per element size:
//   u8  : X itself                (2 rows x 16 lanes: two 16-bit stores)
//   u16 : X itself                (4 rows x  8 lanes: four byte stores)
//   u32 : merge_pair_bpt4 result  (4 rows x  8 lanes of the thread PAIR)
//   u64 : merge_quad_bpt2 result  (4 rows x  8 lanes of the thread QUAD)
// q = rank of the thread's group in row order (rows q*RPG .. q*RPG + RPG-1), j = slice index 0..7.
template <int TBITS>
FLB_HD void scan_store(unsigned char* tile, int q, int j, uint32_t z) {
    // The 4 (u8: 2) rows a thread stores never cross a multiple of 8, so FL_ORDER[r/8] is common to them and
    // consecutive rows are 16 bytes apart (row_bitmap_byte): one base address, immediate offsets.
    if (TBITS == 8) {        // rows 2q + i
        unsigned char* p = tile + row_bitmap_byte(2 * q) + 2 * j;
        for (int i = 0; i < 2; ++i) {
            p[16 * i] = (unsigned char)(z >> (16 * i));
            p[16 * i + 1] = (unsigned char)(z >> (16 * i + 8));
        }
    } else if (TBITS == 16) {  // rows 4q + i
        unsigned char* p = tile + row_bitmap_byte(4 * q) + j;
        for (int i = 0; i < 4; ++i) p[16 * i] = (unsigned char)(z >> (8 * i));
    } else if (TBITS == 32) {  // rows 8q + 4(j&1) + ii
        unsigned char* p = tile + row_bitmap_byte(8 * q + 4 * (j & 1)) + (j >> 1);
        for (int ii = 0; ii < 4; ++ii) p[16 * ii] = (unsigned char)(z >> (8 * ii));
    } else {                   // rows 16q + 8(j&1) + 4((j>>1)&1) + ii
        unsigned char* p = tile + row_bitmap_byte(16 * q + 8 * (j & 1) + 4 * ((j >> 1) & 1)) + (j >> 2);
        for (int ii = 0; ii < 4; ++ii) p[16 * ii] = (unsigned char)(z >> (8 * ii));
    }
}

// ---- delta scan: bitmap in ORIGINAL value order --------------------------------------------------------------------
// After undelta_pack the register tile is in transposed order: lane l walks, row by row, the T consecutive originals
// start(l) .. start(l)+T-1, start(l) = 64*(l%16) + 8*FL_ORDER[l/16] (src/transpose.rs:29-36 composed with
// src/macros.rs:20-24).  A thread's RPG rows of one lane are therefore RPG CONSECUTIVE bits of the original-order
// bitmap, so its 32 predicate bits are kept LANE-major, bit k*RPG + i (k = lane inside the 16-byte slice, i = local
// row): u32 -> 4 whole bytes, u64 -> 2 halfwords; u16 (4 bits per lane) / u8 (2 bits per lane) are completed to whole
// bytes with the threads holding the same lanes in the neighbouring row groups (q^1, q^2) through the same
// merge_pair / merge_quad helpers as above, with the group rank q in the role of j.

// lane-major placement of the SWAR compare result of ONE register (top bit of each lane set = pass):
// u8: lanes 4r..4r+3 -> bits 0,2,4,6 (stride RPG = 2);  u16: lanes 2r, 2r+1 -> bits 0, 4 (stride RPG = 4)
FLB_HD uint32_t top_bits_lane_major_u8(uint32_t le) { return ((((le >> 7) & 0x01010101u) * 0x00041041u) >> 18) & 0x55u; }
FLB_HD uint32_t top_bits_lane_major_u16(uint32_t le) {
    const uint32_t z = (le >> 15) & 0x00010001u;  // bit 0, bit 16
    return (z | (z >> 12)) & 0x11u;
}

// ---- select (fl_scan.cuh, select_warp_kernel): where a thread's slice of a row sits in the block bitmap --------------------
// Thread (q, j) holds, for local row i of its run (global row r = q*RPG + i), the BPT = 128/T values of lanes j*BPT .. of
// that row = BPT CONSECUTIVE original indices starting at index(r, j*BPT).  Rows of one 8-row band share FL_ORDER[r/8], so
// inside band `band` (rows q*RPG + band*8 + 0..7; RPG < 8: the whole run is one band) the first index is
// c0 + (i%8)*128: the bitmap word advances by 4 per row and the bit position inside the word is the same for every row.
template <int TBITS>
FLB_HD int select_band_origin(int q, int band, int j) {
    const int RPG = TBITS / 4, BPT = 128 / TBITS;
    const int r0 = q * RPG + band * 8;  // first global row of the band
    return scan_fl_order(r0 >> 3) * 16 + (r0 & 7) * 128 + j * BPT;
}
// The compaction tests bit k of the slice at register position k + 1 (one R2P then moves bits 1.. into predicates; bit 0
// would cost two extra instructions): the bitmap word is rotated right by s0 - 1, s0 = position of the slice in the word.
FLB_HD uint32_t select_rotation(uint32_t s0) { return (s0 + 31u) & 31u; }
FLB_HD uint32_t select_rotate(uint32_t word, uint32_t rot) { return rot ? ((word >> rot) | (word << (32u - rot))) : word; }
// mask of the bits of the word that precede the slice: their popcount is the slice's rank offset inside the word
FLB_HD uint32_t select_low_mask(uint32_t s0) { return (1u << s0) - 1u; }

// first original index (block-local) of lane l's run
FLB_HD int lane_run_start(int l) { return 64 * (l & 15) + 8 * scan_fl_order(l >> 4); }

// Stores a thread's assembled word of the ORIGINAL-order bitmap into the warp's 128-byte tile.  `z` is
//   u32 : X itself (byte k = lane 4j+k, rows 8q..8q+7)            u64 : X itself (halfword k = lane 2j+k, rows 16q..16q+15)
//   u16 : merge_pair_bpt4(X, X of rank q^1, q)  -> byte ii = lane 8j + 4(q&1) + ii, rows 8(q>>1) .. +7
//   u8  : merge_quad_bpt2(merge_pair_bpt2(X, X of q^1, q), same of q^2, q) -> byte ii = lane 16j + 8(q&1) + 4(q>>1) + ii, rows 0..7
template <int TBITS>
FLB_HD void scan_store_orig(unsigned char* tile, int q, int j, uint32_t z) {
    if (TBITS == 32) {
        for (int k = 0; k < 4; ++k) tile[lane_run_start(4 * j + k) / 8 + q] = (unsigned char)(z >> (8 * k));
    } else if (TBITS == 64) {
        for (int k = 0; k < 2; ++k) {
            unsigned char* p = tile + lane_run_start(2 * j + k) / 8 + 2 * q;
            p[0] = (unsigned char)(z >> (16 * k));
            p[1] = (unsigned char)(z >> (16 * k + 8));
        }
    } else if (TBITS == 16) {
        for (int ii = 0; ii < 4; ++ii) tile[lane_run_start(8 * j + 4 * (q & 1) + ii) / 8 + (q >> 1)] = (unsigned char)(z >> (8 * ii));
    } else {
        for (int ii = 0; ii < 4; ++ii) tile[lane_run_start(16 * j + 8 * (q & 1) + 4 * (q >> 1) + ii) / 8] = (unsigned char)(z >> (8 * ii));
    }
}

}  // namespace flb
