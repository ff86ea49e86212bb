/* fl_oracle.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * C ABI of the CPU oracle: a restatement of spiraldb/fastlanes v0.1.8 (src/macros.rs,
 * src/bitpacking.rs, src/transpose.rs, src/delta.rs, src/ffor.rs).  Used only by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, as the checker
 * and the timed CPU baseline.  The product library never links it.
 *
 * All buffers are HOST memory, contiguous arrays of blocks:
 *   unpacked block b : 1024 elements at  p + b*1024
 *   packed   block b : 1024*width/T elements (= 128*width bytes) at  p + b*(1024*width/T)
 *   base     block b : LANES = 1024/T elements (always 128 bytes) at  base + b*LANES
 * `tbits` is 8/16/32/64.  `n_threads` <= 1 runs on the calling thread; otherwise the block range
 * is split contiguously over std::threads.
 */
#ifndef FL_ORACLE_H
#define FL_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum {
    FLO_OK = 0,
    FLO_ERR_WIDTH = 1, /* width > T  — the reference's unreachable!() (bitpacking.rs:93,126,197) */
    FLO_ERR_TYPE = 2,
    FLO_ERR_INDEX = 3, /* index >= 1024 — the reference's assert! (bitpacking.rs:152) */
    FLO_ERR_NULL = 4
};

enum {
    FLO_OP_PACK = 0,         /* bitpacking.rs:65-96   in: unpacked, out: packed */
    FLO_OP_UNPACK = 1,       /* bitpacking.rs:98-129  in: packed,   out: unpacked */
    FLO_OP_FOR_PACK = 2,     /* ffor.rs:24-36 */
    FLO_OP_UNFOR_PACK = 3,   /* ffor.rs:38-50 */
    FLO_OP_DELTA = 4,        /* delta.rs:24-33  (width ignored) */
    FLO_OP_UNDELTA = 5,      /* delta.rs:36-45  (width ignored) */
    FLO_OP_UNDELTA_PACK = 6, /* delta.rs:48-63 */
    FLO_OP_TRANSPOSE = 7,    /* transpose.rs:11-15 (width ignored) */
    FLO_OP_UNTRANSPOSE = 8,  /* transpose.rs:18-22 (width ignored) */
    FLO_OP_UNFOR_FILTER = 9  /* NOT a reference function: ffor.rs:38-50 into a 1024-value scratch + the caller-side
                                range-predicate loop (README.md:40-41).  base = {lo, hi} (2 elements),
                                out = bitmap, 128 bytes per block, bit i = lo <= value[i] <= hi */
};

/* Generic batched entry point.  base: n_blocks*LANES (delta ops), refs: per-block references or
 * NULL (then ref_scalar is used) for the FoR ops. */
int flo_run(int tbits, int op, unsigned width, size_t n_blocks, const void* in, void* out,
            const void* base, const void* refs, uint64_t ref_scalar, int n_threads);

/* bitpacking.rs:132-200 — unpack_single / unchecked_unpack_single on ONE packed block. */
int flo_unpack_single(int tbits, unsigned width, const void* packed, size_t index, uint64_t* value);

/* Batched random access: global_index[i] = block*1024 + index_in_block; out has n elements of T. */
int flo_unpack_gather(int tbits, unsigned width, const void* packed, const uint64_t* global_index,
                      size_t n, void* out);

/* "x86-64-v2" | "x86-64-v3" | "x86-64-v4": the ISA level the dispatcher selected on this host. */
const char* flo_isa(void);
/* Force an ISA level (2/3/4) if the host supports it; returns the level now in use. */
int flo_set_isa_level(int level);
int flo_hardware_threads(void);

#ifdef __cplusplus
}
#endif
#endif
