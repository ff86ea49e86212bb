/* fastlanes_b200.h — C ABI of the B200-native FastLanes codec (libfastlanes_b200.so).
 *
 * Drop-in boundary for the hot path of spiraldb/fastlanes v0.1.8: the BitPacking / FoR / Delta /
 * Transpose trait surface over 1024-element blocks of u8/u16/u32/u64 at every bit width.  The
 * reference has no FFI of its own; each entry point below replaces one trait method
 * and cites it (paths relative to the reference repo).  INTEGRATION.md shows the Rust `extern "C"`
 * binding + trait impls a maintainer would add; bindings/rust/ ships them as source.
 *
 * Conventions
 *   - Blocks are contiguous:  unpacked block b = 1024 elements at p + b*1024;
 *     packed block b = 1024*width/T elements (exactly 128*width bytes, no padding — wire format)
 *     at p + b*(1024*width/T);  base block b = LANES = 1024/T elements (128 bytes) at base + b*LANES.
 *   - `width` is the runtime bit width of the `unchecked_*` family (src/bitpacking.rs:30,44,58).
 *     width > T returns FL_ERR_WIDTH (the reference's unreachable!(), src/bitpacking.rs:93,126,197).
 *   - Two families per operation:
 *       fl_<op>_<T>        DEVICE pointers, stream-ordered, asynchronous (`stream` is a cudaStream_t
 *                          passed as void*; NULL = the legacy default stream).  Pointers must be
 *                          16-byte aligned (cudaMalloc gives 256) else FL_ERR_ALIGN.  Nothing is
 *                          allocated, nothing is retained after the stream reaches the call.
 *       fl_host_<op>_<T>   HOST pointers, synchronous, on the current CUDA device.  With n_blocks = 1
 *                          these are the reference's single-block trait calls.  Three paths, same
 *                          bytes: calls up to 256 KiB go through a page-locked bounce buffer the
 *                          kernel addresses directly (one launch); calls whose buffers are ALL
 *                          page-locked (fl_host_alloc / fl_host_register / cudaHostAlloc) and that
 *                          move at most FLB_DIRECT_MAX bytes (default 512 MiB) run as one launch on
 *                          the caller's memory; everything else is chunked H2D copy / kernel / D2H
 *                          copy pipelined over internal streams.
 *   - Input and output must not overlap (Rust's & / &mut guarantee in the reference).
 *   - No entry point unwinds, aborts or falls back to the CPU: without a usable CUDA device every
 *     call returns FL_ERR_CUDA.  Thread-safe; fl_last_error_string() is thread-local.
 */
#ifndef FASTLANES_B200_H
#define FASTLANES_B200_H
#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define FL_API __attribute__((visibility("default")))
#else
#define FL_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef int fl_status;
enum {
    FL_OK = 0,
    FL_ERR_WIDTH = 1, /* width > T            (src/bitpacking.rs:93,126,197 unreachable!) */
    FL_ERR_LEN = 2,   /* more than 2^31 blocks in one call; the slice-length debug_asserts (src/bitpacking.rs:78-80,111-113) are the host mirrors' job */
    FL_ERR_INDEX = 3, /* index out of range   (src/bitpacking.rs:152 assert!) */
    FL_ERR_ALIGN = 4, /* device pointer not 16-byte aligned */
    FL_ERR_CUDA = 5,  /* CUDA runtime error or no device; see fl_last_error_string() */
    FL_ERR_NULL = 6,  /* required pointer is NULL */
    FL_ERR_UNSUPPORTED = 7 /* operation not implemented for this element type (reserved; unused in this release) */
};

/* Library / context -------------------------------------------------------------------------- */
FL_API const char* fl_version(void);
FL_API const char* fl_last_error_string(void);       /* thread-local, never NULL */
FL_API const char* fl_status_string(fl_status s);
FL_API int fl_device_count(void);                    /* 0 when no usable CUDA device */
/* Optional: make `device` current for the calling thread and create its host-path context (streams) up front so the
 * first fl_host_* call does not pay for it.  Every entry point also works without it (contexts are created lazily). */
FL_API fl_status fl_init(int device);
/* Host-path tuning, PROCESS-WIDE (every device, every thread): blocks per pipelined chunk (0 = default 16384, clamped to
 * the 2^31-block launch limit) and number of internal streams per pipeline (0 = default 3, at most 16).  Takes effect
 * on the next fl_host_* call. */
FL_API fl_status fl_host_configure(size_t chunk_blocks, int n_streams);
/* Page-locked host memory.  The host family is PCIe-bound and a GPU's DMA runs at full speed only against the memory
 * of its own socket, so fl_host_alloc places the pages on the NUMA node of the CURRENT device: node lookup =
 * FLB_NUMA_MAP="n0,n1,.." (per device ordinal) > cudaDevAttrHostNumaId > /sys/bus/pci/devices/<id>/numa_node; placement =
 * mmap + mbind(MPOL_PREFERRED) + cudaHostRegister (independent of the CPUs the caller is allowed to run on).  When the
 * node is unknown or the kernel refuses, it is a plain cudaHostAlloc.  FLB_NUMA=0 disables.  fl_host_free releases
 * either kind.  fl_device_numa_node: the node found for `device`, or -1.  fl_host_buffer_node: the node that actually
 * holds the page at p (move_pages query), or -1 — to verify a placement. */
FL_API int fl_device_numa_node(int device);
FL_API fl_status fl_host_alloc(void** p, size_t bytes);
FL_API fl_status fl_host_free(void* p);
FL_API int fl_host_buffer_node(const void* p);
FL_API fl_status fl_host_register(void* p, size_t bytes);
FL_API fl_status fl_host_unregister(void* p);
/* The copies of a host call WITHOUT its kernel, through the same chunked multi-stream pipeline: in_bytes_per_block go
 * host->device, out_bytes_per_block come back (contents unspecified).  The time of this call is the PCIe ceiling of the
 * fl_host_* call with the same byte counts on this box (bench.py: e2e.link_ceiling). */
FL_API fl_status fl_host_copy_probe(size_t in_bytes_per_block, size_t out_bytes_per_block, size_t n_blocks, const void* in,
                                    void* out);
/* Release the internal streams and staging buffers of every device used by fl_host_* calls. */
FL_API fl_status fl_shutdown(void);

/* Multi-device context ------------------------------------------------------------------------------
 * every op reads and writes exactly one 1024-value block, so a batch shards by contiguous block ranges
 * with no exchange between devices.  A context owns one worker thread per listed device; fl_ctx_host_<op>_<T> has the
 * signature of fl_host_<op>_<T> plus the context and runs blocks [n*i/G, n*(i+1)/G) of EVERY array of the call (packed,
 * unpacked, base, bitmap, counts shard on the same block index) on device i through that device's own host pipeline:
 * G PCIe links and copy pipelines in one call from one host thread.  The seam it extends is the batched runtime-width
 * family, src/bitpacking.rs:109-129.  devices == NULL or n_devices <= 0: all visible devices.  A device may be listed
 * more than once (each entry is an independent shard worker).  Calls on one context are serialised; use one context
 * per caller thread for concurrency.  No collective library is involved (FL_ERR_NCCL of SURVEY.md §8b does not exist):
 * the only inter-device traffic is the optional scatter/gather below, plain peer copies over NVLink. */
typedef struct fl_ctx fl_ctx;
FL_API fl_status fl_ctx_create(const int* devices, int n_devices, fl_ctx** ctx);
FL_API fl_status fl_ctx_destroy(fl_ctx* ctx);
FL_API int fl_ctx_device_count(const fl_ctx* ctx);
FL_API int fl_ctx_device(const fl_ctx* ctx, int i);         /* CUDA ordinal of shard i, -1 if out of range */
/* shard i owns blocks [*first, *end) of an n_blocks batch */
FL_API fl_status fl_ctx_block_range(const fl_ctx* ctx, size_t n_blocks, int i, size_t* first, size_t* end);
/* Page-locked buffer of n_blocks * bytes_per_block bytes whose pages follow the shards: the byte range of shard i is
 * placed on the NUMA node of device i (see fl_host_alloc).  Free with fl_host_free. */
FL_API fl_status fl_ctx_host_alloc(fl_ctx* ctx, size_t n_blocks, size_t bytes_per_block, void** p);
FL_API fl_status fl_ctx_host_copy_probe(fl_ctx* ctx, size_t in_bytes_per_block, size_t out_bytes_per_block, size_t n_blocks,
                                        const void* in, void* out);
/* The "trivial block shard / gather" between DEVICE buffers: `src` / `dst` holds all n_blocks on context device `root`,
 * shards[i] (device memory of context device i) holds that shard's blocks.  Synchronous peer copies, outside any decode. */
FL_API fl_status fl_ctx_scatter_blocks(fl_ctx* ctx, size_t bytes_per_block, size_t n_blocks, const void* src, int root,
                                       void* const* shards);
FL_API fl_status fl_ctx_gather_blocks(fl_ctx* ctx, size_t bytes_per_block, size_t n_blocks, const void* const* shards, int root,
                                      void* dst);

/* Per-type entry points ------------------------------------------------------------------------ */
#define FL_DECLARE_TYPE(SFX, T)                                                                                     \
    /* BitPacking::pack / unchecked_pack — src/bitpacking.rs:19,30 (impl :65-96) */                                 \
    FL_API fl_status fl_pack_##SFX(unsigned width, size_t n_blocks, const T* in, T* packed, void* stream);                 \
    FL_API fl_status fl_host_pack_##SFX(unsigned width, size_t n_blocks, const T* in, T* packed);                          \
    /* BitPacking::unpack / unchecked_unpack — src/bitpacking.rs:33,44 (impl :98-129) */                            \
    FL_API fl_status fl_unpack_##SFX(unsigned width, size_t n_blocks, const T* packed, T* out, void* stream);              \
    FL_API fl_status fl_host_unpack_##SFX(unsigned width, size_t n_blocks, const T* packed, T* out);                       \
    /* BitPacking::unpack_single / unchecked_unpack_single — src/bitpacking.rs:47,58 (impl :132-200),               \
     * batched: global_index[i] = block*1024 + index_in_block; any index >= n_blocks*1024 -> FL_ERR_INDEX           \
     * is reported by the host variant; the device variant writes 0 for it and sets *oob_flag (may be NULL). */      \
    FL_API fl_status fl_unpack_gather_##SFX(unsigned width, size_t n_blocks, const T* packed,                              \
                                     const uint64_t* global_index, size_t n, T* out, int* oob_flag, void* stream);  \
    FL_API fl_status fl_host_unpack_gather_##SFX(unsigned width, size_t n_blocks, const T* packed,                         \
                                          const uint64_t* global_index, size_t n, T* out);                          \
    FL_API fl_status fl_host_unpack_single_##SFX(unsigned width, const T* packed, size_t index, T* value);                 \
    /* FoR::for_pack — src/ffor.rs:5-10 (impl :24-36).  `_refs`: one reference per block. */                        \
    FL_API fl_status fl_for_pack_##SFX(unsigned width, size_t n_blocks, const T* in, T reference, T* packed, void* stream); \
    FL_API fl_status fl_for_pack_refs_##SFX(unsigned width, size_t n_blocks, const T* in, const T* refs, T* packed,        \
                                     void* stream);                                                                 \
    FL_API fl_status fl_host_for_pack_##SFX(unsigned width, size_t n_blocks, const T* in, T reference, T* packed);         \
    /* FoR::unfor_pack — src/ffor.rs:12-17 (impl :38-50) */                                                         \
    FL_API fl_status fl_unfor_pack_##SFX(unsigned width, size_t n_blocks, const T* packed, T reference, T* out,            \
                                  void* stream);                                                                    \
    FL_API fl_status fl_unfor_pack_refs_##SFX(unsigned width, size_t n_blocks, const T* packed, const T* refs, T* out,     \
                                       void* stream);                                                               \
    FL_API fl_status fl_host_unfor_pack_##SFX(unsigned width, size_t n_blocks, const T* packed, T reference, T* out);      \
    /* Delta::delta — src/delta.rs:7 (impl :24-33); base: n_blocks x LANES */                                       \
    FL_API fl_status fl_delta_##SFX(size_t n_blocks, const T* in, const T* base, T* out, void* stream);                    \
    FL_API fl_status fl_host_delta_##SFX(size_t n_blocks, const T* in, const T* base, T* out);                             \
    /* Delta::undelta — src/delta.rs:9 (impl :36-45) */                                                             \
    FL_API fl_status fl_undelta_##SFX(size_t n_blocks, const T* in, const T* base, T* out, void* stream);                  \
    FL_API fl_status fl_host_undelta_##SFX(size_t n_blocks, const T* in, const T* base, T* out);                           \
    /* Delta::undelta_pack — src/delta.rs:11-17 (impl :48-63): fused unpack + prefix-add */                         \
    FL_API fl_status fl_undelta_pack_##SFX(unsigned width, size_t n_blocks, const T* packed, const T* base, T* out,        \
                                    void* stream);                                                                  \
    FL_API fl_status fl_host_undelta_pack_##SFX(unsigned width, size_t n_blocks, const T* packed, const T* base, T* out);  \
    /* FUSED decode to ORIGINAL order: untranspose(undelta_pack(packed, base)) in one pass — the decode chain of   \
     * src/delta.rs:99 + src/transpose.rs:18-22.                                          */ \
    FL_API fl_status fl_undelta_pack_untranspose_##SFX(unsigned width, size_t n_blocks, const T* packed, const T* base,    \
                                                T* out, void* stream);                                              \
    FL_API fl_status fl_host_undelta_pack_untranspose_##SFX(unsigned width, size_t n_blocks, const T* packed,              \
                                                     const T* base, T* out);                                        \
    /* FUSED encode from ORIGINAL order: pack(delta(transpose(in), base)) in one pass — the encode chain of        \
     * src/delta.rs:88-95.                                                                                      */ \
    FL_API fl_status fl_transpose_delta_pack_##SFX(unsigned width, size_t n_blocks, const T* in, const T* base, T* packed, \
                                            void* stream);                                                          \
    FL_API fl_status fl_host_transpose_delta_pack_##SFX(unsigned width, size_t n_blocks, const T* in, const T* base,       \
                                                 T* packed);                                                        \
    /* Per-block min and max: what a caller needs to pick FoR's `reference` (= min) and W (= bits(max - min))      \
     * before FoR::for_pack, which takes both as givens (src/ffor.rs:5-10).  SURVEY.md §8f rank 3.  mins/maxs:      \
     * n_blocks elements each. */                                                                                   \
    FL_API fl_status fl_block_minmax_##SFX(size_t n_blocks, const T* in, T* mins, T* maxs, void* stream);                  \
    FL_API fl_status fl_host_block_minmax_##SFX(size_t n_blocks, const T* in, T* mins, T* maxs);                           \
    /* cwida/FastLanes row order.  The reference reorders the rows of the bit-packed layout     \
     * (FL_ORDER, src/macros.rs:20-24) and says so: "not binary compatible with original FastLanes" (README.md:49-56,    \
     * src/macros.rs:1-9: "it iterates over the elements respecting the transposed ordering").  These four entry points   \
     * use the ORIGINAL order instead — row r of a block = values r*LANES .. r*LANES+LANES-1 — for the linear           \
     * encodings (bit-packing, FoR), same lane bit-streams, same 128*width bytes per block.  PARITY UNPINNED: the        \
     * reference contains no code, test or vector for that layout; the check is a closed-form oracle written from the  \
     * description above (oracle/cwida.py).  Device pointers only. */                                                   \
    FL_API fl_status fl_pack_cwida_##SFX(unsigned width, size_t n_blocks, const T* in, T* packed, void* stream);           \
    FL_API fl_status fl_unpack_cwida_##SFX(unsigned width, size_t n_blocks, const T* packed, T* out, void* stream);        \
    FL_API fl_status fl_for_pack_cwida_##SFX(unsigned width, size_t n_blocks, const T* in, T reference, T* packed,         \
                                      void* stream);                                                                \
    FL_API fl_status fl_unfor_pack_cwida_##SFX(unsigned width, size_t n_blocks, const T* packed, T reference, T* out,      \
                                        void* stream);                                                              \
    /* FUSED statistics + FoR::for_pack (src/ffor.rs:24-36): reference = the block's own minimum, found in the same   \
     * pass that packs (the warp holds the whole block).  refs_out: n_blocks references (feed them to                 \
     * fl_unfor_pack_refs); spans_out (nullable): n_blocks values max - min — the block is lossless iff                \
     * spans_out[b] < 2^width (for_pack truncates to `width` bits like the reference, src/macros.rs:73).               \
     * Device pointers only. */                                                                                        \
    FL_API fl_status fl_for_pack_auto_##SFX(unsigned width, size_t n_blocks, const T* in, T* refs_out, T* spans_out,       \
                                     T* packed, void* stream);                                                      \
    /* FUSED scan (not a trait method of the reference — README.md:40-41 tells callers to     \
     * unpack the whole block and loop over it): decode in registers, apply a range predicate, never materialise the \
     * block.  value[i] = unfor_pack(packed, reference)[i] (src/ffor.rs:38-50; reference 0 = plain unpack,           \
     * src/bitpacking.rs:98-107); `refs` (nullable) = one reference per block, else the scalar `reference`.          \
     * filter: bit i of block b (bitmap byte b*128 + i/8, bit i%8; i = index in the unpacked block) =                \
     *         lo <= value[i] <= hi (unsigned, inclusive; hi < lo selects nothing).  bitmap: n_blocks*128 bytes;     \
     *         counts (nullable): n_blocks selected-value counts.                                                    \
     * select: out[offsets[b] + k] = the k-th selected value of block b in index order (offsets = exclusive prefix   \
     *         sum of the counts: dense stream compaction).  Device pointers only. */                                \
    FL_API fl_status fl_unpack_filter_##SFX(unsigned width, size_t n_blocks, const T* packed, const T* refs, T reference,  \
                                     T lo, T hi, uint8_t* bitmap, uint32_t* counts, void* stream);                  \
    FL_API fl_status fl_host_unpack_filter_##SFX(unsigned width, size_t n_blocks, const T* packed, T reference, T lo,      \
                                          T hi, uint8_t* bitmap, uint32_t* counts);                                 \
    /* Delta scan: bit i = lo <= untranspose(undelta_pack(packed, base))[i] <= hi — the range scan over a delta-encoded   \
     * column (src/delta.rs:48-63, src/transpose.rs:18-22), answered in ORIGINAL value order, block never materialised. \
     * base: n_blocks * LANES elements as for fl_undelta_pack. */                                                       \
    FL_API fl_status fl_undelta_pack_filter_##SFX(unsigned width, size_t n_blocks, const T* packed, const T* base, T lo,   \
                                           T hi, uint8_t* bitmap, uint32_t* counts, void* stream);                  \
    FL_API fl_status fl_host_undelta_pack_filter_##SFX(unsigned width, size_t n_blocks, const T* packed, const T* base,    \
                                                T lo, T hi, uint8_t* bitmap, uint32_t* counts);                     \
    FL_API fl_status fl_unpack_select_##SFX(unsigned width, size_t n_blocks, const T* packed, const T* refs, T reference,  \
                                     const uint8_t* bitmap, const uint64_t* offsets, T* out, void* stream);         \
    /* Transpose::transpose / untranspose — src/transpose.rs:5-6 (impl :11-22) */                                   \
    FL_API fl_status fl_transpose_##SFX(size_t n_blocks, const T* in, T* out, void* stream);                               \
    FL_API fl_status fl_untranspose_##SFX(size_t n_blocks, const T* in, T* out, void* stream);                             \
    FL_API fl_status fl_host_transpose_##SFX(size_t n_blocks, const T* in, T* out);                                        \
    FL_API fl_status fl_host_untranspose_##SFX(size_t n_blocks, const T* in, T* out);                                      \
    /* Context family: fl_host_<op>_<T> block-sharded over the devices of a context (see fl_ctx above). */          \
    FL_API fl_status fl_ctx_host_pack_##SFX(fl_ctx* ctx, unsigned width, size_t n_blocks, const T* in, T* packed);         \
    FL_API fl_status fl_ctx_host_unpack_##SFX(fl_ctx* ctx, unsigned width, size_t n_blocks, const T* packed, T* out);      \
    FL_API fl_status fl_ctx_host_for_pack_##SFX(fl_ctx* ctx, unsigned width, size_t n_blocks, const T* in, T reference,    \
                                                T* packed);                                                         \
    FL_API fl_status fl_ctx_host_unfor_pack_##SFX(fl_ctx* ctx, unsigned width, size_t n_blocks, const T* packed,           \
                                                  T reference, T* out);                                             \
    FL_API fl_status fl_ctx_host_delta_##SFX(fl_ctx* ctx, size_t n_blocks, const T* in, const T* base, T* out);            \
    FL_API fl_status fl_ctx_host_undelta_##SFX(fl_ctx* ctx, size_t n_blocks, const T* in, const T* base, T* out);          \
    FL_API fl_status fl_ctx_host_undelta_pack_##SFX(fl_ctx* ctx, unsigned width, size_t n_blocks, const T* packed,         \
                                                    const T* base, T* out);                                         \
    FL_API fl_status fl_ctx_host_undelta_pack_untranspose_##SFX(fl_ctx* ctx, unsigned width, size_t n_blocks,              \
                                                                const T* packed, const T* base, T* out);            \
    FL_API fl_status fl_ctx_host_transpose_delta_pack_##SFX(fl_ctx* ctx, unsigned width, size_t n_blocks, const T* in,     \
                                                            const T* base, T* packed);                              \
    FL_API fl_status fl_ctx_host_transpose_##SFX(fl_ctx* ctx, size_t n_blocks, const T* in, T* out);                       \
    FL_API fl_status fl_ctx_host_untranspose_##SFX(fl_ctx* ctx, size_t n_blocks, const T* in, T* out);                     \
    FL_API fl_status fl_ctx_host_block_minmax_##SFX(fl_ctx* ctx, size_t n_blocks, const T* in, T* mins, T* maxs);          \
    FL_API fl_status fl_ctx_host_unpack_filter_##SFX(fl_ctx* ctx, unsigned width, size_t n_blocks, const T* packed,        \
                                                     T reference, T lo, T hi, uint8_t* bitmap, uint32_t* counts);   \
    FL_API fl_status fl_ctx_host_undelta_pack_filter_##SFX(fl_ctx* ctx, unsigned width, size_t n_blocks, const T* packed,  \
                                                           const T* base, T lo, T hi, uint8_t* bitmap,              \
                                                           uint32_t* counts);

FL_DECLARE_TYPE(u8, uint8_t)
FL_DECLARE_TYPE(u16, uint16_t)
FL_DECLARE_TYPE(u32, uint32_t)
FL_DECLARE_TYPE(u64, uint64_t)

#ifdef __cplusplus
}
#endif
#endif /* FASTLANES_B200_H */
