//! Rust host side of the B200 FastLanes codec: the trait surface of `spiraldb/fastlanes` v0.1.8
//! (`BitPacking`, `FoR`, `Delta`, `Transpose` for `u8/u16/u32/u64`) implemented over the
//! `extern "C"` ABI of `libfastlanes_b200.so` (`include/fastlanes_b200.h`).
//!
//! NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Rust toolchain.  The parity of the
//! C ABI itself is tested from Python (`tests/`); this file is the binding a maintainer adds.
//!
//! Two levels:
//!  * the **trait impls** below are drop-in for the reference's single-block calls (host slices in,
//!    host slices out; one `fl_host_*` call = H2D copy + sm_100a kernel + D2H copy).  Where the
//!    reference panics (`unreachable!` on `width > T`, `assert!(index < 1024)`, `debug_assert` on
//!    slice lengths) these panic too, so behaviour is identical;
//!  * the **`device` module** exposes the batched, stream-ordered entry points on device pointers,
//!    which is where the throughput is (a whole column chunk per call, data resident in HBM).
//!
//! Unlike the reference, the const-generic `W` forms of the default traits need no `generic_const_exprs`:
//! the packed length is checked at run time against `1024 * W / T`.  Callers that want the reference's
//! signatures to the letter (`&mut [Self; 1024 * W / Self::T]`, `BitPackWidth<W>: SupportedBitPackWidth<Self>`)
//! enable the cargo feature `nightly-exact` and `use fastlanes_b200::exact::{BitPacking, Delta, FoR, Transpose}`.
#![allow(clippy::missing_safety_doc)]
// `exact` (feature "nightly-exact"): the reference's own signatures, array-typed packed operands included.  They need the
// same nightly feature the reference itself is built with (src/lib.rs:2 there, rust-toolchain.toml).
#![cfg_attr(feature = "nightly-exact", allow(incomplete_features))]
#![cfg_attr(feature = "nightly-exact", feature(generic_const_exprs))]

use core::ffi::c_void;

pub const FL_ORDER: [usize; 8] = [0, 4, 2, 6, 1, 5, 3, 7];

/// Status codes of `include/fastlanes_b200.h`.
pub mod status {
    pub const OK: i32 = 0;
    pub const ERR_WIDTH: i32 = 1;
    pub const ERR_LEN: i32 = 2;
    pub const ERR_INDEX: i32 = 3;
    pub const ERR_ALIGN: i32 = 4;
    pub const ERR_CUDA: i32 = 5;
    pub const ERR_NULL: i32 = 6;
}

extern "C" {
    fn fl_last_error_string() -> *const core::ffi::c_char;
}

fn check(st: i32, what: &str) {
    if st != status::OK {
        let msg = unsafe { std::ffi::CStr::from_ptr(fl_last_error_string()) }.to_string_lossy().into_owned();
        match st {
            status::ERR_WIDTH => unreachable!("Unsupported width ({what}): {msg}"),
            status::ERR_INDEX => panic!("Index must be less than 1024 ({what}): {msg}"),
            _ => panic!("fastlanes_b200 {what} failed with status {st}: {msg}"),
        }
    }
}

pub trait FastLanes: Sized + Copy {
    const T: usize = core::mem::size_of::<Self>() * 8;
    const LANES: usize = 1024 / Self::T;
}

pub trait BitPacking: FastLanes {
    fn pack<const W: usize>(input: &[Self; 1024], output: &mut [Self]);
    unsafe fn unchecked_pack(width: usize, input: &[Self], output: &mut [Self]);
    fn unpack<const W: usize>(input: &[Self], output: &mut [Self; 1024]);
    unsafe fn unchecked_unpack(width: usize, input: &[Self], output: &mut [Self]);
    fn unpack_single<const W: usize>(packed: &[Self], index: usize) -> Self;
    unsafe fn unchecked_unpack_single(width: usize, packed: &[Self], index: usize) -> Self;
}

pub trait FoR: BitPacking {
    fn for_pack<const W: usize>(input: &[Self; 1024], reference: Self, output: &mut [Self]);
    fn unfor_pack<const W: usize>(input: &[Self], reference: Self, output: &mut [Self; 1024]);
}

pub trait Delta: BitPacking {
    fn delta(input: &[Self; 1024], base: &[Self], output: &mut [Self; 1024]);
    fn undelta(input: &[Self; 1024], base: &[Self], output: &mut [Self; 1024]);
    fn undelta_pack<const W: usize>(input: &[Self], base: &[Self], output: &mut [Self; 1024]);
}

pub trait Transpose: FastLanes {
    fn transpose(input: &[Self; 1024], output: &mut [Self; 1024]);
    fn untranspose(input: &[Self; 1024], output: &mut [Self; 1024]);
}

/// `const fn transpose(idx)` of the reference (src/transpose.rs:29-36).
pub const fn transpose(idx: usize) -> usize {
    (idx % 16) * 64 + FL_ORDER[(idx / 16) % 8] * 8 + idx / 128
}

macro_rules! bind_type {
    ($T:ty, $pack:ident, $unpack:ident, $single:ident, $for_pack:ident, $unfor_pack:ident,
     $delta:ident, $undelta:ident, $undelta_pack:ident, $transpose:ident, $untranspose:ident) => {
        extern "C" {
            fn $pack(width: u32, n_blocks: usize, input: *const $T, packed: *mut $T) -> i32;
            fn $unpack(width: u32, n_blocks: usize, packed: *const $T, out: *mut $T) -> i32;
            fn $single(width: u32, packed: *const $T, index: usize, value: *mut $T) -> i32;
            fn $for_pack(width: u32, n_blocks: usize, input: *const $T, reference: $T, packed: *mut $T) -> i32;
            fn $unfor_pack(width: u32, n_blocks: usize, packed: *const $T, reference: $T, out: *mut $T) -> i32;
            fn $delta(n_blocks: usize, input: *const $T, base: *const $T, out: *mut $T) -> i32;
            fn $undelta(n_blocks: usize, input: *const $T, base: *const $T, out: *mut $T) -> i32;
            fn $undelta_pack(width: u32, n_blocks: usize, packed: *const $T, base: *const $T, out: *mut $T) -> i32;
            fn $transpose(n_blocks: usize, input: *const $T, out: *mut $T) -> i32;
            fn $untranspose(n_blocks: usize, input: *const $T, out: *mut $T) -> i32;
        }

        impl FastLanes for $T {}

        impl BitPacking for $T {
            fn pack<const W: usize>(input: &[Self; 1024], output: &mut [Self]) {
                assert_eq!(output.len(), 1024 * W / Self::T, "Output buffer must be of size 1024 * W / T");
                check(unsafe { $pack(W as u32, 1, input.as_ptr(), output.as_mut_ptr()) }, "pack");
            }
            unsafe fn unchecked_pack(width: usize, input: &[Self], output: &mut [Self]) {
                debug_assert_eq!(input.len(), 1024, "Input buffer must be of size 1024");
                debug_assert_eq!(output.len(), 128 * width / core::mem::size_of::<Self>());
                check($pack(width as u32, 1, input.as_ptr(), output.as_mut_ptr()), "unchecked_pack");
            }
            fn unpack<const W: usize>(input: &[Self], output: &mut [Self; 1024]) {
                assert_eq!(input.len(), 1024 * W / Self::T, "Input buffer must be of size 1024 * W / T");
                check(unsafe { $unpack(W as u32, 1, input.as_ptr(), output.as_mut_ptr()) }, "unpack");
            }
            unsafe fn unchecked_unpack(width: usize, input: &[Self], output: &mut [Self]) {
                debug_assert_eq!(output.len(), 1024, "Output buffer must be of size 1024");
                debug_assert_eq!(input.len(), 128 * width / core::mem::size_of::<Self>());
                check($unpack(width as u32, 1, input.as_ptr(), output.as_mut_ptr()), "unchecked_unpack");
            }
            fn unpack_single<const W: usize>(packed: &[Self], index: usize) -> Self {
                assert_eq!(packed.len(), 1024 * W / Self::T);
                let mut v: $T = 0;
                check(unsafe { $single(W as u32, packed.as_ptr(), index, &mut v) }, "unpack_single");
                v
            }
            unsafe fn unchecked_unpack_single(width: usize, packed: &[Self], index: usize) -> Self {
                let mut v: $T = 0;
                check($single(width as u32, packed.as_ptr(), index, &mut v), "unchecked_unpack_single");
                v
            }
        }

        impl FoR for $T {
            fn for_pack<const W: usize>(input: &[Self; 1024], reference: Self, output: &mut [Self]) {
                assert_eq!(output.len(), 1024 * W / Self::T);
                check(unsafe { $for_pack(W as u32, 1, input.as_ptr(), reference, output.as_mut_ptr()) }, "for_pack");
            }
            fn unfor_pack<const W: usize>(input: &[Self], reference: Self, output: &mut [Self; 1024]) {
                assert_eq!(input.len(), 1024 * W / Self::T);
                check(unsafe { $unfor_pack(W as u32, 1, input.as_ptr(), reference, output.as_mut_ptr()) }, "unfor_pack");
            }
        }

        impl Delta for $T {
            fn delta(input: &[Self; 1024], base: &[Self], output: &mut [Self; 1024]) {
                assert_eq!(base.len(), Self::LANES);
                check(unsafe { $delta(1, input.as_ptr(), base.as_ptr(), output.as_mut_ptr()) }, "delta");
            }
            fn undelta(input: &[Self; 1024], base: &[Self], output: &mut [Self; 1024]) {
                assert_eq!(base.len(), Self::LANES);
                check(unsafe { $undelta(1, input.as_ptr(), base.as_ptr(), output.as_mut_ptr()) }, "undelta");
            }
            fn undelta_pack<const W: usize>(input: &[Self], base: &[Self], output: &mut [Self; 1024]) {
                assert_eq!(input.len(), 1024 * W / Self::T);
                assert_eq!(base.len(), Self::LANES);
                check(unsafe { $undelta_pack(W as u32, 1, input.as_ptr(), base.as_ptr(), output.as_mut_ptr()) }, "undelta_pack");
            }
        }

        impl Transpose for $T {
            fn transpose(input: &[Self; 1024], output: &mut [Self; 1024]) {
                check(unsafe { $transpose(1, input.as_ptr(), output.as_mut_ptr()) }, "transpose");
            }
            fn untranspose(input: &[Self; 1024], output: &mut [Self; 1024]) {
                check(unsafe { $untranspose(1, input.as_ptr(), output.as_mut_ptr()) }, "untranspose");
            }
        }
    };
}

bind_type!(u8, fl_host_pack_u8, fl_host_unpack_u8, fl_host_unpack_single_u8, fl_host_for_pack_u8, fl_host_unfor_pack_u8,
           fl_host_delta_u8, fl_host_undelta_u8, fl_host_undelta_pack_u8, fl_host_transpose_u8, fl_host_untranspose_u8);
bind_type!(u16, fl_host_pack_u16, fl_host_unpack_u16, fl_host_unpack_single_u16, fl_host_for_pack_u16, fl_host_unfor_pack_u16,
           fl_host_delta_u16, fl_host_undelta_u16, fl_host_undelta_pack_u16, fl_host_transpose_u16, fl_host_untranspose_u16);
bind_type!(u32, fl_host_pack_u32, fl_host_unpack_u32, fl_host_unpack_single_u32, fl_host_for_pack_u32, fl_host_unfor_pack_u32,
           fl_host_delta_u32, fl_host_undelta_u32, fl_host_undelta_pack_u32, fl_host_transpose_u32, fl_host_untranspose_u32);
bind_type!(u64, fl_host_pack_u64, fl_host_unpack_u64, fl_host_unpack_single_u64, fl_host_for_pack_u64, fl_host_unfor_pack_u64,
           fl_host_delta_u64, fl_host_undelta_u64, fl_host_undelta_pack_u64, fl_host_transpose_u64, fl_host_untranspose_u64);

/// The reference's EXACT trait signatures (src/bitpacking.rs:8-59, src/delta.rs:6-17, src/ffor.rs:4-18,
/// src/transpose.rs:4-7 of spiraldb/fastlanes 0.1.8): packed operands are `[Self; 1024 * W / Self::T]` arrays and the
/// width is bounded at compile time, so existing call sites compile unchanged.  Every method forwards to the
/// slice-based impls above, i.e. to the same `fl_host_*` calls.  Never compiled in this repository (no Rust toolchain).
#[cfg(feature = "nightly-exact")]
pub mod exact {
    use core::mem::size_of;

    pub use super::{transpose, FastLanes, FL_ORDER};

    pub struct Pred<const B: bool>;
    pub trait Satisfied {}
    impl Satisfied for Pred<true> {}

    pub struct BitPackWidth<const W: usize>;
    pub trait SupportedBitPackWidth<T> {}
    impl<const W: usize, T> SupportedBitPackWidth<T> for BitPackWidth<W> where Pred<{ W <= 8 * size_of::<T>() }>: Satisfied {}

    pub trait BitPacking: FastLanes {
        fn pack<const W: usize>(input: &[Self; 1024], output: &mut [Self; 1024 * W / Self::T])
        where
            BitPackWidth<W>: SupportedBitPackWidth<Self>;
        unsafe fn unchecked_pack(width: usize, input: &[Self], output: &mut [Self]);
        fn unpack<const W: usize>(input: &[Self; 1024 * W / Self::T], output: &mut [Self; 1024])
        where
            BitPackWidth<W>: SupportedBitPackWidth<Self>;
        unsafe fn unchecked_unpack(width: usize, input: &[Self], output: &mut [Self]);
        fn unpack_single<const W: usize>(packed: &[Self; 1024 * W / Self::T], index: usize) -> Self
        where
            BitPackWidth<W>: SupportedBitPackWidth<Self>;
        unsafe fn unchecked_unpack_single(width: usize, input: &[Self], index: usize) -> Self;
    }

    pub trait FoR: BitPacking {
        fn for_pack<const W: usize>(input: &[Self; 1024], reference: Self, output: &mut [Self; 1024 * W / Self::T])
        where
            BitPackWidth<W>: SupportedBitPackWidth<Self>;
        fn unfor_pack<const W: usize>(input: &[Self; 1024 * W / Self::T], reference: Self, output: &mut [Self; 1024])
        where
            BitPackWidth<W>: SupportedBitPackWidth<Self>;
    }

    pub trait Delta: BitPacking {
        fn delta(input: &[Self; 1024], base: &[Self; Self::LANES], output: &mut [Self; 1024]);
        fn undelta(input: &[Self; 1024], base: &[Self; Self::LANES], output: &mut [Self; 1024]);
        fn undelta_pack<const W: usize>(input: &[Self; 1024 * W / Self::T], base: &[Self; Self::LANES], output: &mut [Self; 1024])
        where
            BitPackWidth<W>: SupportedBitPackWidth<Self>;
    }

    pub trait Transpose: FastLanes {
        fn transpose(input: &[Self; 1024], output: &mut [Self; 1024]);
        fn untranspose(input: &[Self; 1024], output: &mut [Self; 1024]);
    }

    macro_rules! exact_impl {
        ($T:ty) => {
            impl BitPacking for $T {
                fn pack<const W: usize>(input: &[Self; 1024], output: &mut [Self; 1024 * W / Self::T])
                where
                    BitPackWidth<W>: SupportedBitPackWidth<Self>,
                {
                    <$T as super::BitPacking>::pack::<W>(input, &mut output[..])
                }
                unsafe fn unchecked_pack(width: usize, input: &[Self], output: &mut [Self]) {
                    <$T as super::BitPacking>::unchecked_pack(width, input, output)
                }
                fn unpack<const W: usize>(input: &[Self; 1024 * W / Self::T], output: &mut [Self; 1024])
                where
                    BitPackWidth<W>: SupportedBitPackWidth<Self>,
                {
                    <$T as super::BitPacking>::unpack::<W>(&input[..], output)
                }
                unsafe fn unchecked_unpack(width: usize, input: &[Self], output: &mut [Self]) {
                    <$T as super::BitPacking>::unchecked_unpack(width, input, output)
                }
                fn unpack_single<const W: usize>(packed: &[Self; 1024 * W / Self::T], index: usize) -> Self
                where
                    BitPackWidth<W>: SupportedBitPackWidth<Self>,
                {
                    <$T as super::BitPacking>::unpack_single::<W>(&packed[..], index)
                }
                unsafe fn unchecked_unpack_single(width: usize, input: &[Self], index: usize) -> Self {
                    <$T as super::BitPacking>::unchecked_unpack_single(width, input, index)
                }
            }
            impl FoR for $T {
                fn for_pack<const W: usize>(input: &[Self; 1024], reference: Self, output: &mut [Self; 1024 * W / Self::T])
                where
                    BitPackWidth<W>: SupportedBitPackWidth<Self>,
                {
                    <$T as super::FoR>::for_pack::<W>(input, reference, &mut output[..])
                }
                fn unfor_pack<const W: usize>(input: &[Self; 1024 * W / Self::T], reference: Self, output: &mut [Self; 1024])
                where
                    BitPackWidth<W>: SupportedBitPackWidth<Self>,
                {
                    <$T as super::FoR>::unfor_pack::<W>(&input[..], reference, output)
                }
            }
            impl Delta for $T {
                fn delta(input: &[Self; 1024], base: &[Self; Self::LANES], output: &mut [Self; 1024]) {
                    <$T as super::Delta>::delta(input, &base[..], output)
                }
                fn undelta(input: &[Self; 1024], base: &[Self; Self::LANES], output: &mut [Self; 1024]) {
                    <$T as super::Delta>::undelta(input, &base[..], output)
                }
                fn undelta_pack<const W: usize>(input: &[Self; 1024 * W / Self::T], base: &[Self; Self::LANES], output: &mut [Self; 1024])
                where
                    BitPackWidth<W>: SupportedBitPackWidth<Self>,
                {
                    <$T as super::Delta>::undelta_pack::<W>(&input[..], &base[..], output)
                }
            }
            impl Transpose for $T {
                fn transpose(input: &[Self; 1024], output: &mut [Self; 1024]) {
                    <$T as super::Transpose>::transpose(input, output)
                }
                fn untranspose(input: &[Self; 1024], output: &mut [Self; 1024]) {
                    <$T as super::Transpose>::untranspose(input, output)
                }
            }
        };
    }
    exact_impl!(u8);
    exact_impl!(u16);
    exact_impl!(u32);
    exact_impl!(u64);
}

/// Batched, stream-ordered entry points on DEVICE pointers (the throughput path).
/// `stream` is a `cudaStream_t`; null = the legacy default stream.  Pointers must be 16-byte aligned.
pub mod device {
    use super::c_void;
    extern "C" {
        pub fn fl_unpack_u32(width: u32, n_blocks: usize, packed: *const u32, out: *mut u32, stream: *mut c_void) -> i32;
        pub fn fl_pack_u32(width: u32, n_blocks: usize, input: *const u32, packed: *mut u32, stream: *mut c_void) -> i32;
        pub fn fl_unfor_pack_u32(width: u32, n_blocks: usize, packed: *const u32, reference: u32, out: *mut u32, stream: *mut c_void) -> i32;
        pub fn fl_undelta_pack_u32(width: u32, n_blocks: usize, packed: *const u32, base: *const u32, out: *mut u32, stream: *mut c_void) -> i32;
        pub fn fl_unpack_gather_u32(width: u32, n_blocks: usize, packed: *const u32, global_index: *const u64, n: usize,
                                    out: *mut u32, oob_flag: *mut i32, stream: *mut c_void) -> i32;
        /// Fused scan (not in the reference: `unfor_pack` + the caller-side predicate loop of README.md:40-41 in one
        /// pass).  bitmap: 128 bytes per block, bit i = lo <= unfor_pack(packed, reference)[i] <= hi; counts nullable.
        pub fn fl_unpack_filter_u32(width: u32, n_blocks: usize, packed: *const u32, refs: *const u32, reference: u32,
                                    lo: u32, hi: u32, bitmap: *mut u8, counts: *mut u32, stream: *mut c_void) -> i32;
        /// Delta scan: bit i = lo <= untranspose(undelta_pack(packed, base))[i] <= hi (original value order).
        pub fn fl_undelta_pack_filter_u32(width: u32, n_blocks: usize, packed: *const u32, base: *const u32, lo: u32, hi: u32,
                                          bitmap: *mut u8, counts: *mut u32, stream: *mut c_void) -> i32;
        /// Dense compaction of the selected values: out[offsets[b] + k] = k-th selected value of block b.
        pub fn fl_unpack_select_u32(width: u32, n_blocks: usize, packed: *const u32, refs: *const u32, reference: u32,
                                    bitmap: *const u8, offsets: *const u64, out: *mut u32, stream: *mut c_void) -> i32;
        // ... the same set exists for u8 / u16 / u64 and for pack / for_pack / delta / undelta / (un)transpose:
        // see include/fastlanes_b200.h (FL_DECLARE_TYPE).
    }
}

/// Host-slice scan: `FoR::unfor_pack::<W>` + range predicate over a batch of packed blocks; the decoded values never
/// leave the GPU, only the selection bitmap (128 bytes per block) and the per-block counts come back.
pub mod scan {
    extern "C" {
        fn fl_host_unpack_filter_u32(width: u32, n_blocks: usize, packed: *const u32, reference: u32, lo: u32, hi: u32,
                                     bitmap: *mut u8, counts: *mut u32) -> i32;
    }
    pub fn filter_range_u32(width: usize, packed: &[u32], reference: u32, lo: u32, hi: u32, bitmap: &mut [u8],
                            counts: Option<&mut [u32]>) {
        let n_blocks = bitmap.len() / 128;
        assert_eq!(bitmap.len(), n_blocks * 128);
        assert_eq!(packed.len(), n_blocks * 32 * width);
        let c = match counts {
            Some(c) => { assert_eq!(c.len(), n_blocks); c.as_mut_ptr() }
            None => core::ptr::null_mut(),
        };
        super::check(unsafe { fl_host_unpack_filter_u32(width as u32, n_blocks, packed.as_ptr(), reference, lo, hi,
                                                       bitmap.as_mut_ptr(), c) }, "unpack_filter");
    }
}

/// Multi-GPU in one process: a context owns one worker thread per device and runs contiguous block shards of every
/// array of a call on the devices concurrently (`fl_ctx_*`, include/fastlanes_b200.h).  The batched counterpart of the
/// runtime-width family (`src/bitpacking.rs:109-129` in the reference): whole column chunks, host slices in and out.
pub mod context {
    use core::ffi::c_void;
    #[repr(C)]
    pub struct FlCtx { _private: [u8; 0] }
    extern "C" {
        fn fl_ctx_create(devices: *const i32, n_devices: i32, ctx: *mut *mut FlCtx) -> i32;
        fn fl_ctx_destroy(ctx: *mut FlCtx) -> i32;
        fn fl_ctx_device_count(ctx: *const FlCtx) -> i32;
        fn fl_ctx_block_range(ctx: *const FlCtx, n_blocks: usize, i: i32, first: *mut usize, end: *mut usize) -> i32;
        fn fl_ctx_host_alloc(ctx: *mut FlCtx, n_blocks: usize, bytes_per_block: usize, p: *mut *mut c_void) -> i32;
        fn fl_host_free(p: *mut c_void) -> i32;
        fn fl_ctx_host_unpack_u32(ctx: *mut FlCtx, width: u32, n_blocks: usize, packed: *const u32, out: *mut u32) -> i32;
        fn fl_ctx_host_pack_u32(ctx: *mut FlCtx, width: u32, n_blocks: usize, input: *const u32, packed: *mut u32) -> i32;
        fn fl_ctx_host_undelta_pack_u32(ctx: *mut FlCtx, width: u32, n_blocks: usize, packed: *const u32, base: *const u32,
                                        out: *mut u32) -> i32;
        fn fl_ctx_host_unpack_filter_u32(ctx: *mut FlCtx, width: u32, n_blocks: usize, packed: *const u32, reference: u32,
                                         lo: u32, hi: u32, bitmap: *mut u8, counts: *mut u32) -> i32;
        // ... the same for u8 / u16 / u64 and for for_pack / unfor_pack / delta / undelta / (un)transpose / fused chains
    }

    pub struct Context { raw: *mut FlCtx }
    unsafe impl Send for Context {}

    impl Context {
        /// `devices`: CUDA ordinals; empty = every visible device.
        pub fn new(devices: &[i32]) -> Self {
            let mut raw = core::ptr::null_mut();
            let st = unsafe { fl_ctx_create(if devices.is_empty() { core::ptr::null() } else { devices.as_ptr() }, devices.len() as i32, &mut raw) };
            super::check(st, "fl_ctx_create");
            Context { raw }
        }
        pub fn devices(&self) -> usize { unsafe { fl_ctx_device_count(self.raw) as usize } }
        /// Blocks `[first, end)` of an `n_blocks` batch run on shard `i`.
        pub fn block_range(&self, n_blocks: usize, i: usize) -> (usize, usize) {
            let (mut a, mut b) = (0usize, 0usize);
            super::check(unsafe { fl_ctx_block_range(self.raw, n_blocks, i as i32, &mut a, &mut b) }, "fl_ctx_block_range");
            (a, b)
        }
        /// Page-locked `u32` buffer whose pages sit on the NUMA node of the device that will copy them.
        pub fn pinned_u32(&mut self, n_blocks: usize, elems_per_block: usize) -> PinnedU32 {
            let mut p = core::ptr::null_mut();
            super::check(unsafe { fl_ctx_host_alloc(self.raw, n_blocks, elems_per_block * 4, &mut p) }, "fl_ctx_host_alloc");
            PinnedU32 { ptr: p as *mut u32, len: n_blocks * elems_per_block }
        }
        /// Batched `BitPacking::unchecked_unpack` over `output.len() / 1024` blocks, sharded over the devices.
        pub fn unpack_u32(&mut self, width: usize, packed: &[u32], output: &mut [u32]) {
            let n = output.len() / 1024;
            assert_eq!(output.len(), n * 1024, "Output buffer must be a whole number of 1024-element blocks");
            assert_eq!(packed.len(), n * 32 * width, "Input buffer must be of size n * 1024 * W / T");
            super::check(unsafe { fl_ctx_host_unpack_u32(self.raw, width as u32, n, packed.as_ptr(), output.as_mut_ptr()) }, "ctx unpack");
        }
        pub fn pack_u32(&mut self, width: usize, input: &[u32], packed: &mut [u32]) {
            let n = input.len() / 1024;
            assert_eq!(input.len(), n * 1024);
            assert_eq!(packed.len(), n * 32 * width);
            super::check(unsafe { fl_ctx_host_pack_u32(self.raw, width as u32, n, input.as_ptr(), packed.as_mut_ptr()) }, "ctx pack");
        }
        pub fn undelta_pack_u32(&mut self, width: usize, packed: &[u32], base: &[u32], output: &mut [u32]) {
            let n = output.len() / 1024;
            assert_eq!(packed.len(), n * 32 * width);
            assert_eq!(base.len(), n * 32);
            super::check(unsafe { fl_ctx_host_undelta_pack_u32(self.raw, width as u32, n, packed.as_ptr(), base.as_ptr(), output.as_mut_ptr()) }, "ctx undelta_pack");
        }
        pub fn filter_range_u32(&mut self, width: usize, packed: &[u32], reference: u32, lo: u32, hi: u32, bitmap: &mut [u8], counts: &mut [u32]) {
            let n = counts.len();
            assert_eq!(bitmap.len(), n * 128);
            assert_eq!(packed.len(), n * 32 * width);
            super::check(unsafe { fl_ctx_host_unpack_filter_u32(self.raw, width as u32, n, packed.as_ptr(), reference, lo, hi, bitmap.as_mut_ptr(), counts.as_mut_ptr()) }, "ctx filter");
        }
    }
    impl Drop for Context {
        fn drop(&mut self) { unsafe { fl_ctx_destroy(self.raw) }; }
    }

    pub struct PinnedU32 { ptr: *mut u32, len: usize }
    impl core::ops::Deref for PinnedU32 {
        type Target = [u32];
        fn deref(&self) -> &[u32] { unsafe { core::slice::from_raw_parts(self.ptr, self.len) } }
    }
    impl core::ops::DerefMut for PinnedU32 {
        fn deref_mut(&mut self) -> &mut [u32] { unsafe { core::slice::from_raw_parts_mut(self.ptr, self.len) } }
    }
    impl Drop for PinnedU32 {
        fn drop(&mut self) { unsafe { fl_host_free(self.ptr as *mut c_void) }; }
    }
}

#[cfg(test)]
mod tests {
    //! The reference's own tests, verbatim in spirit (src/bitpacking.rs:249-271, src/lib.rs:71-96).
    use super::*;

    #[test]
    fn pack_u16_into_u3() {
        const W: usize = 3;
        let mut values = [0u16; 1024];
        for i in 0..1024 { values[i] = (i % (1 << W)) as u16; }
        let mut packed = [0u16; 128 * W / 2];
        <u16 as BitPacking>::pack::<W>(&values, &mut packed);
        let mut unpacked = [0u16; 1024];
        <u16 as BitPacking>::unpack::<W>(&packed, &mut unpacked);
        assert_eq!(values, unpacked);
        for i in 0..1024 { assert_eq!(<u16 as BitPacking>::unpack_single::<W>(&packed, i), values[i]); }
    }

    #[test]
    fn unchecked_pack_u32_w10() {
        let input: [u32; 1024] = core::array::from_fn(|i| i as u32);
        let mut packed = [0u32; 320];
        unsafe { <u32 as BitPacking>::unchecked_pack(10, &input, &mut packed) };
        let mut output = [0u32; 1024];
        unsafe { <u32 as BitPacking>::unchecked_unpack(10, &packed, &mut output) };
        assert_eq!(input, output);
    }
}
