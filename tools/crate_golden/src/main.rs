//! crate_golden: feed tests/golden inputs through the real `fastlanes` crate (v0.1.8) and dump what it produces.
//!
//!   cargo run --release -- <in_dir> <out_dir>
//!
//! <in_dir> (written by export_inputs.py), little-endian raw arrays per element type T in {8,16,32,64}:
//!   u{T}_values.bin      N_BLOCKS * 1024 elements (full-range bits: exercises the mask truncation of pack!)
//!   u{T}_base.bin        N_BLOCKS * LANES elements
//!   u{T}_reference.bin   1 element
//! <out_dir>, for every T and every W in 0..=T (all blocks concatenated):
//!   u{T}_w{W}_packed.bin         BitPacking::pack::<W>(values)                       src/bitpacking.rs:65-74
//!   u{T}_w{W}_unpacked.bin       BitPacking::unpack::<W>(packed)                     src/bitpacking.rs:98-107
//!   u{T}_w{W}_rt_packed.bin      unchecked_pack(W, values)  (runtime-width dispatch) src/bitpacking.rs:76-96
//!   u{T}_w{W}_single.bin         unpack_single::<W>(packed block 0, i), i in 0..1024 src/bitpacking.rs:132-179
//!   u{T}_w{W}_for_packed.bin     FoR::for_pack::<W>(values, reference)               src/ffor.rs:24-36
//!   u{T}_w{W}_unfor_pack.bin     FoR::unfor_pack::<W>(packed, reference)             src/ffor.rs:38-50
//!   u{T}_w{W}_undelta_pack.bin   Delta::undelta_pack::<W>(packed, base)              src/delta.rs:48-63
//! and per T:
//!   u{T}_transposed.bin, u{T}_untransposed.bin   Transpose::{transpose,untranspose}(values)   src/transpose.rs:11-22
//!   u{T}_delta_of_transposed.bin                 Delta::delta(transposed, base)                src/delta.rs:24-33
//!   u{T}_undelta_of_values.bin                   Delta::undelta(values, base)                  src/delta.rs:36-45
#![allow(incomplete_features)]
#![feature(generic_const_exprs)]

use fastlanes::{BitPacking, Delta, FoR, Transpose};
use seq_macro::seq;
use std::fs;
use std::path::Path;

fn read_bytes(dir: &Path, name: &str) -> Vec<u8> {
    fs::read(dir.join(name)).unwrap_or_else(|e| panic!("cannot read {name}: {e}"))
}
fn write_bytes(dir: &Path, name: &str, bytes: &[u8]) {
    fs::write(dir.join(name), bytes).unwrap_or_else(|e| panic!("cannot write {name}: {e}"));
}

macro_rules! run_type {
    ($T:ty, $TB:literal, $in_dir:expr, $out_dir:expr) => {{
        const SZ: usize = core::mem::size_of::<$T>();
        const LANES: usize = 1024 / $TB;
        let to_vec = |b: Vec<u8>| -> Vec<$T> {
            b.chunks_exact(SZ).map(|c| <$T>::from_le_bytes(c.try_into().unwrap())).collect()
        };
        let to_bytes = |v: &[$T]| -> Vec<u8> { v.iter().flat_map(|x| x.to_le_bytes()).collect() };
        let values = to_vec(read_bytes($in_dir, &format!("u{}_values.bin", $TB)));
        let base = to_vec(read_bytes($in_dir, &format!("u{}_base.bin", $TB)));
        let reference: $T = to_vec(read_bytes($in_dir, &format!("u{}_reference.bin", $TB)))[0];
        assert_eq!(values.len() % 1024, 0);
        let n_blocks = values.len() / 1024;
        assert_eq!(base.len(), n_blocks * LANES);

        // Transpose / Delta (no width)
        let (mut tr, mut utr, mut dl, mut udl) = (Vec::new(), Vec::new(), Vec::new(), Vec::new());
        for b in 0..n_blocks {
            let v: &[$T; 1024] = values[b * 1024..(b + 1) * 1024].try_into().unwrap();
            let bs: &[$T; LANES] = base[b * LANES..(b + 1) * LANES].try_into().unwrap();
            let mut t = [0 as $T; 1024];
            Transpose::transpose(v, &mut t);
            let mut u = [0 as $T; 1024];
            Transpose::untranspose(v, &mut u);
            let mut d = [0 as $T; 1024];
            Delta::delta(&t, bs, &mut d);
            let mut ud = [0 as $T; 1024];
            Delta::undelta(v, bs, &mut ud);
            tr.extend_from_slice(&t);
            utr.extend_from_slice(&u);
            dl.extend_from_slice(&d);
            udl.extend_from_slice(&ud);
        }
        write_bytes($out_dir, &format!("u{}_transposed.bin", $TB), &to_bytes(&tr));
        write_bytes($out_dir, &format!("u{}_untransposed.bin", $TB), &to_bytes(&utr));
        write_bytes($out_dir, &format!("u{}_delta_of_transposed.bin", $TB), &to_bytes(&dl));
        write_bytes($out_dir, &format!("u{}_undelta_of_values.bin", $TB), &to_bytes(&udl));

        // every width
        seq!(W in 0..=$TB {
            {
                const PL: usize = 1024 * W / $TB;
                let (mut packed_all, mut unpacked_all, mut rt_all) = (Vec::new(), Vec::new(), Vec::new());
                let (mut forp_all, mut unfor_all, mut undp_all) = (Vec::new(), Vec::new(), Vec::new());
                let mut single = Vec::new();
                for b in 0..n_blocks {
                    let v: &[$T; 1024] = values[b * 1024..(b + 1) * 1024].try_into().unwrap();
                    let bs: &[$T; LANES] = base[b * LANES..(b + 1) * LANES].try_into().unwrap();
                    let mut packed = [0 as $T; PL];
                    <$T as BitPacking>::pack::<W>(v, &mut packed);
                    let mut unpacked = [0 as $T; 1024];
                    <$T as BitPacking>::unpack::<W>(&packed, &mut unpacked);
                    let mut rt = vec![0 as $T; PL];
                    unsafe { <$T as BitPacking>::unchecked_pack(W, &v[..], &mut rt[..]) };
                    let mut forp = [0 as $T; PL];
                    <$T as FoR>::for_pack::<W>(v, reference, &mut forp);
                    let mut unfor = [0 as $T; 1024];
                    <$T as FoR>::unfor_pack::<W>(&packed, reference, &mut unfor);
                    let mut undp = [0 as $T; 1024];
                    <$T as Delta>::undelta_pack::<W>(&packed, bs, &mut undp);
                    if b == 0 {
                        for i in 0..1024 {
                            single.push(<$T as BitPacking>::unpack_single::<W>(&packed, i));
                        }
                    }
                    packed_all.extend_from_slice(&packed);
                    unpacked_all.extend_from_slice(&unpacked);
                    rt_all.extend_from_slice(&rt);
                    forp_all.extend_from_slice(&forp);
                    unfor_all.extend_from_slice(&unfor);
                    undp_all.extend_from_slice(&undp);
                }
                let w = W;
                write_bytes($out_dir, &format!("u{}_w{}_packed.bin", $TB, w), &to_bytes(&packed_all));
                write_bytes($out_dir, &format!("u{}_w{}_unpacked.bin", $TB, w), &to_bytes(&unpacked_all));
                write_bytes($out_dir, &format!("u{}_w{}_rt_packed.bin", $TB, w), &to_bytes(&rt_all));
                write_bytes($out_dir, &format!("u{}_w{}_single.bin", $TB, w), &to_bytes(&single));
                write_bytes($out_dir, &format!("u{}_w{}_for_packed.bin", $TB, w), &to_bytes(&forp_all));
                write_bytes($out_dir, &format!("u{}_w{}_unfor_pack.bin", $TB, w), &to_bytes(&unfor_all));
                write_bytes($out_dir, &format!("u{}_w{}_undelta_pack.bin", $TB, w), &to_bytes(&undp_all));
            }
        });
    }};
}

fn main() {
    let args: Vec<String> = std::env::args().collect();
    if args.len() != 3 {
        eprintln!("usage: crate_golden <in_dir> <out_dir>");
        std::process::exit(2);
    }
    let (in_dir, out_dir) = (Path::new(&args[1]), Path::new(&args[2]));
    fs::create_dir_all(out_dir).expect("create out_dir");
    run_type!(u8, 8, in_dir, out_dir);
    run_type!(u16, 16, in_dir, out_dir);
    run_type!(u32, 32, in_dir, out_dir);
    run_type!(u64, 64, in_dir, out_dir);
    println!("crate_golden: wrote outputs of fastlanes 0.1.8 to {}", out_dir.display());
}
