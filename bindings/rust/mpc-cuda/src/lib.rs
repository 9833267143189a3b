//! Thin binding of `include/mpc_cuda.h`.  Safe wrappers take the reference's own types
//! (`ark_bls12_377::{Fr, G1Affine, G2Affine}`), marshal them to Montgomery u64 limbs — the in-memory
//! representation of `Fp256.0.0` / `Fp384.0.0` — and panic on a non-zero status, mirroring the
//! `assert!`/`unwrap` convention of the code they replace.
//!
//! Source only: this repository's build container has no Rust toolchain, so the crate is not compiled
//! there; the same ABI is driven by `zk-mpc_b200/host.py` (ctypes) in the tests.
use ark_bls12_377::{Fq, Fq2, Fr, G1Affine, G2Affine};
use ark_ec::AffineCurve;
use ark_ff::{Fp256, Fp384, Zero};
use std::os::raw::c_char;

#[allow(non_camel_case_types)]
pub mod ffi {
    use std::os::raw::{c_char, c_void};
    extern "C" {
        pub fn mpc_cuda_init(devices: *const i32, n_dev: i32) -> i32;
        pub fn mpc_cuda_set_party(party_id: u32, n_parties: u32) -> i32;
        pub fn mpc_cuda_last_error() -> *const c_char;
        pub fn mpc_cuda_msm_g1(bases_xy: *const u64, inf: *const u8, scalars_mont: *const u64, n: usize,
                               out_xy: *mut u64, out_inf: *mut u8) -> i32;
        pub fn mpc_cuda_msm_g2(bases_xy: *const u64, inf: *const u8, scalars_mont: *const u64, n: usize,
                               out_xy: *mut u64, out_inf: *mut u8) -> i32;
        pub fn mpc_cuda_msm_g1_register_bases(bases_xy: *const u64, inf: *const u8, n: usize, handle: *mut u64) -> i32;
        pub fn mpc_cuda_msm_g1_handle(handle: u64, offset: usize, scalars_mont: *const u64, n: usize,
                                      out_xy: *mut u64, out_inf: *mut u8) -> i32;
        pub fn mpc_cuda_msm_release_bases(handle: u64) -> i32;
        pub fn mpc_cuda_ntt_fr(data: *mut u64, log_n: u32, kind: u32, batch: u32) -> i32;
        pub fn mpc_cuda_divide_by_vanishing_on_coset(data: *mut u64, log_n: u32) -> i32;
        pub fn mpc_cuda_beaver_mask(s: *const u64, x: *const u64, out: *mut u64, n: usize) -> i32;
        pub fn mpc_cuda_beaver_combine(x: *const u64, y: *const u64, z: *const u64, sx: *const u64, oy: *const u64,
                                       out: *mut u64, n: usize, is_leader: u32, spdz: u32) -> i32;
        pub fn mpc_cuda_open_sum(parts: *const u64, n_parties: u32, out: *mut u64, n: usize) -> i32;
        pub fn mpc_cuda_spdz_mac_check(vals: *const u64, macs: *const u64, out: *mut u64, n: usize, is_leader: u32) -> i32;
        pub fn mpc_cuda_vec_op(op: u32, a: *const u64, b: *const u64, c: *const u64, out: *mut u64, n: usize) -> i32;
        pub fn mpc_cuda_stream_sync(stream: *mut c_void) -> i32;
    }
}

fn check(rc: i32) {
    if rc != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(ffi::mpc_cuda_last_error() as *const c_char) };
        panic!("mpc_cuda error {}: {}", rc, msg.to_string_lossy());
    }
}

pub const FFT: u32 = 0;
pub const IFFT: u32 = 1;
pub const COSET_FFT: u32 = 2;
pub const COSET_IFFT: u32 = 3;

/// `mpc_net::MpcNet::party_id()` / `n_parties()`; call once per party thread.
pub fn set_party(party_id: u32, n_parties: u32) { check(unsafe { ffi::mpc_cuda_set_party(party_id, n_parties) }) }

#[inline] fn fr_limbs(v: &[Fr]) -> Vec<u64> { v.iter().flat_map(|f| (f.0).0).collect() }
#[inline] fn fr_from(l: &[u64]) -> Fr { Fp256::new(ark_ff::BigInteger256([l[0], l[1], l[2], l[3]])) }
#[inline] fn fq_from(l: &[u64]) -> Fq { Fp384::new(ark_ff::BigInteger384([l[0], l[1], l[2], l[3], l[4], l[5]])) }

/// Drop-in body of `AffineMsm::<G1Affine>::msm` (mpc-algebra/src/share/msm.rs:33-37).
pub fn msm_g1(bases: &[G1Affine], scalars: &[Fr]) -> G1Affine {
    let n = bases.len().min(scalars.len());                   // variable_base.rs:16-18
    let mut xy = Vec::with_capacity(12 * n);
    let mut inf = Vec::with_capacity(n);
    for b in &bases[..n] { xy.extend_from_slice(&(b.x.0).0); xy.extend_from_slice(&(b.y.0).0); inf.push(b.infinity as u8); }
    let sc = fr_limbs(&scalars[..n]);
    let (mut out, mut oinf) = ([0u64; 12], 0u8);
    check(unsafe { ffi::mpc_cuda_msm_g1(xy.as_ptr(), inf.as_ptr(), sc.as_ptr(), n, out.as_mut_ptr(), &mut oinf) });
    if oinf != 0 { G1Affine::zero() } else { G1Affine::new(fq_from(&out[0..6]), fq_from(&out[6..12]), false) }
}

/// Same over G2 (`b_g2_query`, src/groth16.rs:160).
pub fn msm_g2(bases: &[G2Affine], scalars: &[Fr]) -> G2Affine {
    let n = bases.len().min(scalars.len());
    let mut xy = Vec::with_capacity(24 * n);
    let mut inf = Vec::with_capacity(n);
    for b in &bases[..n] {
        for c in [&b.x.c0, &b.x.c1, &b.y.c0, &b.y.c1] { xy.extend_from_slice(&(c.0).0); }
        inf.push(b.infinity as u8);
    }
    let sc = fr_limbs(&scalars[..n]);
    let (mut out, mut oinf) = ([0u64; 24], 0u8);
    check(unsafe { ffi::mpc_cuda_msm_g2(xy.as_ptr(), inf.as_ptr(), sc.as_ptr(), n, out.as_mut_ptr(), &mut oinf) });
    if oinf != 0 { return G2Affine::zero(); }
    G2Affine::new(Fq2::new(fq_from(&out[0..6]), fq_from(&out[6..12])), Fq2::new(fq_from(&out[12..18]), fq_from(&out[18..24])), false)
}

/// In-place transform of a vector already resized to the domain size (radix2/mod.rs:99-114).
pub fn ntt_in_place(v: &mut [Fr], kind: u32) {
    assert!(v.len().is_power_of_two());
    let mut l = fr_limbs(v);
    check(unsafe { ffi::mpc_cuda_ntt_fr(l.as_mut_ptr(), v.len().trailing_zeros(), kind, 1) });
    for (o, c) in v.iter_mut().zip(l.chunks_exact(4)) { *o = fr_from(c); }
}

/// Local half of `FieldShare::batch_mul` after the two opens (share/field.rs:118-128); additive layout.
pub fn beaver_combine(x: &[Fr], y: &[Fr], z: &[Fr], sx: &[Fr], oy: &[Fr], is_leader: bool) -> Vec<Fr> {
    let n = sx.len();
    let mut out = vec![0u64; 4 * n];
    check(unsafe { ffi::mpc_cuda_beaver_combine(fr_limbs(x).as_ptr(), fr_limbs(y).as_ptr(), fr_limbs(z).as_ptr(),
        fr_limbs(sx).as_ptr(), fr_limbs(oy).as_ptr(), out.as_mut_ptr(), n, is_leader as u32, 0) });
    out.chunks_exact(4).map(fr_from).collect()
}

/// `s_i + x_i` before the open (share/field.rs:108-117).
pub fn beaver_mask(s: &[Fr], x: &[Fr]) -> Vec<Fr> {
    let mut out = vec![0u64; 4 * s.len()];
    check(unsafe { ffi::mpc_cuda_beaver_mask(fr_limbs(s).as_ptr(), fr_limbs(x).as_ptr(), out.as_mut_ptr(), s.len()) });
    out.chunks_exact(4).map(fr_from).collect()
}
