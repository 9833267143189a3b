//! Thin binding of `include/mpc_cuda.h`.  `ffi` (generated from the header by tools/gen_rust_ffi.py) declares
//! every function of the C ABI; the safe wrappers below take the reference's own types
//! (`ark_bls12_377::{Fr, G1Affine, G2Affine}`), marshal them to Montgomery u64 limbs — the in-memory
//! representation of `Fp256.0.0` / `Fp384.0.0` — and panic on a non-zero status, mirroring the
//! `assert!`/`unwrap` convention of the code they replace.  Every wrapper asserts that its slices have the
//! lengths the C side will read.
//!
//! UNCOMPILED SKETCH: this repository's build container has no Rust toolchain, so the crate has never been
//! compiled and its field accessors / constructors are unverified against the reference's arkworks fork.
//! The same ABI is driven by `zk-mpc_b200/host.py` (ctypes) and `tests/c_abi/abi_check.c` (C) in the tests.
use ark_bls12_377::{Fq, Fq2, Fr, G1Affine, G2Affine};
use ark_ec::AffineCurve;
use ark_ff::{Fp256, Fp384, Zero};
use std::os::raw::c_char;

pub mod ffi;

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
#[inline] fn frs_from(l: &[u64]) -> Vec<Fr> { l.chunks_exact(4).map(fr_from).collect() }

fn g1_limbs(bases: &[G1Affine]) -> (Vec<u64>, Vec<u8>) {
    let mut xy = Vec::with_capacity(12 * bases.len());
    let mut inf = Vec::with_capacity(bases.len());
    for b in bases { xy.extend_from_slice(&(b.x.0).0); xy.extend_from_slice(&(b.y.0).0); inf.push(b.infinity as u8); }
    (xy, inf)
}
fn g2_limbs(bases: &[G2Affine]) -> (Vec<u64>, Vec<u8>) {
    let mut xy = Vec::with_capacity(24 * bases.len());
    let mut inf = Vec::with_capacity(bases.len());
    for b in bases {
        for c in [&b.x.c0, &b.x.c1, &b.y.c0, &b.y.c1] { xy.extend_from_slice(&(c.0).0); }
        inf.push(b.infinity as u8);
    }
    (xy, inf)
}
fn g1_from(out: &[u64; 12], inf: u8) -> G1Affine {
    if inf != 0 { G1Affine::zero() } else { G1Affine::new(fq_from(&out[0..6]), fq_from(&out[6..12]), false) }
}
fn g2_from(out: &[u64; 24], inf: u8) -> G2Affine {
    if inf != 0 { return G2Affine::zero(); }
    G2Affine::new(Fq2::new(fq_from(&out[0..6]), fq_from(&out[6..12])), Fq2::new(fq_from(&out[12..18]), fq_from(&out[18..24])), false)
}

/// Drop-in body of `AffineMsm::<G1Affine>::msm` (mpc-algebra/src/share/msm.rs:33-37).
pub fn msm_g1(bases: &[G1Affine], scalars: &[Fr]) -> G1Affine {
    let n = bases.len().min(scalars.len());                   // variable_base.rs:16-18
    let (xy, inf) = g1_limbs(&bases[..n]);
    let sc = fr_limbs(&scalars[..n]);
    assert_eq!(xy.len(), 12 * n); assert_eq!(inf.len(), n); assert_eq!(sc.len(), 4 * n);
    let (mut out, mut oinf) = ([0u64; 12], 0u8);
    check(unsafe { ffi::mpc_cuda_msm_g1(xy.as_ptr(), inf.as_ptr(), sc.as_ptr(), n, out.as_mut_ptr(), &mut oinf) });
    g1_from(&out, oinf)
}

/// Same over G2 (`b_g2_query`, src/groth16.rs:160).
pub fn msm_g2(bases: &[G2Affine], scalars: &[Fr]) -> G2Affine {
    let n = bases.len().min(scalars.len());
    let (xy, inf) = g2_limbs(&bases[..n]);
    let sc = fr_limbs(&scalars[..n]);
    assert_eq!(xy.len(), 24 * n); assert_eq!(inf.len(), n); assert_eq!(sc.len(), 4 * n);
    let (mut out, mut oinf) = ([0u64; 24], 0u8);
    check(unsafe { ffi::mpc_cuda_msm_g2(xy.as_ptr(), inf.as_ptr(), sc.as_ptr(), n, out.as_mut_ptr(), &mut oinf) });
    g2_from(&out, oinf)
}

/// A CRS vector kept resident (pk.*_query, powers_of_g); `parts > 0` shards it over that many GPUs of the process.
pub struct G1Bases { handle: u64, len: usize }
impl G1Bases {
    pub fn register(bases: &[G1Affine], parts: u32, precompute: bool) -> Self {
        let (xy, inf) = g1_limbs(bases);
        let mut handle = 0u64;
        check(unsafe {
            if parts > 0 { ffi::mpc_cuda_msm_g1_register_bases_sharded(xy.as_ptr(), inf.as_ptr(), bases.len(), parts, &mut handle) }
            else { ffi::mpc_cuda_msm_g1_register_bases(xy.as_ptr(), inf.as_ptr(), bases.len(), &mut handle) }
        });
        if precompute { check(unsafe { ffi::mpc_cuda_msm_g1_precompute(handle, 0) }); }
        G1Bases { handle, len: bases.len() }
    }
    /// MSM over `bases[offset .. offset + scalars.len()]`
    pub fn msm(&self, offset: usize, scalars: &[Fr]) -> G1Affine {
        assert!(offset <= self.len && scalars.len() <= self.len - offset);
        let sc = fr_limbs(scalars);
        let (mut out, mut oinf) = ([0u64; 12], 0u8);
        check(unsafe { ffi::mpc_cuda_msm_g1_handle(self.handle, offset, sc.as_ptr(), scalars.len(), out.as_mut_ptr(), &mut oinf) });
        g1_from(&out, oinf)
    }
}
impl Drop for G1Bases { fn drop(&mut self) { unsafe { ffi::mpc_cuda_msm_release_bases(self.handle); } } }

/// In-place transform of a vector already resized to the domain size (radix2/mod.rs:99-114).
pub fn ntt_in_place(v: &mut [Fr], kind: u32) {
    assert!(v.len().is_power_of_two() && kind <= COSET_IFFT);
    let mut l = fr_limbs(v);
    assert_eq!(l.len(), 4 * v.len());
    check(unsafe { ffi::mpc_cuda_ntt_fr(l.as_mut_ptr(), v.len().trailing_zeros(), kind, 1) });
    for (o, c) in v.iter_mut().zip(l.chunks_exact(4)) { *o = fr_from(c); }
}

/// Local half of `FieldShare::batch_mul` after the two opens (share/field.rs:118-128); additive layout.
pub fn beaver_combine(x: &[Fr], y: &[Fr], z: &[Fr], sx: &[Fr], oy: &[Fr], is_leader: bool) -> Vec<Fr> {
    let n = sx.len();
    assert_eq!(x.len(), n); assert_eq!(y.len(), n); assert_eq!(z.len(), n); assert_eq!(oy.len(), n);
    let mut out = vec![0u64; 4 * n];
    let (lx, ly, lz, lsx, loy) = (fr_limbs(x), fr_limbs(y), fr_limbs(z), fr_limbs(sx), fr_limbs(oy));
    check(unsafe { ffi::mpc_cuda_beaver_combine(lx.as_ptr(), ly.as_ptr(), lz.as_ptr(), lsx.as_ptr(), loy.as_ptr(),
        out.as_mut_ptr(), n, is_leader as u32, 0) });
    frs_from(&out)
}

/// `s_i + x_i` before the open (share/field.rs:108-117).
pub fn beaver_mask(s: &[Fr], x: &[Fr]) -> Vec<Fr> {
    assert_eq!(s.len(), x.len());
    let mut out = vec![0u64; 4 * s.len()];
    let (ls, lx) = (fr_limbs(s), fr_limbs(x));
    check(unsafe { ffi::mpc_cuda_beaver_mask(ls.as_ptr(), lx.as_ptr(), out.as_mut_ptr(), s.len()) });
    frs_from(&out)
}

/// The masked vector as the bytes `MpcSerNet::broadcast` sends (channel.rs:12-28): mask + serialise in one pass.
pub fn beaver_mask_wire(s: &[Fr], x: &[Fr]) -> Vec<u8> {
    assert_eq!(s.len(), x.len());
    let mut out = vec![0u8; 8 + 32 * s.len()];
    let (ls, lx) = (fr_limbs(s), fr_limbs(x));
    check(unsafe { ffi::mpc_cuda_beaver_mask_serialize(ls.as_ptr(), lx.as_ptr(), s.len(), out.as_mut_ptr()) });
    out
}

/// `batch_open`'s local half from the received payloads, one per party (share/additive.rs:125-131).
pub fn open_sum_wire(payloads: &[Vec<u8>], n: usize) -> Vec<Fr> {
    let mut flat = Vec::with_capacity(payloads.len() * (8 + 32 * n));
    for p in payloads { assert_eq!(p.len(), 8 + 32 * n); flat.extend_from_slice(p); }
    let mut out = vec![0u64; 4 * n];
    check(unsafe { ffi::mpc_cuda_open_sum_deserialize(flat.as_ptr(), payloads.len() as u32, n, out.as_mut_ptr()) });
    frs_from(&out)
}

/// `(p / (x - z), p(z))` on local share values (share/additive.rs:154-162, kzg10/mod.rs:241-258).
pub fn poly_div_linear(coeffs: &[Fr], z: &Fr) -> (Vec<Fr>, Fr) {
    assert!(!coeffs.is_empty());
    let n = coeffs.len();
    let (lc, lz) = (fr_limbs(coeffs), fr_limbs(std::slice::from_ref(z)));
    let (mut q, mut rem) = (vec![0u64; 4 * (n - 1)], [0u64; 4]);
    check(unsafe { ffi::mpc_cuda_poly_div_linear(lc.as_ptr(), n, lz.as_ptr(), if n > 1 { q.as_mut_ptr() } else { std::ptr::null_mut() }, rem.as_mut_ptr()) });
    (frs_from(&q), fr_from(&rem))
}

/// `(p / (x^m - 1), p mod (x^m - 1))` on local share values: `divide_by_vanishing_poly` on a shared polynomial
/// (poly/src/polynomial/univariate/dense.rs:166-173 -> share/additive.rs:154-162; marlin/src/ahp/prover.rs:352,364,543).
pub fn poly_div_vanishing(coeffs: &[Fr], m: usize) -> (Vec<Fr>, Vec<Fr>) {
    assert!(!coeffs.is_empty() && m >= 1);
    let n = coeffs.len();
    let lc = fr_limbs(coeffs);
    let (mut q, mut rem) = (vec![0u64; 4 * n.saturating_sub(m)], vec![0u64; 4 * m]);
    check(unsafe { ffi::mpc_cuda_poly_div_vanishing(lc.as_ptr(), n, m, if n > m { q.as_mut_ptr() } else { std::ptr::null_mut() }, rem.as_mut_ptr()) });
    (frs_from(&q), frs_from(&rem))
}

/// `p (x^m - 1)` on local share values (`mul_by_vanishing_poly`, dense.rs:155-162; prover.rs:507).
pub fn poly_mul_vanishing(coeffs: &[Fr], m: usize) -> Vec<Fr> {
    assert!(!coeffs.is_empty() && m >= 1);
    let lc = fr_limbs(coeffs);
    let mut out = vec![0u64; 4 * (coeffs.len() + m)];
    check(unsafe { ffi::mpc_cuda_poly_mul_vanishing(lc.as_ptr(), coeffs.len(), m, out.as_mut_ptr()) });
    frs_from(&out)
}

/// `R1CStoQAP::witness_map` (src/groth16.rs:240-307) with everything resident between the calls: the two
/// `batch_open` rounds of the batch product exchange wire payloads only.
pub struct WitnessMap { state: u64, n: usize }
impl WitnessMap {
    /// a = A z, b = B z, c = C z from the party's assignment share, transforms and Beaver masks on the device
    pub fn begin(csr: [u64; 3], assignment: &[Fr], num_inputs: usize, log_n: u32, tx: &[Fr], ty: &[Fr]) -> Self {
        let n = 1usize << log_n;
        assert!(tx.len() == n && ty.len() == n && num_inputs <= assignment.len());
        let (la, lx, ly) = (fr_limbs(assignment), fr_limbs(tx), fr_limbs(ty));
        let mut state = 0u64;
        check(unsafe { ffi::mpc_cuda_witness_map_begin_r1cs(csr[0], csr[1], csr[2], la.as_ptr(), num_inputs, log_n, lx.as_ptr(), ly.as_ptr(), 0,
                                                           std::ptr::null_mut(), std::ptr::null_mut(), &mut state) });
        WitnessMap { state, n }
    }
    /// the payload this party broadcasts for open `which` (0: a' + x, 1: b' + y), `MpcSerNet::broadcast` format
    pub fn masked_payload(&self, which: u32) -> Vec<u8> {
        let mut out = vec![0u8; 8 + 32 * self.n];
        check(unsafe { ffi::mpc_cuda_witness_map_masked_payload(self.state, which, out.as_mut_ptr()) });
        out
    }
    /// every party's payload of open `which`, as received: summed on the device, kept in the state
    pub fn open_payloads(&self, which: u32, payloads: &[Vec<u8>]) {
        assert!(!payloads.is_empty() && payloads.iter().all(|p| p.len() == 8 + 32 * self.n));
        let ptrs: Vec<*const u8> = payloads.iter().map(|p| p.as_ptr()).collect();
        check(unsafe { ffi::mpc_cuda_witness_map_open_payloads(self.state, which, ptrs.as_ptr(), ptrs.len() as u32) });
    }
    /// Beaver combine, - c, / Z_H, coset iFFT; h stays on the device for `G1Bases::msm_resident`
    pub fn finish(&self, tz: &[Fr], is_leader: bool) -> *const u64 {
        assert!(tz.len() == self.n);
        let lz = fr_limbs(tz);
        let mut h: *mut u64 = std::ptr::null_mut();
        check(unsafe { ffi::mpc_cuda_witness_map_finish_dev(self.state, lz.as_ptr(), std::ptr::null(), std::ptr::null(), is_leader as u32, &mut h) });
        h as *const u64
    }
}
impl Drop for WitnessMap { fn drop(&mut self) { unsafe { ffi::mpc_cuda_witness_map_release(self.state); } } }
