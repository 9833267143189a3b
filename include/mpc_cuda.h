/*
 * mpc_cuda.h — C ABI of libmpc_cuda.so, the B200 (sm_100a) implementation of zk-mpc's per-party
 * prover hot path: share MSM, share NTT and the local halves of Beaver multiplication.
 *
 * This is the boundary a thin Rust crate (`mpc-cuda`) binds with
 * `extern "C"`; each entry point names the reference interface it replaces (paths relative to the
 * zk-mpc repository).  Conventions:
 *   - every field element is Montgomery-form little-endian u64 limbs exactly as stored in
 *     arkworks' Fp256.0.0 / Fp384.0.0 (Fr = 4 limbs, Fq = 6 limbs, Fq2 = c0|c1 = 12 limbs);
 *   - G1 affine point = x|y (12 limbs) + one infinity byte; G2 affine = x.c0|x.c1|y.c0|y.c1
 *     (24 limbs) + one infinity byte; affine zero is (0, 1, infinity = 1) as in
 *     arkworks/algebra/ec/src/models/short_weierstrass_jacobian.rs:167-169;
 *   - pointers are HOST memory unless the function name ends in `_dev` (device pointers on the
 *     calling thread's current mpc_cuda device; `stream` is a cudaStream_t passed as void*,
 *     NULL = the library's per-thread stream); the caller owns every buffer;
 *   - return value 0 = OK; non-zero = error (mpc_cuda_last_error() describes it).  The reference
 *     panics on failure (assert!/unwrap), so the Rust shim turns non-zero into panic!;
 *   - host-pointer calls are synchronous; `_dev` calls are asynchronous on `stream`;
 *   - any function may be called concurrently from several host threads (one per party under
 *     mpc-net's LocalTestNet, mpc-net/src/multi.rs:436-441); party/device selection is per thread.
 * There is no CPU fallback: without a CUDA device every compute entry point returns an error.
 */
#ifndef MPC_CUDA_H
#define MPC_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPC_CUDA_OK 0
#define MPC_CUDA_ERR_CUDA 1       /* a CUDA runtime call failed */
#define MPC_CUDA_ERR_ARG 2        /* invalid argument */
#define MPC_CUDA_ERR_NO_DEVICE 3  /* no CUDA device / init not possible */
#define MPC_CUDA_ERR_HANDLE 4     /* unknown handle */
#define MPC_CUDA_ERR_MAC 5        /* an SPDZ MAC check on an opened value failed (spdz.rs:177-196 panics) */

/* ---- context ------------------------------------------------------------------------------- */
/* Select the devices the library may use (NULL/0 = all visible).  Thread-safe.  Idempotent for the same
 * list (or NULL/0); once initialised — explicitly or lazily by any other call — a DIFFERENT explicit list
 * returns MPC_CUDA_ERR_ARG: the per-device caches are keyed by list position. */
int32_t mpc_cuda_init(const int32_t* devices, int32_t n_dev);
/* Party identity of the calling thread: leader = party 0 (mpc-net/src/lib.rs:49-51); the thread's
 * default device becomes devices[party_id % n_dev]. */
int32_t mpc_cuda_set_party(uint32_t party_id, uint32_t n_parties);
/* Explicit device for the calling thread (index into the init list). */
int32_t mpc_cuda_set_device(int32_t dev_index);
int32_t mpc_cuda_device_count(void);
/* Tuning knobs for benchmarks and tests (process-wide atomics; 0 restores the automatic choice):
 *   "msm_window_bits"  Pippenger window width c (3..23)
 *   "msm_task_len"     maximum points one accumulation task adds (bucket splitting)
 *   "msm_affine"       batched-affine pre-reduction of the bucket lists before the XYZZ accumulation: 0 = automatic
 *                      (two rounds when buckets hold >= 8 entries on average), 1 = off, 2 = one round, 3 = two rounds
 *   "msm_affine_split" the pre-reduction as two kernels pipelined over two streams (the memory-bound product pass of
 *                      segment j+1 under the multiplier-bound addition pass of segment j): 0 or 2 = one fused
 *                      kernel (default: the split measured 7% slower at 2^24), 1 = split
 *   "msm_reduce_chunk" buckets per thread of the bucket reduction (a power of two; 0 = automatic)
 *   "msm_reduce_warp_max" the bucket reduction runs one WARP per chunk (warp-cooperative group law: shorter serial
 *                      chains, 32x the threads) while it has at most this many chunks; 0 = automatic
 *   "msm_host_chunks"  point-range chunks a host-buffer MSM is streamed in (copy/compute overlap), 1..16; the chunk
 *                      sizes double towards the middle and halve again (1 2 4 4 2 1), so that neither the first copy
 *                      nor the last chunk's kernels are long
 *   "ntt_occupancy"    NTT pass kernels built for one more resident CTA per SM (64 registers): 0 = automatic (the
 *                      256-row tile shape only), 1 = every shape, 2 = none
 *   "ntt_graph"        1: replay repeated sharded transforms (same blocks, size, kind) from a CUDA graph captured
 *                      across the devices' streams (off by default: measured no faster than direct launches)
 *   "ntt_generic"      1: run every NTT pass through the generic (runtime tile shape) kernel instead of the
 *                      compile-time-shaped ones (A/B measurements; results are identical)
 *   "profile"          1: bracket pipeline stages with CUDA events on the launching stream */
int32_t mpc_cuda_set_option(const char* name, int64_t value);
/* Number of kernels this library has launched so far (all threads). */
uint64_t mpc_cuda_launch_count(void);
/* Device time accumulated by the calling thread for a stage since the last read, and the number of
 * intervals; waits for the pending events.  Stages: "msm_total", "msm_sort", "msm_accumulate",
 * "msm_reduce", "msm_precompute", "ntt". */
int32_t mpc_cuda_profile_read(const char* name, double* ms_total, uint64_t* count);
const char* mpc_cuda_last_error(void);
const char* mpc_cuda_version(void);

/* device memory helpers for resident pipelines (the fused witness-map path, benchmarks) */
int32_t mpc_cuda_malloc(void** dptr, size_t bytes);
int32_t mpc_cuda_free(void* dptr);
/* page-locked host memory for buffers that cross PCIe on every proof (wire payloads, triple shares): copies from
 * pageable memory are staged by the driver at a fraction of the link rate */
int32_t mpc_cuda_host_alloc(void** hptr, size_t bytes);
int32_t mpc_cuda_host_free(void* hptr);
int32_t mpc_cuda_memcpy_h2d(void* dptr, const void* hptr, size_t bytes, void* stream);
int32_t mpc_cuda_memcpy_d2h(void* hptr, const void* dptr, size_t bytes, void* stream);
int32_t mpc_cuda_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream);
int32_t mpc_cuda_memset_zero_dev(void* dptr, size_t bytes, void* stream);
/* strided device-to-device copy (height rows of width bytes, pitches in bytes): e.g. the witness values into the
 * non-input positions of Marlin's domain H (arkworks/marlin/src/ahp/prover.rs:340-349) */
int32_t mpc_cuda_memcpy2d_d2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height,
                              void* stream);
/* extra streams on the calling thread's device, so independent `_dev` calls (the five MSMs of one proof,
 * src/groth16.rs:106-160) overlap; NULL everywhere else means the library's own per-thread stream */
int32_t mpc_cuda_stream_create(void** stream);
int32_t mpc_cuda_stream_destroy(void* stream);
int32_t mpc_cuda_stream_sync(void* stream);

/* ---- Beaver multiplication, local halves ----------------------------------------------------
 * FieldShare::batch_mul (mpc-algebra/src/share/field.rs:97-129) =
 *   mask (local) -> batch_open (network) -> mask (local) -> batch_open -> combine (local). */

/* out[i] = s[i] + x[i]   (share/field.rs:108-117, FieldShare::add additive.rs:132-135 /
 * spdz.rs:197-201).  SPDZ shares: pass the sh plane and the mac plane as one 2n-element call. */
int32_t mpc_cuda_beaver_mask(const uint64_t* s, const uint64_t* x, uint64_t* out, size_t n);
int32_t mpc_cuda_beaver_mask_dev(const uint64_t* s, const uint64_t* x, uint64_t* out, size_t n, void* stream);

/* out = z - y*sx - x*oy (+ sx*oy on the leader)   (share/field.rs:118-128; scale/sub/shift of
 * additive.rs:136-152, spdz.rs:202-219).  x, y, z are the party's triple shares, sx/oy the opened
 * masked values.  spdz = 0: x,y,z,out hold n elements.  spdz = 1: x,y,z,out hold 2n elements laid
 * out as [sh plane | mac plane]; sx/oy always n; the mac plane receives mac_share*sx*oy with
 * mac_share = is_leader ? 1 : 0 (spdz.rs:31-38). */
int32_t mpc_cuda_beaver_combine(const uint64_t* x, const uint64_t* y, const uint64_t* z,
                                const uint64_t* sx_pub, const uint64_t* oy_pub, uint64_t* out,
                                size_t n, uint32_t is_leader, uint32_t spdz);
int32_t mpc_cuda_beaver_combine_dev(const uint64_t* x, const uint64_t* y, const uint64_t* z,
                                    const uint64_t* sx_pub, const uint64_t* oy_pub, uint64_t* out,
                                    size_t n, uint32_t is_leader, uint32_t spdz, void* stream);

/* out[i] = sum_p parts[p*n + i]: the local half of batch_open after the broadcast
 * (share/additive.rs:125-131, spdz.rs:181-184). */
int32_t mpc_cuda_open_sum(const uint64_t* parts, uint32_t n_parties, uint64_t* out, size_t n);
int32_t mpc_cuda_open_sum_dev(const uint64_t* parts, uint32_t n_parties, uint64_t* out, size_t n, void* stream);

/* SPDZ MAC check, local half: out[i] = mac_share*vals[i] - macs[i]   (share/spdz.rs:185-189) */
int32_t mpc_cuda_spdz_mac_check(const uint64_t* vals, const uint64_t* macs, uint64_t* out, size_t n,
                                uint32_t is_leader);
int32_t mpc_cuda_spdz_mac_check_dev(const uint64_t* vals, const uint64_t* macs, uint64_t* out, size_t n,
                                    uint32_t is_leader, void* stream);

/* Elementwise helpers used around the NTTs (src/groth16.rs:298-302, poly/src/domain/mod.rs:183-190,
 * poly/src/polynomial/univariate/dense.rs:345-372):
 *   MPC_CUDA_VEC_SUB        out = a - b
 *   MPC_CUDA_VEC_MUL        out = a * b          (public x share / batch product of publics)
 *   MPC_CUDA_VEC_MUL_CONST  out = a * c[0]
 *   MPC_CUDA_VEC_AXPY       out = a + c[0] * b */
#define MPC_CUDA_VEC_SUB 0
#define MPC_CUDA_VEC_MUL 1
#define MPC_CUDA_VEC_MUL_CONST 2
#define MPC_CUDA_VEC_AXPY 3
/* `out` may alias `a` (in place); it must not alias `b`. */
int32_t mpc_cuda_vec_op(uint32_t op, const uint64_t* a, const uint64_t* b, const uint64_t* c,
                        uint64_t* out, size_t n);
int32_t mpc_cuda_vec_op_dev(uint32_t op, const uint64_t* a, const uint64_t* b, const uint64_t* c_host,
                            uint64_t* out, size_t n, void* stream);
/* out[i] = 1 / a[i] for public values (0 stays 0, as ark_ff::batch_inversion ff/src/fields/mod.rs:597-660 leaves it):
 * the denominators alpha - h of Marlin's r(alpha, .) (arkworks/marlin/src/ahp/mod.rs:357-364).  out may alias a. */
int32_t mpc_cuda_fr_inverse_dev(const uint64_t* a, uint64_t* out, size_t n, void* stream);

/* ---- share NTT ------------------------------------------------------------------------------
 * Radix2EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place over Fr
 * (arkworks/algebra/poly/src/domain/radix2/mod.rs:99-114, radix2/fft.rs:22-35,
 * domain/mod.rs:138-141).  In-order input, in-order output; `data` holds `batch` vectors of
 * 2^log_n elements back to back and is transformed in place.  With MpcField coefficients the
 * transform of a party's local values IS its output share vector. */
#define MPC_CUDA_NTT_FFT 0
#define MPC_CUDA_NTT_IFFT 1
#define MPC_CUDA_NTT_COSET_FFT 2
#define MPC_CUDA_NTT_COSET_IFFT 3
int32_t mpc_cuda_ntt_fr(uint64_t* data, uint32_t log_n, uint32_t kind, uint32_t batch);
int32_t mpc_cuda_ntt_fr_dev(uint64_t* data, uint32_t log_n, uint32_t kind, uint32_t batch, void* stream);
/* Multi-GPU share NTT (one party's vector block-distributed over g = 2^log_g devices): the
 * log_g stages that cross device boundaries.  After an all-to-all the caller holds data[q][l] for q < g and
 * its slice l in [slice_offset, slice_offset + slice_len) of the n/g local offsets; this runs those stages
 * (plus the coset / 1/g scaling of `kind`) in place.  Forward: all-to-all, cross stage, all-to-all back,
 * local mpc_cuda_ntt_fr_dev(kind = FFT, log_n - log_g); device r, local m then holds X[m*g + bitrev(r)].
 * Inverse: local IFFT, all-to-all, cross stage, all-to-all back -> natural block order.
 * zk-mpc_b200/sharding.py drives it over torch.distributed (NCCL). */
int32_t mpc_cuda_ntt_cross_stage_dev(uint64_t* data, uint32_t log_n, uint32_t log_g, size_t slice_offset,
                                     size_t slice_len, uint32_t kind, void* stream);
/* The same transform inside ONE process over g = 2^log_g devices of the init list: blocks[q] is a device
 * pointer on device dev_index[q] (NULL = q) to block q (n/g elements).  The cross-device stages read and write
 * the peers' blocks directly over NVLink inside the butterfly kernel (no separate all-to-all), then every
 * device runs its local transform.  Layouts as above: forward kinds take natural block order and leave device r,
 * local m = X[m*g + bitrev(r)]; inverse kinds take that order and return natural block order.  Asynchronous:
 * the work is ordered on the calling thread's streams of the devices; when it returns, the stream of
 * dev_index[0] is ordered after the whole transform (mpc_cuda_set_device + mpc_cuda_stream_sync(NULL) waits).
 * All dev_index equal = g virtual devices on one GPU (used by the single-GPU parity tests). */
int32_t mpc_cuda_ntt_fr_sharded_dev(uint64_t* const* blocks, const int32_t* dev_index, uint32_t log_n, uint32_t log_g,
                                    uint32_t kind);
/* natural block order <-> transposed order across the devices (out of place; to_transposed = 1: natural in) */
int32_t mpc_cuda_ntt_reorder_sharded_dev(uint64_t* const* in, uint64_t* const* out, const int32_t* dev_index,
                                         uint32_t log_n, uint32_t log_g, uint32_t to_transposed);
/* Host vector, in-order in and out exactly like mpc_cuda_ntt_fr (batch 1), computed on 2^log_g devices: block q
 * crosses its own PCIe link, synchronous. */
int32_t mpc_cuda_ntt_fr_sharded(uint64_t* data, uint32_t log_n, uint32_t kind, uint32_t log_g);
/* evals[i] *= (g^n - 1)^-1, g = 22  (EvaluationDomain::divide_by_vanishing_poly_on_coset_in_place) */
int32_t mpc_cuda_divide_by_vanishing_on_coset(uint64_t* data, uint32_t log_n);
int32_t mpc_cuda_divide_by_vanishing_on_coset_dev(uint64_t* data, uint32_t log_n, void* stream);

/* ---- fused witness map ---------------------------------------------------------------------
 * R1CStoQAP::witness_map on one party's local values (src/groth16.rs:240-307).  The vectors stay in HBM between
 * the calls; only the masked values cross PCIe for the two opens, which stay on mpc-net.
 * begin: a, b, c = the party's evaluations of the A, B, C polynomials over the domain (2^log_n elements each),
 * tx, ty = its Beaver triple shares; computes a' = coset_fft(ifft(a)) (same for b, c) and returns
 * masked_a = a' + tx, masked_b = b' + ty.  finish: tz = triple share, sx / oy = the opened sums of masked_a /
 * masked_b over the parties; returns this party's share of h = coset_ifft((a'*b' - c') / Z_H) and releases the
 * state.
 * _ex, spdz = 1: SPDZ shares (mpc-algebra/src/share/spdz.rs:50-53,197-219): a, b, c, tx, ty, tz, masked_*, h
 * hold two planes [sh | mac] of 2^log_n elements each; sx / oy stay one plane (opened values).
 * _begin_r1cs: starts one step earlier, at the assignment: a = A z, b = B z, c = C z for the public CSR matrices
 * registered with mpc_cuda_csr_register (evaluate_constraint, src/groth16.rs:205-234,263-270,289-293), with
 * a[num_constraints .. num_constraints + num_inputs) = z[0 .. num_inputs) (:272-276) and zero padding to the
 * domain; assignment = instance | witness values, `planes` x cols.
 * _finish_dev: like finish, but h stays on the device (*h_dev, planes x n, valid until _release) so that it feeds
 * the h_query MSM (src/groth16.rs:106) through mpc_cuda_msm_g1_handle_scalars_dev without crossing PCIe.
 *
 * The two opens at the wire level, so that nothing but wire bytes crosses PCIe between begin and finish: pass
 * masked_a = masked_b = NULL to begin, then for which = 0 (masked_a) and 1 (masked_b)
 *   _masked_payload   writes the party's MpcSerNet::broadcast payload of the masked vector (8 + 32 * 2^log_n bytes,
 *                     mpc-algebra/src/channel.rs:12-28; SPDZ: the sh plane, spdz.rs:78-84) into host memory;
 *   _open_payloads    takes every party's received payload (n_parties host pointers, party order irrelevant), sums
 *                     them on the device (share/additive.rs:125-131) and keeps the opened vector in the state; finish
 *                     then takes sx = NULL / oy = NULL for an open that arrived this way;
 *   _mac_payload      SPDZ only, after _open_payloads: the local half dx = mac_share * val - mac of batch_open's MAC
 *                     check (spdz.rs:177-196) as a payload; _mac_verify sums the parties' dx payloads and returns
 *                     MPC_CUDA_ERR_MAC unless every element is zero.
 * _release returns the state's buffer to the device's stream-ordered pool on the calling thread's stream: work the
 * caller put on OTHER streams that still reads h or the assignment must have been synchronised first.
 * _assignment_dev: the assignment _begin_r1cs uploaded (planes x cols, valid until _release), so that the scalar
 * vectors of the a_query / b_query / l_query MSMs (src/groth16.rs:137-160) are device-to-device copies of it. */
int32_t mpc_cuda_witness_map_begin(const uint64_t* a, const uint64_t* b, const uint64_t* c, uint32_t log_n,
                                   const uint64_t* tx, const uint64_t* ty, uint64_t* masked_a, uint64_t* masked_b,
                                   uint64_t* state);
int32_t mpc_cuda_witness_map_begin_ex(const uint64_t* a, const uint64_t* b, const uint64_t* c, uint32_t log_n,
                                      const uint64_t* tx, const uint64_t* ty, uint32_t spdz, uint64_t* masked_a,
                                      uint64_t* masked_b, uint64_t* state);
int32_t mpc_cuda_witness_map_begin_r1cs(uint64_t csr_a, uint64_t csr_b, uint64_t csr_c, const uint64_t* assignment,
                                        size_t num_inputs, uint32_t log_n, const uint64_t* tx, const uint64_t* ty,
                                        uint32_t spdz, uint64_t* masked_a, uint64_t* masked_b, uint64_t* state);
int32_t mpc_cuda_witness_map_finish(uint64_t state, const uint64_t* tz, const uint64_t* sx, const uint64_t* oy,
                                    uint32_t is_leader, uint64_t* h_out);
int32_t mpc_cuda_witness_map_finish_dev(uint64_t state, const uint64_t* tz, const uint64_t* sx, const uint64_t* oy,
                                        uint32_t is_leader, uint64_t** h_dev);
int32_t mpc_cuda_witness_map_masked_payload(uint64_t state, uint32_t which, uint8_t* out);
int32_t mpc_cuda_witness_map_open_payloads(uint64_t state, uint32_t which, const uint8_t* const* payloads,
                                           uint32_t n_parties);
int32_t mpc_cuda_witness_map_mac_payload(uint64_t state, uint32_t which, uint32_t is_leader, uint8_t* out);
int32_t mpc_cuda_witness_map_mac_verify(uint64_t state, const uint8_t* const* payloads, uint32_t n_parties);
int32_t mpc_cuda_witness_map_assignment_dev(uint64_t state, uint64_t** z_dev, size_t* cols);
int32_t mpc_cuda_witness_map_release(uint64_t state);

/* ---- linear steps either side of the path -----------------------------------
 * f2: public sparse matrix x share vector = evaluate_constraint for every row (src/groth16.rs:205-234; Marlin:
 * arkworks/marlin/src/ahp/prover.rs:258-278).  CSR: row_ptr[rows + 1] (u64), col[nnz] (u32 < cols), coeff[nnz]
 * Montgomery Fr.  x holds `planes` vectors of `cols` elements (SPDZ: sh and mac planes), out `planes` x rows. */
int32_t mpc_cuda_csr_register(const uint64_t* row_ptr, const uint32_t* col, const uint64_t* coeff_mont, size_t rows,
                              size_t cols, uint64_t* handle);
int32_t mpc_cuda_csr_release(uint64_t handle);
int32_t mpc_cuda_csr_dims(uint64_t handle, size_t* rows, size_t* cols, size_t* nnz);
int32_t mpc_cuda_csr_spmv(uint64_t handle, const uint64_t* x, uint32_t planes, uint64_t* out);
int32_t mpc_cuda_csr_spmv_dev(uint64_t handle, const uint64_t* x_dev, size_t x_stride, uint32_t planes, uint64_t* out_dev,
                              size_t out_stride, void* stream);
/* f3: the bytes MpcSerNet::broadcast puts on the wire for a Vec<Fr> (mpc-algebra/src/channel.rs:12-28 ->
 * arkworks/algebra/serialize/src/lib.rs:263-272, ff/src/fields/macros.rs:1-110): u64 LE length, then the 32 LE
 * bytes of the canonical integer of every element; `out` / every payload holds 8 + 32 n bytes (8-byte aligned
 * for the _dev entries).  _mask_serialize fuses the Beaver mask s + x (share/field.rs:108-117) with the
 * conversion; _open_sum_deserialize reads the `n_parties` payloads received for a batch_open
 * (share/additive.rs:125-131), back to back, and returns their sum in Montgomery form.  A length prefix other
 * than n or an element >= r is an error (arkworks: SerializationError::InvalidData); the _dev entry reports it
 * through flags_dev[0] (~0 = all valid, else 1 + index of the first bad element) and flags_dev[1] (1 = a bad
 * length prefix). */
int32_t mpc_cuda_fr_serialize(const uint64_t* vals_mont, size_t n, uint8_t* out);
int32_t mpc_cuda_fr_deserialize(const uint8_t* in, size_t n, uint64_t* vals_mont);
int32_t mpc_cuda_beaver_mask_serialize(const uint64_t* s, const uint64_t* x, size_t n, uint8_t* out);
int32_t mpc_cuda_beaver_mask_serialize_dev(const uint64_t* s, const uint64_t* x /* NULL: no mask */, size_t n, uint8_t* out,
                                           void* stream);
int32_t mpc_cuda_open_sum_deserialize(const uint8_t* payloads, uint32_t n_parties, size_t n, uint64_t* out);
int32_t mpc_cuda_open_sum_deserialize_dev(const uint8_t* payloads, uint32_t n_parties, size_t n, uint64_t* out,
                                          uint64_t* flags_dev /*2*/, void* stream);
/* f4: p(x) = q(x) (x - z) + rem for a public point z, on the local share values of p's n coefficients (low degree
 * first): univariate_div_qr (mpc-algebra/src/wire/field.rs:1007-1065 -> share/additive.rs:154-162 ->
 * arkworks/algebra/poly/src/polynomial/univariate/mod.rs:133-172) as KZG10::open divides by (x - point)
 * (arkworks/poly-commit/src/kzg10/mod.rs:241-258).  q_out receives n - 1 coefficients (NULL: evaluation only),
 * rem_out one element = p(z) (DensePolynomial::evaluate, univariate/dense.rs:53-75).  The reference truncates
 * leading zero coefficients of q and drops a zero remainder; this returns the fixed-size arrays. */
int32_t mpc_cuda_poly_div_linear(const uint64_t* coeffs, size_t n, const uint64_t* z_mont, uint64_t* q_out,
                                 uint64_t* rem_out);
int32_t mpc_cuda_poly_div_linear_dev(const uint64_t* coeffs, size_t n, const uint64_t* z_mont_host, uint64_t* q_out,
                                     uint64_t* rem_out, void* stream);
/* f4: p(x) = q(x) (x^m - 1) + rem on the local share values of p's n coefficients, and p(x) (x^m - 1): Marlin's
 * divide_by_vanishing_poly / mul_by_vanishing_poly on shared polynomials (arkworks/algebra/poly/src/polynomial/
 * univariate/dense.rs:155-173 -> univariate/mod.rs:133-143 -> mpc-algebra/src/share/additive.rs:154-162), as
 * AHPForR1CS::prover_first_round / prover_second_round call them (arkworks/marlin/src/ahp/prover.rs:352,364,507,543).
 * q_out receives n - m coefficients (nothing when n <= m; NULL: remainder only), rem_out m coefficients, out n + m.
 * Fixed-size arrays as above (the reference truncates leading zeros). */
int32_t mpc_cuda_poly_div_vanishing(const uint64_t* coeffs, size_t n, size_t m, uint64_t* q_out, uint64_t* rem_out);
int32_t mpc_cuda_poly_div_vanishing_dev(const uint64_t* coeffs, size_t n, size_t m, uint64_t* q_out, uint64_t* rem_out,
                                        void* stream);
int32_t mpc_cuda_poly_mul_vanishing(const uint64_t* coeffs, size_t n, size_t m, uint64_t* out);
int32_t mpc_cuda_poly_mul_vanishing_dev(const uint64_t* coeffs, size_t n, size_t m, uint64_t* out, void* stream);

/* ---- share MSM ------------------------------------------------------------------------------
 * Msm::msm / AffineMsm::msm (mpc-algebra/src/share/msm.rs:6-9,33-37) =
 * AffineCurve::multi_scalar_mul (arkworks/algebra/ec/src/lib.rs:305-314: into_repr each scalar) +
 * VariableBaseMSM::multi_scalar_mul (ec/src/msm/variable_base.rs:12-106) + into affine.
 * n = min(bases, scalars) is decided by the caller (variable_base.rs:16-18).  `inf` may be NULL
 * (no infinity bases).  The result is the affine point in Montgomery limbs + infinity byte. */
int32_t mpc_cuda_msm_g1(const uint64_t* bases_xy, const uint8_t* inf, const uint64_t* scalars_mont,
                        size_t n, uint64_t out_xy[12], uint8_t* out_inf);
int32_t mpc_cuda_msm_g2(const uint64_t* bases_xy, const uint8_t* inf, const uint64_t* scalars_mont,
                        size_t n, uint64_t out_xy[24], uint8_t* out_inf);

/* Keep a CRS vector resident (pk.{a,b_g1,h,l}_query of src/groth16.rs:106-160, powers_of_g of
 * arkworks/poly-commit/src/kzg10/mod.rs:166-170): only scalars cross PCIe afterwards. */
int32_t mpc_cuda_msm_g1_register_bases(const uint64_t* bases_xy, const uint8_t* inf, size_t n, uint64_t* handle);
int32_t mpc_cuda_msm_g1_register_bases_dev(const uint64_t* bases_xy_dev, size_t n, uint64_t* handle);
int32_t mpc_cuda_msm_g2_register_bases(const uint64_t* bases_xy, const uint8_t* inf, size_t n, uint64_t* handle);
int32_t mpc_cuda_msm_g2_register_bases_dev(const uint64_t* bases_xy_dev, size_t n, uint64_t* handle);
/* Multi-GPU inside ONE process: the vector is cut into `parts` point ranges, range k resident on
 * device (k mod n_dev) of the init list.  mpc_cuda_msm_g{1,2}_handle / _precompute on such a handle run every part on its
 * own device concurrently (the caller still sees one msm call, mpc-algebra/src/share/msm.rs:6-9); the Jacobian
 * partials are gathered over NVLink peer copies and added on the calling thread's device. */
int32_t mpc_cuda_msm_g1_register_bases_sharded(const uint64_t* bases_xy, const uint8_t* inf, size_t n, uint32_t parts,
                                               uint64_t* handle);
int32_t mpc_cuda_msm_g2_register_bases_sharded(const uint64_t* bases_xy, const uint8_t* inf, size_t n, uint32_t parts,
                                               uint64_t* handle);
/* Drops the registry's reference; the memory is freed once no concurrent call still uses the vector. */
int32_t mpc_cuda_msm_release_bases(uint64_t handle);
/* Optional, once per registered CRS: build the table 2^(c*w) * P_i for every window w (c = window_bits,
 * 0 = automatic; nwin x the size of the vector in HBM).  Later MSMs over the handle use one shared bucket
 * set with c-bit windows: fewer mixed additions, one bucket reduction, no window doublings.  Results are
 * the same group elements. */
int32_t mpc_cuda_msm_g1_precompute(uint64_t handle, uint32_t window_bits);
int32_t mpc_cuda_msm_g2_precompute(uint64_t handle, uint32_t window_bits);
/* MSM over bases[offset .. offset+n) of a registered vector; scalars on the host */
int32_t mpc_cuda_msm_g1_handle(uint64_t handle, size_t offset, const uint64_t* scalars_mont, size_t n,
                               uint64_t out_xy[12], uint8_t* out_inf);
int32_t mpc_cuda_msm_g2_handle(uint64_t handle, size_t offset, const uint64_t* scalars_mont, size_t n,
                               uint64_t out_xy[24], uint8_t* out_inf);
/* scalars already on the device; writes the Jacobian partial (x,y,z = 18 limbs, z = 0 for infinity)
 * to device memory without normalising, so per-GPU partials can be gathered and added
 * (multi-GPU point-range sharding) */
int32_t mpc_cuda_msm_g1_handle_dev(uint64_t handle, size_t offset, const uint64_t* scalars_mont_dev, size_t n,
                                   uint64_t* out_jac_dev /*18*/, void* stream);
int32_t mpc_cuda_msm_g2_handle_dev(uint64_t handle, size_t offset, const uint64_t* scalars_mont_dev, size_t n,
                                   uint64_t* out_jac_dev /*36*/, void* stream);
/* affine(sum of `count` Jacobian partials); partials on the device, result on the host */
int32_t mpc_cuda_g1_sum_partials_dev(const uint64_t* jac_dev /*count*18*/, uint32_t count,
                                     uint64_t out_xy[12], uint8_t* out_inf, void* stream);
int32_t mpc_cuda_g2_sum_partials_dev(const uint64_t* jac_dev /*count*36*/, uint32_t count,
                                     uint64_t out_xy[24], uint8_t* out_inf, void* stream);
/* scalars already on the device, affine result on the host (e.g. h left resident by mpc_cuda_witness_map_finish_dev) */
int32_t mpc_cuda_msm_g1_handle_scalars_dev(uint64_t handle, size_t offset, const uint64_t* scalars_mont_dev, size_t n,
                                           uint64_t out_xy[12], uint8_t* out_inf);
int32_t mpc_cuda_msm_g2_handle_scalars_dev(uint64_t handle, size_t offset, const uint64_t* scalars_mont_dev, size_t n,
                                           uint64_t out_xy[24], uint8_t* out_inf);
/* whole-vector MSM over a sharded handle with the scalars already resident: scalars_mont_dev[k] is a pointer on
 * device k to the scalars of part k's point range (n/parts (+1) elements, same split as registration) */
int32_t mpc_cuda_msm_g1_handle_sharded_dev(uint64_t handle, const uint64_t* const* scalars_mont_dev, uint32_t parts,
                                           uint64_t out_xy[12], uint8_t* out_inf);

/* Synthetic CRS for benchmarks and tests: bases[i] = k_i * G1 generator with
 * k_i = max(1, mix64(seed + (first+i+1)*0x9E3779B97F4A7C15)), written as affine x|y to device memory. */
int32_t mpc_cuda_g1_generate_dev(uint64_t seed, size_t first, size_t n, uint64_t* out_xy_dev, void* stream);
int32_t mpc_cuda_g2_generate_dev(uint64_t seed, size_t first, size_t n, uint64_t* out_xy_dev, void* stream);

/* ---- diagnostics ----------------------------------------------------------------------------
 * Raw field kernels (one element per thread) used by the parity tests to pin the device
 * arithmetic itself: field 0 = Fr (4 limbs), 1 = Fq (6 limbs); op 0 add, 1 sub, 2 mul, 3 neg,
 * 4 inverse (0 -> 0), 5 Montgomery->canonical, 6 canonical->Montgomery, 7 square, 8 product through the
 * 32-bit-column formulation, 9 inverse through the binary-Euclid routine of the result-emission tail. */
int32_t mpc_cuda_field_op(uint32_t field, uint32_t op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n);
/* Integer-pipe microbenchmark: returns achieved giga-ops/s of `iters` dependent instructions per
 * thread over a full-chip grid.  kind 0 = IMAD.U32 (32-bit), 1 = IMAD.WIDE.U32 with carry chain,
 * 2 = Fq Montgomery products (result in products/s), 3 = Fr Montgomery products, 4 / 5 = the same through the
 * 32-bit-column formulation, 6 = FP64 DFMA. */
int32_t mpc_cuda_microbench(uint32_t kind, uint32_t iters, double* gops);

#ifdef __cplusplus
}
#endif
#endif /* MPC_CUDA_H */
