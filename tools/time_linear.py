#!/usr/bin/env python3
"""Device-resident timings of the f2-f4 kernels (HBM-bound integer/byte work) with their algorithmic bytes:
usage: tools/time_linear.py [log_n]   -> one JSON line per kernel, GB/s against MEASURED_PEAKS.json's copy bandwidth"""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import __graft_entry__ as ge
pkg = ge.load_package(); H, S, L = pkg.host, pkg.synth, pkg._lib
H.init(); H.set_party(0, 1)
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << log_n
peak = None
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")))
except Exception:
    pass


def timed(fn, reps=10):
    fn(); L.call("mpc_cuda_stream_sync", None)
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    L.call("mpc_cuda_stream_sync", None)
    return (time.perf_counter() - t0) / reps * 1e3


def report(name, ms, nbytes, note=""):
    print(json.dumps({"kernel": name, "log_n": log_n, "ms": round(ms, 4), "algorithmic_GB": round(nbytes / 1e9, 3),
                      "GB_s": round(nbytes / ms / 1e6, 1), "note": note}), flush=True)


seed = S.fr_uniform(0x11, 1 << 20)
vec = [H.DeviceBuffer(n * 32) for _ in range(4)]
for b in vec:
    for off in range(0, n, 1 << 20):
        H.dev_upload(b.ptr.value + off * 32, seed[: min(1 << 20, n - off)])
L.call("mpc_cuda_stream_sync", None)
P = lambda b, k=0: b.ptr.value + 32 * k
plen = 8 + 32 * n
pay = H.DeviceBuffer(3 * plen)
flags = H.DeviceBuffer(16)
# f3: serialise (with the Beaver mask fused) and deserialise + open-sum of 3 payloads
ms = timed(lambda: L.call("mpc_cuda_beaver_mask_serialize_dev", H._dp(P(vec[0])), H._dp(P(vec[1])), C.c_size_t(n), H._dp(pay.ptr.value), None))
report("k_serialize<MASK>", ms, n * 96.0, "2 x 32 B in, 32 B out per element")
for p in (1, 2):
    H.dev_copy(pay.ptr.value + p * plen, pay.ptr.value, plen)
ms = timed(lambda: L.call("mpc_cuda_open_sum_deserialize_dev", H._dp(pay.ptr.value), C.c_uint32(3), C.c_size_t(n), H._dp(P(vec[2])),
                          H._dp(flags.ptr.value), None))
report("k_deserialize_sum (3 parties)", ms, n * 128.0, "3 x 32 B in, 32 B out per element")
# f4: division by x - z (blocked scan), by x^m - 1, product with x^m - 1
z = S.fr_uniform(0x12, 1)
ms = timed(lambda: L.call("mpc_cuda_poly_div_linear_dev", H._dp(P(vec[0])), C.c_size_t(n), H._p(z), H._dp(P(vec[2])), H._dp(P(vec[3])), None))
report("poly_div_linear (k_div_*)", ms, n * 96.0, "coefficients read twice (chunk pass + emit), quotient written")
for m in (1 << (log_n - 3), 16):
    ms = timed(lambda: H.dev_poly_div_vanishing(P(vec[0]), n, m, P(vec[2]), P(vec[3])))
    one_pass = (1 << 18) // m <= 1                              # a long divisor needs no run sums: one read, one write
    report("poly_div_vanishing m=2^%d (k_van_*)" % (m.bit_length() - 1), ms, n * (64.0 if one_pass else 96.0),
           "one read, quotient written" if one_pass else "read twice (run sums + emit), quotient written")
ms = timed(lambda: H.dev_poly_mul_vanishing(P(vec[0]), n - (1 << 10), 1 << 10, P(vec[2])))
report("poly_mul_vanishing (k_van_mul)", ms, n * 64.0, "every coefficient read (twice, the second time from L2) and written once")
ms = timed(lambda: H.dev_vec_op("axpy", P(vec[0]), P(vec[1]), z, P(vec[2]), n))
report("k_vec_op<AXPY> (LC accumulation)", ms, n * 96.0, "2 reads + 1 write, one product per element")
m_inv = min(n, 1 << 20)
ms = timed(lambda: H.dev_inverse(P(vec[0]), P(vec[2]), m_inv), reps=3)
report("k_inverse (2^%d elements)" % (m_inv.bit_length() - 1), ms, m_inv * 64.0, "batches of 8 behind one binary-Euclid inversion: bytes are not its limit")
# f2: R1CS-shaped CSR matrix x share vector
rows = min(n, 1 << 22)
mats = S.r1cs_matrices(0x13, rows, rows)
row_ptr, col, coeff = mats[0]
csr = H.CsrMatrix(row_ptr, col, coeff, rows)
nnz = int(row_ptr[-1])
ms = timed(lambda: H.dev_spmv(csr, P(vec[0]), P(vec[2])))
report("k_spmv_rows (2^%d rows, %d nnz)" % (rows.bit_length() - 1, nnz), ms, nnz * (4 + 32 + 32.0) + rows * (8 + 32.0),
       "per term: 4 B column + 32 B coefficient + 32 B gathered value; per row: 8 B pointer + 32 B out")
csr.release()
if peak:
    print(json.dumps({"measured_peaks": peak}))
