"""ctypes/numpy front-end of the CPU oracle (oracle/zkmpc_oracle.c).

ORACLE — TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; the shipped package (zk-mpc_b200/) never
imports this module.

All field elements are Montgomery-form little-endian u64 limbs exactly as arkworks'
`Fp256.0.0` / `Fp384.0.0`: Fr = (n,4) uint64, Fq = (n,6), Fq2 = (n,12),
G1 affine = (n,12) x|y + (n,) uint8 infinity flags, G2 affine = (n,24) x.c0|x.c1|y.c0|y.c1.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libzkmpc_oracle.so")

u64p = C.POINTER(C.c_uint64)
u8p = C.POINTER(C.c_uint8)


def build(force=False):
    """Compile the oracle with the committed Makefile (gcc only, no external deps)."""
    srcs = [os.path.join(_HERE, f) for f in ("zkmpc_oracle.c", "field_tmpl.h", "curve_tmpl.h", "Makefile")]
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(s) for s in srcs)):
        return _LIB_PATH
    r = subprocess.run(["make", "-C", _HERE, "-B"], capture_output=True, text=True)
    if r.returncode != 0:
        # toolchains without libgomp: same code, windows processed serially
        r = subprocess.run(["make", "-C", _HERE, "-B", "OMP="], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_domain_params.restype = C.c_int
        _lib.orc_ntt_fr.restype = C.c_int
        _lib.orc_g1_on_curve.restype = C.c_int
        _lib.orc_g2_on_curve.restype = C.c_int
    return _lib


def _a(x, cols=None):
    x = np.ascontiguousarray(x, dtype=np.uint64)
    if cols is not None:
        assert x.shape[-1] == cols, (x.shape, cols)
    return x


def _p(x):
    return x.ctypes.data_as(u64p)


def _p8(x):
    return x.ctypes.data_as(u8p)


def _inf(inf, n):
    if inf is None:
        return np.zeros(n, dtype=np.uint8)
    return np.ascontiguousarray(inf, dtype=np.uint8)


# ----------------------------------------------------------------------------- fields
_OPS = {"add": 0, "sub": 1, "mul": 2, "neg": 3, "inv": 4, "from_mont": 5, "to_mont": 6, "sqr": 7}


def _vec(fn, limbs, op, a, b=None):
    a = _a(a, limbs)
    out = np.empty_like(a)
    if b is not None:
        b = _a(b, limbs)
        assert b.shape == a.shape
    n = a.size // limbs
    fn(C.c_int(_OPS[op]), _p(a), _p(b) if b is not None else None, _p(out), C.c_size_t(n))
    return out


def fr(op, a, b=None):
    return _vec(lib().orc_fr_vec, 4, op, a, b)


def fq(op, a, b=None):
    return _vec(lib().orc_fq_vec, 6, op, a, b)


def fq2(op, a, b=None):
    return _vec(lib().orc_fq2_vec, 12, op, a, b)


def fr_batch_inv(a):
    a = _a(a, 4).copy()
    lib().orc_fr_batch_inv(_p(a), C.c_size_t(a.size // 4))
    return a


def constants():
    names = ["fr_mod", "fr_r", "fr_r2", "fr_gen", "fr_root", "fq_mod", "fq_r", "fq_r2"]
    bufs = [np.zeros(4 if k.startswith("fr") else 6, dtype=np.uint64) for k in names]
    lib().orc_constants(*[_p(b) for b in bufs])
    return dict(zip(names, bufs))


# ----------------------------------------------------------------------------- G1
def g1_generator():
    out = np.zeros(12, dtype=np.uint64)
    lib().orc_g1_generator(_p(out))
    return out


def g1_on_curve(xy, inf=0):
    xy = _a(xy, 12)
    return bool(lib().orc_g1_on_curve(_p(xy), C.c_uint8(int(inf))))


def g1_scalar_mul(xy, k_limbs, inf=0):
    xy = _a(xy, 12)
    k = _a(k_limbs)
    out = np.zeros(12, dtype=np.uint64)
    oinf = C.c_uint8(0)
    lib().orc_g1_scalar_mul(_p(xy), C.c_uint8(int(inf)), _p(k), C.c_int(k.size), _p(out), C.byref(oinf))
    return out, oinf.value


def g1_add(a_xy, b_xy, a_inf=0, b_inf=0):
    a_xy, b_xy = _a(a_xy, 12), _a(b_xy, 12)
    out = np.zeros(12, dtype=np.uint64)
    oinf = C.c_uint8(0)
    lib().orc_g1_add(_p(a_xy), C.c_uint8(int(a_inf)), _p(b_xy), C.c_uint8(int(b_inf)), _p(out), C.byref(oinf))
    return out, oinf.value


def g1_sum_jac(pts):
    pts = _a(pts, 18)
    out = np.zeros(12, dtype=np.uint64)
    oinf = C.c_uint8(0)
    lib().orc_g1_sum_jac(_p(pts), C.c_size_t(pts.size // 18), _p(out), C.byref(oinf))
    return out, oinf.value


def g1_generate(seed, n, first=0):
    out = np.zeros((n, 12), dtype=np.uint64)
    lib().orc_g1_generate(C.c_uint64(seed), C.c_size_t(first), C.c_size_t(n), _p(out))
    return out


def g1_msm(bases_xy, scalars_mont, inf=None, threads=1):
    bases_xy, scalars_mont = _a(bases_xy, 12), _a(scalars_mont, 4)
    n = min(bases_xy.size // 12, scalars_mont.size // 4)      # variable_base.rs:16-18
    inf = _inf(inf, bases_xy.size // 12)
    out = np.zeros(12, dtype=np.uint64)
    oinf = C.c_uint8(0)
    lib().orc_g1_msm(_p(bases_xy), _p8(inf), _p(scalars_mont), C.c_size_t(n), _p(out), C.byref(oinf),
                     C.c_int(threads))
    return out, oinf.value


def g1_msm_naive(bases_xy, scalars_mont, inf=None):
    bases_xy, scalars_mont = _a(bases_xy, 12), _a(scalars_mont, 4)
    n = min(bases_xy.size // 12, scalars_mont.size // 4)
    inf = _inf(inf, bases_xy.size // 12)
    out = np.zeros(12, dtype=np.uint64)
    oinf = C.c_uint8(0)
    lib().orc_g1_msm_naive(_p(bases_xy), _p8(inf), _p(scalars_mont), C.c_size_t(n), _p(out), C.byref(oinf))
    return out, oinf.value


# ----------------------------------------------------------------------------- G2
def g2_coeff_b():
    out = np.zeros(12, dtype=np.uint64)
    lib().orc_g2_coeff_b(_p(out))
    return out


def g2_on_curve(xy, inf=0):
    xy = _a(xy, 24)
    return bool(lib().orc_g2_on_curve(_p(xy), C.c_uint8(int(inf))))


def g2_scalar_mul(xy, k_limbs, inf=0):
    xy = _a(xy, 24)
    k = _a(k_limbs)
    out = np.zeros(24, dtype=np.uint64)
    oinf = C.c_uint8(0)
    lib().orc_g2_scalar_mul(_p(xy), C.c_uint8(int(inf)), _p(k), C.c_int(k.size), _p(out), C.byref(oinf))
    return out, oinf.value


def g2_add(a_xy, b_xy, a_inf=0, b_inf=0):
    a_xy, b_xy = _a(a_xy, 24), _a(b_xy, 24)
    out = np.zeros(24, dtype=np.uint64)
    oinf = C.c_uint8(0)
    lib().orc_g2_add(_p(a_xy), C.c_uint8(int(a_inf)), _p(b_xy), C.c_uint8(int(b_inf)), _p(out), C.byref(oinf))
    return out, oinf.value


def g2_generate(gen_xy, seed, n, first=0):
    gen_xy = _a(gen_xy, 24)
    out = np.zeros((n, 24), dtype=np.uint64)
    lib().orc_g2_generate(_p(gen_xy), C.c_uint64(seed), C.c_size_t(first), C.c_size_t(n), _p(out))
    return out


def g2_msm(bases_xy, scalars_mont, inf=None, threads=1):
    bases_xy, scalars_mont = _a(bases_xy, 24), _a(scalars_mont, 4)
    n = min(bases_xy.size // 24, scalars_mont.size // 4)
    inf = _inf(inf, bases_xy.size // 24)
    out = np.zeros(24, dtype=np.uint64)
    oinf = C.c_uint8(0)
    lib().orc_g2_msm(_p(bases_xy), _p8(inf), _p(scalars_mont), C.c_size_t(n), _p(out), C.byref(oinf),
                     C.c_int(threads))
    return out, oinf.value


def g2_msm_naive(bases_xy, scalars_mont, inf=None):
    bases_xy, scalars_mont = _a(bases_xy, 24), _a(scalars_mont, 4)
    n = min(bases_xy.size // 24, scalars_mont.size // 4)
    inf = _inf(inf, bases_xy.size // 24)
    out = np.zeros(24, dtype=np.uint64)
    oinf = C.c_uint8(0)
    lib().orc_g2_msm_naive(_p(bases_xy), _p8(inf), _p(scalars_mont), C.c_size_t(n), _p(out), C.byref(oinf))
    return out, oinf.value


# ----------------------------------------------------------------------------- domain / NTT
KIND = {"fft": 0, "ifft": 1, "coset_fft": 2, "coset_ifft": 3}


def domain_params(log_n):
    out = np.zeros((5, 4), dtype=np.uint64)
    ok = lib().orc_domain_params(C.c_uint(log_n), _p(out))
    if not ok:
        raise ValueError("domain too large")
    return dict(group_gen=out[0], group_gen_inv=out[1], size_inv=out[2], generator_inv=out[3], size_as_fe=out[4])


def ntt(data, kind):
    """In-order transform of a (2^k,4) Montgomery Fr array; returns a new array."""
    data = _a(data, 4).copy()
    n = data.size // 4
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    rc = lib().orc_ntt_fr(_p(data), C.c_uint(log_n), C.c_uint(KIND[kind] if isinstance(kind, str) else kind))
    if rc:
        raise ValueError("orc_ntt_fr rc=%d" % rc)
    return data


def divide_by_vanishing_on_coset(data):
    data = _a(data, 4).copy()
    n = data.size // 4
    lib().orc_divide_by_vanishing_on_coset(_p(data), C.c_uint(n.bit_length() - 1))
    return data


def horner(coeffs, point):
    coeffs, point = _a(coeffs, 4), _a(point, 4)
    out = np.zeros(4, dtype=np.uint64)
    lib().orc_fr_horner(_p(coeffs), C.c_size_t(coeffs.size // 4), _p(point), _p(out))
    return out


# ----------------------------------------------------------------------------- Beaver
def beaver_mask(s, x):
    s, x = _a(s, 4), _a(x, 4)
    out = np.empty_like(s)
    lib().orc_beaver_mask(_p(s), _p(x), _p(out), C.c_size_t(s.size // 4))
    return out


def beaver_combine(x, y, z, sx, oy, is_leader, spdz=False):
    """additive: x,y,z (n,4).  SPDZ: x,y,z (2,n,4) = [sh plane, mac plane]; sx,oy always (n,4)."""
    x, y, z, sx, oy = _a(x, 4), _a(y, 4), _a(z, 4), _a(sx, 4), _a(oy, 4)
    n = sx.size // 4
    assert x.size // 4 == (2 * n if spdz else n)
    out = np.empty_like(x)
    lib().orc_beaver_combine(_p(x), _p(y), _p(z), _p(sx), _p(oy), _p(out), C.c_size_t(n),
                             C.c_uint(int(bool(is_leader))), C.c_uint(int(bool(spdz))))
    return out


def open_sum(parts):
    parts = _a(parts, 4)
    P, n = parts.shape[0], parts.shape[1]
    out = np.zeros((n, 4), dtype=np.uint64)
    lib().orc_open_sum(_p(parts), C.c_uint(P), _p(out), C.c_size_t(n))
    return out


def spdz_mac_check(vals, macs, is_leader):
    vals, macs = _a(vals, 4), _a(macs, 4)
    out = np.empty_like(vals)
    lib().orc_spdz_mac_check(_p(vals), _p(macs), _p(out), C.c_size_t(vals.size // 4), C.c_uint(int(bool(is_leader))))
    return out


VEC_OP = {"sub": 0, "mul": 1, "mul_const": 2, "axpy": 3}


def vec_op(op, a, b=None, c=None):
    a = _a(a, 4)
    b = _a(b, 4) if b is not None else None
    c = _a(c, 4) if c is not None else None
    out = np.empty_like(a)
    lib().orc_vec_op(C.c_uint(VEC_OP[op]), _p(a), _p(b) if b is not None else None,
                     _p(c) if c is not None else None, _p(out), C.c_size_t(a.size // 4))
    return out


# ----------------------------------------------------------------------------- next rows
def spmv(row_ptr, col, coeff, x):
    """evaluate_constraint (src/groth16.rs:205-234) per row of a public CSR matrix on local values x (m,4)"""
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint64)
    col = np.ascontiguousarray(col, dtype=np.uint32)
    coeff, x = _a(coeff, 4), _a(x, 4)
    rows = row_ptr.size - 1
    out = np.zeros((rows, 4), dtype=np.uint64)
    lib().orc_spmv(_p(row_ptr), col.ctypes.data_as(C.POINTER(C.c_uint32)), _p(coeff), C.c_size_t(rows), _p(x), _p(out))
    return out


def fr_vec_serialize(vals):
    """the byte string MpcSerNet::broadcast sends for a Vec<Fr>: u64 LE length + 32 LE canonical bytes each"""
    vals = _a(vals, 4)
    n = vals.size // 4
    out = np.zeros(8 + 32 * n, dtype=np.uint8)
    lib().orc_fr_vec_serialize(_p(vals), C.c_size_t(n), _p8(out))
    return out


def fr_vec_deserialize(buf, n):
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    out = np.zeros((n, 4), dtype=np.uint64)
    lib().orc_fr_vec_deserialize.restype = C.c_long
    rc = lib().orc_fr_vec_deserialize(_p8(buf), C.c_size_t(n), _p(out))
    if rc:
        raise ValueError("deserialize failed: rc=%d" % rc)
    return out


def poly_div(num, den):
    """divide_with_q_and_r (univariate/mod.rs:133-172): (quotient, remainder) with leading zeros truncated"""
    num, den = _a(num, 4), _a(den, 4)
    n = num.size // 4
    q = np.zeros((max(n, 1), 4), dtype=np.uint64)
    r = np.zeros((max(n, 1), 4), dtype=np.uint64)
    ql, rl = C.c_size_t(0), C.c_size_t(0)
    rc = lib().orc_poly_div(_p(num), C.c_size_t(n), _p(den), C.c_size_t(den.size // 4), _p(q), C.byref(ql), _p(r),
                            C.byref(rl))
    if rc:
        raise ZeroDivisionError("Dividing by zero polynomial")
    return q[:ql.value].copy(), r[:rl.value].copy()


# ----------------------------------------------------------------------------- the whole prove sequence
def groth16_prove(pkarr, mats, num_inputs, z, log_n, r, s, threads=8):
    """create_proof (src/groth16.rs:68-183) on plain values through the oracle's own MSM / NTT / SpMV"""
    n = 1 << log_n
    nc = len(mats[0][0]) - 1
    ev = []
    for row_ptr, col, coeff in mats:
        v = np.zeros((n, 4), dtype=np.uint64)
        v[:nc] = spmv(row_ptr, col, coeff, z)
        ev.append(v)
    ev[0][nc:nc + num_inputs] = z[:num_inputs]                            # :272-276
    a1, b1, c1 = (ntt(ntt(v, "ifft"), "coset_fft") for v in ev)
    ab = vec_op("sub", vec_op("mul", a1, b1), c1)
    h = ntt(divide_by_vanishing_on_coset(ab), "coset_ifft")
    hq, hinf = pkarr["h_query"]
    h_acc = g1_msm(hq, h[:len(hq)], inf=hinf, threads=threads)
    lq, linf = pkarr["l_query"]
    l_acc = g1_msm(lq, z[num_inputs:], inf=linf, threads=threads)
    from_mont = lambda x: fr("from_mont", x[None])[0]

    def coeff(msm, add, smul, query, vk_param, delta, k):
        pts, inf = query
        acc = smul(delta, from_mont(k))                                     # initial = delta * k
        acc = add(acc[0], pts[0], acc[1], inf[0])                           # + query[0]
        m = msm(pts[1:], z[1:], inf=inf[1:], threads=threads)
        acc = add(acc[0], m[0], acc[1], m[1])
        return add(acc[0], vk_param, acc[1], 0)

    g_a = coeff(g1_msm, g1_add, g1_scalar_mul, pkarr["a_query"], pkarr["alpha_g1"], pkarr["delta_g1"], r)
    g1_b = coeff(g1_msm, g1_add, g1_scalar_mul, pkarr["b_g1_query"], pkarr["beta_g1"], pkarr["delta_g1"], s)
    g2_b = coeff(g2_msm, g2_add, g2_scalar_mul, pkarr["b_g2_query"], pkarr["beta_g2"], pkarr["delta_g2"], s)
    rs = fr("neg", fr("mul", r[None], s[None]))[0]
    t1 = g1_scalar_mul(g_a[0], from_mont(s), g_a[1])
    t2 = g1_scalar_mul(g1_b[0], from_mont(r), g1_b[1])
    t3 = g1_scalar_mul(pkarr["delta_g1"], from_mont(rs))
    g_c = g1_add(t1[0], t2[0], t1[1], t2[1])
    for t in (t3, l_acc, h_acc):
        g_c = g1_add(g_c[0], t[0], g_c[1], t[1])
    return {"a": g_a, "b": g2_b, "c": g_c, "h": h}


# ----------------------------------------------------------------------------- Marlin AHP rounds on plain values
FR_MODULUS = 0x12ab655e9a2ca55660b44d1e5c37b00159aa76fed00000010a11800000000001


def fr_to_ints(a):
    """(n,4) Montgomery limbs -> python ints"""
    c = fr("from_mont", _a(a, 4).reshape(-1, 4))
    return [int(r[0]) | int(r[1]) << 64 | int(r[2]) << 128 | int(r[3]) << 192 for r in c]


def fr_from_ints(vals):
    m = (1 << 64) - 1
    c = np.array([[v & m, (v >> 64) & m, (v >> 128) & m, (v >> 192) & m] for v in vals], dtype=np.uint64).reshape(-1, 4)
    return fr("to_mont", c) if len(c) else c


def _trim(ints):
    ints = list(ints)
    while ints and ints[-1] == 0:
        ints.pop()
    return ints


def _next_pow2(size):
    return 1 << max(size - 1, 0).bit_length()


def _ntt_ints(vals, n, kind):
    v = list(vals) + [0] * (n - len(vals))
    assert len(v) == n
    return fr_to_ints(ntt(fr_from_ints(v), kind))


def _div_ints(num, den):
    q, r = poly_div(fr_from_ints(num), fr_from_ints(den))
    return fr_to_ints(q) if len(q) else [], fr_to_ints(r) if len(r) else []


def _vanishing(n):
    return [FR_MODULUS - 1] + [0] * (n - 1) + [1]


def marlin_rounds(mats, num_constraints, num_inputs, x, w, blinders, mask, alpha, etas):
    """AHPForR1CS::prover_init / prover_first_round / prover_second_round (arkworks/marlin/src/ahp/prover.rs:212-566)
    for a single party holding the opened values, every polynomial a truncated list of python ints.  mats = three
    (row_ptr, col, coeff_ints) square matrices; x, w, blinders, mask, alpha, etas python ints."""
    P = FR_MODULUS
    nh, nx = _next_pow2(num_constraints), _next_pow2(num_inputs)
    assert nx == num_inputs and len(x) + len(w) == num_constraints
    z = list(x) + list(w)

    def inner(mat):                                            # prover.rs:258-278
        row_ptr, col, coeff = mat
        return [sum(coeff[k] * z[col[k]] for k in range(row_ptr[r], row_ptr[r + 1])) % P for r in range(num_constraints)]

    z_a, z_b = inner(mats[0]), inner(mats[1])
    # first round (prover.rs:321-374)
    x_poly = _trim(_ntt_ints(x, nx, "ifft"))
    x_evals = _ntt_ints(x_poly, nh, "fft")
    ratio = nh // nx
    w_ext = list(w) + [0] * (nh - nx - len(w))
    w_evals = [0 if k % ratio == 0 else (w_ext[k - k // ratio - 1] - x_evals[k]) % P for k in range(nh)]

    def blind(evals, r):
        p = _ntt_ints(evals, nh, "ifft") + [0]
        p[nh] = r % P
        p[0] = (p[0] - r) % P
        return _trim(p)

    w_poly, rem = _div_ints(blind(w_evals, blinders[0]), _vanishing(nx))
    assert not rem, "w(x) + r v_H must vanish on the input domain"
    z_a_poly, z_b_poly = blind(z_a, blinders[1]), blind(z_b, blinders[2])
    mask = list(mask)
    assert len(mask) == 3 * nh + 2 * 1 - 2
    _, mrem = _div_ints(mask, _vanishing(nh))
    mask[0] = (mask[0] - (mrem[0] if mrem else 0)) % P
    mask = _trim(mask)
    # second round (prover.rs:461-545)
    def poly_mul(a, b):                                        # dense.rs:567-583
        if not a or not b:
            return []
        n = _next_pow2(len(a) + len(b))
        ea, eb = _ntt_ints(a, n, "fft"), _ntt_ints(b, n, "fft")
        return _trim(_ntt_ints([u * v % P for u, v in zip(ea, eb)], n, "ifft"))

    z_c = poly_mul(z_a_poly, z_b_poly)
    summed = [c * etas[2] % P for c in z_c]
    for i, (a, b) in enumerate(zip(z_a_poly, z_b_poly)):
        if i < len(summed):
            summed[i] = (summed[i] + etas[0] * a + etas[1] * b) % P
    summed = _trim(summed)
    gen = fr_to_ints(domain_params(nh.bit_length() - 1)["group_gen"].reshape(1, 4))[0]
    vanish_alpha = (pow(alpha, nh, P) - 1) % P
    r_evals, h = [], 1
    for _ in range(nh):                                        # ahp/mod.rs:357-364
        r_evals.append(vanish_alpha * pow((alpha - h) % P, P - 2, P) % P)
        h = h * gen % P
    r_alpha = _trim(_ntt_ints(r_evals, nh, "ifft"))
    t_evals = [0] * nh                                         # calculate_t, prover.rs:406-423
    period = nh // nx
    for (row_ptr, col, coeff), eta in zip(mats, etas):
        for r in range(num_constraints):
            for k in range(row_ptr[r], row_ptr[r + 1]):
                c = col[k]
                if c < nx:
                    idx = c * period
                else:
                    i = c - nx
                    idx = i + i // (period - 1) + 1
                t_evals[idx] = (t_evals[idx] + eta * coeff[k] * r_evals[r]) % P
    t_poly = _trim(_ntt_ints(t_evals, nh, "ifft"))
    z_poly = [0] * nx + list(w_poly)                           # mul_by_vanishing_poly, dense.rs:155-162
    for i, c in enumerate(w_poly):
        z_poly[i] = (z_poly[i] - c) % P
    for i, c in enumerate(x_poly):
        z_poly[i] = (z_poly[i] + c) % P
    z_poly = _trim(z_poly)
    mul_n = _next_pow2(max(len(mask), len(r_alpha) + len(summed), len(t_poly) + len(z_poly)))
    ev = lambda p: _ntt_ints(p, mul_n, "fft")
    a, b, c, d = ev(r_alpha), ev(summed), ev(z_poly), ev(t_poly)
    rhs = _trim(_ntt_ints([(a[i] * b[i] - c[i] * d[i]) % P for i in range(mul_n)], mul_n, "ifft"))
    q_1 = [((mask[i] if i < len(mask) else 0) + (rhs[i] if i < len(rhs) else 0)) % P for i in range(max(len(mask), len(rhs)))]
    h_1, x_g_1 = _div_ints(_trim(q_1), _vanishing(nh))
    return dict(z_a=z_a, z_b=z_b, w=w_poly, z_a_poly=z_a_poly, z_b_poly=z_b_poly, mask=mask, z_c=z_c, t=t_poly,
                g_1=x_g_1[1:], h_1=h_1, x_g_1_0=x_g_1[0] if x_g_1 else 0)
