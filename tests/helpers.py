"""Shared helpers of the parity tests (CPU side, exact integer arithmetic)."""
import numpy as np

import pyref as P

_GOLD = np.uint64(0x9E3779B97F4A7C15)


def gen_ks(pkg, seed, n, first=0):
    """the multipliers k_i of the synthetic CRS bases[i] = k_i * G (mpc_cuda_g1_generate_dev / orc_g1_generate)"""
    with np.errstate(over="ignore"):
        idx = np.arange(first + 1, first + n + 1, dtype=np.uint64)
        k = pkg.synth._mix64(np.uint64(seed) + idx * _GOLD)
    k[k == 0] = 1
    return k


def _pieces16(a):
    return [((a >> np.uint64(16 * j)) & np.uint64(0xFFFF)) for j in range(4)]


def dot_mod_r(scalars_mont, ks):
    """(sum_i s_i * k_i) mod r for Montgomery-form Fr limbs (n,4) and uint64 multipliers, exactly:
    16-bit pieces keep every partial dot product below 2^64 for n <= 2^32."""
    scalars_mont = np.ascontiguousarray(scalars_mont, dtype=np.uint64)
    kp = _pieces16(np.ascontiguousarray(ks, dtype=np.uint64))
    total = 0
    for limb in range(4):
        lp = _pieces16(scalars_mont[:, limb])
        for a in range(4):
            for b in range(4):
                total += int(np.dot(lp[a], kp[b])) << (64 * limb + 16 * (a + b))
    return total * pow(P.FR_RR, -1, P.R_MOD) % P.R_MOD


def expected_msm_of_generated(orc, scalars_mont, ks):
    """affine (xy, inf) of (sum s_i k_i) * G through one oracle scalar multiplication"""
    e = dot_mod_r(scalars_mont, ks)
    return orc.g1_scalar_mul(orc.g1_generator(), np.array(P.to_limbs(e, 4), dtype=np.uint64))


# ----------------------------------------------------------------------------- synthetic Groth16 instances
import threading


def synth_r1cs(pkg, seed, num_constraints, num_vars, max_terms=4):
    """three random public CSR matrices shaped like an R1CS: short rows, most coefficients 1, a few rows long"""
    rng = np.random.default_rng(seed)
    mats = []
    for k in range(3):
        lens = rng.integers(0, max_terms + 1, num_constraints)
        if num_constraints > 4:
            lens[rng.integers(0, num_constraints)] = min(num_vars, 200)         # one long row
        row_ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
        nnz = int(row_ptr[-1])
        col = rng.integers(0, num_vars, nnz).astype(np.uint32)
        coeff = np.tile(pkg.synth.FR_R_LIMBS, (nnz, 1))
        other = rng.random(nnz) < 0.3
        coeff[other] = pkg.synth.fr_uniform(seed * 7 + k, int(other.sum()))
        mats.append((row_ptr, col, coeff))
    return mats


def synth_proving_key(orc, seed, num_vars, num_inputs, n):
    """queries of the right lengths with a few infinity entries (variables with zero coefficient), as arrays"""
    import pyref as P
    g2gen, _ = P.g2_to_arr((P.G2_X, P.G2_Y))

    def g1(k, count):
        pts = orc.g1_generate(seed + k, count)
        inf = np.zeros(count, dtype=np.uint8)
        if count > 8:
            inf[[3, count // 2]] = 1
        return pts, inf

    def g2(k, count):
        pts = orc.g2_generate(g2gen, seed + k, count)
        inf = np.zeros(count, dtype=np.uint8)
        if count > 8:
            inf[[5]] = 1
        return pts, inf

    singles = orc.g1_generate(seed + 99, 3)
    singles2 = orc.g2_generate(g2gen, seed + 98, 2)
    return dict(a_query=g1(1, num_vars), b_g1_query=g1(2, num_vars), b_g2_query=g2(3, num_vars), h_query=g1(4, n - 1),
                l_query=g1(5, num_vars - num_inputs), alpha_g1=singles[0], beta_g1=singles[1], delta_g1=singles[2],
                beta_g2=singles2[0], delta_g2=singles2[1])


class ThreadNet:
    """MpcSerNet::broadcast among party threads of one process (the role LocalTestNet plays in the reference)"""

    class _Shared:
        def __init__(self, n):
            self.barrier = threading.Barrier(n)
            self.slots = [None] * n

    def __init__(self, party, n_parties, shared):
        self.party, self.n_parties, self._sh = party, n_parties, shared

    @staticmethod
    def make(n_parties):
        sh = ThreadNet._Shared(n_parties)
        return [ThreadNet(p, n_parties, sh) for p in range(n_parties)]

    def exchange(self, payload):
        self._sh.slots[self.party] = payload
        self._sh.barrier.wait()
        out = list(self._sh.slots)
        self._sh.barrier.wait()
        return out


def oracle_groth16(orc, pkarr, mats, num_inputs, z, log_n, r, s):
    """create_proof (src/groth16.rs:68-183) on plain values through the oracle's own MSM / NTT / SpMV"""
    n = 1 << log_n
    nc = len(mats[0][0]) - 1
    ev = []
    for row_ptr, col, coeff in mats:
        v = np.zeros((n, 4), dtype=np.uint64)
        v[:nc] = orc.spmv(row_ptr, col, coeff, z)
        ev.append(v)
    ev[0][nc:nc + num_inputs] = z[:num_inputs]                            # :272-276
    a1, b1, c1 = (orc.ntt(orc.ntt(v, "ifft"), "coset_fft") for v in ev)
    ab = orc.vec_op("sub", orc.vec_op("mul", a1, b1), c1)
    h = orc.ntt(orc.divide_by_vanishing_on_coset(ab), "coset_ifft")
    hq, hinf = pkarr["h_query"]
    h_acc = orc.g1_msm(hq, h[:len(hq)], inf=hinf, threads=8)
    lq, linf = pkarr["l_query"]
    l_acc = orc.g1_msm(lq, z[num_inputs:], inf=linf, threads=8)
    from_mont = lambda x: orc.fr("from_mont", x[None])[0]

    def coeff(msm, add, smul, query, vk_param, delta, k):
        pts, inf = query
        acc = smul(delta, from_mont(k))                                     # initial = delta * k
        acc = add(acc[0], pts[0], acc[1], inf[0])                           # + query[0]
        m = msm(pts[1:], z[1:], inf=inf[1:], threads=8)
        acc = add(acc[0], m[0], acc[1], m[1])
        return add(acc[0], vk_param, acc[1], 0)

    g_a = coeff(orc.g1_msm, orc.g1_add, orc.g1_scalar_mul, pkarr["a_query"], pkarr["alpha_g1"], pkarr["delta_g1"], r)
    g1_b = coeff(orc.g1_msm, orc.g1_add, orc.g1_scalar_mul, pkarr["b_g1_query"], pkarr["beta_g1"], pkarr["delta_g1"], s)
    g2_b = coeff(orc.g2_msm, orc.g2_add, orc.g2_scalar_mul, pkarr["b_g2_query"], pkarr["beta_g2"], pkarr["delta_g2"], s)
    rs = orc.fr("neg", orc.fr("mul", r[None], s[None]))[0]
    t1 = orc.g1_scalar_mul(g_a[0], from_mont(s), g_a[1])
    t2 = orc.g1_scalar_mul(g1_b[0], from_mont(r), g1_b[1])
    t3 = orc.g1_scalar_mul(pkarr["delta_g1"], from_mont(rs))
    g_c = orc.g1_add(t1[0], t2[0], t1[1], t2[1])
    for t in (t3, l_acc, h_acc):
        g_c = orc.g1_add(g_c[0], t[0], g_c[1], t[1])
    return {"a": g_a, "b": g2_b, "c": g_c, "h": h}
