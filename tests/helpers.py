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
    return pkg.synth.r1cs_matrices(seed, num_constraints, num_vars, max_terms)


def synth_proving_key(orc, seed, num_vars, num_inputs, n):
    import pyref as P
    g2gen, _ = P.g2_to_arr((P.G2_X, P.G2_Y))
    import __graft_entry__ as ge
    return ge.load_package().synth.proving_key_arrays(orc.g1_generate, lambda s, c: orc.g2_generate(g2gen, s, c), seed,
                                                      num_vars, num_inputs, n)


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
    return orc.groth16_prove(pkarr, mats, num_inputs, z, log_n, r, s)


# ----------------------------------------------------------------------------- synthetic Marlin instances
def synth_marlin_instance(pkg, orc, seed, num_constraints, num_inputs, max_terms=3):
    """a satisfied square R1CS: random sparse A, B over a random assignment z = [x | w] with z_0 = 1, and C with one
    entry per row at column 0 carrying (Az)_r (Bz)_r.  Returns (mats as limb arrays, mats as int lists, x, w limbs)."""
    S = pkg.synth
    a, b, _ = S.r1cs_matrices(seed, num_constraints, num_constraints, max_terms)
    z = S.fr_uniform(seed + 1, num_constraints)
    z[0] = S.FR_R_LIMBS
    za, zb = orc.spmv(*a, z), orc.spmv(*b, z)
    c = (np.arange(num_constraints + 1, dtype=np.uint64), np.zeros(num_constraints, dtype=np.uint32), orc.vec_op("mul", za, zb))
    mats = [a, b, c]
    ints = [([int(v) for v in m[0]], [int(v) for v in m[1]], orc.fr_to_ints(m[2])) for m in mats]
    return mats, ints, z[:num_inputs], z[num_inputs:]
