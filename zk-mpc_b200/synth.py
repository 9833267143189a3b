"""Seeded synthetic inputs for the share-MSM / share-NTT / Beaver hot path (numpy, host side).

Convention: SplitMix64 stream -> 4 limbs per element, top limb masked to 253 bits,
rejected while >= r, and the accepted limbs are used *directly as the Montgomery representation*,
which is what the reference's own sampler does (arkworks/algebra/ff/src/fields/arithmetic.rs:200-219).
An additive share of anything is uniform in Fr, so "share-like" == uniform.
"""
import numpy as np

FR_MOD_LIMBS = np.array([725501752471715841, 6461107452199829505, 6968279316240510977, 1345280370688173398],
                        dtype=np.uint64)
FR_R_LIMBS = np.array([9015221291577245683, 8239323489949974514, 1646089257421115374, 958099254763297437],
                      dtype=np.uint64)          # Montgomery form of 1 (fr.rs:58-64)
_GOLD = np.uint64(0x9E3779B97F4A7C15)


def _mix64(z):
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def _stream(seed, start, count):
    with np.errstate(over="ignore"):
        idx = np.arange(start + 1, start + count + 1, dtype=np.uint64)
        return _mix64(np.uint64(seed) + idx * _GOLD)


def _lt_mod(limbs):
    """vectorised limbs < r (most significant limb first)"""
    lt = np.zeros(limbs.shape[0], dtype=bool)
    eq = np.ones(limbs.shape[0], dtype=bool)
    for i in (3, 2, 1, 0):
        lt |= eq & (limbs[:, i] < FR_MOD_LIMBS[i])
        eq &= limbs[:, i] == FR_MOD_LIMBS[i]
    return lt


def fr_uniform(seed, n):
    """(n,4) uint64, uniform in [0, r), to be read as Montgomery limbs."""
    out = np.empty((n, 4), dtype=np.uint64)
    todo = np.arange(n)
    cursor = 0
    while todo.size:
        m = todo.size
        cand = _stream(seed, cursor, 4 * m).reshape(m, 4)
        cursor += 4 * m
        cand[:, 3] &= np.uint64(0xFFFFFFFFFFFFFFFF >> 3)      # REPR_SHAVE_BITS = 3 (fr.rs:56)
        ok = _lt_mod(cand)
        out[todo[ok]] = cand[ok]
        todo = todo[~ok]
    return out


def fr_witness_like(seed, n):
    """Half the entries in {0,1} (the reference short-cuts both: variable_base.rs:19,44-48), rest uniform."""
    out = fr_uniform(seed, n)
    sel = _stream(seed ^ 0xABCDEF, 1 << 40, n) & np.uint64(3)
    out[sel == 0] = 0
    out[sel == 1] = FR_R_LIMBS
    return out


def bench_seed(log_n):
    return 0x5EED0000 + log_n


# ----------------------------------------------------------------------------- synthetic Groth16 instances
def r1cs_matrices(seed, num_constraints, num_vars, max_terms=4):
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
        coeff = np.tile(FR_R_LIMBS, (nnz, 1))
        other = rng.random(nnz) < 0.3
        coeff[other] = fr_uniform(seed * 7 + k, int(other.sum()))
        mats.append((row_ptr, col, coeff))
    return mats


def proving_key_arrays(g1_generate, g2_generate, seed, num_vars, num_inputs, n):
    """queries of the right lengths with a few infinity entries (variables with zero coefficient do produce them,
    SURVEY.md Appendix A).  g1_generate(seed, count) / g2_generate(seed, count) -> affine point arrays."""
    def g1(k, count):
        pts = g1_generate(seed + k, count)
        inf = np.zeros(count, dtype=np.uint8)
        if count > 8:
            inf[[3, count // 2]] = 1
        return pts, inf

    def g2(k, count):
        pts = g2_generate(seed + k, count)
        inf = np.zeros(count, dtype=np.uint8)
        if count > 8:
            inf[[5]] = 1
        return pts, inf

    singles = g1_generate(seed + 99, 3)
    singles2 = g2_generate(seed + 98, 2)
    return dict(a_query=g1(1, num_vars), b_g1_query=g1(2, num_vars), b_g2_query=g2(3, num_vars), h_query=g1(4, n - 1),
                l_query=g1(5, num_vars - num_inputs), alpha_g1=singles[0], beta_g1=singles[1], delta_g1=singles[2],
                beta_g2=singles2[0], delta_g2=singles2[1])


def additive_shares(seed, opened, parties, num_public, sub):
    """split `opened` (m,4) into `parties` additive shares; the first num_public entries are Public values, i.e. held
    by the leader with 0 elsewhere (from_public).  sub(a, b) = a - b on (m,4) arrays (a field subtraction)."""
    shares = [fr_uniform(seed + p, opened.shape[0]) for p in range(parties)]
    rest = np.zeros_like(opened)
    for p in range(1, parties):
        shares[p][:num_public] = 0
    acc = opened.copy()
    for p in range(1, parties):
        acc = sub(acc, shares[p])
    shares[0] = acc
    return shares
