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
