"""Pins the C oracle (oracle/zkmpc_oracle.c) before anything trusts it.

Sources of truth, in order: (1) the reference's constants and KATs,
(2) the independent Python big-int model in tests/pyref.py, (3) algebraic identities.
CPU only.
"""
import random

import numpy as np
import pytest

import pyref as P


def limbs(v, n):
    return np.array(P.to_limbs(v, n), dtype=np.uint64)


# ----------------------------------------------------------------------------- constants (fr.rs / fq.rs)
def test_parameter_self_consistency(orc):
    c = orc.constants()
    assert P.from_limbs(c["fr_mod"]) == P.R_MOD
    assert P.from_limbs(c["fq_mod"]) == P.Q_MOD
    assert P.R_MOD.bit_length() == 253 and P.Q_MOD.bit_length() == 377
    # R = 2^256 mod r, decimal from fr.rs:58 ; R2 = R^2 mod r ; INV = -r^-1 mod 2^64
    assert P.from_limbs(c["fr_r"]) == P.FR_RR % P.R_MOD == P.FR_R_DEC
    assert P.from_limbs(c["fr_r2"]) == P.FR_RR * P.FR_RR % P.R_MOD
    assert (P.R_MOD * P.FR_INV + 1) % (1 << 64) == 0
    assert P.from_limbs(c["fq_r"]) == P.FQ_RR % P.Q_MOD == P.FQ_R_DEC
    assert P.from_limbs(c["fq_r2"]) == P.FQ_RR * P.FQ_RR % P.Q_MOD
    assert (P.Q_MOD * P.FQ_INV + 1) % (1 << 64) == 0
    # GENERATOR = 22 * R (fr.rs:77-85), -5 * R for Fq (fq.rs:64-74)
    assert P.from_limbs(c["fr_gen"]) == 22 * P.FR_RR % P.R_MOD == P.FR_GEN_MONT_DEC
    assert (-5) * P.FQ_RR % P.Q_MOD == P.FQ_GEN_MONT_DEC
    # 2-adicity
    assert (P.R_MOD - 1) % (1 << 47) == 0 and P.FR_T % 2 == 1
    assert (P.Q_MOD - 1) % (1 << 46) == 0 and P.FQ_T % 2 == 1


def test_root_of_unity_kat(orc):
    # fields/tests.rs:352-370 (Fq) and its Fr analogue from fr.rs:9-20
    c = orc.constants()
    root = P.from_limbs(c["fr_root"]) * pow(P.FR_RR, -1, P.R_MOD) % P.R_MOD
    assert root == pow(22, P.FR_T, P.R_MOD)
    assert pow(root, 1 << 47, P.R_MOD) == 1 and pow(root, 1 << 46, P.R_MOD) != 1
    assert pow(22, (P.R_MOD - 1) // 2, P.R_MOD) == P.R_MOD - 1          # generator is a non-residue
    fq_root_limbs = [2022196864061697551, 17419102863309525423, 8564289679875062096,
                     17152078065055548215, 17966377291017729567, 68610905582439508]
    fq_root = P.from_limbs(fq_root_limbs) * pow(P.FQ_RR, -1, P.Q_MOD) % P.Q_MOD
    assert fq_root == pow(-5 % P.Q_MOD, P.FQ_T, P.Q_MOD)
    assert pow(fq_root, 1 << 46, P.Q_MOD) == 1
    assert pow(-5 % P.Q_MOD, (P.Q_MOD - 1) // 2, P.Q_MOD) == P.Q_MOD - 1
    # oracle pow path: domain_params(47).group_gen is the root itself
    d = orc.domain_params(47)
    assert P.fr_from_mont_arr(d["group_gen"])[0] == root


# ----------------------------------------------------------------------------- field arithmetic vs Python ints
def _rand_vals(rng, mod, n):
    edge = [0, 1, 2, mod - 1, mod - 2, (mod - 1) // 2, (1 << 64) - 1, 1 << 64, (1 << 128) + 5]
    return edge + [rng.randrange(mod) for _ in range(n - len(edge))]


@pytest.mark.parametrize("field", ["fr", "fq"])
def test_field_ops_match_python(orc, field):
    rng = random.Random(1234)
    mod = P.R_MOD if field == "fr" else P.Q_MOD
    to_arr = P.fr_to_mont_arr if field == "fr" else P.fq_to_mont_arr
    from_arr = P.fr_from_mont_arr if field == "fr" else P.fq_from_mont_arr
    op = orc.fr if field == "fr" else orc.fq
    nl = 4 if field == "fr" else 6
    a = _rand_vals(rng, mod, 300)
    b = list(reversed(_rand_vals(rng, mod, 300)))
    A, B = to_arr(a), to_arr(b)
    assert from_arr(op("add", A, B)) == [(x + y) % mod for x, y in zip(a, b)]
    assert from_arr(op("sub", A, B)) == [(x - y) % mod for x, y in zip(a, b)]
    assert from_arr(op("mul", A, B)) == [(x * y) % mod for x, y in zip(a, b)]
    assert from_arr(op("sqr", A)) == [(x * x) % mod for x in a]
    assert from_arr(op("neg", A)) == [(-x) % mod for x in a]
    assert from_arr(op("inv", A)) == [pow(x, -1, mod) if x else 0 for x in a]
    # into_repr / from_repr
    canon = op("from_mont", A)
    assert [P.from_limbs(r) for r in canon] == a
    assert np.array_equal(op("to_mont", canon), A)
    # outputs stay fully reduced
    for arr in (op("add", A, B), op("mul", A, B), op("sub", A, B)):
        assert all(P.from_limbs(r) < mod for r in arr.reshape(-1, nl))


def test_fr_batch_inverse(orc):
    rng = random.Random(7)
    a = [0, 5, 0] + [rng.randrange(P.R_MOD) for _ in range(50)] + [0]
    out = orc.fr_batch_inv(P.fr_to_mont_arr(a))
    assert P.fr_from_mont_arr(out) == [pow(x, -1, P.R_MOD) if x else 0 for x in a]


def test_fq2_ops_match_python(orc):
    rng = random.Random(99)
    a = [(rng.randrange(P.Q_MOD), rng.randrange(P.Q_MOD)) for _ in range(64)] + [(0, 0), (1, 0), (0, 1), (P.Q_MOD - 1, 3)]
    b = list(reversed(a))

    def pack(v):
        return P.fq_to_mont_arr([c for pair in v for c in pair]).reshape(-1, 12)

    def unpack(arr):
        flat = P.fq_from_mont_arr(arr.reshape(-1, 6))
        return [(flat[2 * i], flat[2 * i + 1]) for i in range(len(flat) // 2)]

    A, B = pack(a), pack(b)
    assert unpack(orc.fq2("mul", A, B)) == [P.fq2_mul(x, y) for x, y in zip(a, b)]
    assert unpack(orc.fq2("add", A, B)) == [P.fq2_add(x, y) for x, y in zip(a, b)]
    assert unpack(orc.fq2("sub", A, B)) == [P.fq2_sub(x, y) for x, y in zip(a, b)]
    assert unpack(orc.fq2("sqr", A)) == [P.fq2_mul(x, x) for x in a]
    assert unpack(orc.fq2("inv", A)) == [P.fq2_inv(x) if x != (0, 0) else (0, 0) for x in a]


# ----------------------------------------------------------------------------- curve KATs (curves/tests.rs)
def test_g1_generator_kat(orc):
    g = orc.g1_generator()
    assert P.g1_from_arr(g) == (P.G1_X, P.G1_Y)
    assert orc.g1_on_curve(g)
    # in the prime-order subgroup: r * G = O  (curves/tests.rs:35-40)
    out, inf = orc.g1_scalar_mul(g, limbs(P.R_MOD, 4))
    assert inf == 1
    assert P.fq_from_mont_arr(out.reshape(2, 6)) == [0, 1]              # affine zero is (0, 1, inf)
    assert P.ec_mul(P.F1, P.R_MOD, (P.G1_X, P.G1_Y)) is None


def test_g1_generator_raw_kat(orc):
    """curves/tests.rs:92-122: smallest x whose cofactor-cleared point is non-zero is hit at i == 1
    and equals the prime-subgroup generator."""
    i, x = 0, 0
    while True:
        rhs = (x ** 3 + 1) % P.Q_MOD
        if pow(rhs, (P.Q_MOD - 1) // 2, P.Q_MOD) in (0, 1):
            # sqrt via Tonelli-Shanks is overkill: check candidate through the oracle's scalar mul instead
            y = _sqrt_mod_q(rhs)
            y = y if y < (-y) % P.Q_MOD else (-y) % P.Q_MOD
            arr, _ = P.g1_to_arr((x, y))
            out, inf = orc.g1_scalar_mul(arr, limbs(P.G1_COFACTOR, 2))
            if not inf:
                assert i == 1
                assert P.g1_from_arr(out) == (P.G1_X, P.G1_Y)
                assert P.ec_mul(P.F1, P.G1_COFACTOR, (x, y)) == (P.G1_X, P.G1_Y)
                break
        i += 1
        x += 1
        assert i < 10


def _sqrt_mod_q(a):
    """Tonelli-Shanks over Fq (2-adicity 46, non-residue -5)"""
    if a == 0:
        return 0
    q = P.Q_MOD
    s, t = P.FQ_TWO_ADICITY, P.FQ_T
    z = pow(-5 % q, t, q)
    m, c, tt, r = s, z, pow(a, t, q), pow(a, (t + 1) // 2, q)
    while tt != 1:
        i, t2 = 0, tt
        while t2 != 1:
            t2 = t2 * t2 % q
            i += 1
        b = pow(c, 1 << (m - i - 1), q)
        m, c = i, b * b % q
        tt, r = tt * c % q, r * b % q
    assert r * r % q == a
    return r


def test_g2_generator_kat(orc):
    b = orc.g2_coeff_b()
    flat = P.fq_from_mont_arr(b.reshape(2, 6))
    assert (flat[0], flat[1]) == P.G2_B
    g, _ = P.g2_to_arr((P.G2_X, P.G2_Y))
    assert orc.g2_on_curve(g)
    out, inf = orc.g2_scalar_mul(g, limbs(P.R_MOD, 4))
    assert inf == 1
    # small multiples agree with the Python chord-and-tangent model
    for k in (1, 2, 3, 7, 12345678901234567890):
        out, inf = orc.g2_scalar_mul(g, limbs(k, 1))
        assert P.g2_from_arr(out, inf) == P.ec_mul(P.F2, k, (P.G2_X, P.G2_Y))


def test_g1_group_law_branches(orc):
    G = (P.G1_X, P.G1_Y)
    g, _ = P.g1_to_arr(G)
    p5 = P.ec_mul(P.F1, 5, G)
    a5, _ = P.g1_to_arr(p5)
    # generic add, doubling branch (P + P), inverse branch (P + -P), infinity operands
    out, inf = orc.g1_add(g, a5)
    assert P.g1_from_arr(out, inf) == P.ec_mul(P.F1, 6, G)
    out, inf = orc.g1_add(a5, a5)
    assert P.g1_from_arr(out, inf) == P.ec_mul(P.F1, 10, G)
    neg5, _ = P.g1_to_arr((p5[0], (-p5[1]) % P.Q_MOD))
    out, inf = orc.g1_add(a5, neg5)
    assert inf == 1
    zero, _ = P.g1_to_arr(None)
    out, inf = orc.g1_add(zero, a5, a_inf=1)
    assert P.g1_from_arr(out, inf) == p5
    out, inf = orc.g1_add(a5, zero, b_inf=1)
    assert P.g1_from_arr(out, inf) == p5
    for k in (1, 2, 255, (1 << 64) - 1, P.R_MOD - 1):
        out, inf = orc.g1_scalar_mul(g, limbs(k, 4))
        assert P.g1_from_arr(out, inf) == P.ec_mul(P.F1, k, G)


def test_g1_generate_matches_python(orc):
    pts = orc.g1_generate(0x1234, 6)
    G = (P.G1_X, P.G1_Y)
    for i in range(6):
        assert P.g1_from_arr(pts[i]) == P.ec_mul(P.F1, P.gen_scalar_k(0x1234, i), G)
        assert orc.g1_on_curve(pts[i])
    # `first` offsets the schedule
    assert np.array_equal(orc.g1_generate(0x1234, 2, first=3), pts[3:5])


# ----------------------------------------------------------------------------- MSM: Pippenger == naive == Python
def _msm_inputs(orc, pkg, n, seed, witness=False):
    synth = pkg.synth
    bases = orc.g1_generate(seed, n)
    scalars = synth.fr_witness_like(seed, n) if witness else synth.fr_uniform(seed, n)
    return bases, scalars


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 100])
def test_g1_msm_small_vs_python(orc, pkg, n):
    bases, scalars = _msm_inputs(orc, pkg, n, 0xA0 + n, witness=(n % 2 == 0))
    pts = [P.g1_from_arr(b) for b in bases]
    sc = P.fr_from_mont_arr(scalars)
    expect = P.ec_msm(P.F1, pts, sc)
    out, inf = orc.g1_msm(bases, scalars)
    assert P.g1_from_arr(out, inf) == expect
    out2, inf2 = orc.g1_msm_naive(bases, scalars)
    assert P.g1_from_arr(out2, inf2) == expect


def test_g1_msm_pippenger_equals_naive_1024(orc, pkg):
    # test-templates/src/msm.rs:16-33 (2^10 points)
    bases, scalars = _msm_inputs(orc, pkg, 1 << 10, 77)
    a = orc.g1_msm(bases, scalars)
    b = orc.g1_msm_naive(bases, scalars)
    assert np.array_equal(a[0], b[0]) and a[1] == b[1] == 0
    # threaded windows give the identical affine point
    c = orc.g1_msm(bases, scalars, threads=4)
    assert np.array_equal(a[0], c[0])


def test_g1_msm_edge_cases(orc, pkg):
    n = 64
    bases, scalars = _msm_inputs(orc, pkg, n, 5)
    one = pkg.synth.FR_R_LIMBS
    # all-zero scalars -> infinity (0, 1, inf)
    out, inf = orc.g1_msm(bases, np.zeros_like(scalars))
    assert inf == 1 and P.fq_from_mont_arr(out.reshape(2, 6)) == [0, 1]
    # all-one scalars -> plain sum (unit-scalar shortcut variable_base.rs:44-48)
    ones = np.tile(one, (n, 1))
    out, inf = orc.g1_msm(bases, ones)
    acc = None
    for b in bases:
        acc = P.ec_add(P.F1, acc, P.g1_from_arr(b))
    assert P.g1_from_arr(out, inf) == acc
    # infinity bases are skipped; duplicate bases hit the doubling branch; P + (-P) cancels
    b2 = bases.copy()
    b2[1] = b2[0]
    b2[3] = b2[2]
    q = P.g1_from_arr(b2[2])
    b2[3] = P.g1_to_arr((q[0], (-q[1]) % P.Q_MOD))[0]
    s2 = scalars.copy()
    s2[1] = s2[0]
    s2[3] = s2[2]
    infs = np.zeros(n, dtype=np.uint8)
    infs[5] = 1
    pts = [P.g1_from_arr(b, i) for b, i in zip(b2, infs)]
    expect = P.ec_msm(P.F1, [p for p in pts if p is not None],
                      [s for s, p in zip(P.fr_from_mont_arr(s2), pts) if p is not None])
    out, inf = orc.g1_msm(b2, s2, inf=infs)
    assert P.g1_from_arr(out, inf) == expect
    # ragged lengths truncate to min(len) (variable_base.rs:16-18)
    out_a = orc.g1_msm(bases[:40], scalars)
    out_b = orc.g1_msm(bases[:40], scalars[:40])
    assert np.array_equal(out_a[0], out_b[0])
    # linearity: msm(s) + msm(t) == msm(s + t)
    t = pkg.synth.fr_uniform(99, n)
    lhs = orc.g1_add(orc.g1_msm(bases, scalars)[0], orc.g1_msm(bases, t)[0])
    rhs = orc.g1_msm(bases, orc.fr("add", scalars, t))
    assert np.array_equal(lhs[0], rhs[0])


def test_g2_msm_vs_python_and_naive(orc, pkg):
    g, _ = P.g2_to_arr((P.G2_X, P.G2_Y))
    n = 40
    bases = orc.g2_generate(g, 0xB2, n)
    scalars = pkg.synth.fr_witness_like(0xB2, n)
    assert all(orc.g2_on_curve(b) for b in bases)
    pts = [P.g2_from_arr(b) for b in bases]
    assert pts[3] == P.ec_mul(P.F2, P.gen_scalar_k(0xB2, 3), (P.G2_X, P.G2_Y))
    expect = P.ec_msm(P.F2, pts, P.fr_from_mont_arr(scalars))
    out, inf = orc.g2_msm(bases, scalars)
    assert P.g2_from_arr(out, inf) == expect
    out, inf = orc.g2_msm_naive(bases, scalars)
    assert P.g2_from_arr(out, inf) == expect


# ----------------------------------------------------------------------------- NTT
@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 5, 8])
@pytest.mark.parametrize("kind", ["fft", "ifft", "coset_fft", "coset_ifft"])
def test_ntt_matches_naive_dft(orc, pkg, log_n, kind):
    n = 1 << log_n
    data = pkg.synth.fr_uniform(0xF0 + log_n, n)
    vals = P.fr_from_mont_arr(data)
    assert P.fr_from_mont_arr(orc.ntt(data, kind)) == P.dft(vals, kind)


@pytest.mark.parametrize("log_n", [0, 1, 4, 9, 12])
def test_ntt_roundtrips(orc, pkg, log_n):
    # poly/src/test.rs:10-57: fft∘ifft == id, coset variants
    data = pkg.synth.fr_uniform(3 + log_n, 1 << log_n)
    assert np.array_equal(orc.ntt(orc.ntt(data, "fft"), "ifft"), data)
    assert np.array_equal(orc.ntt(orc.ntt(data, "coset_fft"), "coset_ifft"), data)


def test_ntt_is_evaluation(orc, pkg):
    # out[i] = p(w^i): check a few points with Horner (large enough to cross the root-compaction
    # threshold of fft.rs:201-212, i.e. num_chunks >= 128)
    log_n = 11
    n = 1 << log_n
    coeffs = pkg.synth.fr_uniform(21, n)
    ev = orc.ntt(coeffs, "fft")
    cev = orc.ntt(coeffs, "coset_fft")
    w = P.fr_root_of_unity(log_n)
    for i in (0, 1, 2, 777, n - 1):
        pt = P.fr_to_mont_arr([pow(w, i, P.R_MOD)])[0]
        assert np.array_equal(orc.horner(coeffs, pt), ev[i])
        pt = P.fr_to_mont_arr([22 * pow(w, i, P.R_MOD)])[0]
        assert np.array_equal(orc.horner(coeffs, pt), cev[i])


def test_domain_params(orc):
    for log_n in (0, 1, 13, 20, 24):
        d = orc.domain_params(log_n)
        w = P.fr_root_of_unity(log_n)
        assert P.fr_from_mont_arr(d["group_gen"])[0] == w
        assert P.fr_from_mont_arr(d["group_gen_inv"])[0] == pow(w, -1, P.R_MOD)
        assert P.fr_from_mont_arr(d["size_inv"])[0] == pow(1 << log_n, -1, P.R_MOD)
        assert P.fr_from_mont_arr(d["generator_inv"])[0] == pow(22, -1, P.R_MOD)


def test_divide_by_vanishing(orc, pkg):
    n = 64
    data = pkg.synth.fr_uniform(4, n)
    zinv = pow(pow(22, n, P.R_MOD) - 1, -1, P.R_MOD)
    expect = [v * zinv % P.R_MOD for v in P.fr_from_mont_arr(data)]
    assert P.fr_from_mont_arr(orc.divide_by_vanishing_on_coset(data)) == expect


# ----------------------------------------------------------------------------- Beaver
@pytest.mark.parametrize("spdz", [False, True])
@pytest.mark.parametrize("parties", [2, 3])
def test_beaver_batch_mul_protocol(orc, pkg, spdz, parties):
    """Runs FieldShare::batch_mul (share/field.rs:97-129) for all parties in-process, the
    network replaced by open_sum, with RANDOM triples; the opened product must equal a*b
    (the reference's own debug check, share/field.rs:84-93) and SPDZ MACs must verify."""
    n = 50
    S = pkg.synth
    rng = random.Random(5)
    a = [rng.randrange(P.R_MOD) for _ in range(n)]
    b = [rng.randrange(P.R_MOD) for _ in range(n)]
    tx = [rng.randrange(P.R_MOD) for _ in range(n)]
    ty = [rng.randrange(P.R_MOD) for _ in range(n)]
    tz = [x * y % P.R_MOD for x, y in zip(tx, ty)]

    def share(vals, seed):
        parts = [S.fr_uniform(seed + p, n) for p in range(parties - 1)]
        tot = P.fr_to_mont_arr(vals)
        last = tot
        for p in parts:
            last = orc.fr("sub", last, p)
        return parts + [last]

    def planes(parts):
        # MAC key is the constant 1 in the shipped system (spdz.rs:31-47), so mac == sh
        return [np.stack([p, p]) if spdz else p for p in parts]

    xs, ys = planes(share(a, 10)), planes(share(b, 20))
    txs, tys, tzs = planes(share(tx, 30)), planes(share(ty, 40)), planes(share(tz, 50))

    def mask(s, x):
        return np.stack([orc.beaver_mask(s[0], x[0]), orc.beaver_mask(s[1], x[1])]) if spdz else orc.beaver_mask(s, x)

    def open_(parts):
        sh = [p[0] if spdz else p for p in parts]
        vals = orc.open_sum(np.stack(sh))
        if spdz:
            dx = [orc.spdz_mac_check(vals, parts[p][1], p == 0) for p in range(parties)]
            assert not orc.open_sum(np.stack(dx)).any()                 # spdz.rs:190-194
        return vals

    sx = open_([mask(xs[p], txs[p]) for p in range(parties)])
    oy = open_([mask(ys[p], tys[p]) for p in range(parties)])
    outs = [orc.beaver_combine(txs[p], tys[p], tzs[p], sx, oy, p == 0, spdz) for p in range(parties)]
    prod = open_(outs)
    assert P.fr_from_mont_arr(prod) == [x * y % P.R_MOD for x, y in zip(a, b)]


def test_beaver_combine_formula(orc, pkg):
    n = 33
    S = pkg.synth
    x, y, z, sx, oy = (S.fr_uniform(60 + i, n) for i in range(5))
    X, Y, Z, SX, OY = (P.fr_from_mont_arr(v) for v in (x, y, z, sx, oy))
    for leader in (0, 1):
        out = orc.beaver_combine(x, y, z, sx, oy, leader)
        expect = [(Z[i] - Y[i] * SX[i] - X[i] * OY[i] + leader * SX[i] * OY[i]) % P.R_MOD for i in range(n)]
        assert P.fr_from_mont_arr(out) == expect
    # dummy triple source (wire/field.rs:44-63): leader holds (1,1,1), the others (0,0,0)
    one = np.tile(S.FR_R_LIMBS, (n, 1))
    out = orc.beaver_combine(one, one, one, sx, oy, 1)
    assert P.fr_from_mont_arr(out) == [(1 - SX[i] - OY[i] + SX[i] * OY[i]) % P.R_MOD for i in range(n)]
    zero = np.zeros_like(one)
    assert not orc.beaver_combine(zero, zero, zero, sx, oy, 0).any()


def test_vec_ops(orc, pkg):
    n = 17
    S = pkg.synth
    a, b = S.fr_uniform(1, n), S.fr_uniform(2, n)
    c = S.fr_uniform(3, 1)[0]
    A, B = P.fr_from_mont_arr(a), P.fr_from_mont_arr(b)
    Cc = P.fr_from_mont_arr(c)[0]
    assert P.fr_from_mont_arr(orc.vec_op("sub", a, b)) == [(x - y) % P.R_MOD for x, y in zip(A, B)]
    assert P.fr_from_mont_arr(orc.vec_op("mul", a, b)) == [(x * y) % P.R_MOD for x, y in zip(A, B)]
    assert P.fr_from_mont_arr(orc.vec_op("mul_const", a, c=c)) == [x * Cc % P.R_MOD for x in A]
    assert P.fr_from_mont_arr(orc.vec_op("axpy", a, b, c)) == [(x + Cc * y) % P.R_MOD for x, y in zip(A, B)]


def test_synth_sampler(pkg):
    S = pkg.synth
    a = S.fr_uniform(42, 1000)
    assert np.array_equal(a, S.fr_uniform(42, 1000))
    assert all(P.from_limbs(r) < P.R_MOD for r in a)
    w = S.fr_witness_like(42, 4000)
    zeros = (~w.any(axis=1)).sum()
    ones = (w == S.FR_R_LIMBS).all(axis=1).sum()
    assert 800 < zeros < 1200 and 800 < ones < 1200


# ----------------------------------------------------------------------------- next rows: SpMV, wire bytes, division
def test_oracle_spmv_serialize_division_against_bigints(orc, pkg):
    """the oracle's restatements of evaluate_constraint, CanonicalSerialize for Vec<Fr> and
    divide_with_q_and_r against plain Python integers"""
    import helpers
    S = pkg.synth
    rows, cols = 40, 25
    for row_ptr, col, coeff in helpers.synth_r1cs(pkg, 0x77, rows, cols):
        x = S.fr_uniform(0x78, cols)
        xi, ci = P.fr_from_mont_arr(x), P.fr_from_mont_arr(coeff)
        exp = [sum(ci[k] * xi[col[k]] for k in range(int(row_ptr[r]), int(row_ptr[r + 1]))) % P.R_MOD for r in range(rows)]
        assert P.fr_from_mont_arr(orc.spmv(row_ptr, col, coeff, x)) == exp
    v = S.fr_uniform(0x79, 9)
    wire = orc.fr_vec_serialize(v).tobytes()
    assert int.from_bytes(wire[:8], "little") == 9
    assert [int.from_bytes(wire[8 + 32 * i:40 + 32 * i], "little") for i in range(9)] == P.fr_from_mont_arr(v)
    assert np.array_equal(orc.fr_vec_deserialize(np.frombuffer(wire, dtype=np.uint8), 9), v)
    # p = q * d + r with deg r < deg d, for a linear and a cubic divisor
    p = S.fr_uniform(0x7A, 30)
    for d in (S.fr_uniform(0x7B, 2), S.fr_uniform(0x7C, 4)):
        q, r = orc.poly_div(p, d)
        pi, di, qi, ri = (P.fr_from_mont_arr(a) for a in (p, d, q, r))
        assert len(qi) == len(pi) - len(di) + 1 and len(ri) < len(di)
        back = [0] * len(pi)
        for i, a in enumerate(qi):
            for j, b in enumerate(di):
                back[i + j] = (back[i + j] + a * b) % P.R_MOD
        for i, a in enumerate(ri):
            back[i] = (back[i] + a) % P.R_MOD
        assert back == pi
    q0, r0 = orc.poly_div(np.zeros((3, 4), dtype=np.uint64), p[:2])          # zero dividend -> (0, 0)
    assert len(q0) == 0 and len(r0) == 0


def test_oracle_groth16_composition_runs(orc, pkg):
    """helpers.oracle_groth16 (the checker of the GPU prove sequence): r = s = 0 gives c = l_acc + h_acc and the
    A element is linear in the assignment"""
    import helpers
    S = pkg.synth
    nc, ni, nv, log_n = 3, 2, 6, 3
    mats = helpers.synth_r1cs(pkg, 0x7D, nc, nv)
    pkarr = helpers.synth_proving_key(orc, 0x7E, nv, ni, 1 << log_n)
    z = S.fr_uniform(0x7F, nv)
    zero = np.zeros(4, dtype=np.uint64)
    out = helpers.oracle_groth16(orc, pkarr, mats, ni, z, log_n, zero, zero)
    hq, hinf = pkarr["h_query"]
    lq, linf = pkarr["l_query"]
    h_acc = orc.g1_msm_naive(hq, out["h"][:len(hq)], inf=hinf)
    l_acc = orc.g1_msm_naive(lq, z[ni:], inf=linf)
    exp = orc.g1_add(h_acc[0], l_acc[0], h_acc[1], l_acc[1])
    assert np.array_equal(out["c"][0], exp[0]) and out["c"][1] == exp[1]
