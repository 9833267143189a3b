"""GPU parity: device field arithmetic and the Beaver / elementwise kernels vs the oracle,
through the C ABI (host buffers).  Bit-exact (integer arithmetic)."""
import numpy as np
import pytest

import pyref as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def H(pkg):
    pkg.host.init()
    pkg.host.set_party(0, 3)
    return pkg.host


def _edge(field, S, n, seed):
    """uniform elements plus hand-picked extremes (0, 1, p-1, ...) in Montgomery form"""
    mod = P.R_MOD if field == "fr" else P.Q_MOD
    to_arr = P.fr_to_mont_arr if field == "fr" else P.fq_to_mont_arr
    edge = to_arr([0, 1, 2, mod - 1, mod - 2, (mod - 1) // 2, (1 << 64) - 1, 1 << 64, (1 << 200) + 12345])
    if field == "fr":
        rnd = S.fr_uniform(seed, n)
    else:
        rng = np.random.default_rng(seed)
        rnd = to_arr([int.from_bytes(rng.bytes(48), "little") % mod for _ in range(min(n, 4000))])
    return np.concatenate([edge, rnd])


@pytest.mark.parametrize("field", ["fr", "fq"])
def test_device_field_ops(H, orc, pkg, field):
    a = _edge(field, pkg.synth, 20000, 1)
    b = _edge(field, pkg.synth, 20000, 2)[::-1].copy()
    ofn = orc.fr if field == "fr" else orc.fq
    for op in ("add", "sub", "mul"):
        assert np.array_equal(H.field_op(field, op, a, b), ofn(op, a, b)), op
    assert np.array_equal(H.field_op(field, "mul_narrow", a, b), ofn("mul", a, b))
    for op in ("neg", "sqr", "from_mont"):
        assert np.array_equal(H.field_op(field, op, a), ofn(op, a)), op
    canon = ofn("from_mont", a)
    assert np.array_equal(H.field_op(field, "to_mont", canon), a)
    assert np.array_equal(H.field_op(field, "inv", a[:600]), ofn("inv", a[:600]))


@pytest.mark.parametrize("n", [0, 1, 255, 256, 257, 100003, 1 << 18])
def test_beaver_mask(H, orc, pkg, n):
    s, x = pkg.synth.fr_uniform(3, n), pkg.synth.fr_uniform(4, n)
    got = H.beaver_mask(s, x)
    assert got.shape == (n, 4)
    if n:
        assert np.array_equal(got, orc.beaver_mask(s, x))


@pytest.mark.parametrize("spdz", [False, True])
@pytest.mark.parametrize("leader", [False, True])
@pytest.mark.parametrize("n", [1, 257, 70001, 1 << 17])
def test_beaver_combine(H, orc, pkg, n, leader, spdz):
    S = pkg.synth
    m = 2 * n if spdz else n
    x, y, z = S.fr_uniform(10, m), S.fr_uniform(11, m), S.fr_uniform(12, m)
    sx, oy = S.fr_uniform(13, n), S.fr_uniform(14, n)
    assert np.array_equal(H.beaver_combine(x, y, z, sx, oy, leader, spdz),
                          orc.beaver_combine(x, y, z, sx, oy, leader, spdz))


def test_beaver_combine_dummy_triples(H, orc, pkg):
    # DummyFieldTripleSource (wire/field.rs:44-63): leader (1,1,1), others (0,0,0)
    S = pkg.synth
    n = 5000
    sx, oy = S.fr_uniform(21, n), S.fr_uniform(22, n)
    one = np.tile(S.FR_R_LIMBS, (n, 1))
    zero = np.zeros_like(one)
    assert np.array_equal(H.beaver_combine(one, one, one, sx, oy, True), orc.beaver_combine(one, one, one, sx, oy, True))
    assert not H.beaver_combine(zero, zero, zero, sx, oy, False).any()


@pytest.mark.parametrize("parties", [1, 2, 3, 4])
def test_open_sum(H, orc, pkg, parties):
    n = 30011
    parts = np.stack([pkg.synth.fr_uniform(30 + p, n) for p in range(parties)])
    assert np.array_equal(H.open_sum(parts), orc.open_sum(parts))


def test_spdz_mac_check(H, orc, pkg):
    n = 9999
    v, m = pkg.synth.fr_uniform(40, n), pkg.synth.fr_uniform(41, n)
    for leader in (False, True):
        assert np.array_equal(H.spdz_mac_check(v, m, leader), orc.spdz_mac_check(v, m, leader))


def test_vec_ops(H, orc, pkg):
    n = 65537
    a, b = pkg.synth.fr_uniform(50, n), pkg.synth.fr_uniform(51, n)
    c = pkg.synth.fr_uniform(52, 1)[0]
    assert np.array_equal(H.vec_op("sub", a, b), orc.vec_op("sub", a, b))
    assert np.array_equal(H.vec_op("mul", a, b), orc.vec_op("mul", a, b))
    assert np.array_equal(H.vec_op("mul_const", a, c=c), orc.vec_op("mul_const", a, c=c))
    assert np.array_equal(H.vec_op("axpy", a, b, c), orc.vec_op("axpy", a, b, c))


def test_full_beaver_protocol_on_gpu(H, orc, pkg):
    """FieldShare::batch_mul for 3 parties, local halves on the GPU, the network replaced by
    open_sum; the opened product equals a*b (share/field.rs:84-93 debug check)."""
    S = pkg.synth
    n, parties = 4096, 3
    a_sh = [S.fr_uniform(60 + p, n) for p in range(parties)]
    b_sh = [S.fr_uniform(70 + p, n) for p in range(parties)]
    a, b = orc.open_sum(np.stack(a_sh)), orc.open_sum(np.stack(b_sh))
    one, zero = np.tile(S.FR_R_LIMBS, (n, 1)), np.zeros((n, 4), dtype=np.uint64)
    trip = [(one, one, one)] + [(zero, zero, zero)] * (parties - 1)
    sx = H.open_sum(np.stack([H.beaver_mask(a_sh[p], trip[p][0]) for p in range(parties)]))
    oy = H.open_sum(np.stack([H.beaver_mask(b_sh[p], trip[p][1]) for p in range(parties)]))
    outs = [H.beaver_combine(trip[p][0], trip[p][1], trip[p][2], sx, oy, p == 0) for p in range(parties)]
    assert np.array_equal(H.open_sum(np.stack(outs)), orc.fr("mul", a, b))
