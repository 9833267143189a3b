"""GPU parity: share MSM on G1 / G2 (mpc_cuda_msm_*) vs the oracle's restatement of
VariableBaseMSM + AffineMsm, through the C ABI.  Bit-exact affine Montgomery limbs + infinity flag."""
import numpy as np
import pytest

import helpers
import pyref as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def H(pkg):
    pkg.host.init()
    pkg.host.set_party(0, 3)
    return pkg.host


@pytest.fixture(autouse=True)
def _reset_options(H):
    yield
    H.set_option("msm_window_bits", 0)
    H.set_option("msm_task_len", 0)


def _same(a, b):
    return np.array_equal(a[0], b[0]) and a[1] == b[1]


@pytest.fixture(scope="module")
def bases8k(orc):
    return orc.g1_generate(0xC0FFEE, 8192)


@pytest.mark.parametrize("n", [0, 1, 2, 31, 32, 33, 100, 1000, 8191])
def test_g1_msm_matches_oracle(H, orc, pkg, bases8k, n):
    sc = (pkg.synth.fr_witness_like if n % 2 else pkg.synth.fr_uniform)(0xA0 + n, n)
    got = H.msm_g1(bases8k[:n], sc)
    exp = orc.g1_msm(bases8k[:n], sc, threads=8)
    assert _same(got, exp)
    if n == 0:
        assert got[1] == 1 and P.fq_from_mont_arr(got[0].reshape(2, 6)) == [0, 1]


@pytest.mark.parametrize("c", [3, 4, 5, 8, 11, 13, 15, 16])
def test_g1_msm_every_window_width(H, orc, pkg, bases8k, c):
    n = 3000
    sc = pkg.synth.fr_uniform(0xB0 + c, n)
    sc[:4] = P.fr_to_mont_arr([P.R_MOD - 1, P.R_MOD - 2, (1 << 252) + 1, 1 << (c - 1)])
    H.set_option("msm_window_bits", c)
    assert _same(H.msm_g1(bases8k[:n], sc), orc.g1_msm(bases8k[:n], sc, threads=8))


@pytest.mark.parametrize("task_len", [1, 4, 8, 1000])
def test_g1_msm_bucket_splitting(H, orc, pkg, bases8k, task_len):
    # skewed scalars: half of them 0/1 (one bucket holds thousands of points -> split + tree join)
    n = 8000
    sc = pkg.synth.fr_witness_like(0xD0, n)
    H.set_option("msm_task_len", task_len)
    assert _same(H.msm_g1(bases8k[:n], sc), orc.g1_msm(bases8k[:n], sc, threads=8))
    ones = np.tile(pkg.synth.FR_R_LIMBS, (n, 1))
    assert _same(H.msm_g1(bases8k[:n], ones), orc.g1_msm(bases8k[:n], ones, threads=8))


def test_g1_msm_edge_cases(H, orc, pkg, bases8k):
    n = 64
    bases, scalars = bases8k[:n].copy(), pkg.synth.fr_uniform(5, n)
    # all-zero scalars -> affine zero (0, 1, inf)
    out, inf = H.msm_g1(bases, np.zeros_like(scalars))
    assert inf == 1 and P.fq_from_mont_arr(out.reshape(2, 6)) == [0, 1]
    # infinity bases skipped; duplicates hit the doubling branch; P and -P with equal scalars cancel
    bases[1] = bases[0]
    q = P.g1_from_arr(bases[2])
    bases[3] = P.g1_to_arr((q[0], (-q[1]) % P.Q_MOD))[0]
    scalars[1] = scalars[0]
    scalars[3] = scalars[2]
    infs = np.zeros(n, dtype=np.uint8)
    infs[5] = infs[63] = 1
    assert _same(H.msm_g1(bases, scalars, inf=infs), orc.g1_msm(bases, scalars, inf=infs))
    # everything cancels: (P, -P) with the same scalar -> infinity
    assert H.msm_g1(bases[2:4], scalars[2:4])[1] == 1
    # ragged lengths truncate to min(len) (variable_base.rs:16-18)
    assert _same(H.msm_g1(bases[:40], scalars), orc.g1_msm(bases[:40], scalars[:40]))
    # all inputs the same point and scalar
    same_b, same_s = np.tile(bases[7], (50, 1)), np.tile(scalars[7], (50, 1))
    assert _same(H.msm_g1(same_b, same_s), orc.g1_msm(same_b, same_s))


def test_g1_msm_2_16_matches_oracle(H, orc, pkg):
    n = 1 << 16
    dev = H.g1_generate(0x5EED0010, n)
    bases = dev.download().reshape(n, 12)
    dev.free()
    assert np.array_equal(bases[:64], orc.g1_generate(0x5EED0010, 64))
    assert np.array_equal(bases[-64:], orc.g1_generate(0x5EED0010, 64, first=n - 64))
    sc = pkg.synth.fr_uniform(0x5EED0010, n)
    assert _same(H.msm_g1(bases, sc), orc.g1_msm(bases, sc, threads=16))


def test_g1_handle_and_offsets(H, orc, pkg, bases8k):
    h = H.register_bases(bases8k)
    sc = pkg.synth.fr_uniform(0xE0, 8192)
    assert _same(H.msm_handle(h, sc), orc.g1_msm(bases8k, sc, threads=8))
    assert _same(H.msm_handle(h, sc[:1000], offset=5000), orc.g1_msm(bases8k[5000:6000], sc[:1000], threads=8))
    with pytest.raises(H.MpcCudaError):
        H.msm_handle(h, sc, offset=1, n=8192)
    h.release()
    with pytest.raises(H.MpcCudaError):
        H.msm_handle(h if h.handle else H.BaseHandle(12345, 8192, False), sc)


@pytest.mark.parametrize("log_n", [18, 20, 22])
def test_g1_msm_full_size_exact(H, orc, pkg, log_n):
    """BASELINE sizes: bases are k_i*G, so the MSM must equal (sum s_i k_i mod r)*G, which the CPU evaluates
    exactly with integer dot products and one scalar multiplication."""
    n = 1 << log_n
    seed = pkg.synth.bench_seed(log_n)
    dev = H.g1_generate(seed, n)
    h = H.register_bases_dev(dev, n)
    ks = helpers.gen_ks(pkg, seed, n)
    for sc in (pkg.synth.fr_uniform(seed, n), pkg.synth.fr_witness_like(seed + 1, n)):
        assert _same(H.msm_handle(h, sc), helpers.expected_msm_of_generated(orc, sc, ks))
    h.release()
    dev.free()


def test_spdz_multi_scale_pub_group(H, orc, pkg, bases8k):
    n = 500
    sh, mac = pkg.synth.fr_uniform(1, n), pkg.synth.fr_uniform(2, n)
    r_sh, r_mac = H.multi_scale_pub_group(bases8k[:n], np.stack([sh, mac]))
    exp = orc.g1_msm(bases8k[:n], sh)
    assert _same(r_sh, exp) and _same(r_mac, exp)          # spdz.rs:484 quirk: macs built from sh


# ----------------------------------------------------------------------------- G2
@pytest.fixture(scope="module")
def g2bases(orc):
    g, _ = P.g2_to_arr((P.G2_X, P.G2_Y))
    return orc.g2_generate(g, 0xB2, 600)


@pytest.mark.parametrize("n", [0, 1, 33, 600])
def test_g2_msm_matches_oracle(H, orc, pkg, g2bases, n):
    sc = (pkg.synth.fr_witness_like if n % 2 else pkg.synth.fr_uniform)(0xF0 + n, n)
    assert _same(H.msm_g2(g2bases[:n], sc), orc.g2_msm(g2bases[:n], sc, threads=8))


def test_g2_generate_and_handle(H, orc, pkg, g2bases):
    dev = H.g2_generate(0xB2, 600)
    got = dev.download().reshape(600, 24)
    dev.free()
    assert np.array_equal(got, g2bases)
    h = H.register_bases(g2bases, g2=True)
    sc = pkg.synth.fr_uniform(0xF7, 600)
    H.set_option("msm_task_len", 2)
    assert _same(H.msm_handle(h, sc), orc.g2_msm(g2bases, sc, threads=8))
    h.release()


# ----------------------------------------------------------------------------- precomputed window tables
@pytest.mark.parametrize("c", [0, 4, 9, 13, 16])
def test_g1_precomputed_table(H, orc, pkg, bases8k, c):
    infs = np.zeros(8192, dtype=np.uint8)
    infs[[3, 4000]] = 1
    h = H.register_bases(bases8k, inf=infs).precompute(c)
    for seed, sc in ((1, pkg.synth.fr_uniform(0x1A0 + c, 8192)), (2, pkg.synth.fr_witness_like(0x1B0 + c, 8192))):
        sc[:3] = P.fr_to_mont_arr([P.R_MOD - 1, 0, 1])
        assert _same(H.msm_handle(h, sc), orc.g1_msm(bases8k, sc, inf=infs, threads=8))
    sc = pkg.synth.fr_uniform(0x1C0, 777)
    assert _same(H.msm_handle(h, sc, offset=5000), orc.g1_msm(bases8k[5000:5777], sc, inf=infs[5000:5777], threads=8))
    h.release()


def test_g1_precomputed_table_full_size_exact(H, orc, pkg):
    log_n = 20
    n = 1 << log_n
    seed = pkg.synth.bench_seed(log_n)
    dev = H.g1_generate(seed, n)
    h = H.register_bases_dev(dev, n).precompute(0)
    ks = helpers.gen_ks(pkg, seed, n)
    for sc in (pkg.synth.fr_uniform(seed, n), pkg.synth.fr_witness_like(seed + 1, n)):
        assert _same(H.msm_handle(h, sc), helpers.expected_msm_of_generated(orc, sc, ks))
    h.release()
    dev.free()


def test_g2_precomputed_table(H, orc, pkg, g2bases):
    h = H.register_bases(g2bases, g2=True).precompute(7)
    sc = pkg.synth.fr_uniform(0x1D0, 600)
    assert _same(H.msm_handle(h, sc), orc.g2_msm(g2bases, sc, threads=8))
    h.release()
