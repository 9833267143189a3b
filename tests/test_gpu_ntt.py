"""GPU parity: share NTT (mpc_cuda_ntt_fr) vs the oracle's restatement of Radix2EvaluationDomain,
through the C ABI with host buffers.  Bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KINDS = ["fft", "ifft", "coset_fft", "coset_ifft"]


@pytest.fixture(scope="module")
def H(pkg):
    pkg.host.init()
    pkg.host.set_party(0, 3)
    return pkg.host


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 4, 5, 7, 8, 9, 10, 11, 13, 15, 16, 17, 18, 20])
def test_ntt_matches_oracle(H, orc, pkg, log_n, kind):
    v = pkg.synth.fr_uniform(0x100 + log_n, 1 << log_n)
    assert np.array_equal(H.ntt(v, kind), orc.ntt(v, kind))


@pytest.mark.parametrize("kind", KINDS)
def test_ntt_batch(H, orc, pkg, kind):
    # witness_map runs three equal-size transforms back to back (src/groth16.rs:278-282)
    n, batch = 1 << 13, 3
    v = pkg.synth.fr_uniform(0x200, n * batch)
    got = H.ntt(v, kind, batch=batch)
    for b in range(batch):
        assert np.array_equal(got[b * n:(b + 1) * n], orc.ntt(v[b * n:(b + 1) * n], kind))


def test_ntt_edge_values(H, orc, pkg):
    n = 1 << 9
    zero = np.zeros((n, 4), dtype=np.uint64)
    for kind in KINDS:
        assert not H.ntt(zero, kind).any()
    delta = zero.copy()
    delta[0] = pkg.synth.FR_R_LIMBS            # fft(delta_0) = all ones
    assert np.array_equal(H.ntt(delta, "fft"), np.tile(pkg.synth.FR_R_LIMBS, (n, 1)))
    v = pkg.synth.fr_uniform(0x201, n)
    v[::7] = 0
    v[3] = np.array([725501752471715840, 6461107452199829505, 6968279316240510977, 1345280370688173398],
                    dtype=np.uint64)           # r - 1 as raw limbs (a valid Montgomery pattern)
    for kind in KINDS:
        assert np.array_equal(H.ntt(v, kind), orc.ntt(v, kind))


def test_ntt_2_22_matches_oracle(H, orc, pkg):
    v = pkg.synth.fr_uniform(0x122, 1 << 22)
    assert np.array_equal(H.ntt(v, "coset_fft"), orc.ntt(v, "coset_fft"))


@pytest.mark.parametrize("log_n", [22, 24])
def test_ntt_full_size_properties(H, orc, pkg, log_n):
    """BASELINE sizes: round trips and evaluation at points of the domain (size-independent properties)"""
    n = 1 << log_n
    v = pkg.synth.fr_uniform(0x300 + log_n, n)
    f = H.ntt(v, "fft")
    assert np.array_equal(H.ntt(f, "ifft"), v)
    cf = H.ntt(v, "coset_fft")
    assert np.array_equal(H.ntt(cf, "coset_ifft"), v)
    # out[i] = poly(w^i): Horner on the CPU at three indices (w = group_gen of the domain)
    w = orc.domain_params(log_n)["group_gen"]
    g = orc.constants()["fr_gen"]
    for i in (1, 5, n - 1):
        pt = orc.constants()["fr_r"].copy()
        base, e = w.copy(), i
        while e:                                # w^i by square and multiply through the oracle field ops
            if e & 1:
                pt = orc.fr("mul", pt[None], base[None])[0]
            base = orc.fr("sqr", base[None])[0]
            e >>= 1
        assert np.array_equal(orc.horner(v, pt), f[i])
        assert np.array_equal(orc.horner(v, orc.fr("mul", pt[None], g[None])[0]), cf[i])


def test_divide_by_vanishing(H, orc, pkg):
    for log_n in (3, 13, 16):
        v = pkg.synth.fr_uniform(0x400 + log_n, 1 << log_n)
        assert np.array_equal(H.divide_by_vanishing_on_coset(v), orc.divide_by_vanishing_on_coset(v))


def test_witness_map_sequence(H, orc, pkg):
    """R1CStoQAP::witness_map (src/groth16.rs:278-303) on one party's local values, every step on the GPU,
    against the same sequence on the oracle (public Beaver opens replaced by the plain product)."""
    n = 1 << 12
    S = pkg.synth
    a, b, c = S.fr_uniform(1, n), S.fr_uniform(2, n), S.fr_uniform(3, n)

    def seq(X, ntt, vec, div):
        a1, b1 = ntt(ntt(a, "ifft"), "coset_fft"), ntt(ntt(b, "ifft"), "coset_fft")
        ab = vec("mul", a1, b1)
        c1 = ntt(ntt(c, "ifft"), "coset_fft")
        return ntt(div(vec("sub", ab, c1)), "coset_ifft")

    got = seq(H, H.ntt, H.vec_op, H.divide_by_vanishing_on_coset)
    exp = seq(orc, orc.ntt, orc.vec_op, orc.divide_by_vanishing_on_coset)
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("log_n", [3, 13])
def test_fused_witness_map_three_parties(H, orc, pkg, log_n):
    """mpc_cuda_witness_map_begin/finish for 3 parties with the reference's dummy triples (leader 1, others 0;
    mpc-algebra/src/wire/field.rs:44-63), the two opens played by open_sum: the opened h equals the plain
    witness_map of the opened inputs computed by the oracle."""
    S = pkg.synth
    n, parties = 1 << log_n, 3
    sh = [[S.fr_uniform(0x700 + 10 * p + k, n) for k in range(3)] for p in range(parties)]      # a, b, c shares
    one, zero = np.tile(S.FR_R_LIMBS, (n, 1)), np.zeros((n, 4), dtype=np.uint64)
    trip = [(one, one, one)] + [(zero, zero, zero)] * (parties - 1)
    begun = [H.witness_map_begin(sh[p][0], sh[p][1], sh[p][2], trip[p][0], trip[p][1]) for p in range(parties)]
    sx = H.open_sum(np.stack([bg[0] for bg in begun]))
    oy = H.open_sum(np.stack([bg[1] for bg in begun]))
    hs = [H.witness_map_finish(begun[p][2], trip[p][2], sx, oy, p == 0) for p in range(parties)]
    got = H.open_sum(np.stack(hs))
    a, b, c = (orc.open_sum(np.stack([sh[p][k] for p in range(parties)])) for k in range(3))
    a1, b1, c1 = (orc.ntt(orc.ntt(v, "ifft"), "coset_fft") for v in (a, b, c))
    exp = orc.ntt(orc.divide_by_vanishing_on_coset(orc.vec_op("sub", orc.vec_op("mul", a1, b1), c1)), "coset_ifft")
    assert np.array_equal(got, exp)
    with pytest.raises(H.MpcCudaError):
        H.witness_map_finish(begun[0][2], trip[0][2], sx, oy, True)        # state already consumed
