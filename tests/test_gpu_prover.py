"""GPU parity for the rows either side of the NTT / MSM / Beaver kernels, the SPDZ
witness map, the wire-level packing rules (a5, a12) and the end-to-end Groth16 prove sequence (a16 / x1), all
against the same composition on the oracle.  Bit-exact."""
import ctypes as C
import threading

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


def _same(a, b):
    return np.array_equal(a[0], b[0]) and a[1] == b[1]


# ----------------------------------------------------------------------------- f2: constraint rows
@pytest.mark.parametrize("rows,cols", [(0, 3), (1, 1), (37, 50), (6574, 6580)])
def test_spmv_matches_evaluate_constraint(H, orc, pkg, rows, cols):
    for k, (row_ptr, col, coeff) in enumerate(helpers.synth_r1cs(pkg, 0x510 + rows, rows, cols)):
        m = H.CsrMatrix(row_ptr, col, coeff, cols)
        x = pkg.synth.fr_uniform(0x520 + k, cols)
        assert np.array_equal(m.spmv(x), orc.spmv(row_ptr, col, coeff, x))
        planes = np.stack([x, pkg.synth.fr_uniform(0x530 + k, cols)])          # SPDZ: sh and mac planes
        got = m.spmv(planes)
        assert np.array_equal(got[0], orc.spmv(row_ptr, col, coeff, planes[0]))
        assert np.array_equal(got[1], orc.spmv(row_ptr, col, coeff, planes[1]))
        m.release()


def test_spmv_dense_rows_use_the_warp_kernel(H, orc, pkg):
    rows, cols = 8, 4000
    rng = np.random.default_rng(3)
    row_ptr = (np.arange(rows + 1) * 1000).astype(np.uint64)
    col = rng.integers(0, cols, 8000).astype(np.uint32)
    coeff = pkg.synth.fr_uniform(0x540, 8000)
    coeff[::3] = pkg.synth.FR_R_LIMBS
    x = pkg.synth.fr_uniform(0x541, cols)
    m = H.CsrMatrix(row_ptr, col, coeff, cols)
    assert np.array_equal(m.spmv(x), orc.spmv(row_ptr, col, coeff, x))
    m.release()
    with pytest.raises(H.MpcCudaError):
        H.CsrMatrix(row_ptr, col + np.uint32(cols), coeff, cols)                 # column index out of range


# ----------------------------------------------------------------------------- f3: wire bytes
@pytest.mark.parametrize("n", [0, 1, 5, 1000, 1 << 16])
def test_serialize_matches_canonical_serialize(H, orc, pkg, n):
    v = pkg.synth.fr_uniform(0x610 + n, n)
    if n >= 5:
        v[:4] = P.fr_to_mont_arr([0, 1, P.R_MOD - 1, 1 << 200])
    wire = H.fr_serialize(v)
    assert np.array_equal(wire, orc.fr_vec_serialize(v))
    if n >= 5:      # the bytes are the canonical little-endian integers behind a u64 length
        assert int.from_bytes(wire[:8].tobytes(), "little") == n
        assert int.from_bytes(wire[8 + 64:8 + 96].tobytes(), "little") == P.R_MOD - 1
    assert np.array_equal(H.fr_deserialize(wire, n), v)
    assert np.array_equal(orc.fr_vec_deserialize(wire, n), v)
    x = pkg.synth.fr_uniform(0x620 + n, n)
    assert np.array_equal(H.beaver_mask_serialize(v, x), orc.fr_vec_serialize(orc.beaver_mask(v, x)))


def test_open_sum_from_wire_payloads(H, orc, pkg):
    n, parties = 3000, 3
    shares = np.stack([pkg.synth.fr_uniform(0x630 + p, n) for p in range(parties)])
    payloads = np.stack([orc.fr_vec_serialize(s) for s in shares])
    assert np.array_equal(H.open_sum_deserialize(payloads, n), orc.open_sum(shares))
    bad = payloads.copy()
    bad[1, 8 + 32 * 7:8 + 32 * 8] = 0xFF                                        # element 7 of party 1 >= r
    with pytest.raises(H.MpcCudaError) as e:
        H.open_sum_deserialize(bad, n)
    assert "element 7" in str(e.value)
    with pytest.raises(ValueError):
        orc.fr_vec_deserialize(bad[1], n)
    bad = payloads.copy()
    bad[2, 0] ^= 1                                                              # wrong length prefix
    with pytest.raises(H.MpcCudaError):
        H.open_sum_deserialize(bad, n)


# ----------------------------------------------------------------------------- f4: division by (x - z)
@pytest.mark.parametrize("n", [1, 2, 3, 17, 4096, 4097, 70000, (1 << 20) + 5])
def test_poly_div_linear_matches_divide_with_q_and_r(H, orc, pkg, n):
    p = pkg.synth.fr_uniform(0x710 + (n & 0xFF), n)
    z = pkg.synth.fr_uniform(0x720, 1)[0]
    den = np.stack([orc.fr("neg", z[None])[0], pkg.synth.FR_R_LIMBS])          # x - z
    q, rem = H.poly_div_linear(p, z)
    eq, er = orc.poly_div(p, den)
    assert np.array_equal(q[:len(eq)], eq) and not q[len(eq):].any()
    assert np.array_equal(rem, er[0] if len(er) else np.zeros(4, dtype=np.uint64))
    assert np.array_equal(rem, orc.horner(p, z))                                # remainder = p(z)
    assert np.array_equal(H.poly_evaluate(p, z), rem)


def test_poly_div_linear_edge_points(H, orc, pkg):
    n = 1000
    p = pkg.synth.fr_uniform(0x730, n)
    one = pkg.synth.FR_R_LIMBS
    for z in (np.zeros(4, dtype=np.uint64), one, P.fr_to_mont_arr([P.R_MOD - 1])[0]):
        den = np.stack([orc.fr("neg", z[None])[0], one])
        q, rem = H.poly_div_linear(p, z)
        eq, er = orc.poly_div(p, den)
        assert np.array_equal(q, eq)
        assert np.array_equal(rem, er[0] if len(er) else np.zeros(4, dtype=np.uint64))
    # exact division: p = (x - z) * q0 has remainder 0 (arkworks returns the empty remainder)
    z = pkg.synth.fr_uniform(0x731, 1)[0]
    q0 = pkg.synth.fr_uniform(0x732, 50)
    prod = np.zeros((51, 4), dtype=np.uint64)
    prod[1:] = q0
    prod[:50] = orc.vec_op("sub", prod[:50], orc.vec_op("mul_const", q0, c=z))
    q, rem = H.poly_div_linear(prod, z)
    assert np.array_equal(q, q0) and not rem.any()
    assert len(orc.poly_div(prod, np.stack([orc.fr("neg", z[None])[0], one]))[1]) == 0


# ----------------------------------------------------------------------------- SPDZ witness map
@pytest.mark.parametrize("log_n", [3, 12])
def test_spdz_witness_map_three_parties(H, orc, pkg, log_n):
    """[sh | mac] planes through mpc_cuda_witness_map_begin_ex / finish: both planes of the opened h equal the
    plain witness_map of the opened inputs; MAC key 1 shared as (1, 0, 0) (share/spdz.rs:31-47)"""
    S = pkg.synth
    n, parties = 1 << log_n, 3
    sh = [[S.fr_uniform(0x800 + 10 * p + k, n) for k in range(3)] for p in range(parties)]
    mac = [[S.fr_uniform(0x900 + 10 * p + k, n) for k in range(3)] for p in range(parties)]
    # make the macs consistent: sum_p mac_p = sum_p sh_p (key 1): fix party 0's mac plane
    for k in range(3):
        tot_sh = orc.open_sum(np.stack([sh[p][k] for p in range(parties)]))
        rest = orc.open_sum(np.stack([mac[p][k] for p in range(1, parties)]))
        mac[0][k] = orc.vec_op("sub", tot_sh, rest)
    one, zero = np.tile(S.FR_R_LIMBS, (n, 1)), np.zeros((n, 4), dtype=np.uint64)
    trip = [np.stack([one, one])] + [np.stack([zero, zero])] * (parties - 1)     # dummy triple, sh = mac = (1,0,0)
    begun = [H.witness_map_begin(*(np.stack([sh[p][k], mac[p][k]]) for k in range(3)), trip[p], trip[p], spdz=True)
             for p in range(parties)]
    sx = H.open_sum(np.stack([bg[0][0] for bg in begun]))
    oy = H.open_sum(np.stack([bg[1][0] for bg in begun]))
    # the MAC check of the two opens (share/spdz.rs:185-189): sum_p (mac_share_p * val - mac_p) == 0
    for val, idx in ((sx, 0), (oy, 1)):
        dx = np.stack([H.spdz_mac_check(val, begun[p][idx][1], p == 0) for p in range(parties)])
        assert not H.open_sum(dx).any()
    hs = [H.witness_map_finish(begun[p][2], trip[p], sx, oy, p == 0) for p in range(parties)]
    a, b, c = (orc.open_sum(np.stack([sh[p][k] for p in range(parties)])) for k in range(3))
    a1, b1, c1 = (orc.ntt(orc.ntt(v, "ifft"), "coset_fft") for v in (a, b, c))
    exp = orc.ntt(orc.divide_by_vanishing_on_coset(orc.vec_op("sub", orc.vec_op("mul", a1, b1), c1)), "coset_ifft")
    assert np.array_equal(H.open_sum(np.stack([h[0] for h in hs])), exp)         # sh plane
    assert np.array_equal(H.open_sum(np.stack([h[1] for h in hs])), exp)         # mac plane (key 1)
    # per party the result equals the oracle's Beaver semantics on its planes
    p = 1
    a1p, b1p, c1p = (np.stack([orc.ntt(orc.ntt(v, "ifft"), "coset_fft") for v in (sh[p][k], mac[p][k])]) for k in range(3))
    comb = orc.beaver_combine(trip[p], trip[p], trip[p], sx, oy, False, spdz=True)
    for plane in range(2):
        e = orc.ntt(orc.divide_by_vanishing_on_coset(orc.vec_op("sub", comb[plane], c1p[plane])), "coset_ifft")
        assert np.array_equal(hs[p][plane], e)


# ----------------------------------------------------------------------------- a5 / a12: Public / Shared packing
def test_wire_rules_for_fft_and_msm(H, orc, pkg):
    W, S = pkg.wire, pkg.synth
    n = 1 << 8
    pub = S.fr_uniform(0xA10, n)
    # all-Public stays Public: the same plain transform on every party
    for leader in (True, False):
        out = W.fft(W.MpcVec.public(pub), "coset_fft", leader)
        assert not out.shared.any() and np.array_equal(out.val, orc.ntt(pub, "coset_fft"))
    # mixed vector: Public(x) counts as x on the leader and 0 elsewhere; the opened transform is the transform
    # of the opened vector
    tags = np.arange(n) % 3 != 0
    shares = [S.fr_uniform(0xA20 + p, n) for p in range(3)]
    opened_in = orc.open_sum(np.stack(shares))
    opened_in[~tags] = pub[~tags]
    outs = []
    for p in range(3):
        v = shares[p].copy()
        v[~tags] = pub[~tags]
        out = W.fft(W.MpcVec(v, tags), "ifft", p == 0)
        assert out.shared.all()
        outs.append(out.val)
    assert np.array_equal(orc.open_sum(np.stack(outs)), orc.ntt(opened_in, "ifft"))
    # SPDZ: Public(x) -> sh as above, mac = x * mac_share; planes transform independently
    v = shares[1].copy()
    v[~tags] = pub[~tags]
    macs = S.fr_uniform(0xA30, n)
    for leader in (True, False):
        planes = W.force_shared(W.MpcVec(v, tags, macs), leader, True)
        exp_sh, exp_mac = v.copy(), macs.copy()
        exp_mac[~tags] = pub[~tags] if leader else 0
        if not leader:
            exp_sh[~tags] = 0
        assert np.array_equal(planes[0], exp_sh) and np.array_equal(planes[1], exp_mac)
        out = W.fft(W.MpcVec(v, tags, macs), "fft", leader, spdz=True)
        assert np.array_equal(out.val, orc.ntt(exp_sh, "fft")) and np.array_equal(out.mac, orc.ntt(exp_mac, "fft"))
    # MSM dispatch (wire/pairing.rs:714-777)
    bases = orc.g1_generate(0xA40, n)
    plain = orc.g1_msm(bases, pub)
    lead = W.multi_scalar_mul(bases, W.MpcVec.public(pub), True)
    other = W.multi_scalar_mul(bases, W.MpcVec.public(pub), False)
    assert _same(lead.sh, plain) and other.sh[1] == 1                            # from_public: leader holds it
    assert P.fq_from_mont_arr(other.sh[0].reshape(2, 6)) == [0, 1]
    parts = []
    for p in range(3):
        v = shares[p].copy()
        v[~tags] = pub[~tags]
        parts.append(W.multi_scalar_mul(bases, W.MpcVec(v, tags), p == 0).sh)
    acc = parts[0]
    for q in parts[1:]:
        acc = orc.g1_add(acc[0], q[0], acc[1], q[1])
    assert _same(acc, orc.g1_msm(bases, opened_in))
    sp = W.multi_scalar_mul(bases, W.MpcVec(v, tags, macs), False, spdz=True)     # SPDZ quirk: mac built from sh
    assert _same(sp.sh, sp.mac)
    empty = W.multi_scalar_mul(bases[:0], W.MpcVec.share(np.zeros((0, 4), dtype=np.uint64)), True)
    assert empty.sh[1] == 1


# ----------------------------------------------------------------------------- the whole prove sequence
def _run_parties(fn, n_parties):
    out, errs = [None] * n_parties, []

    def runner(p):
        try:
            out[p] = fn(p)
        except Exception as e:      # noqa: BLE001
            errs.append(e)
            raise

    ts = [threading.Thread(target=runner, args=(p,)) for p in range(n_parties)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    return out


@pytest.mark.parametrize("shape", ["run_groth16_zsh", "my_secret_input_circuit"])
@pytest.mark.parametrize("zk", [False, True])
def test_groth16_prove_sequence_three_parties(H, orc, pkg, shape, zk):
    """create_proof (src/groth16.rs:68-183) for 3 parties as 3 threads (LocalTestNet), additive shares, dummy
    triples: witness map from the assignment (SpMV on the device), the two opens as wire payloads, h resident into
    the h_query MSM, 4 G1 + 1 G2 MSMs.  The opened proof must equal the oracle's plain prover on the opened
    assignment, bit for bit.  Shapes: run_groth16.zsh's circuit (domain 8, MSMs < 32 points) and
    MySecretInputCircuit (6574 constraints + 5 inputs -> domain 2^13)."""
    G, S = pkg.groth16, pkg.synth
    if shape == "run_groth16_zsh":
        nc, ni, nv = 3, 2, 6
    else:
        nc, ni, nv = 6574, 5, 6600
    parties = 3
    mats = helpers.synth_r1cs(pkg, 0xB10 + nc, nc, nv)
    log_n = max(nc + ni - 1, 0).bit_length()
    pkarr = helpers.synth_proving_key(orc, 0xB20 + nc, nv, ni, 1 << log_n)
    # shares of the assignment: the constant 1 and the public inputs are Public -> leader holds them
    z_open = S.fr_uniform(0xB30, nv)
    z_open[0] = S.FR_R_LIMBS
    shares = [S.fr_uniform(0xB40 + p, nv) for p in range(parties)]
    rest = orc.open_sum(np.stack(shares[1:]))
    shares[0] = orc.vec_op("sub", z_open, rest)
    for p in range(1, parties):
        shares[p][:ni] = 0
    shares[0][:ni] = z_open[:ni]
    assert np.array_equal(orc.open_sum(np.stack(shares)), z_open)
    r, s = (S.fr_uniform(0xB50, 2) if zk else np.zeros((2, 4), dtype=np.uint64))
    nets = helpers.ThreadNet.make(parties)
    ready = threading.Barrier(parties)
    state = {}

    def party(p):
        H.set_party(p, parties)
        H.set_device(0)
        if p == 0:
            state["pk"] = G.ProvingKey(**pkarr)
            state["r1cs"] = G.R1CS(*mats, num_inputs=ni, num_vars=nv)
        ready.wait()
        return G.prove_party(state["pk"], state["r1cs"], shares[p], nets[p], r=r, s=s)

    proofs = _run_parties(party, parties)
    H.set_party(0, 3)
    exp = helpers.oracle_groth16(orc, pkarr, mats, ni, z_open, log_n, r, s)
    for key, add in (("a", orc.g1_add), ("b", orc.g2_add), ("c", orc.g1_add)):
        acc = proofs[0][key]
        for q in proofs[1:]:
            acc = add(acc[0], q[key][0], acc[1], q[key][1])
        assert _same(acc, exp[key]), key
    state["pk"].release()
    state["r1cs"].release()


# ----------------------------------------------------------------------------- Marlin-side compositions (a16)
class _SeqNet:
    """three parties played one after the other: exchange() replays payloads recorded in a first pass"""

    def __init__(self, party, store):
        self.party, self.n_parties, self.store, self.k = party, 3, store, 0

    def exchange(self, payload):
        slot = self.store.setdefault(self.k, {})
        slot[self.party] = payload
        self.k += 1
        if len(slot) < 3:
            raise _NeedOthers()
        return [slot[p] for p in range(3)]


class _NeedOthers(Exception):
    pass


def _shares_of(orc, pkg, seed, opened, parties=3):
    sh = [pkg.synth.fr_uniform(seed + p, len(opened)) for p in range(parties)]
    sh[0] = orc.vec_op("sub", opened, orc.open_sum(np.stack(sh[1:])))
    return sh


def _sum_points(orc, pts):
    acc = pts[0]
    for q in pts[1:]:
        acc = orc.g1_add(acc[0], q[0], acc[1], q[1])
    return acc


def test_kzg_commit_and_open_on_shares(H, orc, pkg):
    """KZG10::commit / open (poly-commit/src/kzg10/mod.rs:140-290) with shared coefficients, a shared blinding
    polynomial and a public point: the sum of the parties' results equals the plain computation on the oracle"""
    K, S = pkg.kzg, pkg.synth
    deg = 1500
    pg, pgg = orc.g1_generate(0xD10, deg + 1), orc.g1_generate(0xD11, 8)
    powers = K.Powers(pg, pgg)
    p_open, blind_open = S.fr_uniform(0xD20, deg + 1), S.fr_uniform(0xD21, 4)
    z = S.fr_uniform(0xD22, 1)[0]
    p_sh, b_sh = _shares_of(orc, pkg, 0xD30, p_open), _shares_of(orc, pkg, 0xD40, blind_open)
    # commit
    got = _sum_points(orc, [K.commit(powers, p_sh[i], b_sh[i]) for i in range(3)])
    c, rc = orc.g1_msm(pg, p_open, threads=8), orc.g1_msm(pgg[:4], blind_open)
    assert _same(got, orc.g1_add(c[0], rc[0], c[1], rc[1]))
    assert _same(_sum_points(orc, [K.commit(powers, p_sh[i]) for i in range(3)]), c)
    # open
    one = S.FR_R_LIMBS
    den = np.stack([orc.fr("neg", z[None])[0], one])
    wit, _ = orc.poly_div(p_open, den)
    rwit, _ = orc.poly_div(blind_open, den)
    w = orc.g1_msm(pg[:len(wit)], wit, threads=8)
    rw = orc.g1_msm(pgg[:len(rwit)], rwit)
    exp_w = orc.g1_add(w[0], rw[0], w[1], rw[1])
    res = [K.open(powers, p_sh[i], z, b_sh[i]) for i in range(3)]
    assert _same(_sum_points(orc, [r[0] for r in res]), exp_w)
    assert np.array_equal(orc.open_sum(np.stack([r[1][None] for r in res]))[0], orc.horner(blind_open, z))
    assert _same(_sum_points(orc, [K.open(powers, p_sh[i], z)[0] for i in range(3)]), w)
    powers.release()


def test_dense_polynomial_mul_on_shares(H, orc, pkg):
    """DensePolynomial::mul (dense.rs:567-583): public x shared is local; shared x shared is a Beaver batch product
    on the evaluations (two opens).  Lengths chosen so the domain is not tight (300 + 211 -> 512)."""
    K, S = pkg.kzg, pkg.synth
    a_open, b_open = S.fr_uniform(0xE10, 300), S.fr_uniform(0xE11, 211)
    n = 512

    def plain(a, b):
        pa, pb = np.zeros((n, 4), dtype=np.uint64), np.zeros((n, 4), dtype=np.uint64)
        pa[:len(a)], pb[:len(b)] = a, b
        return orc.ntt(orc.vec_op("mul", orc.ntt(pa, "fft"), orc.ntt(pb, "fft")), "ifft")

    exp = plain(a_open, b_open)
    assert not exp[510:].any()                                   # degree 509: the top coefficients are zero
    b_sh = _shares_of(orc, pkg, 0xE20, b_open)
    got = orc.open_sum(np.stack([K.poly_mul_public(a_open, b_sh[i]) for i in range(3)]))
    assert np.array_equal(got, exp)
    # shared x shared with a random (consistent) triple
    a_sh = _shares_of(orc, pkg, 0xE30, a_open)
    tx_o, ty_o = S.fr_uniform(0xE40, n), S.fr_uniform(0xE41, n)
    tz_o = orc.vec_op("mul", tx_o, ty_o)
    tx, ty, tz = (_shares_of(orc, pkg, 0xE50 + 8 * k, v) for k, v in enumerate((tx_o, ty_o, tz_o)))
    store, outs = {}, [None] * 3
    for rnd in range(3):                                         # each pass gets one exchange further
        for i in range(3):
            try:
                outs[i] = K.poly_mul_shared(a_sh[i], b_sh[i], _SeqNet(i, store), (tx[i], ty[i], tz[i]))
            except _NeedOthers:
                pass
    assert all(o is not None for o in outs)
    assert np.array_equal(orc.open_sum(np.stack(outs)), exp)
    # LC accumulation (dense.rs:345-372): acc += (f, other) with ragged lengths
    f = S.fr_uniform(0xE60, 1)[0]
    acc = K.add_assign_scaled(a_open, f, b_open[:100])
    exp_acc = a_open.copy()
    exp_acc[:100] = orc.vec_op("axpy", a_open[:100], b_open[:100], c=f)
    assert np.array_equal(acc, exp_acc)
    assert len(K.add_assign_scaled(b_open[:10], f, a_open)) == 300


def test_groth16_prove_sequence_spdz(H, orc, pkg):
    """the malicious backend (BASELINE config 5): [sh | mac] planes, MAC key 1 shared as (1, 0, 0), both opens MAC
    checked; the opened proof equals the plain prover's, the mac components repeat the sh ones (spdz.rs:482-488), and
    a corrupted MAC plane is caught by the check"""
    G, S = pkg.groth16, pkg.synth
    nc, ni, nv, parties = 200, 3, 230, 3
    mats = helpers.synth_r1cs(pkg, 0xF10, nc, nv)
    log_n = max(nc + ni - 1, 0).bit_length()
    pkarr = helpers.synth_proving_key(orc, 0xF20, nv, ni, 1 << log_n)
    z_open = S.fr_uniform(0xF30, nv)
    z_open[0] = S.FR_R_LIMBS
    sh = S.additive_shares(0xF40, z_open, parties, ni, lambda a, b: orc.vec_op("sub", a, b))
    mac = S.additive_shares(0xF50, z_open, parties, ni, lambda a, b: orc.vec_op("sub", a, b))     # key 1: macs open to z too
    zero = np.zeros(4, dtype=np.uint64)

    def run(mac_planes):
        nets = helpers.ThreadNet.make(parties)
        ready = threading.Barrier(parties)
        state = {}

        def party(p):
            H.set_party(p, parties)
            H.set_device(0)
            if p == 0:
                state["pk"] = G.ProvingKey(**pkarr)
                state["r1cs"] = G.R1CS(*mats, num_inputs=ni, num_vars=nv)
            ready.wait()
            sess = G.ProverSession(state["pk"], state["r1cs"])
            try:
                return sess.prove(np.stack([sh[p], mac_planes[p]]), nets[p], spdz=True)
            except H.MpcCudaError as e:
                return e
            finally:
                sess.close()

        out = _run_parties(party, parties)
        H.set_party(0, 3)
        state["pk"].release()
        state["r1cs"].release()
        return out

    proofs = run(mac)
    exp = helpers.oracle_groth16(orc, pkarr, mats, ni, z_open, log_n, zero, zero)
    for key, add in (("a", orc.g1_add), ("b", orc.g2_add), ("c", orc.g1_add)):
        acc = proofs[0][key]
        for q in proofs[1:]:
            acc = add(acc[0], q[key][0], acc[1], q[key][1])
        assert _same(acc, exp[key]), key
        assert all(_same(pr["mac"][key], pr[key]) for pr in proofs)
    bad = [m.copy() for m in mac]
    bad[1][ni + 7, 0] ^= np.uint64(1)                       # one party's MAC share of one witness value is off
    res = run(bad)
    assert all(isinstance(r, H.MpcCudaError) and "MAC check" in str(r) for r in res)


def test_poly_div_and_mul_by_vanishing(H, orc, pkg):
    """divide_by_vanishing_poly / mul_by_vanishing_poly on local share values (dense.rs:155-173 through
    univariate_div_qr): against the oracle's literal long division, for short divisors (long columns, chunked scan),
    divisors longer than the dividend, ragged lengths, and (q, r) recombined"""
    S = pkg.synth
    one = S.FR_R_LIMBS
    for n, m in ((1, 1), (5, 8), (8, 8), (9, 8), (1000, 1), (1000, 2), (4097, 16), (3 * 1024, 1024), (70001, 4), (5000, 4096),
                 (300000, 2)):
        p = S.fr_uniform(0x4A00 + n + m, n)
        den = np.zeros((m + 1, 4), dtype=np.uint64)
        den[m] = one
        den[0] = orc.fr("neg", one[None])[0]
        q, r = H.poly_div_vanishing(p, m)
        eq, er = orc.poly_div(p, den)
        assert q.shape == (max(n - m, 0), 4) and r.shape == (m, 4)
        assert np.array_equal(q[:len(eq)], eq) and not q[len(eq):].any(), (n, m)
        assert np.array_equal(r[:len(er)], er) and not r[len(er):].any(), (n, m)
        prod = H.poly_mul_vanishing(p, m)
        shifted = np.zeros((n + m, 4), dtype=np.uint64)
        shifted[m:] = p
        low = np.zeros((n + m, 4), dtype=np.uint64)
        low[:n] = p
        assert np.array_equal(prod, orc.vec_op("sub", shifted, low)), (n, m)
        if n > m:                                               # q v + r = p
            back = H.poly_mul_vanishing(q, m)
            back[:m] = orc.fr("add", back[:m], r)
            assert np.array_equal(back, p), (n, m)
    with pytest.raises(pkg._lib.MpcCudaError):
        pkg._lib.call("mpc_cuda_poly_div_vanishing", None, 4, 2, None, None)


@pytest.mark.parametrize("nc,ni", [(50, 2), (256, 4), (1000, 8), (64, 1)])
def test_marlin_rounds_three_parties(H, orc, pkg, nc, ni):
    """AHPForR1CS::prover_init / first / second round (arkworks/marlin/src/ahp/prover.rs:212-566) for 3 parties as 3
    threads on additive shares of the witness, the blinders and the mask polynomial, z_A * z_B through a Beaver batch
    product with wire-level opens.  Every oracle the prover would commit to, summed over the parties, must equal the
    plain prover's (oracle.marlin_rounds) coefficient for coefficient; the extra high coefficients the shared prover
    carries (it cannot truncate shared polynomials) must open to zero; t is public and identical at every party."""
    M, S = pkg.marlin, pkg.synth
    parties = 3
    mats, ints, x, w = helpers.synth_marlin_instance(pkg, orc, 0x5A00 + nc, nc, ni)
    nh = 1 << max(nc - 1, 0).bit_length()
    rnd = S.fr_uniform(0x5B00 + nc, 3 * nh + 16)
    blinders, alpha, etas, mask = rnd[:3], rnd[3], rnd[4:7], rnd[8:8 + 3 * nh]
    I = orc.fr_to_ints
    exp = orc.marlin_rounds(ints, nc, ni, I(x), I(w), I(blinders), I(mask), I(alpha[None])[0], I(etas))
    w_sh, bl_sh, mask_sh = (_shares_of(orc, pkg, 0x5C00 + 16 * k, v) for k, v in enumerate((w, blinders, mask)))
    tx_o, ty_o = S.fr_uniform(0x5D00, 4 * nh), S.fr_uniform(0x5D01, 4 * nh)
    tz_o = orc.vec_op("mul", tx_o, ty_o)
    tx, ty, tz = (_shares_of(orc, pkg, 0x5E00 + 16 * k, v) for k, v in enumerate((tx_o, ty_o, tz_o)))
    nets = helpers.ThreadNet.make(parties)
    ready = threading.Barrier(parties)
    state = {}

    def party(p):
        H.set_party(p, parties)
        H.set_device(0)
        if p == 0:
            state["index"] = M.Index(mats, nc, ni)
        ready.wait()
        index, leader = state["index"], p == 0
        z_a, z_b = M.prover_init(index, x, w_sh[p], leader)
        first = M.prover_first_round(index, x, w_sh[p], z_a, z_b, bl_sh[p], mask_sh[p], leader)
        second = M.prover_second_round(index, first, x, alpha, etas, nets[p], (tx[p], ty[p], tz[p]), leader)
        return dict(z_a_evals=z_a, z_b_evals=z_b, **first, **second)

    outs = _run_parties(party, parties)
    H.set_party(0, 3)
    state["index"].release()

    def opened(key):
        return I(orc.open_sum(np.stack([o[key] for o in outs])))

    def same_poly(got, want, what):
        assert got[:len(want)] == want, what
        assert not any(got[len(want):]), what + ": high coefficients must open to zero"

    same_poly(opened("z_a_evals"), exp["z_a"], "z_A evaluations")
    same_poly(opened("z_b_evals"), exp["z_b"], "z_B evaluations")
    same_poly(opened("w"), exp["w"], "w")
    same_poly(opened("z_a"), exp["z_a_poly"], "z_A")
    same_poly(opened("z_b"), exp["z_b_poly"], "z_B")
    same_poly(opened("mask"), exp["mask"], "mask")
    same_poly(opened("z_c"), exp["z_c"], "z_A z_B")
    same_poly(opened("g_1"), exp["g_1"], "g_1")
    same_poly(opened("h_1"), exp["h_1"], "h_1")
    for o in outs:
        assert I(o["t"]) == exp["t"]
        assert o["mul_domain"] == (8 * nh).bit_length() - 1     # untruncated shared lengths: 8|H| (plain prover: 4|H|)
    assert len(outs[0]["w"]) == nh + 1 - ni and len(outs[0]["z_c"]) == 4 * nh and len(outs[0]["h_1"]) == 7 * nh


# ----------------------------------------------------------------------------- the opens at the wire level
@pytest.mark.parametrize("spdz", [False, True])
def test_witness_map_resident_opens_match_host_route(H, orc, pkg, spdz):
    """masked=False route (_masked_payload / _open_payloads / _mac_payload / _mac_verify / finish with sx = oy = NULL /
    _assignment_dev) against the host-array route of the same calls: same payload bytes, same h, page-locked and
    pageable payload buffers alike; then the error cases"""
    S, G = pkg.synth, pkg.groth16
    nc, ni, nv, parties = 300, 3, 320, 3
    log_n = 9
    n = 1 << log_n
    mats = helpers.synth_r1cs(pkg, 0x6A0, nc, nv)
    r1cs = G.R1CS(*mats, num_inputs=ni, num_vars=nv)
    z_open = S.fr_uniform(0x6B0, nv)
    sub = lambda a, b: orc.vec_op("sub", a, b)
    sh = S.additive_shares(0x6C0, z_open, parties, ni, sub)
    mac = S.additive_shares(0x6D0, z_open, parties, ni, sub)
    assign = [np.stack([sh[p], mac[p]]) if spdz else sh[p] for p in range(parties)]
    trip = []
    for p in range(parties):
        v = np.tile(S.FR_R_LIMBS, (n, 1)) if p == 0 else np.zeros((n, 4), dtype=np.uint64)
        trip.append(np.stack([v, v]) if spdz else v)
    # host-array route
    begun = [H.witness_map_begin_r1cs(r1cs.A, r1cs.B, r1cs.C, assign[p], ni, log_n, trip[p], trip[p], spdz=spdz)
             for p in range(parties)]
    first = lambda v: v[0] if spdz else v
    sx = H.open_sum(np.stack([first(bg[0]) for bg in begun]))
    oy = H.open_sum(np.stack([first(bg[1]) for bg in begun]))
    h_host = [H.witness_map_finish(begun[p][2], trip[p], sx, oy, p == 0) for p in range(parties)]
    # resident route
    pinned = H.PinnedBuffer(2 * (8 + 32 * n))
    states = [H.witness_map_begin_r1cs(r1cs.A, r1cs.B, r1cs.C, assign[p], ni, log_n, trip[p], trip[p], spdz=spdz,
                                       masked=False) for p in range(parties)]
    assert all(st[0] is None and st[1] is None for st in states)
    states = [st[2] for st in states]
    for which in (0, 1):
        pays = [H.witness_map_masked_payload(states[p], which,
                                             out=pinned.array(np.uint8, 8 + 32 * n, which * (8 + 32 * n)) if p == 0 else None).copy()
                for p in range(parties)]
        for p in range(parties):
            assert np.array_equal(pays[p], H.fr_serialize(first(begun[p][which]))), (which, p)
        for p in range(parties):
            H.witness_map_open_payloads(states[p], which, pays[::-1] if p == 1 else pays)     # party order is irrelevant
        if spdz:
            dxs = [H.witness_map_mac_payload(states[p], which, p == 0) for p in range(parties)]
            val = sx if which == 0 else oy
            for p in range(parties):
                assert np.array_equal(dxs[p], H.fr_serialize(H.spdz_mac_check(val, begun[p][which][1], p == 0)))
                H.witness_map_mac_verify(states[p], dxs)
            bad = [d.copy() for d in dxs]
            bad[1][8] ^= 1
            with pytest.raises(H.MpcCudaError, match="MAC check"):
                H.witness_map_mac_verify(states[0], bad)
    for p in range(parties):
        ptr, cols = H.witness_map_assignment_dev(states[p])
        assert cols == nv
        back = np.empty((nv, 4), dtype=np.uint64)
        pkg._lib.call("mpc_cuda_memcpy_d2h", back.ctypes.data_as(C.c_void_p), C.c_void_p(ptr), C.c_size_t(nv * 32), None)
        pkg._lib.call("mpc_cuda_stream_sync", None)
        assert np.array_equal(back, sh[p])
        h_ptr = H.witness_map_finish_dev(states[p], trip[p], None, None, p == 0)
        planes = 2 if spdz else 1
        h = np.empty((planes * n, 4), dtype=np.uint64)
        pkg._lib.call("mpc_cuda_memcpy_d2h", h.ctypes.data_as(C.c_void_p), C.c_void_p(h_ptr), C.c_size_t(h.nbytes), None)
        pkg._lib.call("mpc_cuda_stream_sync", None)
        assert np.array_equal(h.reshape(h_host[p].shape), h_host[p]), p
        states[p].release()
    # error cases: finish without the opens, MAC payload on an additive state / before the open, malformed payloads
    st = H.witness_map_begin_r1cs(r1cs.A, r1cs.B, r1cs.C, assign[0], ni, log_n, trip[0], trip[0], spdz=spdz, masked=False)[2]
    with pytest.raises(H.MpcCudaError):
        H.witness_map_finish_dev(st, trip[0], None, None, True)
    with pytest.raises(H.MpcCudaError):
        H.witness_map_mac_payload(st, 0, True)
    pay = H.witness_map_masked_payload(st, 0)
    wrong_len = pay.copy()
    wrong_len[0] ^= 1
    with pytest.raises(H.MpcCudaError, match="length prefix"):
        H.witness_map_open_payloads(st, 0, [pay, wrong_len])
    not_reduced = pay.copy()
    not_reduced[8:40] = 0xFF
    with pytest.raises(H.MpcCudaError, match="modulus"):
        H.witness_map_open_payloads(st, 0, [not_reduced])
    with pytest.raises(H.MpcCudaError):
        pkg._lib.call("mpc_cuda_witness_map_masked_payload", C.c_uint64(st.state), C.c_uint32(2), pay.ctypes.data_as(C.c_void_p))
    H.witness_map_open_payloads(st, 0, [pay])
    H.witness_map_open_payloads(st, 1, [H.witness_map_masked_payload(st, 1)])
    H.witness_map_finish_dev(st, trip[0], None, None, True)
    with pytest.raises(H.MpcCudaError):                         # already finished
        H.witness_map_open_payloads(st, 0, [pay])
    st.release()
    with pytest.raises(H.MpcCudaError):
        H.witness_map_masked_payload(st, 0)
    pinned.free()
    r1cs.release()


@pytest.mark.parametrize("nc,ni", [(50, 2), (300, 4), (64, 1)])
def test_marlin_resident_prover_matches_host_route(H, orc, pkg, nc, ni):
    """marlin.ResidentProver (every vector on the device, only the two Beaver payloads cross PCIe) against the
    host-array functions, share by share and bit for bit, for 3 party threads; the commitments over powers_of_g summed
    over the parties equal the plain commitments of the opened oracles"""
    M, S, K = pkg.marlin, pkg.synth, pkg.kzg
    parties = 3
    mats, ints, x, w = helpers.synth_marlin_instance(pkg, orc, 0x7A00 + nc, nc, ni)
    nh = 1 << max(nc - 1, 0).bit_length()
    rnd = S.fr_uniform(0x7B00 + nc, 3 * nh + 16)
    blinders, alpha, etas, mask = rnd[:3], rnd[3], rnd[4:7], rnd[8:8 + 3 * nh]
    w_sh, bl_sh, mask_sh = (_shares_of(orc, pkg, 0x7C00 + 16 * k, v) for k, v in enumerate((w, blinders, mask)))
    tx_o, ty_o = S.fr_uniform(0x7D00, 4 * nh), S.fr_uniform(0x7D01, 4 * nh)
    tz_o = orc.vec_op("mul", tx_o, ty_o)
    tx, ty, tz = (_shares_of(orc, pkg, 0x7E00 + 16 * k, v) for k, v in enumerate((tx_o, ty_o, tz_o)))
    pg = orc.g1_generate(0x7F0, 7 * nh)
    nets = helpers.ThreadNet.make(parties)
    ready = threading.Barrier(parties)
    state = {}

    def party(p):
        H.set_party(p, parties)
        H.set_device(0)
        if p == 0:
            state["index"] = M.Index(mats, nc, ni)
            state["powers"] = K.Powers(pg, pg[:2])
        ready.wait()
        index, leader = state["index"], p == 0
        z_a, z_b = M.prover_init(index, x, w_sh[p], leader)
        first = M.prover_first_round(index, x, w_sh[p], z_a, z_b, bl_sh[p], mask_sh[p], leader)
        second = M.prover_second_round(index, first, x, alpha, etas, nets[p], (tx[p], ty[p], tz[p]), leader)
        rp = M.ResidentProver(index, parties)
        res = [rp.rounds(x, w_sh[p], bl_sh[p], mask_sh[p], alpha, etas, nets[p], (tx[p], ty[p], tz[p]), leader,
                         powers=state["powers"]) for _ in range(2)]          # twice: the buffers are reused
        # the opening phase on the resident oracles against the host-array route: LC, witness, evaluation, commitment
        terms = [("w", etas[0]), ("z_b", etas[1]), ("h_1", etas[2]), ("g_1", alpha), ("t", blinders[0])]
        opened = rp.open_combination(terms, alpha, state["powers"])
        host_polys = dict(**first, **second)
        lc = np.zeros((0, 4), dtype=np.uint64)
        for name, coeff in terms:
            lc = K.add_assign_scaled(lc, coeff, host_polys[name])
        host_open = K.open(state["powers"], lc, alpha)
        rp.close()
        return dict(host=host_polys, resident=res, opened=opened, host_open=(host_open[0], H.poly_evaluate(lc, alpha)), lc=lc)

    outs = _run_parties(party, parties)
    H.set_party(0, 3)
    for o in outs:
        for res in o["resident"]:
            for key in ("w", "z_a", "z_b", "mask", "z_c", "t", "g_1", "h_1"):
                assert np.array_equal(res[key], o["host"][key]), key
            assert res["mul_domain"] == o["host"]["mul_domain"]
    for key, length in (("w", nh + 1 - ni), ("h_1", 7 * nh), ("g_1", nh - 1), ("mask", 3 * nh)):
        opened = orc.open_sum(np.stack([o["host"][key] for o in outs]))
        assert len(opened) == length
        exp = orc.g1_msm(pg[:length], opened, threads=8)
        got = _sum_points(orc, [o["resident"][1]["commitments"][key] for o in outs])
        assert _same(got, exp), key
    for o in outs:                                               # resident opening = host-array opening, per share
        assert _same(o["opened"][0], o["host_open"][0]) and np.array_equal(o["opened"][1], o["host_open"][1])
    lc_open = orc.open_sum(np.stack([o["lc"] for o in outs]))
    assert np.array_equal(orc.open_sum(np.stack([o["opened"][1][None] for o in outs]))[0], orc.horner(lc_open, alpha))
    t_len = len(outs[0]["host"]["t"])
    assert _same(outs[1]["resident"][0]["commitments"]["t"], orc.g1_msm(pg[:t_len], outs[0]["host"]["t"], threads=8))
    state["index"].release()
    state["powers"].release()
