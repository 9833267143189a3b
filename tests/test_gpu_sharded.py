"""GPU parity for the paths bench.py and the multi-GPU entry points run, all on ONE GPU so the driver's
single-GPU box executes them: Jacobian partials + mpc_cuda_g1_sum_partials_dev, the in-library sharded MSM
(parts placed round robin), the cross-device NTT stages with the exchange emulated by a local transpose, the
in-library sharded NTT with g virtual devices on one GPU, the c = 23 table at 2^22, G2 at 2^16 / 2^20, NTT 2^22
of every kind, and three party threads calling concurrently.  Bit-exact against the oracle."""
import threading

import numpy as np
import pytest

import helpers
import pyref as P

pytestmark = pytest.mark.gpu
KINDS = ["fft", "ifft", "coset_fft", "coset_ifft"]


@pytest.fixture(scope="module")
def H(pkg):
    pkg.host.init()
    pkg.host.set_party(0, 3)
    return pkg.host


@pytest.fixture(autouse=True)
def _reset_options(H):
    yield
    for name in ("msm_window_bits", "msm_task_len", "msm_host_chunks", "msm_affine", "msm_affine_split", "msm_reduce_chunk", "msm_reduce_warp_max", "ntt_graph"):
        H.set_option(name, 0)


def _same(a, b):
    return np.array_equal(a[0], b[0]) and a[1] == b[1]


@pytest.fixture(scope="module")
def bases8k(orc):
    return orc.g1_generate(0xC0FFEE, 8192)


# ----------------------------------------------------------------------------- MSM partials (bench.py's step)
@pytest.mark.parametrize("k", [1, 3, 8])
@pytest.mark.parametrize("table", [False, True])
def test_g1_partials_sum_to_the_whole_msm(H, orc, pkg, bases8k, k, table):
    """mpc_cuda_msm_g1_handle_dev per point range (Jacobian partial, k_emit mode 1) + mpc_cuda_g1_sum_partials_dev
    == the oracle's MSM of the whole vector; one range carries only zero scalars (infinity partial)."""
    n = 8000
    sc = pkg.synth.fr_uniform(0x2A0 + k, n)
    cuts = [n * i // k for i in range(k + 1)]
    if k >= 3:
        sc[cuts[1]:cuts[2]] = 0                              # this range's partial is the point at infinity
    h = H.register_bases(bases8k[:n])
    if table:
        h.precompute(0)
    sbuf = H.DeviceBuffer(n * 32).upload(sc)
    parts = H.DeviceBuffer(k * 18 * 8)
    for i in range(k):
        view = H.DeviceBuffer.__new__(H.DeviceBuffer)
        view.ptr = type(parts.ptr)(parts.ptr.value + i * 18 * 8)
        view.nbytes = 18 * 8
        H.msm_handle_dev(h, sbuf, cuts[i + 1] - cuts[i], offset=cuts[i], scalar_offset=cuts[i], out=view)
    got = H.sum_partials(parts, k)
    assert _same(got, orc.g1_msm(bases8k[:n], sc, threads=8))
    if k >= 3:
        jac = parts.download().reshape(k, 18)
        assert not jac[1, 12:].any()                         # z = 0 for the infinity partial
    # zero partials -> affine zero
    out, inf = H.sum_partials(parts, 0)
    assert inf == 1 and P.fq_from_mont_arr(out.reshape(2, 6)) == [0, 1]
    h.release(); sbuf.free(); parts.free()


@pytest.mark.parametrize("parts", [1, 2, 5])
def test_g1_sharded_handle_in_library(H, orc, pkg, bases8k, parts):
    """mpc_cuda_msm_g1_register_bases_sharded + mpc_cuda_msm_g1_handle: one msm call, the library drives the
    parts (here round robin on the visible devices), gathers the Jacobian partials by peer copies and adds them"""
    n = 8191
    infs = np.zeros(n, dtype=np.uint8)
    infs[[0, 4097, n - 1]] = 1
    h = H.register_bases(bases8k[:n], inf=infs, parts=parts)
    sc = pkg.synth.fr_witness_like(0x2B0 + parts, n)
    assert _same(H.msm_handle(h, sc), orc.g1_msm(bases8k[:n], sc, inf=infs, threads=8))
    # sub-range crossing part boundaries
    assert _same(H.msm_handle(h, sc[:5000], offset=1500), orc.g1_msm(bases8k[1500:6500], sc[:5000], inf=infs[1500:6500], threads=8))
    h.precompute(9)
    assert _same(H.msm_handle(h, sc), orc.g1_msm(bases8k[:n], sc, inf=infs, threads=8))
    h.release()


def test_g2_sharded_handle_in_library(H, orc, pkg):
    g, _ = P.g2_to_arr((P.G2_X, P.G2_Y))
    bases = orc.g2_generate(g, 0xB3, 300)
    h = H.register_bases(bases, g2=True, parts=3)
    sc = pkg.synth.fr_uniform(0x2C0, 300)
    assert _same(H.msm_handle(h, sc), orc.g2_msm(bases, sc, threads=8))
    h.release()


@pytest.mark.parametrize("chunks", [1, 2, 7, 16])
def test_host_msm_streamed_in_chunks(H, orc, pkg, bases8k, chunks):
    """mpc_cuda_msm_g1 streams the points as chunks that add into one bucket set (copy/compute overlap)"""
    n = 8000
    infs = np.zeros(n, dtype=np.uint8)
    infs[[7, 3999, 4000]] = 1
    H.set_option("msm_host_chunks", chunks)
    for sc in (pkg.synth.fr_uniform(0x2D0 + chunks, n), pkg.synth.fr_witness_like(0x2E0 + chunks, n)):
        assert _same(H.msm_g1(bases8k[:n], sc, inf=infs), orc.g1_msm(bases8k[:n], sc, inf=infs, threads=8))
    H.set_option("msm_task_len", 3)                            # split buckets merge with the earlier chunks' sums
    ones = np.tile(pkg.synth.FR_R_LIMBS, (n, 1))
    assert _same(H.msm_g1(bases8k[:n], ones), orc.g1_msm(bases8k[:n], ones, threads=8))


def test_g1_table_c23_at_2_22_exact(H, orc, pkg):
    """the widest table path (c = 23, 12 windows in one bucket set of 2^22 buckets) through the
    sum s_i (k_i G) = (sum s_i k_i mod r) G identity, resident scalars and the Jacobian partial route"""
    log_n = 22
    n = 1 << log_n
    seed = pkg.synth.bench_seed(log_n)
    dev = H.g1_generate(seed, n)
    h = H.register_bases_dev(dev, n).precompute(23)
    ks = helpers.gen_ks(pkg, seed, n)
    sc = pkg.synth.fr_uniform(seed, n)
    sbuf = H.DeviceBuffer(n * 32).upload(sc)
    part = H.msm_handle_dev(h, sbuf, n)
    exp = helpers.expected_msm_of_generated(orc, sc, ks)
    assert _same(H.sum_partials(part, 1), exp)
    assert _same(H.msm_handle(h, sc), exp)
    h.release(); dev.free(); sbuf.free(); part.free()


# ----------------------------------------------------------------------------- G2 at the sizes SURVEY 8d names
def test_g2_msm_2_16_matches_oracle(H, orc, pkg):
    n = 1 << 16
    dev = H.g2_generate(0x5EED0110, n)
    bases = dev.download().reshape(n, 24)
    g, _ = P.g2_to_arr((P.G2_X, P.G2_Y))
    assert np.array_equal(bases[:16], orc.g2_generate(g, 0x5EED0110, 16))
    sc = pkg.synth.fr_uniform(0x5EED0110, n)
    assert _same(H.msm_g2(bases, sc), orc.g2_msm(bases, sc, threads=16))
    h = H.register_bases_dev(dev, n, g2=True)
    assert _same(H.msm_handle(h, sc), orc.g2_msm(bases, sc, threads=16))
    h.release(); dev.free()


def test_g2_msm_2_20_exact(H, orc, pkg):
    log_n = 20
    n = 1 << log_n
    seed = pkg.synth.bench_seed(log_n) + 0x200
    dev = H.g2_generate(seed, n)
    h = H.register_bases_dev(dev, n, g2=True)
    ks = helpers.gen_ks(pkg, seed, n)
    g, _ = P.g2_to_arr((P.G2_X, P.G2_Y))
    for sc in (pkg.synth.fr_uniform(seed, n), pkg.synth.fr_witness_like(seed + 1, n)):
        e = helpers.dot_mod_r(sc, ks)
        exp = orc.g2_scalar_mul(g, np.array(P.to_limbs(e, 4), dtype=np.uint64))
        assert _same(H.msm_handle(h, sc), exp)
    h.release(); dev.free()


# ----------------------------------------------------------------------------- NTT: cross-device stages on one GPU
def _bitrev(x, bits):
    return int(format(x, "0%db" % bits)[::-1], 2) if bits else 0


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("log_g", [1, 2, 3])
def test_cross_stage_with_emulated_all_to_all(H, orc, pkg, log_g, kind):
    """sharding.dist_ntt's schedule for g devices played on one GPU: the two all-to-alls are numpy transposes,
    mpc_cuda_ntt_cross_stage_dev (k_ntt_cross<LOG_G>) and the local transforms run on the device."""
    g = 1 << log_g
    for log_n in (2 * log_g, 9, 13):
        n = 1 << log_n
        m, sl = n // g, n // g // g
        full = pkg.synth.fr_uniform(0xA00 + log_n, n)
        inverse = kind in ("ifft", "coset_ifft")
        expect = orc.ntt(full, kind)
        if inverse:        # the inverse kinds take the transposed order the forward kinds produce
            blocks = [full[np.arange(m) * g + _bitrev(r, log_g)].copy() for r in range(g)]
        else:
            blocks = [full[r * m:(r + 1) * m].copy() for r in range(g)]
        bufs = [H.DeviceBuffer(m * 32).upload(b) for b in blocks]
        if inverse:
            for b in bufs:
                H.ntt_dev(b.ptr.value, log_n - log_g, "ifft")
            blocks = [b.download().reshape(m, 4) for b in bufs]
        # all-to-all: device r receives slice r of every block
        gathered = [np.stack([blocks[q][r * sl:(r + 1) * sl] for q in range(g)]) for r in range(g)]
        gb = [H.DeviceBuffer(g * sl * 32).upload(x) for x in gathered]
        for r in range(g):
            H.ntt_cross_stage_dev(gb[r].ptr.value, log_n, log_g, r * sl, sl, kind)
        gathered = [b.download().reshape(g, sl, 4) for b in gb]
        blocks = [np.concatenate([gathered[r][q] for r in range(g)]) for q in range(g)]      # all-to-all back
        if not inverse:
            for q in range(g):
                bufs[q].upload(blocks[q])
                H.ntt_dev(bufs[q].ptr.value, log_n - log_g, "fft")
            blocks = [b.download().reshape(m, 4) for b in bufs]
            for r in range(g):
                assert np.array_equal(blocks[r], expect[np.arange(m) * g + _bitrev(r, log_g)]), (log_n, r)
        else:
            assert np.array_equal(np.concatenate(blocks), expect), log_n
        for b in bufs + gb:
            b.free()


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("log_g", [1, 2, 3])
def test_sharded_ntt_in_library_virtual_devices(H, orc, pkg, log_g, kind):
    """mpc_cuda_ntt_fr_sharded_dev with all g blocks on one GPU: the fused cross kernel reads and writes the
    blocks through their own pointers (the peer loads / stores of the multi-GPU run) — no exchange step at all."""
    g = 1 << log_g
    for log_n in (2 * log_g, 10, 16):
        n = 1 << log_n
        m = n // g
        full = pkg.synth.fr_uniform(0xB00 + log_n, n)
        expect = orc.ntt(full, kind)
        inverse = kind in ("ifft", "coset_ifft")
        if inverse:
            blocks = [full[np.arange(m) * g + _bitrev(r, log_g)].copy() for r in range(g)]
        else:
            blocks = [full[r * m:(r + 1) * m].copy() for r in range(g)]
        bufs = [H.DeviceBuffer(m * 32).upload(b) for b in blocks]
        H.ntt_sharded_dev([b.ptr.value for b in bufs], log_n, kind, dev_index=[0] * g)
        out = [b.download().reshape(m, 4) for b in bufs]
        if inverse:
            assert np.array_equal(np.concatenate(out), expect), log_n
        else:
            for r in range(g):
                assert np.array_equal(out[r], expect[np.arange(m) * g + _bitrev(r, log_g)]), (log_n, r)
            # reorder to natural block order and back
            nat = [H.DeviceBuffer(m * 32) for _ in range(g)]
            H.ntt_reorder_sharded_dev([b.ptr.value for b in bufs], [b.ptr.value for b in nat], log_n, False, dev_index=[0] * g)
            assert np.array_equal(np.concatenate([b.download().reshape(m, 4) for b in nat]), expect)
            H.ntt_reorder_sharded_dev([b.ptr.value for b in nat], [b.ptr.value for b in bufs], log_n, True, dev_index=[0] * g)
            for r in range(g):
                assert np.array_equal(bufs[r].download().reshape(m, 4), out[r])
            for b in nat:
                b.free()
        for b in bufs:
            b.free()


@pytest.mark.parametrize("kind", KINDS)
def test_ntt_2_22_every_kind_matches_oracle(H, orc, pkg, kind):
    v = pkg.synth.fr_uniform(0x422, 1 << 22)
    assert np.array_equal(H.ntt(v, kind), orc.ntt(v, kind))


# ----------------------------------------------------------------------------- several party threads, one device
def test_three_party_threads_concurrently(H, orc, pkg, bases8k):
    """LocalTestNet runs the parties as threads of one process (mpc-net/src/multi.rs:436-441): three threads call
    witness_map + MSM + NTTs of growing coset sizes at the same time on one device; every result must equal the
    serial run bit for bit (per-thread streams, shared twiddle / base registries)."""
    S = pkg.synth
    n = 1 << 12
    h = H.register_bases(bases8k).precompute(0)
    plain = H.register_bases(bases8k)
    one, zero = np.tile(S.FR_R_LIMBS, (n, 1)), np.zeros((n, 4), dtype=np.uint64)

    def work(party, rounds=3):
        H.set_party(party, 3)
        H.set_device(0)                        # the registered vectors live on device 0 (LocalTestNet: one GPU)
        out = []
        for it in range(rounds):
            a, b, c = (S.fr_uniform(0xC00 + 16 * party + 3 * it + k, n) for k in range(3))
            tx = one if party == 0 else zero
            ma, mb, st = H.witness_map_begin(a, b, c, tx, tx)
            hh = H.witness_map_finish(st, tx, ma, mb, party == 0)
            sc = S.fr_uniform(0xD00 + party + 7 * it, 8192)
            out += [ma, mb, hh, H.msm_handle(h, sc)[0], H.msm_handle(plain, sc[:5000], offset=100)[0],
                    H.msm_g1(bases8k[:3000], sc[:3000])[0]]
            for log_n in (13 + party, 15 + it, 17):                      # growing coset tables while others run
                v = S.fr_uniform(0xE00 + log_n + party, 1 << log_n)
                out.append(H.ntt(H.ntt(v, "coset_fft"), "coset_ifft")[:64].copy())
                out.append(H.ntt(v, "coset_fft")[-64:].copy())
        return out

    serial = [work(p) for p in range(3)]
    got, errs = [None] * 3, []

    def runner(p):
        try:
            got[p] = work(p)
        except Exception as e:          # noqa: BLE001
            errs.append(e)

    threads = [threading.Thread(target=runner, args=(p,)) for p in range(3)]
    for t in threads:
        t.start()
    # table replacement and release by a fourth thread while the parties are launching
    tmp = H.register_bases(bases8k[:2048])
    tmp.precompute(5)
    tmp.precompute(7)
    tmp.release()
    for t in threads:
        t.join()
    H.set_party(0, 3)
    assert not errs, errs
    for p in range(3):
        assert len(got[p]) == len(serial[p])
        for x, y in zip(got[p], serial[p]):
            assert np.array_equal(x, y), p
    h.release(); plain.release()


def test_init_rejects_a_different_device_list(H):
    H.init()                                       # same (implicit) list: fine
    with pytest.raises(H.MpcCudaError):
        H.init([0, 0])                             # duplicates are never valid
    n = H.device_count()
    if n >= 2:
        with pytest.raises(H.MpcCudaError):
            H.init([1])                            # initialised with all devices: a different list is refused


def test_euclid_inverse_matches_fermat(H, orc, pkg):
    fr = np.concatenate([P.fr_to_mont_arr([0, 1, 2, P.R_MOD - 1, (P.R_MOD - 1) // 2]), pkg.synth.fr_uniform(0xF10, 500)])
    assert np.array_equal(H.field_op("fr", "inv_euclid", fr), orc.fr("inv", fr))
    rng = np.random.default_rng(5)
    fq = P.fq_to_mont_arr([0, 1, 2, P.Q_MOD - 1] + [int.from_bytes(rng.bytes(48), "little") % P.Q_MOD for _ in range(300)])
    assert np.array_equal(H.field_op("fq", "inv_euclid", fq), orc.fq("inv", fq))


# ----------------------------------------------------------------------------- real multi-GPU (skipped on 1 GPU)
@pytest.mark.parametrize("log_g", [1, 2, 3])
def test_sharded_ntt_and_msm_on_real_devices(H, orc, pkg, bases8k, log_g):
    g = 1 << log_g
    if H.device_count() < g:
        pytest.skip("needs %d GPUs in this process" % g)
    log_n = 16
    v = pkg.synth.fr_uniform(0xF20 + log_g, 1 << log_n)
    for kind in KINDS:
        assert np.array_equal(H.ntt_sharded(v, kind, log_g), orc.ntt(v, kind)), kind
    h = H.register_bases(bases8k, parts=g).precompute(0)
    sc = pkg.synth.fr_uniform(0xF30, 8192)
    assert _same(H.msm_handle(h, sc), orc.g1_msm(bases8k, sc, threads=8))
    h.release()


# ----------------------------------------------------------------------------- batched-affine pre-reduction
@pytest.mark.parametrize("split", [2, 1])                # option msm_affine_split: 2 = fused kernel, 1 = two kernels / two streams
@pytest.mark.parametrize("rounds_opt", [2, 3])          # option msm_affine: 2 = one round, 3 = two rounds
def test_g1_affine_prereduction_matches_oracle(H, orc, pkg, bases8k, rounds_opt, split):
    """k_affine_pairs (pairwise affine additions with one shared inversion per warp) in front of the XYZZ
    accumulation: forced on for small inputs, every window width class, duplicates (tangent), P + (-P) (infinity),
    infinity bases, skewed scalars (one huge bucket), table mode, chunked streaming"""
    H.set_option("msm_affine", rounds_opt)
    H.set_option("msm_affine_split", split)
    try:
        n = 8000
        for seed, sc in ((1, pkg.synth.fr_uniform(0x3A0, n)), (2, pkg.synth.fr_witness_like(0x3B0, n))):
            assert _same(H.msm_g1(bases8k[:n], sc), orc.g1_msm(bases8k[:n], sc, threads=8)), seed
        for c in (4, 9, 13):
            H.set_option("msm_window_bits", c)
            sc = pkg.synth.fr_uniform(0x3C0 + c, 3000)
            assert _same(H.msm_g1(bases8k[:3000], sc), orc.g1_msm(bases8k[:3000], sc, threads=8)), c
        H.set_option("msm_window_bits", 0)
        # the same point many times with the same scalar (tangent case in every pair), P and -P, infinity bases
        bases, scalars = bases8k[:64].copy(), pkg.synth.fr_uniform(0x3D0, 64)
        bases[1:9] = bases[0]
        scalars[1:9] = scalars[0]
        q = P.g1_from_arr(bases[20])
        bases[21] = P.g1_to_arr((q[0], (-q[1]) % P.Q_MOD))[0]
        scalars[21] = scalars[20]
        infs = np.zeros(64, dtype=np.uint8)
        infs[[5, 40]] = 1
        assert _same(H.msm_g1(bases, scalars, inf=infs), orc.g1_msm(bases, scalars, inf=infs))
        same_b, same_s = np.tile(bases[7], (57, 1)), np.tile(scalars[7], (57, 1))
        assert _same(H.msm_g1(same_b, same_s), orc.g1_msm(same_b, same_s))
        assert H.msm_g1(bases[20:22], scalars[20:22])[1] == 1
        ones = np.tile(pkg.synth.FR_R_LIMBS, (n, 1))
        H.set_option("msm_task_len", 5)
        assert _same(H.msm_g1(bases8k[:n], ones), orc.g1_msm(bases8k[:n], ones, threads=8))
        H.set_option("msm_task_len", 0)
        # table mode and chunked streaming (later chunks merge into the buckets of the earlier ones)
        h = H.register_bases(bases8k).precompute(9)
        sc = pkg.synth.fr_uniform(0x3E0, 8192)
        assert _same(H.msm_handle(h, sc), orc.g1_msm(bases8k, sc, threads=8))
        h.release()
        H.set_option("msm_host_chunks", 3)
        assert _same(H.msm_g1(bases8k[:n], sc[:n]), orc.g1_msm(bases8k[:n], sc[:n], threads=8))
    finally:
        H.set_option("msm_affine", 0)
        H.set_option("msm_affine_split", 0)


def test_g2_affine_prereduction_matches_oracle(H, orc, pkg):
    g, _ = P.g2_to_arr((P.G2_X, P.G2_Y))
    bases = orc.g2_generate(g, 0xB4, 500)
    bases[1:5] = bases[0]
    H.set_option("msm_affine", 3)
    try:
        for split in (2, 1):
            H.set_option("msm_affine_split", split)
            for sc in (pkg.synth.fr_uniform(0x3F0, 500), pkg.synth.fr_witness_like(0x3F1, 500)):
                sc[1:5] = sc[0]
                assert _same(H.msm_g2(bases, sc), orc.g2_msm(bases, sc, threads=8))
    finally:
        H.set_option("msm_affine", 0)
        H.set_option("msm_affine_split", 0)


def test_g1_affine_prereduction_full_size_exact(H, orc, pkg):
    """2^20 points, where the automatic rule turns the pre-reduction on (plain and table mode)"""
    log_n = 20
    n = 1 << log_n
    seed = pkg.synth.bench_seed(log_n) + 0x300
    dev = H.g1_generate(seed, n)
    h = H.register_bases_dev(dev, n)
    ks = helpers.gen_ks(pkg, seed, n)
    sc = pkg.synth.fr_uniform(seed, n)
    exp = helpers.expected_msm_of_generated(orc, sc, ks)
    assert _same(H.msm_handle(h, sc), exp)
    H.set_option("msm_affine", 1)
    assert _same(H.msm_handle(h, sc), exp)                 # and the same without it
    H.set_option("msm_affine", 0)
    h.precompute(0)
    assert _same(H.msm_handle(h, sc), exp)
    wl = pkg.synth.fr_witness_like(seed + 1, n)
    assert _same(H.msm_handle(h, wl), helpers.expected_msm_of_generated(orc, wl, ks))
    h.release(); dev.free()


@pytest.mark.parametrize("log_g", [1, 3])
def test_sharded_ntt_replayed_from_a_cuda_graph(H, orc, pkg, log_g):
    """repeated sharded transforms over the same blocks: first call direct, second captured into a CUDA graph,
    later ones replayed — every call bit-exact on fresh data (virtual devices, option ntt_graph = 1)"""
    g = 1 << log_g
    H.set_option("ntt_graph", 1)
    try:
        for log_n in (10, 17):               # one local pass / several local passes (ping-pong buffer in the graph)
            n = 1 << log_n
            m = n // g
            bufs = [H.DeviceBuffer(m * 32) for _ in range(g)]
            for kind in ("fft", "coset_ifft"):
                inverse = kind == "coset_ifft"
                for rep in range(4):
                    full = pkg.synth.fr_uniform(0xC00 + 16 * log_n + rep, n)
                    expect = orc.ntt(full, kind)
                    for r in range(g):
                        bufs[r].upload(full[np.arange(m) * g + _bitrev(r, log_g)] if inverse else full[r * m:(r + 1) * m])
                    H.ntt_sharded_dev([b.ptr.value for b in bufs], log_n, kind, dev_index=[0] * g)
                    out = [b.download().reshape(m, 4) for b in bufs]
                    if inverse:
                        assert np.array_equal(np.concatenate(out), expect), (log_n, kind, rep)
                    else:
                        for r in range(g):
                            assert np.array_equal(out[r], expect[np.arange(m) * g + _bitrev(r, log_g)]), (log_n, kind, rep, r)
            for b in bufs:
                b.free()
    finally:
        H.set_option("ntt_graph", 0)
