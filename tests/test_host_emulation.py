"""CPU-only checks of the device templates (fp.cuh, fp2.cuh, ec.cuh, msm_digits.cuh) compiled with
g++ through the carry-flag emulation in ptx.cuh, against the oracle.  This pins the *logic* of the
CUDA field / curve code on a box without a GPU; the same comparisons run on the device in the
-m gpu tests.  Bit-exact."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import pyref as P

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "zk-mpc_b200", "csrc")
EMU_SRC = os.path.join(HERE, "emu", "host_emu.cpp")
EMU_LIB = os.path.join(HERE, "emu", "libhost_emu.so")

u32p = C.POINTER(C.c_uint32)
u8p = C.POINTER(C.c_uint8)


@pytest.fixture(scope="module")
def emu():
    deps = [EMU_SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    if not os.path.exists(EMU_LIB) or os.path.getmtime(EMU_LIB) < max(os.path.getmtime(d) for d in deps):
        gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([gxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-I", CSRC, "-o", EMU_LIB, EMU_SRC],
                       check=True)
    return C.CDLL(EMU_LIB)


def _p32(a):
    return a.ctypes.data_as(u32p)


def _vec(emu, name, op, a, b=None):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    out = np.empty_like(a)
    limbs = a.shape[-1]
    getattr(emu, name)(C.c_int(op), _p32(a), _p32(np.ascontiguousarray(b, dtype=np.uint64)) if b is not None else None,
                       _p32(out), C.c_size_t(a.size // limbs))
    return out


OPS = {"add": 0, "sub": 1, "mul": 2, "neg": 3, "inv": 4, "from_mont": 5, "to_mont": 6, "sqr": 7}


def _edge_fr(pkg, n, seed):
    edge = P.fr_to_mont_arr([0, 1, 2, P.R_MOD - 1, P.R_MOD - 2, (P.R_MOD - 1) // 2, (1 << 64) - 1, 1 << 64])
    return np.concatenate([edge, pkg.synth.fr_uniform(seed, n)])


def _edge_fq(n, seed):
    rng = np.random.default_rng(seed)
    vals = [0, 1, 2, P.Q_MOD - 1, P.Q_MOD - 2, (P.Q_MOD - 1) // 2, (1 << 64) - 1, 1 << 64]
    vals += [int.from_bytes(rng.bytes(48), "little") % P.Q_MOD for _ in range(n)]
    return P.fq_to_mont_arr(vals)


def test_emulated_fr_fq(emu, orc, pkg):
    for name, ofn, a, b in (("emu_fr_vec", orc.fr, _edge_fr(pkg, 500, 1), _edge_fr(pkg, 500, 2)[::-1].copy()),
                            ("emu_fq_vec", orc.fq, _edge_fq(500, 3), _edge_fq(500, 4)[::-1].copy())):
        for op in ("add", "sub", "mul"):
            assert np.array_equal(_vec(emu, name, OPS[op], a, b), ofn(op, a, b)), (name, op)
        for op in ("neg", "sqr", "from_mont"):
            assert np.array_equal(_vec(emu, name, OPS[op], a), ofn(op, a)), (name, op)
        assert np.array_equal(_vec(emu, name, 10, a), ofn("sqr", a)), (name, "dedicated squaring")
        assert np.array_equal(_vec(emu, name, OPS["to_mont"], ofn("from_mont", a)), a)
        assert np.array_equal(_vec(emu, name, OPS["inv"], a[:40]), ofn("inv", a[:40]))


def test_emulated_fq2(emu, orc):
    a = np.concatenate([_edge_fq(200, 5), _edge_fq(200, 6)[::-1]], axis=1)
    b = np.concatenate([_edge_fq(200, 7)[::-1], _edge_fq(200, 8)], axis=1)
    for op in ("add", "sub", "mul"):
        assert np.array_equal(_vec(emu, "emu_fq2_vec", OPS[op], a, b), orc.fq2(op, a, b)), op
    for op in ("neg", "sqr"):
        assert np.array_equal(_vec(emu, "emu_fq2_vec", OPS[op], a), orc.fq2(op, a)), op
    assert np.array_equal(_vec(emu, "emu_fq2_vec", OPS["inv"], a[:30]), orc.fq2("inv", a[:30]))


def _sums(emu, fn, pts, negate):
    limbs = pts.shape[1]
    out = np.zeros((3, limbs), dtype=np.uint64)
    inf = np.zeros(3, dtype=np.uint8)
    neg = np.ascontiguousarray(negate, dtype=np.uint8)
    getattr(emu, fn)(_p32(np.ascontiguousarray(pts)), neg.ctypes.data_as(u8p), C.c_size_t(len(pts)), _p32(out),
                     inf.ctypes.data_as(u8p))
    return out, inf


def _py_sum(F, pts):
    acc = None
    for p in pts:
        acc = P.ec_add(F, acc, p)
    return acc


def _neg(F, p):
    if F is P.F1:
        return (p[0], (-p[1]) % P.Q_MOD)
    return (p[0], tuple((-c) % P.Q_MOD for c in p[1]))


@pytest.mark.parametrize("group", ["g1", "g2"])
def test_emulated_group_law(emu, orc, group):
    if group == "g1":
        F, fn, from_arr = P.F1, "emu_g1_sums", P.g1_from_arr
        bases = orc.g1_generate(0xE1, 24)
    else:
        F, fn, from_arr = P.F2, "emu_g2_sums", P.g2_from_arr
        g, _ = P.g2_to_arr((P.G2_X, P.G2_Y))
        bases = orc.g2_generate(g, 0xE2, 24)
    rng = np.random.default_rng(9)
    negate = rng.integers(0, 2, len(bases))
    pts = [from_arr(b) for b in bases]
    signed = [_neg(F, p) if s else p for p, s in zip(pts, negate)]
    expect = _py_sum(F, signed)
    out, inf = _sums(emu, fn, bases, negate)
    assert not inf.any()
    assert from_arr(out[0]) == expect and from_arr(out[1]) == expect
    assert from_arr(out[2]) == P.ec_add(F, expect, expect)
    # exceptional cases: P + P (doubling inside madd), P + P - P - P = infinity, then continue
    dup = np.stack([bases[0], bases[0], bases[0], bases[0], bases[1]])
    out, inf = _sums(emu, fn, dup, [0, 0, 1, 1, 0])
    assert not inf.any() and from_arr(out[0]) == pts[1] and from_arr(out[1]) == pts[1]
    out, inf = _sums(emu, fn, dup[:4], [0, 0, 1, 1])
    assert inf.all()
    zero = out[0].reshape(2, -1)
    if group == "g1":
        assert P.fq_from_mont_arr(zero) == [0, 1]          # affine zero is (0, 1)
    out, inf = _sums(emu, fn, dup[:2], [0, 0])
    assert from_arr(out[0]) == P.ec_add(F, pts[0], pts[0])
    # empty sum
    out, inf = _sums(emu, fn, dup[:0], [])
    assert inf.all()


@pytest.mark.parametrize("group", ["g1", "g2"])
def test_emulated_mul_small(emu, orc, group):
    if group == "g1":
        F, fn, from_arr, limbs = P.F1, emu.emu_g1_mul_small, P.g1_from_arr, 12
        base = orc.g1_generate(0xE3, 1)[0]
    else:
        F, fn, from_arr, limbs = P.F2, emu.emu_g2_mul_small, P.g2_from_arr, 24
        g, _ = P.g2_to_arr((P.G2_X, P.G2_Y))
        base = orc.g2_generate(g, 0xE4, 1)[0]
    for k in (0, 1, 2, 3, 32, 1023, 32768 - 32, (1 << 40) + 12345):
        out = np.zeros(limbs, dtype=np.uint64)
        inf = C.c_uint8(0)
        fn(_p32(base), C.c_uint64(k), _p32(out), C.byref(inf))
        expect = P.ec_mul(F, k, from_arr(base)) if k else None
        assert from_arr(out, inf.value) == expect


@pytest.mark.parametrize("c", [3, 4, 7, 11, 13, 15, 16, 20, 23])
def test_signed_digit_recoding(emu, pkg, c):
    nwin = 253 // c + 1
    vals = [0, 1, 2, P.R_MOD - 1, P.R_MOD - 2, (1 << (c - 1)), (1 << (c - 1)) + 1, (1 << c) - 1, 1 << c,
            (1 << 252) + (1 << (c - 1)), int("5" * 76) % P.R_MOD, ((1 << 253) - 1) % P.R_MOD]
    vals += P.fr_from_mont_arr(pkg.synth.fr_uniform(c, 300))
    canon = np.array([P.to_limbs(v, 4) for v in vals], dtype=np.uint64)
    out = np.zeros((nwin, len(vals)), dtype=np.uint32)
    emu.emu_signed_digits(_p32(canon), C.c_size_t(len(vals)), C.c_uint32(c), C.c_uint32(nwin), _p32(out))
    half = 1 << (c - 1)
    for i, v in enumerate(vals):
        total = 0
        for w in range(nwin):
            d = int(out[w, i])
            mag = d & 0x7FFFFFFF
            assert mag <= half
            assert not (d >> 31 and mag == 0)
            total += (-mag if d >> 31 else mag) << (c * w)
        assert total == v


def test_emulated_euclid_inverse(emu, orc, pkg):
    """inv_euclid (the binary extended Euclid of ff/src/fields/macros.rs:389-443) == the Fermat ladder == oracle"""
    a = _edge_fr(pkg, 300, 21)
    assert np.array_equal(_vec(emu, "emu_fr_vec", 11, a), orc.fr("inv", a))
    q = _edge_fq(300, 22)
    assert np.array_equal(_vec(emu, "emu_fq_vec", 11, q), orc.fq("inv", q))
    q2 = np.concatenate([_edge_fq(60, 23), _edge_fq(60, 24)[::-1]], axis=1)
    assert np.array_equal(_vec(emu, "emu_fq2_vec", 11, q2), orc.fq2("inv", q2))


@pytest.mark.parametrize("group", ["g1", "g2"])
def test_emulated_jacobian_doubling_chain(emu, orc, group):
    """jac_dbl (window-table builder): 2^k P for k up to a full window, against the big-int model"""
    if group == "g1":
        F, fn, from_arr, limbs = P.F1, emu.emu_g1_jac_dbl_chain, P.g1_from_arr, 12
        base = orc.g1_generate(0xE5, 1)[0]
    else:
        F, fn, from_arr, limbs = P.F2, emu.emu_g2_jac_dbl_chain, P.g2_from_arr, 24
        g, _ = P.g2_to_arr((P.G2_X, P.G2_Y))
        base = orc.g2_generate(g, 0xE6, 1)[0]
    for k in (0, 1, 2, 13, 23):
        out = np.zeros(limbs, dtype=np.uint64)
        fn(_p32(base), C.c_uint32(k), _p32(out))
        assert from_arr(out) == P.ec_mul(F, 1 << k, from_arr(base))


def test_emulated_lazy_reduction(emu, orc, pkg):
    """add_lazy / sub_lazy / mul_lazy / reduce_full (values in [0, 2p), the NTT butterflies) agree with the strict
    arithmetic, including the extreme operands"""
    a, b = _edge_fr(pkg, 400, 31), _edge_fr(pkg, 400, 32)[::-1].copy()
    exp = orc.fr("add", orc.fr("mul", orc.fr("sub", a, b), b), orc.fr("add", a, b))
    assert np.array_equal(_vec(emu, "emu_fr_vec", 12, a, b), exp)
