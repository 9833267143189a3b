"""CPU checks of the oracle's restatement of the Marlin AHP prover rounds (oracle.marlin_rounds, following
arkworks/marlin/src/ahp/prover.rs:212-566): the algebraic identities the verifier relies on must hold for a satisfied
instance, evaluated at a random point with independent python big-int arithmetic."""
import numpy as np
import pytest

import helpers

P = None


def _eval(poly, x):
    acc = 0
    for c in reversed(poly):
        acc = (acc * x + c) % P
    return acc


@pytest.mark.parametrize("nc,ni", [(16, 4), (50, 2), (64, 1), (37, 8)])
def test_marlin_rounds_identities(orc, pkg, nc, ni):
    global P
    P = orc.FR_MODULUS
    S = pkg.synth
    mats, ints, x, w = helpers.synth_marlin_instance(pkg, orc, 0x3A00 + nc, nc, ni)
    nh = 1 << max(nc - 1, 0).bit_length()
    rnd = orc.fr_to_ints(S.fr_uniform(0x3B00 + nc, 3 * nh + 16))
    blinders, alpha, etas, xi = rnd[:3], rnd[3], rnd[4:7], rnd[7]
    mask = rnd[8:8 + 3 * nh]
    xs, ws = orc.fr_to_ints(x), orc.fr_to_ints(w)
    out = orc.marlin_rounds(ints, nc, ni, xs, ws, blinders, mask, alpha, etas)
    v_h = lambda y: (pow(y, nh, P) - 1) % P
    gen = orc.fr_to_ints(orc.domain_params(nh.bit_length() - 1)["group_gen"].reshape(1, 4))[0]
    # z_A interpolates A z on H (blinded by a multiple of v_H): check at two domain points
    for k in (0, nh - 1, nh // 2):
        h = pow(gen, k, P)
        assert _eval(out["z_a_poly"], h) == (out["z_a"][k] if k < nc else 0)
        assert _eval(out["z_b_poly"], h) == (out["z_b"][k] if k < nc else 0)
    # z(x) = w(x) v_X(x) + x(x) takes the assignment's values on H in subdomain order: input j sits at index j |H|/|X|
    x_poly = orc.fr_to_ints(orc.ntt(x, "ifft"))
    z_at = lambda y: (_eval(out["w"], y) * ((pow(y, ni, P) - 1) % P) + _eval(x_poly, y)) % P
    period = nh // ni
    for j in range(ni):
        assert z_at(pow(gen, j * period, P)) == xs[j]
    if period > 1:
        assert z_at(gen) == ws[0]                               # index 1 of H is the first witness slot
    # the mask sums to zero over H, so its remainder mod v_H has no constant term
    rem = [sum(out["mask"][i::nh]) % P for i in range(nh)]
    assert rem[0] == 0
    # sumcheck identity at a random point; satisfied constraints make the remainder's constant term vanish
    assert out["x_g_1_0"] == 0
    z_c_xi = _eval(out["z_a_poly"], xi) * _eval(out["z_b_poly"], xi) % P
    assert _eval(out["z_c"], xi) == z_c_xi
    r_alpha_xi = (v_h(alpha) - v_h(xi)) * pow((alpha - xi) % P, P - 2, P) % P
    summed = (etas[0] * _eval(out["z_a_poly"], xi) + etas[1] * _eval(out["z_b_poly"], xi) + etas[2] * z_c_xi) % P
    lhs = (_eval(out["mask"], xi) + r_alpha_xi * summed - _eval(out["t"], xi) * z_at(xi)) % P
    rhs = (_eval(out["h_1"], xi) * v_h(xi) + xi * _eval(out["g_1"], xi)) % P
    assert lhs == rhs
    assert len(out["g_1"]) <= nh - 1 and len(out["h_1"]) <= 2 * nh + 2 * 1 - 1   # prover.rs:549-550
    # t(x) from its definition: t(h) = sum_M eta_M sum_r M[r][c(h)] r(alpha, h_r), at the domain point of column 0
    t0 = 0
    for (row_ptr, col, coeff), eta in zip(ints, etas):
        for r in range(nc):
            for k in range(row_ptr[r], row_ptr[r + 1]):
                if col[k] == 0:
                    h_r = pow(gen, r, P)
                    t0 += eta * coeff[k] * (v_h(alpha) * pow((alpha - h_r) % P, P - 2, P))
    assert _eval(out["t"], 1) == t0 % P


def test_host_side_index_helpers(pkg):
    """the numpy helpers of marlin.py that run on the host (no GPU): reindex_by_subdomain against the reference's scalar
    definition (poly/src/domain/mod.rs:195-217), domain sizes, public truncation"""
    M = pkg.marlin

    def scalar(nh, nx, index):
        period = nh // nx
        if index < nx:
            return index * period
        i = index - nx
        return i + i // (period - 1) + 1

    for nh, nx in ((8, 1), (8, 2), (64, 4), (1024, 8), (16, 16)):
        idx = np.arange(nh)
        got = M.reindex_by_subdomain(nh, nx, idx)
        want = [scalar(nh, nx, int(i)) if (nh > nx or i < nx) else None for i in idx]
        assert [int(g) for g in got] == want
        assert sorted(int(g) for g in got) == list(range(nh))          # a permutation of H
    assert [M._domain(k) for k in (0, 1, 2, 3, 4, 5, 1000, 1024, 1025)] == [1, 1, 2, 4, 4, 8, 1024, 1024, 2048]
    v = np.zeros((6, 4), dtype=np.uint64)
    v[1, 2] = 7
    v[3, 0] = 1
    assert len(M._trim_public(v)) == 4 and len(M._trim_public(np.zeros((3, 4), dtype=np.uint64))) == 0
