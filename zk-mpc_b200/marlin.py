"""The share-side op sequence of Marlin's AHP prover rounds, composed from the C ABI:

  prover_init          arkworks/marlin/src/ahp/prover.rs:212-308   z_A = A z, z_B = B z (public CSR x share vector)
  prover_first_round   prover.rs:311-404                           w, z_A, z_B polynomials (iFFTs on shares, blinding by
                                                                   r * v_H, division by v_X) and the mask polynomial
  prover_second_round  prover.rs:438-566                           z_A * z_B (Beaver batch product between two FFTs),
                                                                   r(alpha, .), t, z, the sumcheck quotient h_1 and g_1
  calculate_t          prover.rs:406-423                           public: sum_M eta_M * M^T r(alpha, .), re-indexed

The third round works on public index polynomials only (prover.rs:583-716) and is not part of the share path.

Shared polynomials are fixed-length coefficient arrays: `is_zero` on a Shared value is false
(mpc-algebra/src/wire/field.rs:598-608), so the reference never truncates them and the multiplication domain of the
second round follows the untruncated lengths (8 |H| where the plain prover uses 4 |H|).  Public polynomials are
truncated like DensePolynomial::from_coefficients_vec does.  Shared - Public and Shared + Public touch the leader's share
only (`shift`, share/additive.rs:147-152).  Arrays are (n, 4) Montgomery limbs, low degree first.
"""
import numpy as np

from . import host as H
from . import kzg
from .synth import FR_R_LIMBS

ZK_BOUND = 1                                                  # prover.rs:283


def _fr(v):
    return np.ascontiguousarray(v, dtype=np.uint64).reshape(-1, 4)


def _pad(v, n):
    v = _fr(v)
    if len(v) > n:
        raise ValueError("vector of %d elements does not fit a domain of %d" % (len(v), n))
    out = np.zeros((n, 4), dtype=np.uint64)
    out[:len(v)] = v
    return out


def _trim_public(v):
    """DensePolynomial::from_coefficients_vec on public coefficients: drop leading zeros"""
    v = _fr(v)
    nz = np.flatnonzero(v.any(axis=1))
    return v[:nz[-1] + 1] if len(nz) else v[:0]


def _domain(size):
    return 1 << max(size - 1, 0).bit_length()               # GeneralEvaluationDomain::new


def _add(a, b):
    return H.vec_op("axpy", a, b, FR_R_LIMBS.reshape(1, 4))


def reindex_by_subdomain(nh, nx, index):
    """EvaluationDomain::reindex_by_subdomain (poly/src/domain/mod.rs:195-217), vectorised over `index`"""
    index = np.asarray(index, dtype=np.int64)
    period = nh // nx
    i = index - nx
    rest = i + i // max(period - 1, 1) + 1
    return np.where(index < nx, index * period, rest)


class Index:
    """The public square matrices A, B, C kept resident, plus their transposes re-indexed for calculate_t"""

    def __init__(self, mats, num_constraints, num_inputs):
        self.nh = _domain(num_constraints)
        self.nx = _domain(num_inputs)
        if self.nx != num_inputs:
            raise ValueError("formatted public input must have a power-of-two length (prover.rs:253)")
        self.num_constraints, self.num_inputs = num_constraints, num_inputs
        self.m, self.mt = [], []
        for row_ptr, col, coeff in mats:
            row_ptr, col, coeff = np.asarray(row_ptr, np.uint64), np.asarray(col, np.uint32), _fr(coeff)
            if len(row_ptr) - 1 != num_constraints:
                raise ValueError("matrix must have num_constraints rows")
            self.m.append(H.CsrMatrix(row_ptr, col, coeff, num_constraints))
            # transpose: entry (r, c) -> row reindex(c), column r (calculate_t: t[reindex(c)] += eta * coeff * r_alpha[r])
            rows = np.repeat(np.arange(num_constraints, dtype=np.int64), np.diff(row_ptr.astype(np.int64)))
            tr = reindex_by_subdomain(self.nh, self.nx, col.astype(np.int64))
            order = np.argsort(tr, kind="stable")
            t_ptr = np.zeros(self.nh + 1, dtype=np.uint64)
            np.cumsum(np.bincount(tr, minlength=self.nh), out=t_ptr[1:])
            self.mt.append(H.CsrMatrix(t_ptr, rows[order].astype(np.uint32), coeff[order], num_constraints))

    def release(self):
        for m in self.m + self.mt:
            m.release()


def prover_init(index, x_public, w_share, is_leader):
    """z_A, z_B on the local assignment [x | w]: public entries live on the leader (from_public)"""
    x, w = _fr(x_public), _fr(w_share)
    if len(x) != index.num_inputs or len(x) + len(w) != index.num_constraints:
        raise ValueError("instance does not match index")          # Error::InstanceDoesNotMatchIndex
    z = np.concatenate([x if is_leader else np.zeros_like(x), w])
    return index.m[0].spmv(z), index.m[1].spmv(z)


def _x_poly(index, x_public):
    return H.ntt(_fr(x_public), "ifft")


def _blind(poly, r):
    """poly + r * v_H for a poly of |H| coefficients: degree |H| gets r, degree 0 loses it"""
    r = _fr(r)
    out = np.concatenate([poly, r])
    out[0:1] = H.vec_op("sub", poly[0:1], r)
    return out


def prover_first_round(index, x_public, w_share, z_a, z_b, blinders, mask_share, is_leader):
    """blinders = this party's shares of the three F::rand values (w, z_A, z_B order); mask_share = its share of the
    random mask polynomial (3|H| + 2 zk - 2 coefficients).  Returns the four first-round oracles."""
    nh, nx = index.nh, index.nx
    x_evals = H.ntt(_pad(_x_poly(index, x_public), nh), "fft")
    ratio = nh // nx
    w_ext = _pad(w_share, nh - nx)
    k = np.arange(nh)
    keep = k % ratio != 0
    w_evals = np.zeros((nh, 4), dtype=np.uint64)
    w_evals[keep] = w_ext[(k - k // ratio - 1)[keep]]
    if is_leader:                                             # Shared - Public: the leader's share moves
        w_evals[keep] = H.vec_op("sub", w_evals[keep], x_evals[keep])
    w_full = _blind(H.ntt(w_evals, "ifft"), blinders[0])
    w_poly, _ = H.poly_div_vanishing(w_full, nx)              # remainder is a sharing of zero (prover.rs:353)
    z_a_poly = _blind(H.ntt(_pad(z_a, nh), "ifft"), blinders[1])
    z_b_poly = _blind(H.ntt(_pad(z_b, nh), "ifft"), blinders[2])
    mask = _fr(mask_share).copy()
    if len(mask) != 3 * nh + 2 * ZK_BOUND - 2:
        raise ValueError("mask polynomial must have degree 3|H| + 2 zk - 3")
    _, rem = H.poly_div_vanishing(mask, nh)
    mask[0:1] = H.vec_op("sub", mask[0:1], rem[0:1])          # mask_poly[0] -= scaled_sigma_1
    return dict(w=w_poly, z_a=z_a_poly, z_b=z_b_poly, mask=mask)


def r_alpha_x_evals(nh, alpha):
    """batch_eval_unnormalized_bivariate_lagrange_poly_with_diff_inputs (ahp/mod.rs:357-364): v_H(alpha) / (alpha - h)"""
    alpha = _fr(alpha)
    e1 = np.zeros((nh, 4), dtype=np.uint64)
    if nh > 1:
        e1[1] = FR_R_LIMBS
        elements = H.ntt(e1, "fft")                           # the polynomial x on the domain: its elements
    else:
        elements = FR_R_LIMBS.reshape(1, 4).copy()
    inv = H.field_op("fr", "inv", H.vec_op("sub", np.repeat(alpha, nh, axis=0), elements))
    v_h = np.zeros((nh + 1, 4), dtype=np.uint64)
    v_h[nh] = FR_R_LIMBS
    v_h[0:1] = H.field_op("fr", "neg", FR_R_LIMBS.reshape(1, 4))
    return H.vec_op("mul_const", inv, None, H.poly_evaluate(v_h, alpha).reshape(1, 4))


def calculate_t(index, etas, r_alpha_evals):
    r = _pad(r_alpha_evals, index.nh)[:index.num_constraints]
    t = H.vec_op("mul_const", index.mt[0].spmv(r), None, _fr(etas[0]))
    for m, eta in zip(index.mt[1:], etas[1:]):
        t = H.vec_op("axpy", t, m.spmv(r), _fr(eta))
    return _trim_public(H.ntt(t, "ifft"))


def prover_second_round(index, first, x_public, alpha, etas, net, triple, is_leader):
    """etas = (eta_a, eta_b, eta_c); triple = Beaver (x, y, z) shares of 4|H| elements for z_A * z_B.
    Returns (t [public], g_1, h_1 [shares]) and the intermediate z_c for inspection."""
    nh, nx = index.nh, index.nx
    z_a, z_b = first["z_a"], first["z_b"]
    z_c = kzg.poly_mul_shared(z_a, z_b, net, triple)
    summed = H.vec_op("mul_const", z_c, None, _fr(etas[2]))
    low = H.vec_op("axpy", H.vec_op("axpy", summed[:len(z_a)], z_a, _fr(etas[0])), z_b, _fr(etas[1]))
    summed[:len(z_a)] = low
    r_evals = r_alpha_x_evals(nh, alpha)
    r_alpha = _trim_public(H.ntt(r_evals, "ifft"))
    t_poly = calculate_t(index, etas, r_evals)
    z_poly = H.poly_mul_vanishing(first["w"], nx)
    if is_leader:                                             # Shared + Public
        z_poly[:nx] = _add(z_poly[:nx], _x_poly(index, x_public))
    mask = first["mask"]
    mul_n = _domain(max(len(mask), len(r_alpha) + len(summed), len(t_poly) + len(z_poly)))
    log_n = mul_n.bit_length() - 1
    ev = lambda p: H.ntt(_pad(p, mul_n), "fft")
    lhs = H.vec_op("mul", ev(summed), ev(r_alpha))            # Shared * Public: local scale
    rhs = H.vec_op("mul", ev(z_poly), ev(t_poly))
    q_1 = H.ntt(H.vec_op("sub", lhs, rhs), "ifft")
    q_1[:len(mask)] = _add(q_1[:len(mask)], mask)
    h_1, x_g_1 = H.poly_div_vanishing(q_1, nh)
    return dict(t=t_poly, g_1=x_g_1[1:], h_1=h_1, z_c=z_c, mul_domain=log_n)
