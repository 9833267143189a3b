"""The share-side operations of Marlin's polynomial-commitment and polynomial arithmetic, composed from the C ABI
(SURVEY.md §8 a16): what marlin_pc / KZG10 and DensePolynomial do with MpcField coefficients.

  KZG10::commit                     arkworks/poly-commit/src/kzg10/mod.rs:140-203   MSM over powers_of_g (+ the blinding
                                                                                    polynomial over powers_of_gamma_g)
  compute_witness_polynomial, open  kzg10/mod.rs:205-290                            p / (x - z) on shares, then the same MSMs;
                                                                                    random_v = blinding(z)
  DensePolynomial::mul              arkworks/algebra/poly/src/polynomial/univariate/dense.rs:567-583
                                                                                    2 FFT + batch product + iFFT on a domain of
                                                                                    len(a) + len(b); the product of two SHARED
                                                                                    polynomials is a Beaver batch product
                                                                                    (evaluations/univariate/mod.rs:69-77 ->
                                                                                    wire/field.rs:917-958)
  AddAssign<(F, &DensePolynomial)>  dense.rs:345-372                                the LC accumulation of marlin/mod.rs:275

Everything is linear in the shares (a party's output is its share of the result) except the shared x shared product,
whose two opens go through `net.exchange` like the prover's.  Coefficient vectors are (n, 4) Montgomery limb arrays,
low degree first; leading zero coefficients are harmless to every operation here (the reference truncates them).
"""
import numpy as np

from . import host as H
from .synth import FR_R_LIMBS


class Powers:
    """`Powers { powers_of_g, powers_of_gamma_g }` (kzg10/data_structures.rs) kept resident as two base vectors"""

    def __init__(self, powers_of_g, powers_of_gamma_g, precompute=False):
        self.g = H.register_bases(powers_of_g)
        self.gamma_g = H.register_bases(powers_of_gamma_g)
        if precompute:
            self.g.precompute(0)
            self.gamma_g.precompute(0)

    def release(self):
        self.g.release()
        self.gamma_g.release()


def _point_add(p, q):
    """p + q for two affine (xy, inf) pairs through a 2-point MSM with unit scalars (add_assign_mixed, mod.rs:199)"""
    pts = np.stack([p[0], q[0]])
    inf = np.array([p[1], q[1]], dtype=np.uint8)
    return H.msm_g1(pts, np.stack([FR_R_LIMBS, FR_R_LIMBS]), inf=inf)


def commit(powers, coeffs, blinding=None):
    """KZG10::commit on one party's share of the coefficients; `blinding` = its share of the blinding polynomial's
    coefficients (hiding_bound given) or None.  Returns the party's share of the commitment (affine point, flag)."""
    coeffs = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 4)
    c = H.msm_handle(powers.g, coeffs, offset=0, n=len(coeffs))              # num_leading_zeros = 0 (mod.rs:159)
    if blinding is None:
        return c
    blinding = np.ascontiguousarray(blinding, dtype=np.uint64).reshape(-1, 4)
    rc = H.msm_handle(powers.gamma_g, blinding, offset=0, n=len(blinding))
    return _point_add(c, rc)


def open(powers, coeffs, point, blinding=None):
    """KZG10::open: witness polynomial p / (x - point) (the remainder p(point) is dropped, mod.rs:205-217), committed
    over powers_of_g; with a blinding polynomial also its witness over powers_of_gamma_g and random_v = blinding(point).
    Returns (share of w, share of random_v or None)."""
    coeffs = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 4)
    witness, _ = H.poly_div_linear(coeffs, point)
    w = H.msm_handle(powers.g, witness, offset=0, n=len(witness)) if len(witness) else H.msm_g1(np.zeros((0, 12), np.uint64), witness)
    if blinding is None:
        return w, None
    blinding = np.ascontiguousarray(blinding, dtype=np.uint64).reshape(-1, 4)
    rw, random_v = H.poly_div_linear(blinding, point)
    if len(rw):
        w = _point_add(w, H.msm_handle(powers.gamma_g, rw, offset=0, n=len(rw)))
    return w, random_v


def _domain_log(size):
    return max(size - 1, 0).bit_length()                       # GeneralEvaluationDomain::new: next power of two


def _evaluate(coeffs, log_n):
    padded = np.zeros((1 << log_n, 4), dtype=np.uint64)
    padded[:len(coeffs)] = coeffs
    return H.ntt(padded, "fft")


def poly_mul_public(a_public, b_share):
    """&a * &b with a Public and b Shared: the batch product of a public and a shared evaluation vector is a local
    scale of every share (MpcField Mul, wire/field.rs:414-436)"""
    a, b = (np.ascontiguousarray(v, dtype=np.uint64).reshape(-1, 4) for v in (a_public, b_share))
    if not len(a) or not len(b):
        return np.zeros((0, 4), dtype=np.uint64)
    log_n = _domain_log(len(a) + len(b))
    return H.ntt(H.vec_op("mul", _evaluate(a, log_n), _evaluate(b, log_n)), "ifft")


def poly_mul_shared(a_share, b_share, net, triple):
    """&a * &b with both Shared (DensePolynomial::mul -> Evaluations *= -> batch_product_in_place -> Beaver,
    share/field.rs:97-129): triple = (x, y, z) shares of length 2^log_n; the two opens are sums of what
    net.exchange returns (wire payloads, as in the prover)."""
    a, b = (np.ascontiguousarray(v, dtype=np.uint64).reshape(-1, 4) for v in (a_share, b_share))
    log_n = _domain_log(len(a) + len(b))
    n = 1 << log_n
    tx, ty, tz = triple
    ea, eb = _evaluate(a, log_n), _evaluate(b, log_n)
    sx = H.open_sum_deserialize(np.stack(net.exchange(H.beaver_mask_serialize(ea, tx))), n)
    oy = H.open_sum_deserialize(np.stack(net.exchange(H.beaver_mask_serialize(eb, ty))), n)
    prod = H.beaver_combine(tx, ty, tz, sx, oy, net.party == 0)
    return H.ntt(prod, "ifft")


def add_assign_scaled(acc, f, other):
    """acc += (f, &other) for a public scalar f (dense.rs:345-372): resize to the longer operand, acc + f * other"""
    acc, other = (np.ascontiguousarray(v, dtype=np.uint64).reshape(-1, 4) for v in (acc, other))
    if not len(other):
        return acc
    n = max(len(acc), len(other))
    a = np.zeros((n, 4), dtype=np.uint64)
    a[:len(acc)] = acc
    b = np.zeros((n, 4), dtype=np.uint64)
    b[:len(other)] = other
    return H.vec_op("axpy", a, b, np.asarray(f, dtype=np.uint64).reshape(1, 4))
