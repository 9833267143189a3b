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
import ctypes as C

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


# ----------------------------------------------------------------------------- the same rounds, resident
class ResidentProver:
    """prover_init + first + second round of one party with every vector resident on the device: between the upload of
    the party's inputs (assignment share, blinders, mask share, triple shares) and the oracles only the two wire
    payloads of the z_A * z_B Beaver product cross PCIe.  Same arithmetic, in the same order, as the host-array
    functions above (tests compare the two share by share); buffers are allocated once per index and reused."""

    def __init__(self, index, n_parties=3):
        self.index, self.parties = index, n_parties
        nh = index.nh
        self.n4, self.n8 = 4 * nh, 8 * nh
        E = lambda count: H.DeviceBuffer(count * 32)
        self.z = E(index.num_constraints)
        self.za, self.zb, self.we = E(nh + 1), E(nh + 1), E(nh + 1)
        self.xe, self.ra, self.rp, self.tp, self.t2 = E(nh), E(nh), E(nh), E(nh), E(nh)
        self.w, self.zp, self.rem, self.ws = E(nh + 1), E(nh + 1), E(nh), E(nh)
        self.mk = E(3 * nh)
        self.ea, self.eb, self.sx, self.oy = E(self.n4), E(self.n4), E(self.n4), E(self.n4)
        self.tx, self.ty, self.tz = E(self.n4), E(self.n4), E(self.n4)
        self.f = [E(self.n8) for _ in range(4)]
        self.h1, self.xg = E(7 * nh), E(nh)
        self.small = E(16)                                       # blinders, x_poly head, scratch elements
        self.flags = H.DeviceBuffer(16)
        plen = 8 + 32 * self.n4
        self.plen = plen
        self.pay_dev = H.DeviceBuffer(n_parties * plen)
        self.pinned = H.PinnedBuffer(2 * plen)
        self.payload = [self.pinned.array(np.uint8, plen, k * plen) for k in range(2)]
        # public per-index vectors: ones, and the 0/1 mask that clears the input positions of H (k % ratio == 0)
        ones = np.tile(FR_R_LIMBS, (nh, 1))
        self.ones = E(nh)
        H.dev_upload(self.ones.ptr.value, ones)
        keep = ones.copy()
        keep[::nh // index.nx] = 0
        self.keep = E(nh)
        H.dev_upload(self.keep.ptr.value, keep)
        H._lib.call("mpc_cuda_stream_sync", None)
        self.bufs = [self.z, self.za, self.zb, self.we, self.xe, self.ra, self.rp, self.tp, self.t2, self.w, self.zp, self.rem, self.ws,
                     self.mk, self.ea, self.eb, self.sx, self.oy, self.tx, self.ty, self.tz, self.h1, self.xg, self.small,
                     self.flags, self.pay_dev, self.ones, self.keep] + self.f

    def close(self):
        for b in self.bufs:
            b.free()
        self.bufs = []
        self.pinned.free()

    def _open(self, vec, mask, out, slot, net):
        """batch_open of vec + mask: payload from the device, exchange, received payloads summed on the device"""
        n4, plen = self.n4, self.plen
        H._lib.call("mpc_cuda_beaver_mask_serialize_dev", H._dp(vec), H._dp(mask), C.c_size_t(n4), H._dp(self.pay_dev.ptr.value), None)
        H._lib.call("mpc_cuda_memcpy_d2h", self.payload[slot].ctypes.data_as(C.c_void_p), H._dp(self.pay_dev.ptr.value),
                    C.c_size_t(plen), None)
        H._lib.call("mpc_cuda_stream_sync", None)
        got = net.exchange(self.payload[slot])
        if len(got) != self.parties:
            raise ValueError("expected %d payloads" % self.parties)
        for p, pay in enumerate(got):
            pay = np.ascontiguousarray(pay, dtype=np.uint8)
            if pay.size != plen:
                raise ValueError("payload of %d bytes, expected %d" % (pay.size, plen))
            H._lib.call("mpc_cuda_memcpy_h2d", H._dp(self.pay_dev.ptr.value + p * plen), pay.ctypes.data_as(C.c_void_p),
                        C.c_size_t(plen), None)
        H._lib.call("mpc_cuda_open_sum_deserialize_dev", H._dp(self.pay_dev.ptr.value), C.c_uint32(self.parties), C.c_size_t(n4),
                    H._dp(out), H._dp(self.flags.ptr.value), None)
        flags = H.dev_download(self.flags.ptr.value, 2)          # synchronises: the peers' buffers are free again
        if flags[1] or flags[0] != np.uint64(0xFFFFFFFFFFFFFFFF):
            raise H.MpcCudaError("malformed payload in a Beaver open")

    def rounds(self, x_public, w_share, blinders, mask_share, alpha, etas, net, triple, is_leader, powers=None, download=True):
        """Returns the oracles as host arrays (download=True) and, with `powers` (a kzg.Powers), this party's shares of
        their commitments over powers_of_g (`commitments`, without the hiding terms)."""
        ix = self.index
        nh, nx, nc, n4, n8 = ix.nh, ix.nx, ix.num_constraints, self.n4, self.n8
        log_h, ratio = nh.bit_length() - 1, nh // nx
        P = lambda b, k=0: b.ptr.value + 32 * k
        x, w = _fr(x_public), _fr(w_share)
        if len(x) != ix.num_inputs or len(x) + len(w) != nc:
            raise ValueError("instance does not match index")
        mask_share, bl = _fr(mask_share), _fr(blinders)
        if len(mask_share) != 3 * nh + 2 * ZK_BOUND - 2 or len(bl) != 3:
            raise ValueError("mask polynomial / blinders of the wrong length")
        keepalive = []
        up = lambda dst, arr: keepalive.append(H.dev_upload(dst, arr))
        one = FR_R_LIMBS.reshape(1, 4)
        # ---- prover_init: z_A = A z, z_B = B z
        up(P(self.z), np.concatenate([x if is_leader else np.zeros_like(x), w]))
        for buf, m in ((self.za, ix.m[0]), (self.zb, ix.m[1])):
            H.dev_zero(P(buf), (nh + 1) * 32)
            H.dev_spmv(m, P(self.z), P(buf))
        # ---- first round
        x_poly = _x_poly(ix, x)                                   # |X| public elements: host-side call
        H.dev_zero(P(self.xe), nh * 32)
        up(P(self.xe), x_poly)
        H.ntt_dev(P(self.xe), log_h, "fft")
        H.dev_vec_op("mul", P(self.xe), P(self.keep), None, P(self.xe), nh)           # no input positions
        # witness values into the non-input positions of H: position k <- w_ext[k - k / ratio - 1], i.e. every run of
        # ratio - 1 values lands behind one input slot: a strided copy
        H.dev_zero(P(self.ws), nh * 32)
        H.dev_zero(P(self.we), (nh + 1) * 32)
        if len(w):
            up(P(self.ws), w)
        H.dev_copy2d(P(self.we, 1), ratio * 32, P(self.ws), (ratio - 1) * 32, (ratio - 1) * 32, nx)
        up(P(self.small), bl)
        if is_leader:
            H.dev_vec_op("sub", P(self.we), P(self.xe), None, P(self.we), nh)

        def blind(buf, j):                                        # + r v_H
            H.ntt_dev(P(buf), log_h, "ifft")
            H.dev_copy(P(buf, nh), P(self.small, j), 32)
            H.dev_vec_op("sub", P(buf), P(self.small, j), None, P(buf), 1)

        blind(self.we, 0)
        H.dev_poly_div_vanishing(P(self.we), nh + 1, nx, P(self.w), P(self.rem))       # w: nh + 1 - nx coefficients
        blind(self.za, 1)
        blind(self.zb, 2)
        up(P(self.mk), mask_share)
        H.dev_poly_div_vanishing(P(self.mk), 3 * nh, nh, None, P(self.rem))
        H.dev_vec_op("sub", P(self.mk), P(self.rem), None, P(self.mk), 1)
        # ---- second round: z_A * z_B on a domain of 4|H| through Beaver
        tx, ty, tz = (_fr(v) for v in triple)
        if any(len(v) != n4 for v in (tx, ty, tz)):
            raise ValueError("triple shares must hold 4|H| elements")
        for buf, src in ((self.tx, tx), (self.ty, ty), (self.tz, tz)):
            up(P(buf), src)
        for buf, src in ((self.ea, self.za), (self.eb, self.zb)):
            H.dev_zero(P(buf), n4 * 32)
            H.dev_copy(P(buf), P(src), (nh + 1) * 32)
            H.ntt_dev(P(buf), log_h + 2, "fft")
        self._open(P(self.ea), P(self.tx), P(self.sx), 0, net)
        self._open(P(self.eb), P(self.ty), P(self.oy), 1, net)
        zc = self.f[0]                                            # 4|H| of its 8|H|: becomes summed, then padded
        H._lib.call("mpc_cuda_beaver_combine_dev", H._dp(P(self.tx)), H._dp(P(self.ty)), H._dp(P(self.tz)), H._dp(P(self.sx)),
                    H._dp(P(self.oy)), H._dp(P(zc)), C.c_size_t(n4), C.c_uint32(int(bool(is_leader))), C.c_uint32(0), None)
        H.ntt_dev(P(zc), log_h + 2, "ifft")
        z_c = H.dev_download(P(zc), n4 * 4).reshape(n4, 4) if download else None
        H.dev_vec_op("mul_const", P(zc), None, etas[2], P(zc), n4)
        H.dev_vec_op("axpy", P(zc), P(self.za), etas[0], P(zc), nh + 1)
        H.dev_vec_op("axpy", P(zc), P(self.zb), etas[1], P(zc), nh + 1)
        H.dev_zero(P(zc, n4), (n8 - n4) * 32)
        # r(alpha, .) on H: v_H(alpha) / (alpha - h); public, computed by every party
        H.dev_zero(P(self.ra), nh * 32)
        if nh > 1:
            H.dev_copy(P(self.ra, 1), P(self.ones), 32)
            H.ntt_dev(P(self.ra), log_h, "fft")                   # the elements of H
        else:
            H.dev_copy(P(self.ra), P(self.ones), 32)
        neg_alpha = H.field_op("fr", "neg", _fr(alpha))
        H.dev_vec_op("axpy", P(self.ra), P(self.ones), neg_alpha, P(self.ra), nh)      # h - alpha
        H.dev_inverse(P(self.ra), P(self.ra), nh)
        v_h = np.zeros((nh + 1, 4), dtype=np.uint64)
        v_h[nh] = FR_R_LIMBS
        v_h[0:1] = H.field_op("fr", "neg", one)
        scale = H.field_op("fr", "neg", H.poly_evaluate(v_h, _fr(alpha)[0]).reshape(1, 4))   # -v_H(alpha)
        H.dev_vec_op("mul_const", P(self.ra), None, scale, P(self.ra), nh)
        H.dev_copy(P(self.rp), P(self.ra), nh * 32)
        H.ntt_dev(P(self.rp), log_h, "ifft")
        # t = sum_M eta_M M^T r(alpha, .), re-indexed
        H.dev_spmv(ix.mt[0], P(self.ra), P(self.tp))
        H.dev_vec_op("mul_const", P(self.tp), None, etas[0], P(self.tp), nh)
        for m, eta in zip(ix.mt[1:], etas[1:]):
            H.dev_spmv(m, P(self.ra), P(self.t2))
            H.dev_vec_op("axpy", P(self.tp), P(self.t2), eta, P(self.tp), nh)
        H.ntt_dev(P(self.tp), log_h, "ifft")
        # z = w v_X + x
        H.dev_poly_mul_vanishing(P(self.w), nh + 1 - nx, nx, P(self.zp))
        if is_leader:
            if len(x_poly) > 8:
                raise ValueError("the resident composition stages at most 8 public inputs")
            up(P(self.small, 8), x_poly)
            H.dev_vec_op("axpy", P(self.zp), P(self.small, 8), one, P(self.zp), nx)
        # the multiplication domain of the untruncated shared polynomials: 8|H|
        def public_len(buf):                                      # from_coefficients_vec: only the top element as a rule
            if H.dev_download(P(buf, nh - 1), 4).any():
                return nh
            return len(_trim_public(H.dev_download(P(buf), nh * 4).reshape(nh, 4)))

        r_len, t_len = public_len(self.rp), public_len(self.tp)
        mul_n = _domain(max(3 * nh, r_len + n4, t_len + nh + 1))
        if mul_n != n8:
            raise H.MpcCudaError("unexpected multiplication domain %d" % mul_n)
        for buf, src, count in ((self.f[1], self.rp, nh), (self.f[2], self.zp, nh + 1), (self.f[3], self.tp, nh)):
            H.dev_zero(P(buf), n8 * 32)
            H.dev_copy(P(buf), P(src), count * 32)
        for buf in self.f:
            H.ntt_dev(P(buf), log_h + 3, "fft")
        H.dev_vec_op("mul", P(self.f[0]), P(self.f[1]), None, P(self.f[0]), n8)
        H.dev_vec_op("mul", P(self.f[2]), P(self.f[3]), None, P(self.f[2]), n8)
        H.dev_vec_op("sub", P(self.f[0]), P(self.f[2]), None, P(self.f[0]), n8)
        H.ntt_dev(P(self.f[0]), log_h + 3, "ifft")
        H.dev_vec_op("axpy", P(self.f[0]), P(self.mk), one, P(self.f[0]), 3 * nh)
        H.dev_poly_div_vanishing(P(self.f[0]), n8, nh, P(self.h1), P(self.xg))
        out = {"mul_domain": log_h + 3}
        if powers is not None:
            jac = H.DeviceBuffer(7 * 18 * 8)
            com = {}
            for j, (name, buf, off, count) in enumerate((("w", self.w, 0, nh + 1 - nx), ("z_a", self.za, 0, nh + 1),
                                                         ("z_b", self.zb, 0, nh + 1), ("mask", self.mk, 0, 3 * nh),
                                                         ("t", self.tp, 0, t_len), ("g_1", self.xg, 1, nh - 1),
                                                         ("h_1", self.h1, 0, 7 * nh))):
                if count:
                    H._lib.call("mpc_cuda_msm_g1_handle_dev", C.c_uint64(powers.g.handle), C.c_size_t(0), H._dp(P(buf, off)),
                                C.c_size_t(count), H._dp(jac.ptr.value + j * 144), None)
                else:
                    H.dev_zero(jac.ptr.value + j * 144, 144)
                xy, inf = np.zeros(12, dtype=np.uint64), C.c_uint8(0)
                H._lib.call("mpc_cuda_g1_sum_partials_dev", H._dp(jac.ptr.value + j * 144), C.c_uint32(1), H._p(xy), C.byref(inf), None)
                com[name] = (xy, inf.value)
            jac.free()
            out["commitments"] = com
        if download:
            D = lambda buf, count, off=0: H.dev_download(P(buf, off), count * 4).reshape(count, 4)
            out.update(w=D(self.w, nh + 1 - nx), z_a=D(self.za, nh + 1), z_b=D(self.zb, nh + 1), mask=D(self.mk, 3 * nh),
                       z_c=z_c, t=D(self.tp, t_len), g_1=D(self.xg, nh - 1, 1), h_1=D(self.h1, 7 * nh))
        H._lib.call("mpc_cuda_stream_sync", None)
        del keepalive
        return out

    def open_combination(self, terms, point, powers):
        """The opening phase on the resident oracles of the last `rounds` call: the linear combination
        sum_i coeff_i * poly_i (marlin/mod.rs:213-306 accumulates `poly += (coeff, poly)`, dense.rs:345-372), its witness
        polynomial / (x - point) and evaluation (kzg10/mod.rs:205-290, univariate_div_qr on shares), and the witness
        commitment over powers_of_g.  terms = [(name, public coeff)], names from w z_a z_b mask t g_1 h_1.
        Returns (this party's share of the witness commitment, its share of the evaluation)."""
        nh, nx = self.index.nh, self.index.nx
        P = lambda b, k=0: b.ptr.value + 32 * k
        polys = {"w": (self.w, 0, nh + 1 - nx), "z_a": (self.za, 0, nh + 1), "z_b": (self.zb, 0, nh + 1), "mask": (self.mk, 0, 3 * nh),
                 "t": (self.tp, 0, nh), "g_1": (self.xg, 1, nh - 1), "h_1": (self.h1, 0, 7 * nh)}
        if not terms:
            raise ValueError("empty combination")
        acc, wit = self.f[1], self.f[2]                           # 8|H| scratch of the second round
        n = max(polys[name][2] for name, _ in terms)
        H.dev_zero(P(acc), n * 32)
        for name, coeff in terms:
            buf, off, count = polys[name]
            if count:
                H.dev_vec_op("axpy", P(acc), P(buf, off), _fr(coeff), P(acc), count)
        H._lib.call("mpc_cuda_poly_div_linear_dev", H._dp(P(acc)), C.c_size_t(n), H._p(_fr(point)), H._dp(P(wit)) if n > 1 else None,
                    H._dp(P(self.small, 12)), None)
        value = H.dev_download(P(self.small, 12), 4)
        xy, inf = np.zeros(12, dtype=np.uint64), C.c_uint8(0)
        if n > 1:
            jac = H.DeviceBuffer(144)
            H._lib.call("mpc_cuda_msm_g1_handle_dev", C.c_uint64(powers.g.handle), C.c_size_t(0), H._dp(P(wit)), C.c_size_t(n - 1),
                        H._dp(jac.ptr.value), None)
            H._lib.call("mpc_cuda_g1_sum_partials_dev", H._dp(jac.ptr.value), C.c_uint32(1), H._p(xy), C.byref(inf), None)
            jac.free()
        else:
            inf = C.c_uint8(1)
        return (xy, inf.value), value
