"""The hot path of one party's collaborative Groth16 proof, composed from the C ABI.

Mirrors `create_proof` of the reference (src/groth16.rs:68-183) between constraint synthesis and reveal:

    witness_map            src/groth16.rs:100-104,240-307   A z, B z, C z -> 3 iFFT, 3 coset FFT, Beaver batch
                                                            product (2 opens over the network), - c, / Z_H, coset iFFT
    h_acc, l_aux_acc       :106-113                         MSMs over pk.h_query / pk.l_query
    g_a, g1_b, g2_b        :137-160 + calculate_coeff :185-201   initial + query[0] + MSM(query[1..], assignment) + vk_param
    g_c                    :165-171                         s*g_a + r*g1_b - r*s*delta_g1 + l_aux_acc + h_acc

Everything stays on the device between the steps: the assignment is uploaded once and read by three MSMs, h goes
from the witness map straight into the h_query MSM, the five MSMs run on separate streams, and the public points of
calculate_coeff are folded into the registered vectors so each proof element is ONE MSM.  The two network rounds
stay with the caller (`net.exchange`, the reference's MpcSerNet::broadcast): the masked vectors leave the library
as the wire bytes of a Vec<Fr> and the received payloads are summed on the device (§8 f3).

Share semantics (additive backend; leader = party 0): a public group element enters a share only on the leader
(from_public, mpc-algebra/src/share/additive.rs:387-392), so the folded public points get scalar 0 on the other
parties.  r, s are PUBLIC scalars here (0 = create_proof_no_zk, :49-64); the reference's create_random_proof samples
them as shares, whose product with the shared g_a needs a group-by-field Beaver round that stays in Rust.
"""
import ctypes as C

import numpy as np

from . import _lib
from . import host as H
from .synth import FR_R_LIMBS


def _cat(*parts):
    return np.ascontiguousarray(np.concatenate([np.asarray(p, dtype=np.uint64).reshape(-1, parts[0].shape[-1]) for p in parts]))


class ProvingKey:
    """pk.{a,b_g1,b_g2,h,l}_query + the vk points, resident on the calling thread's device.

    Each query is (points (m, 12|24) uint64, infinity flags (m,) uint8 or None).  Layout of the registered vectors:
      A   = a_query[1:]    | a_query[0]    | alpha_g1 | delta_g1        scalars: assignment | 1 | 1 | r
      B1  = b_g1_query[1:] | b_g1_query[0] | beta_g1  | delta_g1        scalars: assignment | 1 | 1 | s
      B2  = b_g2_query[1:] | b_g2_query[0] | beta_g2  | delta_g2  (G2)  scalars: assignment | 1 | 1 | s
      LH  = l_query        | h_query                                    scalars: witness | h[: len(h_query)]
    """

    def __init__(self, a_query, b_g1_query, b_g2_query, h_query, l_query, alpha_g1, beta_g1, beta_g2, delta_g1, delta_g2,
                 precompute=True, table_bits_g1=0, table_bits_g2=0):
        def flags(q):
            pts, inf = q
            return np.zeros(len(pts), dtype=np.uint8) if inf is None else np.asarray(inf, dtype=np.uint8)

        def coeff_vector(q, vk_param, delta):
            pts, inf = np.asarray(q[0], dtype=np.uint64), flags(q)
            bases = _cat(pts[1:], pts[:1], vk_param.reshape(1, -1), delta.reshape(1, -1))
            return bases, np.concatenate([inf[1:], inf[:1], np.zeros(2, dtype=np.uint8)])

        self.num_vars = len(a_query[0])                     # instance (with the constant 1) + witness variables
        if len(b_g1_query[0]) != self.num_vars or len(b_g2_query[0]) != self.num_vars:
            raise ValueError("a_query, b_g1_query and b_g2_query must have one point per variable")
        self.n_l, self.n_h = len(l_query[0]), len(h_query[0])
        self.delta_g1 = np.asarray(delta_g1, dtype=np.uint64)
        ba, ia = coeff_vector(a_query, alpha_g1, delta_g1)
        bb, ib = coeff_vector(b_g1_query, beta_g1, delta_g1)
        b2, i2 = coeff_vector(b_g2_query, beta_g2, delta_g2)
        blh = _cat(np.asarray(l_query[0], dtype=np.uint64), np.asarray(h_query[0], dtype=np.uint64))
        self.A = H.register_bases(ba, inf=ia)
        self.B1 = H.register_bases(bb, inf=ib)
        self.B2 = H.register_bases(b2, inf=i2, g2=True)
        self.LH = H.register_bases(blh, inf=np.concatenate([flags(l_query), flags(h_query)]))
        if precompute:                                       # the CRS is fixed per circuit: window tables pay off
            for h in (self.A, self.B1, self.B2, self.LH):
                h.precompute(table_bits_g2 if h.g2 else table_bits_g1)     # 0 = the library's choice for the size

    def release(self):
        for h in (self.A, self.B1, self.B2, self.LH):
            h.release()


class R1CS:
    """the public constraint matrices (cs.to_matrices(), src/groth16.rs:244) kept resident as CSR"""

    def __init__(self, a, b, c, num_inputs, num_vars):
        """a, b, c: (row_ptr, col, coeff) triples; num_inputs = instance variables including the constant 1"""
        self.A, self.B, self.C = (H.CsrMatrix(*m, cols=num_vars) for m in (a, b, c))
        self.num_constraints, self.num_inputs, self.num_vars = self.A.rows, int(num_inputs), int(num_vars)
        size = self.num_constraints + self.num_inputs         # D::new(num_constraints + num_inputs), :258-260
        self.log_n = max(size - 1, 0).bit_length()
        self.n = 1 << self.log_n

    def release(self):
        for m in (self.A, self.B, self.C):
            m.release()


def dummy_triple(n, is_leader):
    """DummyFieldTripleSource (mpc-algebra/src/wire/field.rs:44-63): shares of (1, 1, 1) = 1 on the leader, 0 elsewhere"""
    v = np.tile(FR_R_LIMBS, (n, 1)) if is_leader else np.zeros((n, 4), dtype=np.uint64)
    return v, v, v


class ProverSession:
    """Per-party working set reused across proofs of one circuit: the scalar buffers of the four MSMs, their
    Jacobian partials and four streams.  Allocating these per proof costs more than the kernels at Groth16 sizes
    (cudaMalloc / cudaFree synchronise the device, which also stalls the other parties' streams)."""

    def __init__(self, pk, r1cs):
        if pk.num_vars != r1cs.num_vars:
            raise ValueError("proving key and R1CS disagree on the number of variables")
        if pk.n_h > r1cs.n or pk.n_l != r1cs.num_vars - r1cs.num_inputs:
            raise ValueError("h_query / l_query do not fit the domain / witness size")
        self.pk, self.r1cs = pk, r1cs
        m = r1cs.num_vars - 1
        self.za = H.DeviceBuffer((m + 3) * 32)
        self.zb = H.DeviceBuffer((m + 3) * 32)
        self.lh = H.DeviceBuffer((pk.n_l + pk.n_h) * 32)
        self.parts = [H.DeviceBuffer(18 * 8), H.DeviceBuffer(18 * 8), H.DeviceBuffer(36 * 8), H.DeviceBuffer(18 * 8)]
        self.streams = []
        for _ in range(4):
            p = C.c_void_p(0)
            _lib.call("mpc_cuda_stream_create", C.byref(p))
            self.streams.append(p)
        # page-locked staging: the four wire payloads of a proof (masked a, masked b, and the SPDZ dx of each), the
        # six tail scalars, and the dummy triple when the caller brings none
        plen = 8 + 32 * r1cs.n
        self.pinned = [H.PinnedBuffer(4 * plen + 6 * 32)]
        self.payload = [self.pinned[0].array(np.uint8, plen, k * plen) for k in range(4)]
        self.tails = self.pinned[0].array(np.uint64, 24, 4 * plen).reshape(6, 4)
        self._dummy = {}

    def _dummy_triple(self, leader, spdz):
        """dummy_triple in page-locked memory, built once per session (SPDZ: MAC key 1 shared as (1, 0, 0), so the
        mac plane equals the sh plane)"""
        key = (bool(leader), bool(spdz))
        if key not in self._dummy:
            n, planes = self.r1cs.n, 2 if spdz else 1
            buf = H.PinnedBuffer(planes * n * 32)
            self.pinned.append(buf)
            v = buf.array(np.uint64, planes * n * 4).reshape((planes, n, 4) if spdz else (n, 4))
            v[...] = FR_R_LIMBS if leader else 0
            self._dummy[key] = v
        v = self._dummy[key]
        return v, v, v

    def close(self):
        for p in self.streams:
            _lib.call("mpc_cuda_stream_destroy", p)
        self.streams = []
        self.payload, self.tails, self._dummy = [], None, {}
        for b in self.pinned:
            b.free()
        self.pinned = []
        for b in [self.za, self.zb, self.lh] + self.parts:
            b.free()

    def prove(self, assignment, net, triple=None, r=None, s=None, spdz=False):
        """One party's share of the proof (a, b, c) from its share of the full assignment (instance | witness
        values, (num_vars, 4) Montgomery limbs; the constant 1 and public inputs already lifted with from_public).

        net: .party, .n_parties, .exchange(uint8 array) -> list of every party's array in party order (a
        broadcast).  Returns {"a": (xy, inf), "b": (xy, inf) over G2, "c": (xy, inf)}: the shares
        MpcPairingEngine would reveal.

        spdz (the malicious backend, mpc-algebra/src/share/spdz.rs): assignment and triple are (2, ., 4) = [sh, mac]
        planes; the witness map runs on both planes, each of the two opens is followed by its MAC check (the local
        halves dx = mac_share * val - mac are exchanged and must sum to zero, spdz.rs:177-196), and — the reference's
        own quirk — every MSM reads the sh plane for BOTH components (spdz.rs:482-488), so the mac component of each
        proof element equals its sh component: returned under "mac"."""
        pk, r1cs = self.pk, self.r1cs
        leader = net.party == 0
        if spdz:
            planes = np.ascontiguousarray(assignment, dtype=np.uint64)
            if planes.ndim != 3 or planes.shape[0] != 2:
                raise ValueError("SPDZ assignments are (2, num_vars, 4) = [sh, mac] planes")
            z = planes[0]
        else:
            z = np.ascontiguousarray(assignment, dtype=np.uint64).reshape(-1, 4)
        if z.shape[0] != r1cs.num_vars:
            raise ValueError("assignment has %d values, the circuit %d variables" % (z.shape[0], r1cs.num_vars))
        if triple is not None:
            tx, ty, tz = triple
        zero, one = np.zeros(4, dtype=np.uint64), FR_R_LIMBS
        r = zero if r is None else np.asarray(r, dtype=np.uint64)
        s = zero if s is None else np.asarray(s, dtype=np.uint64)

        # ---- witness map: the vectors stay on the device; only wire payloads cross PCIe for the two opens
        if triple is None:
            tx, ty, tz = self._dummy_triple(leader, spdz)
        st = H.witness_map_begin_r1cs(r1cs.A, r1cs.B, r1cs.C, planes if spdz else z, r1cs.num_inputs, r1cs.log_n,
                                      tx, ty, spdz=spdz, masked=False)[2]
        try:
            for which in (0, 1):
                pay = H.witness_map_masked_payload(st, which, out=self.payload[which])
                H.witness_map_open_payloads(st, which, net.exchange(pay))
                if spdz:                            # batch_open's MAC check, local half + zero test of the sum
                    dx = H.witness_map_mac_payload(st, which, leader, out=self.payload[2 + which])
                    H.witness_map_mac_verify(st, net.exchange(dx))
            h_ptr = H.witness_map_finish_dev(st, tz, None, None, leader)    # SPDZ: [sh | mac] planes, sh first
            z_ptr, _ = H.witness_map_assignment_dev(st)                     # the sh plane of the assignment, resident

            # ---- scalars of the four MSMs, resident: [assignment[1:] | 1 | 1 | r or s] (leader) and [witness | h]:
            # device-to-device copies of the assignment the witness map uploaded, plus three host elements each
            m = r1cs.num_vars - 1
            sa, sb1, sb2, slh = (p.value for p in self.streams)
            tails = self.tails
            tails[:] = 0
            if leader:
                tails[0:2], tails[2], tails[3:5], tails[5] = one, r, one, s
            h2d = lambda dst, src, nbytes, stream: _lib.call("mpc_cuda_memcpy_h2d", C.c_void_p(dst), src.ctypes.data_as(C.c_void_p),
                                                             C.c_size_t(nbytes), C.c_void_p(stream))
            d2d = lambda dst, src, nbytes, stream: _lib.call("mpc_cuda_memcpy_d2d", C.c_void_p(dst), C.c_void_p(src),
                                                             C.c_size_t(nbytes), C.c_void_p(stream))
            d2d(self.za.ptr.value, z_ptr + 32, m * 32, sa)
            h2d(self.za.ptr.value + m * 32, tails[0:3], 3 * 32, sa)
            d2d(self.zb.ptr.value, z_ptr + 32, m * 32, sb1)
            h2d(self.zb.ptr.value + m * 32, tails[3:6], 3 * 32, sb1)
            _lib.call("mpc_cuda_stream_sync", C.c_void_p(sb1))   # B2 reads zb on its own stream
            d2d(self.lh.ptr.value, z_ptr + r1cs.num_inputs * 32, pk.n_l * 32, slh)
            d2d(self.lh.ptr.value + pk.n_l * 32, h_ptr, pk.n_h * 32, slh)
            H.msm_handle_dev(pk.A, self.za, m + 3, out=self.parts[0], stream=sa)
            H.msm_handle_dev(pk.B1, self.zb, m + 3, out=self.parts[1], stream=sb1)
            H.msm_handle_dev(pk.B2, self.zb, m + 3, out=self.parts[2], stream=sb2)
            H.msm_handle_dev(pk.LH, self.lh, pk.n_l + pk.n_h, out=self.parts[3], stream=slh)
            g_a = H.sum_partials(self.parts[0], 1, stream=sa)
            g1_b = H.sum_partials(self.parts[1], 1, stream=sb1)
            g2_b = H.sum_partials(self.parts[2], 1, g2=True, stream=sb2)
            lh_acc = H.sum_partials(self.parts[3], 1, stream=slh)
        finally:
            st.release()

        # ---- g_c = s*g_a + r*g1_b - r*s*delta_g1 + l_aux_acc + h_acc   (src/groth16.rs:165-171)
        if not r.any() and not s.any():
            g_c = lh_acc
        else:
            rs = H.field_op("fr", "neg", H.field_op("fr", "mul", r[None], s[None]))[0] if leader else zero
            pts = np.stack([g_a[0], g1_b[0], pk.delta_g1, lh_acc[0]])
            inf = np.array([g_a[1], g1_b[1], 0, lh_acc[1]], dtype=np.uint8)
            g_c = H.msm_g1(pts, np.stack([s, r, rs, one]), inf=inf)
        out = {"a": g_a, "b": g2_b, "c": g_c}
        if spdz:
            out["mac"] = {"a": g_a, "b": g2_b, "c": g_c}
        return out


def prove_party(pk, r1cs, assignment, net, triple=None, r=None, s=None):
    """one-shot form of ProverSession.prove (allocates and frees the working set around one proof)"""
    session = ProverSession(pk, r1cs)
    try:
        return session.prove(assignment, net, triple, r, s)
    finally:
        session.close()
