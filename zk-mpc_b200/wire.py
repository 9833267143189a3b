"""Host mirror of the wire-level rules that decide WHAT the GPU kernels are handed when the reference's
MpcField / MpcGroup values reach the two linear seams.

A `Vec<MpcField<Fr, S>>` of one party is an MpcVec: per element a tag (Public / Shared), the local value and,
for the SPDZ backend, the local MAC value.  Rules mirrored (paths relative to the zk-mpc repository):

  from_public            additive: value on the leader, 0 elsewhere        mpc-algebra/src/share/additive.rs:89-93
                         SPDZ: sh as above, mac = value * mac_share        mpc-algebra/src/share/spdz.rs:135-140
                         (mac_share = 1 on the leader, 0 elsewhere         spdz.rs:31-38)
  all_public_or_shared   all Public -> the public values; mixed -> publics lifted with from_public; all Shared
                         or empty -> the shares                             mpc-algebra/src/wire/field.rs:75-99
  MpcField as FftField   roots are Public, so a transform is the plain NTT of the party's local values; an
                         all-Public vector stays Public                     wire/field.rs:1068-1082, :339-492
  multi_scalar_mul       bases must be Public; all-Public scalars -> plain MSM, result lifted with from_public;
                         otherwise multi_scale_pub_group on the (forced) shares   wire/pairing.rs:714-777

Only tag bookkeeping and limb packing happen here; every field or curve operation goes through the C ABI.
"""
import numpy as np

from . import host as H


class MpcVec:
    """local view of a Vec<MpcField>: val (n,4) Montgomery limbs, shared (n,) bool tags, mac (n,4) or None"""

    def __init__(self, val, shared, mac=None):
        self.val = np.ascontiguousarray(val, dtype=np.uint64).reshape(-1, 4)
        self.shared = np.broadcast_to(np.asarray(shared, dtype=bool), (self.val.shape[0],)).copy()
        self.mac = None if mac is None else np.ascontiguousarray(mac, dtype=np.uint64).reshape(-1, 4)

    @staticmethod
    def public(val):
        return MpcVec(val, False)

    @staticmethod
    def share(val, mac=None):
        return MpcVec(val, True, mac)

    def __len__(self):
        return self.val.shape[0]

    @property
    def spdz(self):
        return self.mac is not None


def force_shared(vec, is_leader, spdz):
    """every element as a share: Shared entries as they are, Public(x) through from_public.  Returns the planes
    the kernels consume: (n,4) additive, (2,n,4) = [sh, mac] for SPDZ."""
    pub = ~vec.shared
    sh = vec.val.copy()
    if not is_leader:
        sh[pub] = 0                                     # additive.rs:89-93
    if not spdz:
        return sh
    mac = vec.mac.copy() if vec.mac is not None else np.zeros_like(sh)
    mac[pub] = vec.val[pub] if is_leader else 0          # value * mac_share (spdz.rs:31-38,135-140)
    return np.stack([sh, mac])


def all_public_or_shared(vec, is_leader, spdz):
    """("public", values) or ("shared", planes) exactly as MpcField::all_public_or_shared decides"""
    n_pub, n_sh = int((~vec.shared).sum()), int(vec.shared.sum())
    if n_pub and not n_sh:
        return "public", vec.val
    return "shared", force_shared(vec, is_leader, spdz)


def fft(vec, kind, is_leader, spdz=False):
    """EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place over MpcField coefficients"""
    tag, data = all_public_or_shared(vec, is_leader, spdz)
    if len(vec) == 0:
        return vec
    if tag == "public":
        return MpcVec.public(H.ntt(data, kind))          # every party computes the same public transform
    if spdz:
        out = H.ntt(data.reshape(-1, 4), kind, batch=2).reshape(2, -1, 4)      # sh and mac planes: linear
        return MpcVec.share(out[0], out[1])
    return MpcVec.share(H.ntt(data, kind))


class GroupShare:
    """one party's share of a group element: affine limbs + infinity flag (+ the mac component for SPDZ)"""

    def __init__(self, sh, mac=None):
        self.sh, self.mac = sh, mac


def _zero_point(limbs):
    out = np.zeros(limbs, dtype=np.uint64)
    out[limbs // 2:limbs // 2 + 6] = H_ONE_FQ       # affine zero is (0, 1, infinity)
    return out, 1


H_ONE_FQ = np.array([202099033278250856, 5854854902718660529, 11492539364873682930, 8885205928937022213,
                     5545221690922665192, 39800542322357402], dtype=np.uint64)       # Fq::one() (fq.rs R)


def multi_scalar_mul(bases_xy, scalars, is_leader, spdz=False, inf=None, g2=False):
    """MpcG{1,2}Affine::multi_scalar_mul for Public bases (asserted by the reference, pairing.rs:716)"""
    limbs = 24 if g2 else 12
    fn = H.msm_g2 if g2 else H.msm_g1
    tag, data = all_public_or_shared(scalars, is_leader, spdz)
    if tag == "public":
        r = fn(bases_xy, data, inf)                      # plain MSM on every party ...
        mine = r if is_leader else _zero_point(limbs)    # ... lifted with from_public: the leader holds it
        return GroupShare(mine, mine if spdz else None)  # SPDZ: mac = value * mac_share
    if spdz:
        sh, mac = H.multi_scale_pub_group(bases_xy, data, inf, g2)
        return GroupShare(sh, mac)
    return GroupShare(fn(bases_xy, data, inf))
