"""Host-side mirror of the reference's operator interface for the hot path, over libmpc_cuda.so.

The reference is Rust; this image has no Rust toolchain, so the host layer above the C ABI is
written in Python with the reference's names and argument meaning (the Rust crate that binds the
same ABI is in bindings/rust/mpc-cuda and INTEGRATION.md).  Arrays are numpy uint64 in the
reference's in-memory layout (Montgomery limbs): Fr (n,4), G1Affine (n,12)+(n,) infinity bytes,
G2Affine (n,24)+(n,).  Every function goes through the C ABI; nothing here computes field or
curve arithmetic on the CPU.

Reference interfaces mirrored:
  Msm::msm / AffineMsm                 mpc-algebra/src/share/msm.rs:6-9,33-37
  GroupShare::multi_scale_pub_group    mpc-algebra/src/share/additive.rs:517-520, spdz.rs:482-488
  Radix2EvaluationDomain               arkworks/algebra/poly/src/domain/radix2/mod.rs:51-114
  FieldShare::{batch_open,batch_mul}   mpc-algebra/src/share/field.rs:40-42,97-129
  Field::batch_product_in_place        mpc-algebra/src/wire/field.rs:917-958
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import MpcCudaError, u64p, u8p  # noqa: F401


def _a(x, cols):
    x = np.ascontiguousarray(x, dtype=np.uint64)
    if x.shape[-1] != cols:
        raise ValueError("expected last dimension %d, got shape %s" % (cols, x.shape))
    return x


def _p(x):
    return x.ctypes.data_as(u64p) if x is not None else None


# ----------------------------------------------------------------------------- context
def init(devices=None):
    if devices:
        arr = (C.c_int32 * len(devices))(*devices)
        _lib.call("mpc_cuda_init", arr, C.c_int32(len(devices)))
    else:
        _lib.call("mpc_cuda_init", None, C.c_int32(0))


def set_party(party_id, n_parties):
    _lib.call("mpc_cuda_set_party", C.c_uint32(party_id), C.c_uint32(n_parties))


def set_device(index):
    _lib.call("mpc_cuda_set_device", C.c_int32(index))


def device_count():
    return _lib.lib().mpc_cuda_device_count()


# ----------------------------------------------------------------------------- diagnostics
_FOPS = {"add": 0, "sub": 1, "mul": 2, "neg": 3, "inv": 4, "from_mont": 5, "to_mont": 6, "sqr": 7, "mul_narrow": 8}


def field_op(field, op, a, b=None):
    limbs = 4 if field == "fr" else 6
    a = _a(a, limbs)
    b = _a(b, limbs) if b is not None else None
    out = np.empty_like(a)
    _lib.call("mpc_cuda_field_op", C.c_uint32(0 if field == "fr" else 1), C.c_uint32(_FOPS[op]), _p(a), _p(b),
              _p(out), C.c_size_t(a.size // limbs))
    return out


def microbench(kind, iters=2000):
    g = C.c_double(0)
    _lib.call("mpc_cuda_microbench", C.c_uint32(kind), C.c_uint32(iters), C.byref(g))
    return g.value


# ----------------------------------------------------------------------------- Beaver local halves
def beaver_mask(s, x):
    s, x = _a(s, 4), _a(x, 4)
    if s.shape != x.shape:
        raise ValueError("shape mismatch")
    out = np.empty_like(s)
    _lib.call("mpc_cuda_beaver_mask", _p(s), _p(x), _p(out), C.c_size_t(s.size // 4))
    return out


def beaver_combine(x, y, z, sx, oy, is_leader, spdz=False):
    x, y, z, sx, oy = _a(x, 4), _a(y, 4), _a(z, 4), _a(sx, 4), _a(oy, 4)
    n = sx.size // 4
    if x.size // 4 != (2 * n if spdz else n) or y.shape != x.shape or z.shape != x.shape or oy.shape != sx.shape:
        raise ValueError("shape mismatch")
    out = np.empty_like(x)
    _lib.call("mpc_cuda_beaver_combine", _p(x), _p(y), _p(z), _p(sx), _p(oy), _p(out), C.c_size_t(n),
              C.c_uint32(int(bool(is_leader))), C.c_uint32(int(bool(spdz))))
    return out


def open_sum(parts):
    parts = _a(parts, 4)
    P, n = parts.shape[0], parts.shape[1]
    out = np.empty((n, 4), dtype=np.uint64)
    _lib.call("mpc_cuda_open_sum", _p(parts), C.c_uint32(P), _p(out), C.c_size_t(n))
    return out


def spdz_mac_check(vals, macs, is_leader):
    vals, macs = _a(vals, 4), _a(macs, 4)
    out = np.empty_like(vals)
    _lib.call("mpc_cuda_spdz_mac_check", _p(vals), _p(macs), _p(out), C.c_size_t(vals.size // 4),
              C.c_uint32(int(bool(is_leader))))
    return out


VEC_OP = {"sub": 0, "mul": 1, "mul_const": 2, "axpy": 3}


def vec_op(op, a, b=None, c=None):
    a = _a(a, 4)
    b = _a(b, 4) if b is not None else None
    c = _a(c, 4) if c is not None else None
    out = np.empty_like(a)
    _lib.call("mpc_cuda_vec_op", C.c_uint32(VEC_OP[op]), _p(a), _p(b), _p(c), _p(out), C.c_size_t(a.size // 4))
    return out
