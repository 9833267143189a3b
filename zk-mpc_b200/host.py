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
_FOPS = {"add": 0, "sub": 1, "mul": 2, "neg": 3, "inv": 4, "from_mont": 5, "to_mont": 6, "sqr": 7, "mul_narrow": 8,
         "inv_euclid": 9}


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
    if parts.ndim != 3:
        raise ValueError("open_sum expects (parties, n, 4), got shape %s" % (parts.shape,))
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


def set_option(name, value):
    _lib.call("mpc_cuda_set_option", name.encode(), C.c_int64(int(value)))


# ----------------------------------------------------------------------------- share NTT
NTT_KIND = {"fft": 0, "ifft": 1, "coset_fft": 2, "coset_ifft": 3}


def ntt(data, kind, batch=1):
    """Radix2EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place on `batch` back-to-back vectors
    (host arrays; returns a new array).  Like the reference, the caller zero-pads to the domain size."""
    data = _a(data, 4).copy()
    total = data.size // 4
    if batch < 1 or total % batch:
        raise ValueError("batch does not divide the element count")
    n = total // batch
    log_n = n.bit_length() - 1
    if n == 0 or (1 << log_n) != n:
        raise ValueError("domain size must be a power of two, got %d" % n)
    _lib.call("mpc_cuda_ntt_fr", _p(data), C.c_uint32(log_n), C.c_uint32(NTT_KIND[kind] if isinstance(kind, str) else kind),
              C.c_uint32(batch))
    return data


def divide_by_vanishing_on_coset(data):
    data = _a(data, 4).copy()
    n = data.size // 4
    if n == 0 or n & (n - 1):
        raise ValueError("domain size must be a power of two, got %d" % n)
    _lib.call("mpc_cuda_divide_by_vanishing_on_coset", _p(data), C.c_uint32(n.bit_length() - 1))
    return data


# ----------------------------------------------------------------------------- share MSM
def _inf(inf, n):
    if inf is None:
        return None, None
    inf = np.ascontiguousarray(inf, dtype=np.uint8)
    if inf.size < n:
        raise ValueError("infinity flags shorter than bases")
    return inf, inf.ctypes.data_as(u8p)


def _msm(fn, limbs, bases_xy, scalars_mont, inf):
    bases_xy, scalars_mont = _a(bases_xy, limbs), _a(scalars_mont, 4)
    n = min(bases_xy.size // limbs, scalars_mont.size // 4)       # variable_base.rs:16-18
    keep, pinf = _inf(inf, n)
    out = np.zeros(limbs, dtype=np.uint64)
    oinf = C.c_uint8(0)
    _lib.call(fn, _p(bases_xy), pinf, _p(scalars_mont), C.c_size_t(n), _p(out), C.byref(oinf))
    return out, oinf.value


def msm_g1(bases_xy, scalars_mont, inf=None):
    """AffineMsm::msm over G1 (mpc-algebra/src/share/msm.rs:33-37): affine result + infinity flag"""
    return _msm("mpc_cuda_msm_g1", 12, bases_xy, scalars_mont, inf)


def msm_g2(bases_xy, scalars_mont, inf=None):
    return _msm("mpc_cuda_msm_g2", 24, bases_xy, scalars_mont, inf)


class BaseHandle:
    """A CRS vector kept resident on the calling thread's device (pk.*_query, powers_of_g)."""

    def __init__(self, handle, n, g2):
        self.handle, self.n, self.g2 = handle, n, g2

    def precompute(self, window_bits=0):
        """build the 2^(c*w)*P table on the device (one shared bucket set per MSM afterwards)"""
        _lib.call("mpc_cuda_msm_g2_precompute" if self.g2 else "mpc_cuda_msm_g1_precompute", C.c_uint64(self.handle),
                  C.c_uint32(window_bits))
        return self

    def release(self):
        if self.handle:
            _lib.call("mpc_cuda_msm_release_bases", C.c_uint64(self.handle))
            self.handle = 0


def register_bases(bases_xy, inf=None, g2=False, parts=0):
    """parts > 0: point-range sharding over the devices of the init list inside this process (range k on
    device k mod n_dev); msm_handle / precompute then drive every part concurrently."""
    limbs = 24 if g2 else 12
    bases_xy = _a(bases_xy, limbs)
    n = bases_xy.size // limbs
    keep, pinf = _inf(inf, n)
    h = C.c_uint64(0)
    if parts:
        _lib.call("mpc_cuda_msm_g2_register_bases_sharded" if g2 else "mpc_cuda_msm_g1_register_bases_sharded",
                  _p(bases_xy), pinf, C.c_size_t(n), C.c_uint32(parts), C.byref(h))
    else:
        _lib.call("mpc_cuda_msm_g2_register_bases" if g2 else "mpc_cuda_msm_g1_register_bases", _p(bases_xy), pinf,
                  C.c_size_t(n), C.byref(h))
    return BaseHandle(h.value, n, g2)


def msm_handle(handle, scalars_mont, offset=0, n=None):
    scalars_mont = _a(scalars_mont, 4)
    if n is None:
        n = min(handle.n - offset, scalars_mont.size // 4)
    limbs = 24 if handle.g2 else 12
    out = np.zeros(limbs, dtype=np.uint64)
    oinf = C.c_uint8(0)
    _lib.call("mpc_cuda_msm_g2_handle" if handle.g2 else "mpc_cuda_msm_g1_handle", C.c_uint64(handle.handle),
              C.c_size_t(offset), _p(scalars_mont), C.c_size_t(n), _p(out), C.byref(oinf))
    return out, oinf.value


def multi_scale_pub_group(bases_xy, share_planes, inf=None, g2=False):
    """GroupShare::multi_scale_pub_group.  Additive shares (mpc-algebra/src/share/additive.rs:517-520):
    `share_planes` is (n,4) and the result one point.  SPDZ shares (share/spdz.rs:482-488): `share_planes`
    is (2,n,4) = [sh, mac]; the reference builds BOTH scalar vectors from `sh` (line 484 reads s.sh.val for
    the macs), so the mac component of the result equals the sh component — reproduced here by
    computing the MSM once."""
    planes = np.ascontiguousarray(share_planes, dtype=np.uint64)
    fn = msm_g2 if g2 else msm_g1
    if planes.ndim == 2:
        return fn(bases_xy, planes, inf)
    sh = fn(bases_xy, planes[0], inf)
    return sh, sh


# ----------------------------------------------------------------------------- device-resident helpers
class DeviceBuffer:
    def __init__(self, nbytes):
        self.ptr = C.c_void_p(0)
        self.nbytes = nbytes
        _lib.call("mpc_cuda_malloc", C.byref(self.ptr), C.c_size_t(nbytes))

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        _lib.call("mpc_cuda_memcpy_h2d", self.ptr, arr.ctypes.data_as(C.c_void_p), C.c_size_t(arr.nbytes), None)
        _lib.call("mpc_cuda_stream_sync", None)
        return self

    def download(self, dtype=np.uint64, count=None):
        out = np.empty(self.nbytes // np.dtype(dtype).itemsize if count is None else count, dtype=dtype)
        _lib.call("mpc_cuda_memcpy_d2h", out.ctypes.data_as(C.c_void_p), self.ptr, C.c_size_t(out.nbytes), None)
        _lib.call("mpc_cuda_stream_sync", None)
        return out

    def u64(self):
        return C.cast(self.ptr, u64p)

    def free(self):
        if self.ptr:
            _lib.call("mpc_cuda_free", self.ptr)
            self.ptr = C.c_void_p(0)


def g1_generate(seed, n, first=0):
    """synthetic CRS on the device: bases[i] = k_i * G1 generator (same schedule as the oracle's generator)"""
    buf = DeviceBuffer(max(n, 1) * 96)
    _lib.call("mpc_cuda_g1_generate_dev", C.c_uint64(seed), C.c_size_t(first), C.c_size_t(n), buf.u64(), None)
    _lib.call("mpc_cuda_stream_sync", None)
    return buf


def g2_generate(seed, n, first=0):
    buf = DeviceBuffer(max(n, 1) * 192)
    _lib.call("mpc_cuda_g2_generate_dev", C.c_uint64(seed), C.c_size_t(first), C.c_size_t(n), buf.u64(), None)
    _lib.call("mpc_cuda_stream_sync", None)
    return buf


def register_bases_dev(buf, n, g2=False):
    h = C.c_uint64(0)
    _lib.call("mpc_cuda_msm_g2_register_bases_dev" if g2 else "mpc_cuda_msm_g1_register_bases_dev", buf.u64(),
              C.c_size_t(n), C.byref(h))
    return BaseHandle(h.value, n, g2)


def msm_handle_dev(handle, scalars_buf, n, offset=0, scalar_offset=0, out=None, stream=None):
    """resident scalars -> Jacobian partial left on the device (DeviceBuffer of 3 x 6 (G1) / 3 x 12 (G2) limbs);
    asynchronous on `stream` (None = the library's per-thread stream)"""
    limbs = 36 if handle.g2 else 18
    out = out or DeviceBuffer(limbs * 8)
    sc = C.cast(C.c_void_p(scalars_buf.ptr.value + 32 * scalar_offset), u64p)
    _lib.call("mpc_cuda_msm_g2_handle_dev" if handle.g2 else "mpc_cuda_msm_g1_handle_dev", C.c_uint64(handle.handle),
              C.c_size_t(offset), sc, C.c_size_t(n), out.u64(), C.c_void_p(stream) if stream else None)
    return out


def sum_partials(buf, count, g2=False, stream=None):
    """affine(sum of `count` Jacobian partials resident in `buf`)"""
    limbs = 24 if g2 else 12
    out = np.zeros(limbs, dtype=np.uint64)
    oinf = C.c_uint8(0)
    _lib.call("mpc_cuda_g2_sum_partials_dev" if g2 else "mpc_cuda_g1_sum_partials_dev", buf.u64(), C.c_uint32(count),
              _p(out), C.byref(oinf), C.c_void_p(stream) if stream else None)
    return out, oinf.value


def launch_count():
    return int(_lib.lib().mpc_cuda_launch_count())


def profile_read(name):
    """(total device ms, intervals) of a profiled stage since the last read (option "profile" = 1)"""
    ms, cnt = C.c_double(0), C.c_uint64(0)
    _lib.call("mpc_cuda_profile_read", name.encode(), C.byref(ms), C.byref(cnt))
    return ms.value, cnt.value


def ntt_dev(ptr, log_n, kind, batch=1, stream=None):
    """in-place NTT of device-resident data (ptr: integer device address)"""
    _lib.call("mpc_cuda_ntt_fr_dev", C.cast(ptr, u64p), C.c_uint32(log_n), C.c_uint32(NTT_KIND[kind]), C.c_uint32(batch),
              C.c_void_p(stream) if stream else None)


def ntt_sharded_dev(block_ptrs, log_n, kind, dev_index=None):
    """one party's 2^log_n transform over len(block_ptrs) = 2^log_g blocks resident on the devices dev_index[q]
    (None = q) of this process; asynchronous, ordered before later work on the first device's stream"""
    g = len(block_ptrs)
    log_g = g.bit_length() - 1
    if (1 << log_g) != g:
        raise ValueError("the number of blocks must be a power of two")
    arr = (u64p * g)(*[C.cast(C.c_void_p(int(p)), u64p) for p in block_ptrs])
    dv = (C.c_int32 * g)(*dev_index) if dev_index is not None else None
    _lib.call("mpc_cuda_ntt_fr_sharded_dev", arr, dv, C.c_uint32(log_n), C.c_uint32(log_g), C.c_uint32(NTT_KIND[kind]))


def ntt_reorder_sharded_dev(in_ptrs, out_ptrs, log_n, to_transposed, dev_index=None):
    g = len(in_ptrs)
    log_g = g.bit_length() - 1
    a = (u64p * g)(*[C.cast(C.c_void_p(int(p)), u64p) for p in in_ptrs])
    b = (u64p * g)(*[C.cast(C.c_void_p(int(p)), u64p) for p in out_ptrs])
    dv = (C.c_int32 * g)(*dev_index) if dev_index is not None else None
    _lib.call("mpc_cuda_ntt_reorder_sharded_dev", a, b, dv, C.c_uint32(log_n), C.c_uint32(log_g),
              C.c_uint32(int(bool(to_transposed))))


def ntt_sharded(data, kind, log_g):
    """host vector in natural order in and out, computed on 2^log_g devices of this process"""
    data = _a(data, 4).copy()
    n = data.size // 4
    log_n = n.bit_length() - 1
    if n == 0 or (1 << log_n) != n:
        raise ValueError("domain size must be a power of two, got %d" % n)
    _lib.call("mpc_cuda_ntt_fr_sharded", _p(data), C.c_uint32(log_n), C.c_uint32(NTT_KIND[kind]), C.c_uint32(log_g))
    return data


def ntt_cross_stage_dev(ptr, log_n, log_g, slice_offset, slice_len, kind, stream=None):
    _lib.call("mpc_cuda_ntt_cross_stage_dev", C.cast(ptr, u64p), C.c_uint32(log_n), C.c_uint32(log_g),
              C.c_size_t(slice_offset), C.c_size_t(slice_len), C.c_uint32(NTT_KIND[kind]),
              C.c_void_p(stream) if stream else None)


# ----------------------------------------------------------------------------- fused witness map
def _planes(v, spdz):
    """(n,4) additive or (2,n,4) SPDZ [sh, mac] -> contiguous array, n"""
    v = _a(v, 4)
    if spdz:
        if v.ndim != 3 or v.shape[0] != 2:
            raise ValueError("SPDZ share vectors are (2, n, 4) = [sh plane, mac plane], got %s" % (v.shape,))
        return v, v.shape[1]
    if v.ndim != 2:
        raise ValueError("additive share vectors are (n, 4), got %s" % (v.shape,))
    return v, v.shape[0]


class WitnessMapState:
    """handle of a begun witness_map: remembers the domain size so finish can validate its inputs"""

    def __init__(self, state, n, spdz):
        self.state, self.n, self.spdz = state, n, spdz
        self.h_ptr = None

    def release(self):
        if self.state:
            _lib.call("mpc_cuda_witness_map_release", C.c_uint64(self.state))
            self.state = 0


def witness_map_begin(a, b, c, tx, ty, spdz=False):
    """first half of R1CStoQAP::witness_map (src/groth16.rs:278-285) on the device; returns
    (masked_a, masked_b, state) — the masked vectors are what the party broadcasts for the two opens.
    spdz: every vector is (2,n,4) = [sh, mac] planes (mpc-algebra/src/share/spdz.rs:50-53)."""
    (a, n), (b, _), (c, _), (tx, _), (ty, _) = (_planes(v, spdz) for v in (a, b, c, tx, ty))
    log_n = n.bit_length() - 1
    if n == 0 or (1 << log_n) != n or any(v.shape != a.shape for v in (b, c, tx, ty)):
        raise ValueError("witness_map needs five equal power-of-two vectors")
    ma, mb = np.empty_like(a), np.empty_like(a)
    st = C.c_uint64(0)
    _lib.call("mpc_cuda_witness_map_begin_ex", _p(a), _p(b), _p(c), C.c_uint32(log_n), _p(tx), _p(ty),
              C.c_uint32(int(bool(spdz))), _p(ma), _p(mb), C.byref(st))
    return ma, mb, WitnessMapState(st.value, n, bool(spdz))


def witness_map_begin_r1cs(csr_a, csr_b, csr_c, assignment, num_inputs, log_n, tx, ty, spdz=False, masked=True):
    """the same from the assignment: a = A z, b = B z, c = C z on the device (evaluate_constraint,
    src/groth16.rs:205-234,263-276,289-293); assignment = instance | witness local values, (cols,4) or (2,cols,4).
    masked=False: the masked vectors stay on the device (returned as None); the opens then go through
    witness_map_masked_payload / witness_map_open_payloads."""
    z, cols = _planes(assignment, spdz)
    (tx, n), (ty, _) = _planes(tx, spdz), _planes(ty, spdz)
    if n != 1 << log_n or ty.shape != tx.shape or cols != csr_a.cols:
        raise ValueError("witness_map_begin_r1cs: triple planes must have 2^log_n elements and the assignment %d" % csr_a.cols)
    ma, mb = (np.empty_like(tx), np.empty_like(tx)) if masked else (None, None)
    st = C.c_uint64(0)
    _lib.call("mpc_cuda_witness_map_begin_r1cs", C.c_uint64(csr_a.handle), C.c_uint64(csr_b.handle), C.c_uint64(csr_c.handle),
              _p(z), C.c_size_t(num_inputs), C.c_uint32(log_n), _p(tx), _p(ty), C.c_uint32(int(bool(spdz))), _p(ma), _p(mb),
              C.byref(st))
    return ma, mb, WitnessMapState(st.value, n, bool(spdz))


def _finish_args(state, tz, sx, oy):
    if not isinstance(state, WitnessMapState) or not state.state:
        raise MpcCudaError("witness_map state already consumed")
    (tz, n), sx, oy = _planes(tz, state.spdz), (_a(sx, 4) if sx is not None else None), (_a(oy, 4) if oy is not None else None)
    if n != state.n or any(v is not None and v.shape != (state.n, 4) for v in (sx, oy)):
        raise ValueError("witness_map_finish: tz / sx / oy must match the begun domain of %d elements" % state.n)
    return tz, sx, oy                         # sx / oy None: opened through witness_map_open_payloads


def witness_map_finish(state, tz, sx, oy, is_leader):
    """second half (src/groth16.rs:285-303): Beaver combine, - c, / Z_H on the coset, coset iFFT"""
    tz, sx, oy = _finish_args(state, tz, sx, oy)
    h = np.empty_like(tz)
    st, state.state = state.state, 0            # finish always releases the state
    _lib.call("mpc_cuda_witness_map_finish", C.c_uint64(st), _p(tz), _p(sx), _p(oy),
              C.c_uint32(int(bool(is_leader))), _p(h))
    return h


def witness_map_finish_dev(state, tz, sx, oy, is_leader):
    """like witness_map_finish, but h stays on the device: returns its device address (planes x n elements,
    valid until state.release()) for mpc_cuda_msm_g1_handle_scalars_dev / msm_handle_dev"""
    tz, sx, oy = _finish_args(state, tz, sx, oy)
    ptr = u64p()
    _lib.call("mpc_cuda_witness_map_finish_dev", C.c_uint64(state.state), _p(tz), _p(sx), _p(oy),
              C.c_uint32(int(bool(is_leader))), C.byref(ptr))
    state.h_ptr = C.cast(ptr, C.c_void_p).value
    return state.h_ptr


class PinnedBuffer:
    """page-locked host memory (mpc_cuda_host_alloc) viewed as numpy arrays: wire payloads and triple shares cross
    PCIe at the link rate from here, pageable arrays at a fraction of it"""

    def __init__(self, nbytes):
        self.nbytes = int(nbytes)
        self.ptr = C.c_void_p(0)
        _lib.call("mpc_cuda_host_alloc", C.byref(self.ptr), C.c_size_t(max(self.nbytes, 1)))

    def array(self, dtype=np.uint8, count=None, offset=0):
        item = np.dtype(dtype).itemsize
        count = (self.nbytes - offset) // item if count is None else count
        if offset + count * item > self.nbytes:
            raise ValueError("view exceeds the pinned buffer")
        raw = (C.c_uint8 * (count * item)).from_address(self.ptr.value + offset)
        return np.frombuffer(raw, dtype=dtype, count=count)

    def free(self):
        if self.ptr.value:
            _lib.call("mpc_cuda_host_free", self.ptr)
            self.ptr = C.c_void_p(0)


def _payload_out(n, out):
    if out is None:
        return np.zeros(8 + 32 * n, dtype=np.uint8)
    if out.dtype != np.uint8 or out.size != 8 + 32 * n or not out.flags.c_contiguous:
        raise ValueError("payload buffer must be %d contiguous bytes" % (8 + 32 * n))
    return out


def _payload_ptrs(payloads, n):
    arrs = [np.ascontiguousarray(p, dtype=np.uint8) for p in payloads]
    if not arrs or any(a.size != 8 + 32 * n for a in arrs):
        raise ValueError("expected payloads of %d bytes" % (8 + 32 * n))
    return arrs, (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])


def witness_map_masked_payload(state, which, out=None):
    """the party's broadcast payload of masked_a (which=0) / masked_b (1), straight from the device"""
    out = _payload_out(state.n, out)
    _lib.call("mpc_cuda_witness_map_masked_payload", C.c_uint64(state.state), C.c_uint32(which), out.ctypes.data_as(u8p))
    return out


def witness_map_open_payloads(state, which, payloads):
    """every party's payload of open `which` summed on the device; the opened vector stays in the state"""
    arrs, ptrs = _payload_ptrs(payloads, state.n)
    _lib.call("mpc_cuda_witness_map_open_payloads", C.c_uint64(state.state), C.c_uint32(which), ptrs, C.c_uint32(len(arrs)))


def witness_map_mac_payload(state, which, is_leader, out=None):
    out = _payload_out(state.n, out)
    _lib.call("mpc_cuda_witness_map_mac_payload", C.c_uint64(state.state), C.c_uint32(which), C.c_uint32(int(bool(is_leader))),
              out.ctypes.data_as(u8p))
    return out


def witness_map_mac_verify(state, payloads):
    """raises MpcCudaError unless the parties' dx payloads sum to zero everywhere (spdz.rs:177-196)"""
    arrs, ptrs = _payload_ptrs(payloads, state.n)
    _lib.call("mpc_cuda_witness_map_mac_verify", C.c_uint64(state.state), ptrs, C.c_uint32(len(arrs)))


def witness_map_assignment_dev(state):
    """device address and length (elements per plane) of the assignment uploaded by witness_map_begin_r1cs"""
    ptr, cols = u64p(), C.c_size_t(0)
    _lib.call("mpc_cuda_witness_map_assignment_dev", C.c_uint64(state.state), C.byref(ptr), C.byref(cols))
    return C.cast(ptr, C.c_void_p).value, cols.value


def msm_handle_scalars_dev(handle, scalars_ptr, n, offset=0):
    """scalars resident at device address `scalars_ptr`, affine result + infinity flag on the host"""
    limbs = 24 if handle.g2 else 12
    out = np.zeros(limbs, dtype=np.uint64)
    oinf = C.c_uint8(0)
    _lib.call("mpc_cuda_msm_g2_handle_scalars_dev" if handle.g2 else "mpc_cuda_msm_g1_handle_scalars_dev",
              C.c_uint64(handle.handle), C.c_size_t(offset), C.cast(C.c_void_p(int(scalars_ptr)), u64p), C.c_size_t(n),
              _p(out), C.byref(oinf))
    return out, oinf.value


# ----------------------------------------------------------------------------- linear steps
class CsrMatrix:
    """a public sparse matrix kept resident (the A, B, C matrices of the R1CS, Marlin's index matrices)"""

    def __init__(self, row_ptr, col, coeff, cols):
        row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint64)
        col = np.ascontiguousarray(col, dtype=np.uint32)
        coeff = _a(coeff, 4)
        if row_ptr.ndim != 1 or row_ptr.size < 1 or col.size != int(row_ptr[-1]) or coeff.size // 4 != col.size:
            raise ValueError("malformed CSR arrays")
        self.rows, self.cols, self.nnz = row_ptr.size - 1, int(cols), int(col.size)
        h = C.c_uint64(0)
        _lib.call("mpc_cuda_csr_register", _p(row_ptr), col.ctypes.data_as(C.POINTER(C.c_uint32)), _p(coeff),
                  C.c_size_t(self.rows), C.c_size_t(self.cols), C.byref(h))
        self.handle = h.value

    def spmv(self, x):
        """rows of evaluate_constraint on local values: x (cols,4) -> (rows,4); (planes,cols,4) -> (planes,rows,4)"""
        x = _a(x, 4)
        planes = 1 if x.ndim == 2 else x.shape[0]
        if x.shape[-2] != self.cols:
            raise ValueError("vector length %d != matrix columns %d" % (x.shape[-2], self.cols))
        out = np.empty(x.shape[:-2] + (self.rows, 4), dtype=np.uint64)
        _lib.call("mpc_cuda_csr_spmv", C.c_uint64(self.handle), _p(x), C.c_uint32(planes), _p(out))
        return out

    def release(self):
        if self.handle:
            _lib.call("mpc_cuda_csr_release", C.c_uint64(self.handle))
            self.handle = 0


def fr_serialize(vals):
    """CanonicalSerialize of a Vec<Fr> (what MpcSerNet::broadcast sends): uint8 array of 8 + 32 n bytes"""
    vals = _a(vals, 4)
    n = vals.size // 4
    out = np.zeros(8 + 32 * n, dtype=np.uint8)
    _lib.call("mpc_cuda_fr_serialize", _p(vals), C.c_size_t(n), out.ctypes.data_as(u8p))
    return out


def fr_deserialize(buf, n):
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    if buf.size != 8 + 32 * n:
        raise ValueError("payload of %d bytes cannot hold %d elements" % (buf.size, n))
    out = np.empty((n, 4), dtype=np.uint64)
    _lib.call("mpc_cuda_fr_deserialize", buf.ctypes.data_as(u8p), C.c_size_t(n), _p(out))
    return out


def beaver_mask_serialize(s, x):
    """wire payload of the masked vector s + x (share/field.rs:108-117 + channel.rs:12-28) in one pass"""
    s, x = _a(s, 4), _a(x, 4)
    if s.shape != x.shape:
        raise ValueError("shape mismatch")
    n = s.size // 4
    out = np.zeros(8 + 32 * n, dtype=np.uint8)
    _lib.call("mpc_cuda_beaver_mask_serialize", _p(s), _p(x), C.c_size_t(n), out.ctypes.data_as(u8p))
    return out


def open_sum_deserialize(payloads, n):
    """batch_open's local half straight from the received wire payloads (one per party, 8 + 32 n bytes each)"""
    payloads = np.ascontiguousarray(payloads, dtype=np.uint8)
    if payloads.ndim != 2 or payloads.shape[1] != 8 + 32 * n:
        raise ValueError("expected (parties, %d) bytes, got %s" % (8 + 32 * n, payloads.shape))
    out = np.empty((n, 4), dtype=np.uint64)
    _lib.call("mpc_cuda_open_sum_deserialize", payloads.ctypes.data_as(u8p), C.c_uint32(payloads.shape[0]), C.c_size_t(n),
              _p(out))
    return out


def poly_div_linear(coeffs, z):
    """(quotient, remainder) of p(x) / (x - z) on local share values: quotient (n-1,4), remainder (4,) = p(z)"""
    coeffs, z = _a(coeffs, 4), _a(z, 4)
    n = coeffs.size // 4
    if n < 1:
        raise ValueError("empty polynomial")
    q = np.empty((n - 1, 4), dtype=np.uint64)
    rem = np.empty(4, dtype=np.uint64)
    _lib.call("mpc_cuda_poly_div_linear", _p(coeffs), C.c_size_t(n), _p(z), _p(q) if n > 1 else None, _p(rem))
    return q, rem


def poly_evaluate(coeffs, z):
    coeffs, z = _a(coeffs, 4), _a(z, 4)
    rem = np.empty(4, dtype=np.uint64)
    _lib.call("mpc_cuda_poly_div_linear", _p(coeffs), C.c_size_t(coeffs.size // 4), _p(z), None, _p(rem))
    return rem


def poly_div_vanishing(coeffs, m):
    """(quotient, remainder) of p(x) / (x^m - 1) on local share values: quotient (max(n-m,0),4), remainder (m,4)"""
    coeffs = _a(coeffs, 4)
    n, m = coeffs.size // 4, int(m)
    if n < 1 or m < 1:
        raise ValueError("empty polynomial or domain")
    q = np.empty((max(n - m, 0), 4), dtype=np.uint64)
    rem = np.empty((m, 4), dtype=np.uint64)
    _lib.call("mpc_cuda_poly_div_vanishing", _p(coeffs), C.c_size_t(n), C.c_size_t(m), _p(q) if n > m else None, _p(rem))
    return q, rem


def poly_mul_vanishing(coeffs, m):
    """p(x) (x^m - 1) on local share values: (n + m, 4)"""
    coeffs = _a(coeffs, 4)
    n, m = coeffs.size // 4, int(m)
    if n < 1 or m < 1:
        raise ValueError("empty polynomial or domain")
    out = np.empty((n + m, 4), dtype=np.uint64)
    _lib.call("mpc_cuda_poly_mul_vanishing", _p(coeffs), C.c_size_t(n), C.c_size_t(m), _p(out))
    return out


# ----------------------------------------------------------------------------- device-pointer forms (resident sessions)
def _dp(ptr):
    return C.c_void_p(int(ptr)) if ptr is not None else None


def dev_zero(ptr, nbytes, stream=None):
    _lib.call("mpc_cuda_memset_zero_dev", _dp(ptr), C.c_size_t(nbytes), stream)


def dev_copy(dst, src, nbytes, stream=None):
    _lib.call("mpc_cuda_memcpy_d2d", _dp(dst), _dp(src), C.c_size_t(nbytes), stream)


def dev_copy2d(dst, dpitch, src, spitch, width, height, stream=None):
    _lib.call("mpc_cuda_memcpy2d_d2d", _dp(dst), C.c_size_t(dpitch), _dp(src), C.c_size_t(spitch), C.c_size_t(width),
              C.c_size_t(height), stream)


def dev_upload(dst, arr, stream=None):
    arr = np.ascontiguousarray(arr)
    _lib.call("mpc_cuda_memcpy_h2d", _dp(dst), arr.ctypes.data_as(C.c_void_p), C.c_size_t(arr.nbytes), stream)
    return arr                                  # keep the source alive until the caller synchronises


def dev_download(src, count, dtype=np.uint64, stream=None):
    out = np.empty(count, dtype=dtype)
    _lib.call("mpc_cuda_memcpy_d2h", out.ctypes.data_as(C.c_void_p), _dp(src), C.c_size_t(out.nbytes), stream)
    _lib.call("mpc_cuda_stream_sync", stream)
    return out


def dev_vec_op(op, a, b, c, out, n, stream=None):
    c = _a(c, 4) if c is not None else None
    _lib.call("mpc_cuda_vec_op_dev", C.c_uint32(VEC_OP[op]), _dp(a), _dp(b), _p(c), _dp(out), C.c_size_t(n), stream)


def dev_inverse(a, out, n, stream=None):
    _lib.call("mpc_cuda_fr_inverse_dev", _dp(a), _dp(out), C.c_size_t(n), stream)


def dev_spmv(csr, x_ptr, out_ptr, stream=None):
    _lib.call("mpc_cuda_csr_spmv_dev", C.c_uint64(csr.handle), _dp(x_ptr), C.c_size_t(csr.cols), C.c_uint32(1), _dp(out_ptr),
              C.c_size_t(csr.rows), stream)


def dev_poly_div_vanishing(p, n, m, q, rem, stream=None):
    _lib.call("mpc_cuda_poly_div_vanishing_dev", _dp(p), C.c_size_t(n), C.c_size_t(m), _dp(q), _dp(rem), stream)


def dev_poly_mul_vanishing(p, n, m, out, stream=None):
    _lib.call("mpc_cuda_poly_mul_vanishing_dev", _dp(p), C.c_size_t(n), C.c_size_t(m), _dp(out), stream)


def dev_poly_evaluate(p, n, z, rem, stream=None):
    _lib.call("mpc_cuda_poly_div_linear_dev", _dp(p), C.c_size_t(n), _p(_a(z, 4)), None, _dp(rem), stream)
