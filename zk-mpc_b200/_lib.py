"""ctypes binding of libmpc_cuda.so (the C ABI declared in include/mpc_cuda.h).

No fallback of any kind: a missing library raises at load time and every non-zero status
raises MpcCudaError, mirroring the reference's panic-on-failure convention.
"""
import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmpc_cuda.so")
HEADER = os.path.join(HERE, "..", "include", "mpc_cuda.h")

u64p = C.POINTER(C.c_uint64)
u8p = C.POINTER(C.c_uint8)


class MpcCudaError(RuntimeError):
    pass


_lib = None


def declared_symbols():
    """every function name include/mpc_cuda.h declares (used by the CPU export test)"""
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mpc_cuda_\w+)\s*\(", text)))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MpcCudaError(
                "libmpc_cuda.so is missing (%s). Build it with `python zk-mpc_b200/build.py` "
                "or __graft_entry__.build(); there is no CPU fallback." % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.mpc_cuda_last_error.restype = C.c_char_p
        _lib.mpc_cuda_version.restype = C.c_char_p
        _lib.mpc_cuda_launch_count.restype = C.c_uint64
        for name in declared_symbols():
            fn = getattr(_lib, name, None)
            if fn is not None and name not in ("mpc_cuda_last_error", "mpc_cuda_version", "mpc_cuda_launch_count"):
                fn.restype = C.c_int32
    return _lib


def check(rc):
    if rc != 0:
        raise MpcCudaError("mpc_cuda error %d: %s" % (rc, lib().mpc_cuda_last_error().decode()))


def call(name, *args):
    check(getattr(lib(), name)(*args))
