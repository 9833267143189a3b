"""CPU-only checks of the drop-in boundary: libmpc_cuda.so builds for sm_100a, loads, exports every
function include/mpc_cuda.h declares, carries sm_100a SASS, and fails loudly without a GPU (no CPU
fallback).  No compute call is made here."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest


@pytest.fixture(scope="module")
def built(pkg):
    pkg.build_recipe.build()
    return pkg


def test_library_exports_every_declared_symbol(built):
    lib = built._lib.lib()
    names = built._lib.declared_symbols()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert b"sm_100a" in lib.mpc_cuda_version()


def test_library_contains_sm100a_sass_only(built):
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", built._lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = {line.split(".")[-2] for line in out.splitlines() if ".cubin" in line}
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the compute path is exercised by the -m gpu tests")
    H = built.host
    a = np.zeros((4, 4), dtype=np.uint64)
    for call in (lambda: H.beaver_mask(a, a), lambda: H.ntt(a, "fft"), lambda: H.msm_g1(np.zeros((4, 12), np.uint64), a)):
        with pytest.raises(H.MpcCudaError) as e:
            call()
        assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_argument_validation_messages(built):
    lib = built._lib.lib()
    # option names are validated before any device work
    rc = lib.mpc_cuda_set_option(b"no_such_option", C.c_int64(1))
    assert rc != 0 and b"unknown option" in lib.mpc_cuda_last_error()


def test_product_package_never_imports_the_oracle(built):
    import re
    root = os.path.dirname(os.path.abspath(built.__file__))
    for dirpath, _, files in os.walk(root):
        for f in files:
            text = open(os.path.join(dirpath, f), errors="ignore").read() if f.endswith((".py", ".cu", ".cuh")) else ""
            assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), f
            assert not re.search(r"#include\s+[\"<].*oracle", text), f
            assert "libzkmpc_oracle" not in text, f


def test_rust_ffi_block_covers_the_whole_header(built):
    """bindings/rust/mpc-cuda/src/ffi.rs is generated from include/mpc_cuda.h (tools/gen_rust_ffi.py): it must
    declare every function of the C ABI, and be up to date with the header"""
    import re
    import sys
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    ffi = open(os.path.join(root, "bindings", "rust", "mpc-cuda", "src", "ffi.rs")).read()
    declared = set(re.findall(r"pub fn (mpc_cuda_\w+)\(", ffi))
    assert declared == set(built._lib.declared_symbols())
    sys.path.insert(0, os.path.join(root, "tools"))
    import gen_rust_ffi
    before = ffi
    gen_rust_ffi.main()
    assert open(gen_rust_ffi.OUT).read() == before, "ffi.rs is stale: run tools/gen_rust_ffi.py"
