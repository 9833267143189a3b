"""The C ABI consumed from C: include/mpc_cuda.h compiles as plain C (gcc -std=c99 -pedantic), the program links
against libmpc_cuda.so, and on a GPU box it runs MSM / NTT / Beaver through the boundary against the oracle."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_abi", "abi_check.c")
EXE = os.path.join(ROOT, "tests", "c_abi", "abi_check")


def _build(pkg, orc):
    pkg.build_recipe.build()
    lib_dir = os.path.dirname(pkg._lib.LIB_PATH)
    orc_dir = os.path.join(ROOT, "oracle")
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    cmd = [gcc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"), SRC, "-o", EXE,
           "-L", lib_dir, "-l:libmpc_cuda.so", "-L", orc_dir, "-l:libzkmpc_oracle.so",
           "-Wl,-rpath," + lib_dir, "-Wl,-rpath," + orc_dir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return EXE


def test_header_compiles_and_links_from_c(pkg, orc):
    exe = _build(pkg, orc)
    assert os.path.exists(exe)


@pytest.mark.gpu
def test_c_program_runs_the_hot_path(pkg, orc):
    exe = _build(pkg, orc)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failure(s)" in r.stdout
