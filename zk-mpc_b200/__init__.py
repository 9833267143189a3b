"""zk-mpc_b200 — B200-native share-MSM / share-NTT / Beaver kernels behind a C ABI.

Layout: csrc/ (hand-written sm_100a CUDA + the extern "C" boundary declared in
include/mpc_cuda.h), build.py (nvcc recipe), _lib.py (ctypes binding that fails loudly
when libmpc_cuda.so is absent), host.py (host-side mirror of the reference's operator
interface), wire.py (Public / Shared packing rules of MpcField / MpcGroup), groth16.py (the prover's hot path
composed from the ABI), synth.py (seeded synthetic inputs).  The directory name carries a hyphen, so
import it through `__graft_entry__.load_package()` (module name `zk_mpc_b200`).
"""
from . import synth  # noqa: F401
from . import build as build_recipe  # noqa: F401
from . import _lib  # noqa: F401
from . import host  # noqa: F401
from . import sharding  # noqa: F401
from . import wire  # noqa: F401
from . import groth16  # noqa: F401
from . import kzg  # noqa: F401
from . import marlin  # noqa: F401
