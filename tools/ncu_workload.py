#!/usr/bin/env python3
"""The kernels of the headline step, once each at the bench size, for ncu captures (tools/ncu_capture.sh):
usage: tools/ncu_workload.py msm|ntt|combine [log_n]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import __graft_entry__ as ge
pkg = ge.load_package(); H, S, L = pkg.host, pkg.synth, pkg._lib
H.init(); H.set_party(0, 1)
what = sys.argv[1]
for env, opt in (("MSM_AFFINE", "msm_affine"), ("MSM_AFFINE_SPLIT", "msm_affine_split")):      # A/B captures
    if os.environ.get(env):
        H.set_option(opt, int(os.environ[env]))
log_n = int(sys.argv[2]) if len(sys.argv) > 2 else 24
n = 1 << log_n
seed = S.bench_seed(log_n)
if what == "msm":
    dev = H.g1_generate(seed, n)
    h = H.register_bases_dev(dev, n).precompute(0)
    sc = H.DeviceBuffer(n * 32).upload(S.fr_uniform(seed, n))
    out = H.DeviceBuffer(144)
    for _ in range(2):          # the capture skips the first call's launches (-s), profiling a warm one
        L.call("mpc_cuda_msm_g1_handle_dev", C.c_uint64(h.handle), C.c_size_t(0), sc.u64(), C.c_size_t(n), out.u64(), None)
        L.call("mpc_cuda_stream_sync", None)
elif what == "ntt":
    buf = H.DeviceBuffer(n * 32).upload(S.fr_uniform(seed, n))
    for kind in (0, 0, 3):
        L.call("mpc_cuda_ntt_fr_dev", buf.u64(), C.c_uint32(log_n), C.c_uint32(kind), C.c_uint32(1), None)
        L.call("mpc_cuda_stream_sync", None)
else:
    bufs = [H.DeviceBuffer(n * 32).upload(S.fr_uniform(seed + k, n)) for k in range(5)]
    out = H.DeviceBuffer(n * 32)
    for leader in (1, 1, 0):
        L.call("mpc_cuda_beaver_combine_dev", *[b.u64() for b in bufs], out.u64(), C.c_size_t(n), C.c_uint32(leader), C.c_uint32(0), None)
        L.call("mpc_cuda_stream_sync", None)
print("workload done")
