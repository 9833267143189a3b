#!/usr/bin/env python3
"""Host-buffer MSM (mpc_cuda_msm_g1: bases + scalars over PCIe, chunk-streamed) against the chunk count and the
batched-affine option: wall-clock ms per call from page-locked buffers.  usage: tools/e2e_ab.py [log_n]"""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import __graft_entry__ as ge
pkg = ge.load_package(); H, S, L = pkg.host, pkg.synth, pkg._lib
H.init(); H.set_party(0, 1)
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << log_n
seed = S.bench_seed(log_n)
dev = H.g1_generate(seed, n)
pb, ps = H.PinnedBuffer(n * 96), H.PinnedBuffer(n * 32)
bases = pb.array(np.uint64, n * 12)
L.call("mpc_cuda_memcpy_d2h", bases.ctypes.data_as(C.c_void_p), dev.ptr, C.c_size_t(n * 96), None)
L.call("mpc_cuda_stream_sync", None)
dev.free()
scal = ps.array(np.uint64, n * 4)
scal[:] = S.fr_uniform(seed, n).reshape(-1)
out = np.zeros(12, dtype=np.uint64); inf = C.c_uint8(0)
ref = None
STAGES = ("msm_total", "msm_sort", "msm_accumulate", "msm_reduce")
for chunks in [int(x) for x in os.environ.get("CHUNKS", "0,1,3,4,5,8").split(",")]:
    for aff in [int(x) for x in os.environ.get("AFFINE", "0").split(",")]:
        H.set_option("msm_host_chunks", chunks)
        H.set_option("msm_affine", aff)
        ts = []
        for it in range(4):
            t0 = time.perf_counter()
            L.call("mpc_cuda_msm_g1", bases.ctypes.data_as(L.u64p), None, scal.ctypes.data_as(L.u64p), C.c_size_t(n),
                   out.ctypes.data_as(L.u64p), C.byref(inf))
            ts.append((time.perf_counter() - t0) * 1e3)
        if ref is None: ref = out.copy()
        H.set_option("profile", 1)                   # one more call with the library's stage events on
        for nm in STAGES: H.profile_read(nm)
        L.call("mpc_cuda_msm_g1", bases.ctypes.data_as(L.u64p), None, scal.ctypes.data_as(L.u64p), C.c_size_t(n),
               out.ctypes.data_as(L.u64p), C.byref(inf))
        st = {nm: round(H.profile_read(nm)[0], 2) for nm in STAGES}
        H.set_option("profile", 0)
        print(json.dumps({"log_n": log_n, "chunks": chunks, "msm_affine": aff, "ms": round(min(ts[1:]), 2),
                          "Mpts_s": round(n / min(ts[1:]) / 1e3, 1), "same_result": bool(np.array_equal(out, ref)), "stage_ms_sum": st}), flush=True)
