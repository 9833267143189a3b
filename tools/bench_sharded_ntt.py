#!/usr/bin/env python3
"""Strong scaling of ONE party's NTT over the GPUs of this process (mpc_cuda_ntt_fr_sharded_dev), with and without
the CUDA-graph replay; also checks the forward result against the single-GPU transform.
usage: tools/bench_sharded_ntt.py [log_n ...]"""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import __graft_entry__ as ge
pkg = ge.load_package(); H, S, L = pkg.host, pkg.synth, pkg._lib
H.init(); H.set_party(0, 1)
ndev = H.device_count()
g = 1
while g * 2 <= min(ndev, 8):
    g *= 2
log_g = g.bit_length() - 1


def bitrev(x, bits):
    return int(format(x, "0%db" % bits)[::-1], 2) if bits else 0


for log_n in [int(x) for x in sys.argv[1:]] or [24, 26]:
    n = 1 << log_n
    m = n // g
    chunk = S.fr_uniform(0x5EED + log_n, min(m, 1 << 20))
    bufs = []
    for q in range(g):
        H.set_device(q)
        b = H.DeviceBuffer(m * 32)
        for off in range(0, m, chunk.shape[0]):
            k = min(chunk.shape[0], m - off)
            L.call("mpc_cuda_memcpy_h2d", C.c_void_p(b.ptr.value + off * 32), chunk.ctypes.data_as(C.c_void_p), C.c_size_t(k * 32), None)
        L.call("mpc_cuda_stream_sync", None)
        bufs.append(b)
    H.set_device(0)
    ptrs = [b.ptr.value for b in bufs]
    row = {"log_n": log_n, "devices": g}
    # single-GPU reference time and (for sizes that fit) result check on a sample of outputs
    one = H.DeviceBuffer(n * 32)
    for q in range(g):
        for off in range(0, m, chunk.shape[0]):
            k = min(chunk.shape[0], m - off)
            L.call("mpc_cuda_memcpy_h2d", C.c_void_p(one.ptr.value + (q * m + off) * 32), chunk.ctypes.data_as(C.c_void_p), C.c_size_t(k * 32), None)
    L.call("mpc_cuda_stream_sync", None)
    H.ntt_dev(one.ptr.value, log_n, "fft")
    L.call("mpc_cuda_stream_sync", None)
    ref_head = np.empty((64, 4), dtype=np.uint64)
    L.call("mpc_cuda_memcpy_d2h", ref_head.ctypes.data_as(C.c_void_p), one.ptr, C.c_size_t(64 * 32), None)
    L.call("mpc_cuda_stream_sync", None)
    H.set_option("profile", 1)
    H.profile_read("ntt")
    for _ in range(3):
        H.ntt_dev(one.ptr.value, log_n, "fft")
    L.call("mpc_cuda_stream_sync", None)
    t, cnt = H.profile_read("ntt")
    H.set_option("profile", 0)
    row["ms_1gpu_fft"] = round(t / cnt, 4)
    one.free()
    H.ntt_sharded_dev(ptrs, log_n, "fft")
    L.call("mpc_cuda_stream_sync", None)
    # X[i] for i < 64 lives on device bitrev(i % g), local i // g
    got = np.empty((64, 4), dtype=np.uint64)
    for i in range(64):
        H.set_device(bitrev(i % g, log_g))
        L.call("mpc_cuda_memcpy_d2h", got[i].ctypes.data_as(C.c_void_p), C.c_void_p(bufs[bitrev(i % g, log_g)].ptr.value + (i // g) * 32), C.c_size_t(32), None)
        L.call("mpc_cuda_stream_sync", None)
    H.set_device(0)
    row["matches_1gpu"] = bool(np.array_equal(got, ref_head))
    for mode, name in ((2, "direct"), (0, "graph")):
        H.set_option("ntt_graph", mode)
        for kind in ("fft", "coset_ifft"):
            for _ in range(3):
                H.ntt_sharded_dev(ptrs, log_n, kind)
            L.call("mpc_cuda_stream_sync", None)
            reps = 10
            t0 = time.perf_counter()
            for _ in range(reps):
                H.ntt_sharded_dev(ptrs, log_n, kind)
            L.call("mpc_cuda_stream_sync", None)
            row["ms_%s_%s" % (kind, name)] = round((time.perf_counter() - t0) / reps * 1e3, 4)
    H.set_option("ntt_graph", 0)
    row["speedup_fft_graph"] = round(row["ms_1gpu_fft"] / row["ms_fft_graph"], 3)
    row["efficiency_fft_graph"] = round(row["speedup_fft_graph"] / g, 3)
    row["nvlink_bytes_per_gpu"] = 2 * (g - 1) * (m // g) * 32
    for b in bufs:
        b.free()
    print(json.dumps(row), flush=True)
