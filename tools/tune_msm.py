#!/usr/bin/env python3
"""Sweep the MSM window width / task length on the current GPU: prints ms per stage.
usage: tools/tune_msm.py <log_n> [c_lo c_hi] [task_len ...]"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import __graft_entry__ as ge
pkg = ge.load_package(); H, S, L = pkg.host, pkg.synth, pkg._lib
H.init(); H.set_party(0, 1)
log_n = int(sys.argv[1]); n = 1 << log_n
c_lo, c_hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (log_n - 8, log_n - 2)
task_lens = [int(x) for x in sys.argv[4:]] or [0]
seed = S.bench_seed(log_n)
G2 = bool(os.environ.get("MSM_G2"))
dev = (H.g2_generate if G2 else H.g1_generate)(seed, n); h = H.register_bases_dev(dev, n, g2=G2)
if os.environ.get("MSM_TABLE"): h.precompute(int(os.environ["MSM_TABLE"]))
sc = H.DeviceBuffer(n * 32).upload(S.fr_uniform(seed, n))
out = H.DeviceBuffer(288 if G2 else 144)
ref = None
H.set_option("profile", 1)
if os.environ.get("MSM_AFFINE"): H.set_option("msm_affine", int(os.environ["MSM_AFFINE"]))
if os.environ.get("MSM_REDUCE_CHUNK"): H.set_option("msm_reduce_chunk", int(os.environ["MSM_REDUCE_CHUNK"]))
for tl in task_lens:
    H.set_option("msm_task_len", tl)
    for c in range(max(3, c_lo), min(23, c_hi) + 1):
        H.set_option("msm_window_bits", c)
        for it in range(3):
            L.call("mpc_cuda_msm_g2_handle_dev" if G2 else "mpc_cuda_msm_g1_handle_dev", C.c_uint64(h.handle), C.c_size_t(0), sc.u64(), C.c_size_t(n), out.u64(), None)
            L.call("mpc_cuda_stream_sync", None)
            if it == 0:
                for nm in ("msm_total", "msm_sort", "msm_accumulate", "msm_reduce"): H.profile_read(nm)
        xy = np.zeros(24 if G2 else 12, dtype=np.uint64); inf = C.c_uint8(0)
        L.call("mpc_cuda_g2_sum_partials_dev" if G2 else "mpc_cuda_g1_sum_partials_dev", out.u64(), C.c_uint32(1), xy.ctypes.data_as(L.u64p), C.byref(inf), None)
        if ref is None: ref = xy.copy()
        st = {nm: round(H.profile_read(nm)[0] / 2, 3) for nm in ("msm_total", "msm_sort", "msm_accumulate", "msm_reduce")}
        print(json.dumps({"g2": G2, "log_n": log_n, "c": c, "task_len": tl, "same_result": bool(np.array_equal(xy, ref)), **st}), flush=True)
