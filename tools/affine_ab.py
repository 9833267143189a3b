#!/usr/bin/env python3
"""A/B of the batched-affine pre-reduction (option msm_affine: 1 = off, 2 = one round, 3 = two rounds, 0 = auto)
on resident MSMs: stage times per size, plain and table mode.  usage: tools/affine_ab.py [log_n ...]"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import __graft_entry__ as ge
pkg = ge.load_package(); H, S, L = pkg.host, pkg.synth, pkg._lib
H.init(); H.set_party(0, 1)
if os.environ.get("L2_FETCH"):                      # cudaLimitMaxL2FetchGranularity hint: 32, 64 or 128
    H.set_option("l2_fetch_granularity", int(os.environ["L2_FETCH"]))
MODES = ((False, True) if not os.environ.get("TABLE_ONLY") else (True,))
STAGES = ("msm_total", "msm_sort", "msm_accumulate", "msm_reduce")
for log_n in [int(x) for x in sys.argv[1:]] or [18, 20, 22, 24]:
    n = 1 << log_n
    seed = S.bench_seed(log_n)
    dev = H.g1_generate(seed, n)
    sc = H.DeviceBuffer(n * 32).upload(S.fr_uniform(seed, n))
    out = H.DeviceBuffer(144)
    for table in MODES:
        h = H.register_bases_dev(dev, n)
        if table:
            h.precompute(0)
        ref = None
        for mode, name, split in ((1, "off", 0), (2, "one_round", 0), (3, "two_rounds", 0), (3, "two_rounds_split", 1), (0, "auto", 0)):
            H.set_option("msm_affine", mode)
            H.set_option("msm_affine_split", split)
            H.set_option("profile", 1)
            for it in range(4):
                L.call("mpc_cuda_msm_g1_handle_dev", C.c_uint64(h.handle), C.c_size_t(0), sc.u64(), C.c_size_t(n), out.u64(), None)
                L.call("mpc_cuda_stream_sync", None)
                if it == 0:
                    for nm in STAGES: H.profile_read(nm)
            res = H.sum_partials(out, 1)
            if ref is None: ref = res
            st = {nm: round(H.profile_read(nm)[0] / 3, 3) for nm in STAGES}
            H.set_option("profile", 0)
            print(json.dumps({"log_n": log_n, "table": table, "affine": name, "same_result": bool(np.array_equal(res[0], ref[0])), **st}), flush=True)
        H.set_option("msm_affine", 0)
        H.set_option("msm_affine_split", 0)
        h.release()
    dev.free(); sc.free(); out.free()
