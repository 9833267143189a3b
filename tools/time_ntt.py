#!/usr/bin/env python3
"""NTT timings on the current GPU (device-resident data, profile events): ms per transform, for the compile-time-shaped
pass kernels and the generic one (option ntt_generic)."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import __graft_entry__ as ge
pkg = ge.load_package(); H, S, L = pkg.host, pkg.synth, pkg._lib
H.init(); H.set_party(0, 1)
H.set_option("profile", 1)
for log_n in [int(x) for x in sys.argv[1:]] or [16, 20, 22, 24]:
    n = 1 << log_n
    buf = H.DeviceBuffer(n * 32).upload(S.fr_uniform(log_n, n))
    for mb in (0, 1, 2):
        H.set_option("ntt_generic", 1 if mb == 1 else 0)
        H.set_option("ntt_occupancy", 1 if mb == 2 else 2 if mb == 0 else 0)
        out = {}
        for kind, name in ((0, "fft"), (1, "ifft"), (2, "coset_fft"), (3, "coset_ifft")):
            for it in range(6):
                L.call("mpc_cuda_ntt_fr_dev", buf.u64(), C.c_uint32(log_n), C.c_uint32(kind), C.c_uint32(1), None)
                L.call("mpc_cuda_stream_sync", None)
                if it == 1: H.profile_read("ntt")
            t, cnt = H.profile_read("ntt")
            out[name] = round(t / cnt, 4)
        print(json.dumps({"log_n": log_n, "variant": ["shaped_80regs", "generic", "shaped_64regs"][mb], **out, "Melem_s_fft": round(n / out["fft"] / 1e3, 1)}), flush=True)
    H.set_option("ntt_generic", 0)
    H.set_option("ntt_occupancy", 0)
    buf.free()
