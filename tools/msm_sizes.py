#!/usr/bin/env python3
"""MSM timings across the BASELINE sizes on the current GPU: resident scalars with / without the window table,
the table build time, and the host-buffer call (bases + scalars over PCIe, streamed in chunks).
usage: tools/msm_sizes.py [log_n ...]   (JSON lines on stdout)"""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import __graft_entry__ as ge
pkg = ge.load_package(); H, S, L = pkg.host, pkg.synth, pkg._lib
H.init(); H.set_party(0, 1)
STAGES = ("msm_total", "msm_sort", "msm_accumulate", "msm_reduce")


def timed(fn, reps=3):
    fn()
    H.set_option("profile", 1)
    for nm in STAGES:
        H.profile_read(nm)
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    wall = (time.perf_counter() - t0) / reps * 1e3
    st = {nm: round(H.profile_read(nm)[0] / reps, 3) for nm in STAGES}
    H.set_option("profile", 0)
    return round(wall, 3), st


for log_n in [int(x) for x in sys.argv[1:]] or [13, 16, 18, 20, 22, 24]:
    n = 1 << log_n
    seed = S.bench_seed(log_n)
    dev = H.g1_generate(seed, n)
    plain = H.register_bases_dev(dev, n)
    sc_host = S.fr_uniform(seed, n)
    wl_host = S.fr_witness_like(seed + 1, n)
    sc = H.DeviceBuffer(n * 32).upload(sc_host)
    wl = H.DeviceBuffer(n * 32).upload(wl_host)
    out = H.DeviceBuffer(144)
    row = {"log_n": log_n}

    def resident(h, buf):
        L.call("mpc_cuda_msm_g1_handle_dev", C.c_uint64(h.handle), C.c_size_t(0), buf.u64(), C.c_size_t(n), out.u64(), None)
        H.sum_partials(out, 1)

    w, st = timed(lambda: resident(plain, sc))
    row["plain_uniform"] = {"wall_ms": w, **st}
    w, st = timed(lambda: resident(plain, wl))
    row["plain_witness_like"] = {"wall_ms": w, **st}
    tab = H.register_bases_dev(dev, n)
    H.set_option("profile", 1)
    t0 = time.perf_counter()
    tab.precompute(0)
    row["precompute_wall_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
    row["precompute_ms"] = round(H.profile_read("msm_precompute")[0], 2)
    H.set_option("profile", 0)
    w, st = timed(lambda: resident(tab, sc))
    row["table_uniform"] = {"wall_ms": w, **st}
    w, st = timed(lambda: resident(tab, wl))
    row["table_witness_like"] = {"wall_ms": w, **st}
    tab.release()
    # host-buffer call: pinned staging through torch if available
    bases_host = dev.download().reshape(n, 12)
    try:
        import torch
        bh = torch.from_numpy(bases_host.view(np.int64)).pin_memory().numpy().view(np.uint64)
        sh = torch.from_numpy(sc_host.view(np.int64)).pin_memory().numpy().view(np.uint64)
    except Exception:
        bh, sh = bases_host, sc_host
    for chunks in ([1, 2, 4, 8] if log_n >= 20 else [1]):
        for aff in ((1, 2, 3) if log_n >= 22 else (0,)):
            H.set_option("msm_host_chunks", chunks)
            H.set_option("msm_affine", aff)
            w, st = timed(lambda: H.msm_g1(bh, sh), reps=2)
            row["host_chunks_%d_affine_%d" % (chunks, aff)] = {"wall_ms": w, "Mpts_s": round(n / w / 1e3, 2), **st}
    H.set_option("msm_host_chunks", 0)
    H.set_option("msm_affine", 0)
    plain.release(); dev.free(); sc.free(); wl.free(); out.free()
    print(json.dumps(row), flush=True)
