#!/usr/bin/env python
"""bench.py — share-MSM throughput of the B200 hot path (BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--log-n L] [--impl reference]

One "step" = one share MSM over G1: every rank owns n = 2^L points of an (N * n)-point MSM (point-range
sharding), computes its Jacobian partial with the Pippenger pipeline of libmpc_cuda.so,
the partials are all-gathered over NCCL and rank 0 adds and normalises them.  Scalars and bases are
resident in HBM when the timed region starts (`value`); `e2e` is the same MSM through the host-buffer
C ABI call (mpc_cuda_msm_g1: bases + scalars cross PCIe inside the timed region).  The JSON line also
carries the roofline of the dominant kernel (bucket accumulation, integer-pipe bound), the NTT and
Beaver-combine kernels against the HBM roofline (`extra`), and the CPU restatement of the reference's
arkworks path timed on this box (`cpu_baseline`).

`--impl reference` times that CPU restatement (oracle/, all host threads) on a bounded sample of the
same workload; the reference itself is Rust and cannot be built in this image.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "share_msm_g1_throughput"
UNIT = "Mpts/s"
IMAD_PER_POINT = 52800.0        # 16 windows x 1 mixed add x 11 Fq-mul-eq x 300 IMAD
NTT_BYTES_PER_ELEM = 64.0       # one read + one write of the vector
COMBINE_BYTES_PER_ELEM = 192.0  # x, y, z, sx, oy in + out, additive layout


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class Traffic:
    """DRAM bytes per launch of a kernel: bytes per unit from the committed ncu --set full capture
    (profiles/traffic.json, dram__bytes_read.sum + dram__bytes_write.sum) x the units of this launch"""

    def __init__(self):
        path = os.path.join(ROOT, "profiles", "traffic.json")
        self.table = json.load(open(path)) if os.path.exists(path) else {}

    def get(self, kernel, units=0):
        e = self.table.get(kernel)
        return e["bytes_per_unit"] * units if e else None


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) >= 7 and p[0].isdigit():
                    rows.append(p)
            os.unlink(self.path)
        except Exception:
            pass
        if rows:
            sm = [int(r[0]) for r in rows]
            busy = [s for s, r in zip(sm, rows) if float(r[2] or 0) > 250.0] or sm
            out["sm_mhz"] = statistics.median(busy)
            out["sm_max_mhz"] = int(rows[0][1])
            out["power_w_max"] = max(float(r[2] or 0) for r in rows)
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            out["reasons"] = [nm for k, nm in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in rows)]
            out["samples"] = len(rows)
        return out


def cpu_msm_baseline(oracle, bases, scalars, threads):
    t0 = time.perf_counter()
    oracle.g1_msm(bases, scalars, threads=threads)
    return time.perf_counter() - t0


# ======================================================================================= reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle
    import __graft_entry__ as ge
    pkg = ge.load_package()
    oracle.build()
    cores = os.cpu_count() or 1
    log_s = args.ref_log_n
    n = 1 << log_s
    seed = pkg.synth.bench_seed(log_s)
    bases = oracle.g1_generate(seed, n)
    scalars = pkg.synth.fr_uniform(seed, n)
    nwin = (253 + (log_s * 69 // 100 + 2) - 1) // (log_s * 69 // 100 + 2)
    threads = min(cores, nwin)
    for _ in range(args.warmup):
        cpu_msm_baseline(oracle, bases, scalars, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_msm_baseline(oracle, bases, scalars, threads)
    dt = (time.perf_counter() - t0) / args.steps
    value = n / dt / 1e6
    sample = ("C restatement of arkworks VariableBaseMSM (oracle/zkmpc_oracle.c), %d-point sample of the 2^%d workload, "
              "windows processed by %d threads (the shipped reference is single-threaded)" % (n, args.log_n, threads))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 limbs (Montgomery Fq/Fr)", "data": "synthetic",
        "config": {"workload": "share MSM G1, 2^%d points per GPU, uniform share scalars" % args.log_n,
                   "sample_points": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ======================================================================================= B200 arm
def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d does not match WORLD_SIZE %d" % (args.gpus, world))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the first communicator is created: send fd 1 to
        # stderr until that has happened, so stdout carries exactly the one JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        warm = torch.zeros(1, device="cuda")
        dist.all_reduce(warm)
        torch.cuda.synchronize()
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)

    pkg = ge.load_package()
    if not os.path.exists(pkg._lib.LIB_PATH):
        ge.build()
    H, S, L = pkg.host, pkg.synth, pkg._lib
    H.init([local_rank])
    H.set_party(0, 1)
    # a dedicated non-default stream: the library treats a NULL stream argument as "use my own stream",
    # and CUDA events must be recorded on the stream the kernels are launched on
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sptr = C.c_void_p(stream.cuda_stream)
    assert sptr.value

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    log_n = args.log_n
    n = 1 << log_n
    seed = S.bench_seed(log_n)

    # ---- synthetic inputs: bases k_i*G generated on the device, uniform ("share-like") scalars
    bases_dev = H.g1_generate(seed, n, first=rank * n)
    plain = H.register_bases_dev(bases_dev, n)          # CRS resident, no table: 20-bit windows, 13 bucket sets
    handle = H.register_bases_dev(bases_dev, n)
    if not args.no_table:
        handle.precompute(0)                            # + 2^(cw)*P table (12 x 1.6 GB): one shared bucket set
    scalars_host = torch.from_numpy(S.fr_uniform(seed + 1000 * rank, n).view(np.int64)).pin_memory()
    scalars_dev = scalars_host.to("cuda", non_blocking=True)
    partial = torch.zeros(18, dtype=torch.int64, device="cuda")
    gathered = torch.zeros(18 * world, dtype=torch.int64, device="cuda")
    out_xy = np.zeros(12, dtype=np.uint64)
    out_inf = C.c_uint8(0)
    torch.cuda.synchronize()

    def step(hd=None):
        L.call("mpc_cuda_msm_g1_handle_dev", C.c_uint64((hd or handle).handle), C.c_size_t(0),
               C.cast(scalars_dev.data_ptr(), L.u64p), C.c_size_t(n), C.cast(partial.data_ptr(), L.u64p), sptr)
        if world > 1:
            dist.all_gather_into_tensor(gathered, partial)
            src = gathered
        else:
            src = partial
        if rank == 0:
            L.call("mpc_cuda_g1_sum_partials_dev", C.cast(src.data_ptr(), L.u64p), C.c_uint32(world),
                   out_xy.ctypes.data_as(L.u64p), C.byref(out_inf), sptr)

    for _ in range(args.warmup):
        step()
    barrier()
    H.set_option("profile", 1)
    launches0 = H.launch_count()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    barrier()
    clock_info = clocks.stop() if rank == 0 else None
    launches = H.launch_count() - launches0
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    H.set_option("profile", 0)
    stage_ms = {}
    for nm in ("msm_total", "msm_sort", "msm_accumulate", "msm_reduce"):
        t, cnt = H.profile_read(nm)
        stage_ms[nm] = t / max(cnt, 1)
    ms_per_step = ms_total / args.steps
    value = world * n / (ms_per_step * 1e-3) / 1e6

    # the same step without the table (what a host-buffer call can use)
    step(plain)
    barrier()
    H.set_option("profile", 1)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record(stream)
    for _ in range(args.steps):
        step(plain)
    p1.record(stream)
    barrier()
    H.set_option("profile", 0)
    plain_ms = max_over_ranks(p0.elapsed_time(p1)) / args.steps
    plain_stage = {}
    for nm in ("msm_total", "msm_sort", "msm_accumulate", "msm_reduce"):
        t, cnt = H.profile_read(nm)
        plain_stage[nm] = t / max(cnt, 1)

    # ---- integer-pipe roofline of the dominant kernel (k_accumulate), IMAD peak measured live
    imad_peak = max(H.microbench(0, 20000) for _ in range(2))            # G IMAD/s
    acc_ms = max_over_ranks(stage_ms["msm_accumulate"])
    achieved = IMAD_PER_POINT * n / (acc_ms * 1e-3) / 1e9
    traffic = Traffic()
    roofline = {"kernel": "k_accumulate<Fq> (Pippenger bucket accumulation, XYZZ mixed additions)",
                "bound": "int32-pipe", "achieved": achieved, "peak": imad_peak, "unit": "GIMAD/s",
                "frac": achieved / imad_peak, "traffic": traffic.get("k_accumulate", n),
                "kernel_ms": acc_ms, "share_of_step": acc_ms / ms_per_step,
                "note": "algorithmic work 52800 IMAD/point (SURVEY.md 8d); peak = dependent-free 32-bit IMAD "
                        "microbenchmark on this GPU in this run; MSM is integer-pipe bound, not HBM or tensor"}

    # ---- end to end through the host-buffer C ABI (bases + scalars cross PCIe every step)
    bases_host = torch.empty(n * 12, dtype=torch.int64).pin_memory()
    L.call("mpc_cuda_memcpy_d2h", C.c_void_p(bases_host.data_ptr()), bases_dev.ptr, C.c_size_t(n * 96), None)
    L.call("mpc_cuda_stream_sync", None)
    e_xy = np.zeros(12, dtype=np.uint64)
    e_inf = C.c_uint8(0)
    jac = torch.zeros(18, dtype=torch.int64, device="cuda")
    one_mont = torch.from_numpy(np.array([202099033278250856, 5854854902718660529, 11492539364873682930,
                                          8885205928937022213, 5545221690922665192, 39800542322357402],
                                         dtype=np.uint64).view(np.int64))

    def e2e_step():
        L.call("mpc_cuda_msm_g1", C.cast(bases_host.data_ptr(), L.u64p), None,
               C.cast(scalars_host.data_ptr(), L.u64p), C.c_size_t(n), e_xy.ctypes.data_as(L.u64p), C.byref(e_inf))
        if world > 1:
            # affine partial -> Jacobian (x, y, 1) (or z = 0), gathered and folded on rank 0
            z = torch.zeros(6, dtype=torch.int64) if e_inf.value else one_mont
            jac.copy_(torch.cat([torch.from_numpy(e_xy.view(np.int64)), z]))
            dist.all_gather_into_tensor(gathered, jac)
            if rank == 0:
                L.call("mpc_cuda_g1_sum_partials_dev", C.cast(gathered.data_ptr(), L.u64p), C.c_uint32(world),
                       out_xy.ctypes.data_as(L.u64p), C.byref(out_inf), sptr)

    e2e_steps = max(1, min(args.steps, 3))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    same = bool(np.array_equal(e_xy, out_xy)) if world == 1 else None     # resident path == host path
    e2e = {"value": world * n / e2e_s / 1e6, "unit": UNIT, "h2d_bytes_per_step": n * (96 + 32),
           "d2h_bytes_per_step": 100, "ms_per_step": e2e_s * 1e3,
           "api": "mpc_cuda_msm_g1 (host bases + host scalars, pinned)", "matches_resident_result": same}
    # the deployment shape: CRS registered once (pk.*_query / powers_of_g), only the share scalars move
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        L.call("mpc_cuda_msm_g1_handle", C.c_uint64(handle.handle), C.c_size_t(0),
               C.cast(scalars_host.data_ptr(), L.u64p), C.c_size_t(n), e_xy.ctypes.data_as(L.u64p), C.byref(e_inf))
    e2e_resident_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    e2e["resident_crs"] = {"value": world * n / e2e_resident_s / 1e6, "unit": UNIT,
                           "h2d_bytes_per_step": n * 32, "api": "mpc_cuda_msm_g1_handle (scalars only)"}

    # ---- the HBM-bound kernels of the path: share NTT and Beaver combine
    hbm_peak, hbm_src = load_peaks()
    extra = {}
    vec = torch.from_numpy(S.fr_uniform(seed + 7, n).view(np.int64)).to("cuda")
    others = [vec.clone() for _ in range(5)]
    outv = torch.empty_like(vec)
    H.set_option("profile", 1)
    for kind, name in ((0, "fft"), (3, "coset_ifft")):
        for it in range(args.warmup + args.steps):
            if it == args.warmup:
                barrier()
                H.profile_read("ntt")
            L.call("mpc_cuda_ntt_fr_dev", C.cast(vec.data_ptr(), L.u64p), C.c_uint32(log_n), C.c_uint32(kind),
                   C.c_uint32(1), sptr)
        barrier()
        t, cnt = H.profile_read("ntt")
        ms = max_over_ranks(t / max(cnt, 1))
        gbs = NTT_BYTES_PER_ELEM * n / (ms * 1e-3) / 1e9
        gimad = (n / 2) * log_n * 136.0 / (ms * 1e-3) / 1e9           # 136 IMAD per butterfly
        extra["ntt_" + name] = {"value": world * n / (ms * 1e-3) / 1e6, "unit": "Melem/s", "ms": ms, "log_n": log_n,
                                "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                                             "frac": gbs / hbm_peak, "traffic": traffic.get("k_ntt_pass", n),
                                             "passes": (log_n + 7) // 8,
                                             "int_pipe": {"achieved": gimad, "peak": imad_peak, "unit": "GIMAD/s",
                                                          "frac": gimad / imad_peak},
                                             "note": "64 B/element algorithmic (traffic is per pass); butterflies cost "
                                                     "~136 IMAD each, so the integer pipe, not HBM, is the ceiling "
                                                     "(SURVEY.md 8d): int_pipe is the binding fraction"}}
    H.set_option("profile", 0)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for it in range(args.warmup + args.steps):
        if it == args.warmup:
            barrier()
            evs[0].record(stream)
        L.call("mpc_cuda_beaver_combine_dev", *[C.cast(t.data_ptr(), L.u64p) for t in others],
               C.cast(outv.data_ptr(), L.u64p), C.c_size_t(n), C.c_uint32(1), C.c_uint32(0), sptr)
    evs[1].record(stream)
    barrier()
    ms = max_over_ranks(evs[0].elapsed_time(evs[1]) / args.steps)
    gbs = COMBINE_BYTES_PER_ELEM * n / (ms * 1e-3) / 1e9
    extra["beaver_combine"] = {"value": world * n / (ms * 1e-3) / 1e6, "unit": "Melem/s", "ms": ms,
                               "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                                            "frac": gbs / hbm_peak, "traffic": traffic.get("k_combine", n)}}
    # SPDZ layout of the same kernel: [sh | mac] planes, 320 B per share pair
    sp_in = [torch.cat([t, t]) for t in others[:3]]
    sp_out = torch.empty_like(sp_in[0])
    for it in range(args.warmup + args.steps):
        if it == args.warmup:
            barrier()
            evs[0].record(stream)
        L.call("mpc_cuda_beaver_combine_dev", *[C.cast(t.data_ptr(), L.u64p) for t in sp_in],
               C.cast(others[3].data_ptr(), L.u64p), C.cast(others[4].data_ptr(), L.u64p),
               C.cast(sp_out.data_ptr(), L.u64p), C.c_size_t(n), C.c_uint32(1), C.c_uint32(1), sptr)
    evs[1].record(stream)
    barrier()
    ms = max_over_ranks(evs[0].elapsed_time(evs[1]) / args.steps)
    gbs = 320.0 * n / (ms * 1e-3) / 1e9
    extra["beaver_combine_spdz"] = {"value": world * n / (ms * 1e-3) / 1e6, "unit": "Melem/s", "ms": ms,
                                    "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                                                 "frac": gbs / hbm_peak, "traffic": None}}
    del sp_in, sp_out

    # share MSM over G2 (b_g2_query, src/groth16.rs:160): 2^20 points per GPU, resident CRS, no table
    g2_log = min(20, log_n)
    g2_n = 1 << g2_log
    g2_dev = H.g2_generate(seed, g2_n, first=rank * g2_n)
    g2_host = g2_dev.download().reshape(g2_n, 24)
    g2_dev.free()
    g2_handle = H.register_bases(g2_host, g2=True)
    g2_sc = scalars_host.numpy().view(np.uint64).reshape(n, 4)[:g2_n]
    H.msm_handle(g2_handle, g2_sc)
    t0 = time.perf_counter()
    for _ in range(3):
        H.msm_handle(g2_handle, g2_sc)
    g2_s = max_over_ranks((time.perf_counter() - t0) / 3)
    g2_handle.release()
    extra["msm_g2"] = {"value": world * g2_n / g2_s / 1e6, "unit": "Mpts/s", "ms": g2_s * 1e3, "log_n": g2_log,
                       "api": "mpc_cuda_msm_g2_handle (host scalars, resident CRS)"}

    if world > 1 and (world & (world - 1)) == 0 and world <= 8:
        # one party's 2^log_n NTT block-distributed over the ranks (strong scaling; two NCCL all-to-alls
        # around the cross-device stages, sharding.dist_ntt)
        log_g = world.bit_length() - 1
        m = n // world
        blockv = vec[: m].clone()
        cross = lambda data, l0, kind: H.ntt_cross_stage_dev(data.data_ptr(), log_n, log_g, l0, data.shape[1], kind, sptr.value)
        local_ntt = lambda blk, kind: H.ntt_dev(blk.data_ptr(), log_n - log_g, kind, 1, sptr.value)
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for it in range(args.warmup + args.steps):
            if it == args.warmup:
                barrier()
                d0.record(stream)
            pkg.sharding.dist_ntt(dist, rank, world, blockv, log_n, "fft", cross, local_ntt)
        d1.record(stream)
        barrier()
        ms = max_over_ranks(d0.elapsed_time(d1) / args.steps)
        extra["ntt_fft_sharded"] = {"value": n / (ms * 1e-3) / 1e6, "unit": "Melem/s", "ms": ms, "log_n": log_n,
                                    "scaling": "strong", "exchange": "2 x all_to_all_single of (g-1)/g of each block (NCCL)",
                                    "nvlink_bytes_per_gpu": 2 * (world - 1) * (m // world) * 32}
    extra["hbm_peak_source"] = hbm_src
    extra["msm_without_table"] = {"value": world * n / (plain_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": plain_ms,
                                  "stage_ms": plain_stage, "note": "resident CRS, no precomputed window table"}

    # ---- CPU restatement of the reference's path on this box's host cores (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle
        oracle.build()
        cores = os.cpu_count() or 1
        ls = min(args.ref_log_n, log_n)
        m = 1 << ls
        cb = bases_host.numpy().view(np.uint64).reshape(n, 12)[:m]
        cs = scalars_host.numpy().view(np.uint64).reshape(n, 4)[:m]
        cwin = ls * 69 // 100 + 2
        threads = min(cores, (253 + cwin - 1) // cwin)
        dt = cpu_msm_baseline(oracle, cb, cs, threads)
        m1 = 1 << min(16, ls)
        dt1 = cpu_msm_baseline(oracle, cb[:m1], cs[:m1], 1)
        cpu = {"value": m / dt / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "oracle/zkmpc_oracle.c (C restatement of arkworks VariableBaseMSM + into_repr + affine), first %d "
                         "of the 2^%d points, %d window threads; the shipped reference runs 1 thread: %.4f Mpts/s on a "
                         "%d-point sample" % (m, log_n, threads, m1 / dt1 / 1e6, m1),
               "value_1_thread": m1 / dt1 / 1e6, "host_cores": cores}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 limbs (Montgomery Fq 12x32 / Fr 8x32)", "data": "synthetic",
            "config": {"workload": "share MSM G1, 2^%d points per GPU, uniform share scalars" % log_n,
                       "points_per_gpu": n, "total_points": world * n,
                       "crs": "registered on the device" + ("" if args.no_table else " with the 2^(22w)*P window table (one bucket set)"), "sharding": "point range + NCCL all-gather of Jacobian partials",
                       "l2": "inputs (%.1f GB of bases + scalars, %.1f GB of sort scratch) exceed the 126 MB L2" % (
                           n * 128 / 1e9, n * 16 * 8 / 1e9)},
            "stage_ms": stage_ms, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clock_info, "extra": extra,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--log-n", type=int, default=24, help="log2 of the points per GPU")
    ap.add_argument("--ref-log-n", type=int, default=20, help="log2 of the CPU baseline's sample")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-table", action="store_true", help="do not precompute the window table of the CRS")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    return run_reference(args) if args.impl == "reference" else run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
