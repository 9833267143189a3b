#!/usr/bin/env python
"""bench.py — share-MSM throughput of the B200 hot path (BASELINE.json configs[1]) + the co-headline numbers.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--log-n L] [--impl reference]

One "step" = one share MSM over G1: every rank owns n = 2^L points of an (N * n)-point MSM (point-range
sharding), computes its Jacobian partial with the Pippenger pipeline of libmpc_cuda.so, the
partials are all-gathered over NCCL and rank 0 adds and normalises them (weak scaling).  Scalars and bases are
resident in HBM when the timed region starts (`value`); `e2e` is the same MSM through the host-buffer C ABI call
(mpc_cuda_msm_g1: bases + scalars cross PCIe inside the timed region, streamed in chunks).

The JSON line also carries
  roofline      the dominant kernel (bucket accumulation) against the integer-pipe peak measured in the same run
  cpu_baseline  the CPU restatement of the reference's arkworks path on this box's cores (a sample, stated)
  extra.sweep   MSM and NTT at 2^16 .. 2^24: resident and host-buffer (e2e), uniform and witness-like scalars
  extra.table   what the window table behind `value` costs: build time, bytes, break-even MSM count
  extra.ntt_* / extra.beaver_*   the share NTT and Beaver kernels: device-timed, e2e through the host calls, rooflines
  extra.msm_g2  G2 MSM with its own roofline (158 400 IMAD/point)
  extra.prove   the Groth16 prove sequence (witness map + 4 G1 + 1 G2 MSM, 3 parties as threads; additive and SPDZ)
                beside the same composition on the CPU restatement
  extra.marlin  Marlin's AHP rounds 1-2 and the seven commitment MSMs on a 2^20-constraint R1CS, vectors resident
  extra.prove_one_party_per_gpu  (N > 1) the SPDZ prove sequence with party p on device p mod N (BASELINE config 5)
  extra.strong  (N > 1) ONE party's 2^24 MSM and 2^24 / 2^26 NTT sharded over all N GPUs by the library itself
                (single process, NVLink peer loads/stores inside the cross-stage kernel), strong scaling

`--impl reference` times the CPU restatement (oracle/, all host threads) on a bounded sample of the same
workload; the reference itself is Rust and cannot be built in this image.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "share_msm_g1_throughput"
UNIT = "Mpts/s"
IMAD_PER_POINT = 52800.0        # 16 windows x 1 mixed add x 11 Fq-mul-eq x 300 IMAD
IMAD_PER_POINT_G2 = 158400.0    # Fq2 product = 3 Fq products
NTT_BYTES_PER_ELEM = 64.0       # one read + one write of the vector
NTT_IMAD_PER_BFLY = 136.0
COMBINE_BYTES_PER_ELEM = 192.0  # x, y, z, sx, oy in + out, additive layout
FR_MUL_WIDE_MADS = 128.0        # 2 * 8^2 IMAD.WIDE per Fr Montgomery product
STAGES = ("msm_total", "msm_sort", "msm_accumulate", "msm_reduce")


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

    def get(self, kernel, units=0, all_launches=False):
        e = self.table.get(kernel)
        if not e:
            return None
        return e.get("bytes_per_unit_all_launches" if all_launches else "bytes_per_unit", e["bytes_per_unit"]) * units


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


def bind_to_gpu_numa_node(local_rank):
    """pin this rank's host threads (and so its pinned staging buffers, first-touch) to the NUMA node of its GPU:
    with 8 ranks pulling 2 GiB each through host memory the cross-socket hops are what the e2e number loses"""
    try:
        q = subprocess.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                           capture_output=True, text=True).stdout.strip()
        bus = q[-12:].lower()                      # "00000000:1B:00.0" -> sysfs "0000:1b:00.0"
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return None
        cpus = []
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus += list(range(int(lo), int(hi or lo) + 1))
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def cpu_msm_baseline(oracle, bases, scalars, threads):
    t0 = time.perf_counter()
    oracle.g1_msm(bases, scalars, threads=threads)
    return time.perf_counter() - t0


def workload_name(log_n):
    return "share MSM G1, 2^%d points per GPU, uniform share scalars" % log_n


def shared_config(log_n, world, ref_log_n):
    """the `config` object both arms print, key for key: the workload, and the statement that the CPU arm (and the
    cpu_baseline leg) times a bounded sample of it"""
    n = 1 << log_n
    return {"workload": workload_name(log_n), "points_per_gpu": n, "total_points": world * n,
            "l2": "GPU arm: the inputs of a step (%.1f GB of bases + scalars per GPU, or the %d-window table) exceed the "
                  "126 MB L2, so no flush between timed iterations" % (n * 128 / 1e9, 12),
            "cpu_arm_sample": "the CPU restatement is timed on the first 2^%d points of this workload (a full 2^%d MSM "
                              "takes ~25 s per step on 16 cores)" % (min(ref_log_n, log_n), log_n)}


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
    cwin = log_s * 69 // 100 + 2
    threads = min(cores, (253 + cwin - 1) // cwin)
    for _ in range(args.warmup):
        cpu_msm_baseline(oracle, bases, scalars, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_msm_baseline(oracle, bases, scalars, threads)
    dt = (time.perf_counter() - t0) / args.steps
    value = n / dt / 1e6
    sample = ("C restatement of arkworks VariableBaseMSM (oracle/zkmpc_oracle.c), %d-point sample of the 2^%d workload "
              "(the reference's own window rule gives c = %d here, c = %d at 2^%d: ~10-15 %% fewer additions per point at "
              "full size), windows processed by %d threads (the shipped reference is single-threaded)"
              % (n, args.log_n, cwin, args.log_n * 69 // 100 + 2, args.log_n, threads))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 limbs (Montgomery Fq/Fr)", "data": "synthetic",
        "config": shared_config(args.log_n, max(args.gpus, 1), args.ref_log_n),
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
    numa_node = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    torch.cuda.set_device(local_rank)
    host_group = None
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
        host_group = dist.new_group(backend="gloo")       # CPU-side barrier: no kernel spins on the idle GPUs
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)

    pkg = ge.load_package()
    if not os.path.exists(pkg._lib.LIB_PATH):
        ge.build()
    H, S, L = pkg.host, pkg.synth, pkg._lib
    # this rank's GPU first; rank 0 also lists the others for the single-process strong-scaling entries
    n_visible = torch.cuda.device_count()
    devices = [local_rank] + ([d for d in range(min(world, n_visible)) if d != local_rank] if rank == 0 else [])
    H.init(devices)
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

    def dev_ms(fn, reps, warm=1):
        """average device time of fn() on `stream` (CUDA events), after `warm` untimed calls"""
        for _ in range(warm):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def wall_ms(fn, reps, warm=1):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) / reps * 1e3

    def pinned(arr):
        return torch.from_numpy(np.ascontiguousarray(arr).view(np.int64)).pin_memory()

    def u64(t):
        return C.cast(t.data_ptr(), L.u64p)

    log_n = args.log_n
    n = 1 << log_n
    seed = S.bench_seed(log_n)

    # ---- synthetic inputs: bases k_i*G generated on the device, uniform ("share-like") scalars
    bases_dev = H.g1_generate(seed, n, first=rank * n)
    plain = H.register_bases_dev(bases_dev, n)          # CRS resident, no table: 20-bit windows, 13 bucket sets
    handle = H.register_bases_dev(bases_dev, n)
    table = None
    if not args.no_table:
        H.set_option("profile", 1)
        handle.precompute(0)                            # + 2^(cw)*P table: one shared bucket set
        pre_ms, _ = H.profile_read("msm_precompute")
        H.set_option("profile", 0)
        tc = 22 if log_n >= 22 else 20 if log_n >= 20 else 17 if log_n >= 18 else 16 if log_n >= 16 else 12
        table = {"window_bits": tc, "windows": 253 // tc + 1, "bytes": (253 // tc + 1) * n * 96, "precompute_ms": pre_ms}
    scalars_host = pinned(S.fr_uniform(seed + 1000 * rank, n))
    scalars_dev = scalars_host.to("cuda", non_blocking=True)
    partial = torch.zeros(18, dtype=torch.int64, device="cuda")
    gathered = torch.zeros(18 * world, dtype=torch.int64, device="cuda")
    out_xy = np.zeros(12, dtype=np.uint64)
    out_inf = C.c_uint8(0)
    torch.cuda.synchronize()

    def step(hd=None, sc=None):
        L.call("mpc_cuda_msm_g1_handle_dev", C.c_uint64((hd or handle).handle), C.c_size_t(0),
               u64(sc if sc is not None else scalars_dev), C.c_size_t(n), u64(partial), sptr)
        if world > 1:
            dist.all_gather_into_tensor(gathered, partial)
            src = gathered
        else:
            src = partial
        if rank == 0:
            L.call("mpc_cuda_g1_sum_partials_dev", u64(src), C.c_uint32(world), out_xy.ctypes.data_as(L.u64p),
                   C.byref(out_inf), sptr)

    def read_stages(reps):
        return {nm: H.profile_read(nm)[0] / max(reps, 1) for nm in STAGES}

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
    stage_ms = read_stages(args.steps)
    ms_per_step = ms_total / args.steps
    value = world * n / (ms_per_step * 1e-3) / 1e6
    resident_xy = out_xy.copy()

    # the same step without the table (what a one-shot call over a fresh CRS can use)
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
    plain_stage = read_stages(args.steps)

    # ---- integer-pipe roofline of the dominant kernel (k_accumulate), peaks measured live
    imad_peak = max(H.microbench(0, 20000) for _ in range(2))            # G IMAD/s (32-bit)
    wide_peak = max(H.microbench(1, 20000) for _ in range(2))            # G IMAD.WIDE/s with carry chains
    acc_ms = max_over_ranks(stage_ms["msm_accumulate"])
    achieved = IMAD_PER_POINT * n / (acc_ms * 1e-3) / 1e9
    traffic = Traffic()
    roofline = {"kernel": "bucket accumulation: k_affine_pairs<Fq> x2 (batched-affine pre-reduction) + k_accumulate<Fq> (XYZZ mixed additions)",
                "bound": "int32-pipe", "achieved": achieved, "peak": imad_peak, "unit": "GIMAD/s",
                "frac": achieved / imad_peak,
                "traffic": (traffic.get("k_accumulate", n) or 0) + (traffic.get("k_affine_pairs", n, all_launches=True) or 0) or None,
                "traffic_note": "k_accumulate + both k_affine_pairs launches (the pre-reduction gathers every operand twice: DRAM bytes "
                                "traded for multiplier cycles; the stage runs at ~30 % of DRAM bandwidth)",
                "kernel_ms": acc_ms, "share_of_step": acc_ms / ms_per_step,
                "imad_wide_peak": wide_peak,
                "note": "algorithmic work 52800 IMAD/point (SURVEY.md 8d); peak = dependent-free 32-bit IMAD "
                        "microbenchmark on this GPU in this run; the kernel issues IMAD.WIDE (half rate, imad_wide_peak): "
                        "MSM is integer-pipe bound, not HBM or tensor"}

    # ---- end to end through the host-buffer C ABI (bases + scalars cross PCIe every step, chunked overlap)
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
        L.call("mpc_cuda_msm_g1", u64(bases_host), None, u64(scalars_host), C.c_size_t(n), e_xy.ctypes.data_as(L.u64p),
               C.byref(e_inf))
        if world > 1:
            # affine partial -> Jacobian (x, y, 1) (or z = 0), gathered and folded on rank 0
            z = torch.zeros(6, dtype=torch.int64) if e_inf.value else one_mont
            jac.copy_(torch.cat([torch.from_numpy(e_xy.view(np.int64)), z]))
            dist.all_gather_into_tensor(gathered, jac)
            if rank == 0:
                L.call("mpc_cuda_g1_sum_partials_dev", u64(gathered), C.c_uint32(world), out_xy.ctypes.data_as(L.u64p),
                       C.byref(out_inf), sptr)

    e2e_steps = max(1, min(args.steps, 3))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    same = bool(np.array_equal(e_xy, resident_xy)) if world == 1 else None     # resident path == host path
    e2e = {"value": world * n / e2e_s / 1e6, "unit": UNIT, "h2d_bytes_per_step": n * (96 + 32),
           "d2h_bytes_per_step": 100, "ms_per_step": e2e_s * 1e3,
           "api": "mpc_cuda_msm_g1 (host bases + host scalars, pinned; streamed as point-range chunks so the copy of "
                  "chunk j+1 overlaps the sort + accumulation of chunk j)",
           "matches_resident_result": same, "numa_node": numa_node}
    # the deployment shape: CRS registered once (pk.*_query / powers_of_g), only the share scalars move
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        L.call("mpc_cuda_msm_g1_handle", C.c_uint64(handle.handle), C.c_size_t(0), u64(scalars_host), C.c_size_t(n),
               e_xy.ctypes.data_as(L.u64p), C.byref(e_inf))
    e2e_resident_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    e2e["resident_crs"] = {"value": world * n / e2e_resident_s / 1e6, "unit": UNIT,
                           "h2d_bytes_per_step": n * 32, "api": "mpc_cuda_msm_g1_handle (scalars only)"}

    hbm_peak, hbm_src = load_peaks()
    extra = {"hbm_peak_source": hbm_src}
    extra["msm_without_table"] = {"value": world * n / (plain_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": plain_ms,
                                  "stage_ms": plain_stage, "note": "resident CRS, no precomputed window table"}
    if table:
        saved = plain_ms - ms_per_step
        table["break_even_msms"] = (table["precompute_ms"] / saved) if saved > 0 else None
        table["note"] = ("built once per registered CRS (pk.*_query / powers_of_g are fixed per circuit); 253 Jacobian "
                         "doublings per base bound the build (~1.15 s of field products at 2^24)")
        extra["table"] = table

    # witness-like scalars (half of them 0 / 1: the reference short-cuts both, variable_base.rs:19,44-48)
    wl_dev = pinned(S.fr_witness_like(seed + 77, n)).to("cuda")
    wl_ms = max_over_ranks(dev_ms(lambda: step(handle, wl_dev), max(2, args.steps // 2)))
    extra["msm_witness_like"] = {"value": world * n / (wl_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": wl_ms}
    del wl_dev

    # ---- the share NTT and Beaver kernels at the same size
    vec = pinned(S.fr_uniform(seed + 7, n)).to("cuda")
    vec_host = pinned(S.fr_uniform(seed + 7, n))
    others = [vec.clone() for _ in range(5)]
    outv = torch.empty_like(vec)
    H.set_option("profile", 1)
    for kind, name in ((0, "fft"), (3, "coset_ifft")):
        for it in range(args.warmup + args.steps):
            if it == args.warmup:
                barrier()
                H.profile_read("ntt")
            L.call("mpc_cuda_ntt_fr_dev", u64(vec), C.c_uint32(log_n), C.c_uint32(kind), C.c_uint32(1), sptr)
        barrier()
        t, cnt = H.profile_read("ntt")
        ms = max_over_ranks(t / max(cnt, 1))
        gbs = NTT_BYTES_PER_ELEM * n / (ms * 1e-3) / 1e9
        gimad = (n / 2) * log_n * NTT_IMAD_PER_BFLY / (ms * 1e-3) / 1e9
        entry = {"value": world * n / (ms * 1e-3) / 1e6, "unit": "Melem/s", "ms": ms, "log_n": log_n,
                 "roofline": {"bound": "int32-pipe (nominally hbm)", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                              "frac": gbs / hbm_peak, "traffic": traffic.get("k_ntt_pass", n),
                              "passes": (log_n + 7) // 8,
                              "int_pipe": {"achieved": gimad, "peak": imad_peak, "unit": "GIMAD/s",
                                           "frac": gimad / imad_peak},
                              "note": "64 B/element algorithmic (traffic is per pass); butterflies cost ~136 IMAD each, so "
                                      "the integer pipe, not HBM, is the ceiling (SURVEY.md 8d): int_pipe is the binding "
                                      "fraction"}}
        extra["ntt_" + name] = entry
    H.set_option("profile", 0)
    # e2e: host vector in and out through mpc_cuda_ntt_fr (512 MiB each way at 2^24)
    ms = max_over_ranks(wall_ms(lambda: L.call("mpc_cuda_ntt_fr", u64(vec_host), C.c_uint32(log_n), C.c_uint32(0),
                                               C.c_uint32(1)), 2))
    extra["ntt_fft"]["e2e"] = {"value": world * n / (ms * 1e-3) / 1e6, "unit": "Melem/s", "ms": ms,
                               "h2d_bytes_per_step": n * 32, "d2h_bytes_per_step": n * 32, "api": "mpc_cuda_ntt_fr"}

    def combine(inputs, out, leader, spdz):
        L.call("mpc_cuda_beaver_combine_dev", *[u64(t) for t in inputs], u64(out), C.c_size_t(n), C.c_uint32(leader),
               C.c_uint32(spdz), sptr)

    def combine_entry(ms, bytes_per_elem, products):
        gbs = bytes_per_elem * n / (ms * 1e-3) / 1e9
        gwide = products * FR_MUL_WIDE_MADS * n / (ms * 1e-3) / 1e9
        return {"value": world * n / (ms * 1e-3) / 1e6, "unit": "Melem/s", "ms": ms,
                "roofline": {"bound": "int32-pipe (IMAD.WIDE); HBM is the second ceiling", "achieved": gwide,
                             "peak": wide_peak, "unit": "G IMAD.WIDE/s", "frac": gwide / wide_peak,
                             "hbm": {"achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak},
                             "products_per_element": products, "traffic": traffic.get("k_combine", n)}}

    ms = max_over_ranks(dev_ms(lambda: combine(others, outv, 1, 0), args.steps, args.warmup))
    extra["beaver_combine"] = combine_entry(ms, COMBINE_BYTES_PER_ELEM, 3)
    ms = max_over_ranks(dev_ms(lambda: combine(others, outv, 0, 0), args.steps, args.warmup))
    extra["beaver_combine_non_leader"] = combine_entry(ms, COMBINE_BYTES_PER_ELEM, 2)
    sp_in = [torch.cat([t, t]) for t in others[:3]] + others[3:]
    sp_out = torch.empty_like(sp_in[0])
    ms = max_over_ranks(dev_ms(lambda: combine(sp_in, sp_out, 1, 1), args.steps, args.warmup))
    extra["beaver_combine_spdz"] = combine_entry(ms, 320.0, 5)
    del sp_in, sp_out
    hosts = [pinned(S.fr_uniform(seed + 20 + k, n)) for k in range(5)]
    host_out = torch.empty(n * 4, dtype=torch.int64).pin_memory()
    ms = max_over_ranks(wall_ms(lambda: L.call("mpc_cuda_beaver_combine", *[u64(t) for t in hosts], u64(host_out),
                                               C.c_size_t(n), C.c_uint32(1), C.c_uint32(0)), 2))
    extra["beaver_combine"]["e2e"] = {"value": world * n / (ms * 1e-3) / 1e6, "unit": "Melem/s", "ms": ms,
                                      "h2d_bytes_per_step": 5 * n * 32, "d2h_bytes_per_step": n * 32,
                                      "api": "mpc_cuda_beaver_combine"}
    del hosts, host_out, others, outv

    # ---- share MSM over G2 (b_g2_query, src/groth16.rs:160): 2^20 points per GPU, resident CRS and scalars
    g2_log = min(20, log_n)
    g2_n = 1 << g2_log
    g2_dev = H.g2_generate(seed, g2_n, first=rank * g2_n)
    g2_handle = H.register_bases_dev(g2_dev, g2_n, g2=True)
    g2_part = H.DeviceBuffer(36 * 8)
    g2_sc = H.DeviceBuffer(g2_n * 32).upload(scalars_host.numpy().view(np.uint64).reshape(n, 4)[:g2_n])

    def g2_step():
        L.call("mpc_cuda_msm_g2_handle_dev", C.c_uint64(g2_handle.handle), C.c_size_t(0), g2_sc.u64(), C.c_size_t(g2_n),
               g2_part.u64(), sptr)

    g2_step()
    torch.cuda.synchronize()
    H.set_option("profile", 1)
    read_stages(1)
    g2_ms = max_over_ranks(dev_ms(g2_step, 3, 0))
    g2_stage = read_stages(3)
    H.set_option("profile", 0)
    g2_acc = max_over_ranks(g2_stage["msm_accumulate"])
    g2_ach = IMAD_PER_POINT_G2 * g2_n / (g2_acc * 1e-3) / 1e9
    extra["msm_g2"] = {"value": world * g2_n / (g2_ms * 1e-3) / 1e6, "unit": "Mpts/s", "ms": g2_ms, "log_n": g2_log,
                       "stage_ms": g2_stage, "api": "mpc_cuda_msm_g2_handle_dev (resident CRS and scalars)",
                       "roofline": {"kernel": "k_accumulate<Fq2>", "bound": "int32-pipe", "achieved": g2_ach, "peak": imad_peak,
                                    "unit": "GIMAD/s", "frac": g2_ach / imad_peak, "kernel_ms": g2_acc,
                                    "note": "158400 IMAD/point (SURVEY.md 8d: Fq2 product = 3 Fq products)"}}
    g2_handle.release(); g2_dev.free(); g2_part.free(); g2_sc.free()

    # ---- size sweep (N = 1): BASELINE names 2^16 - 2^24 for MSM and NTT
    if world == 1 and not args.no_sweep:
        sweep = []
        for ln in (16, 18, 20, 22, 24):
            if ln > log_n:
                break
            m = 1 << ln
            row = {"log_n": ln}
            sub = H.register_bases_dev(bases_dev, m)            # a prefix of the resident CRS
            sc = scalars_dev                                     # the first m scalars
            part = torch.zeros(18, dtype=torch.int64, device="cuda")

            def resident(hd, scalars):
                L.call("mpc_cuda_msm_g1_handle_dev", C.c_uint64(hd.handle), C.c_size_t(0), u64(scalars), C.c_size_t(m),
                       u64(part), sptr)
                L.call("mpc_cuda_g1_sum_partials_dev", u64(part), C.c_uint32(1), out_xy.ctypes.data_as(L.u64p),
                       C.byref(out_inf), sptr)

            ms = dev_ms(lambda: resident(sub, sc), 3)
            row["msm_resident"] = {"ms": ms, "Mpts_s": m / ms / 1e3}
            wl = pinned(S.fr_witness_like(seed + ln, m)).to("cuda")
            ms = dev_ms(lambda: resident(sub, wl), 3)
            row["msm_resident_witness_like"] = {"ms": ms, "Mpts_s": m / ms / 1e3}
            if ln < log_n:                                       # the full size is the headline above
                H.set_option("profile", 1)
                sub.precompute(0)
                pre, _ = H.profile_read("msm_precompute")
                H.set_option("profile", 0)
                ms = dev_ms(lambda: resident(sub, sc), 3)
                row["msm_resident_table"] = {"ms": ms, "Mpts_s": m / ms / 1e3, "precompute_ms": pre}
                ms = wall_ms(lambda: L.call("mpc_cuda_msm_g1", u64(bases_host), None, u64(scalars_host), C.c_size_t(m),
                                            e_xy.ctypes.data_as(L.u64p), C.byref(e_inf)), 2)
                row["msm_e2e_host"] = {"ms": ms, "Mpts_s": m / ms / 1e3, "h2d_bytes": m * 128}
            else:
                row["msm_resident_table"] = {"ms": ms_per_step, "Mpts_s": value, "precompute_ms": table["precompute_ms"] if table else None}
                row["msm_e2e_host"] = {"ms": e2e_s * 1e3, "Mpts_s": n / e2e_s / 1e6, "h2d_bytes": n * 128}
            sub.release()
            del wl
            H.set_option("profile", 1)
            H.profile_read("ntt")
            for kind in (0, 0, 0, 0):
                L.call("mpc_cuda_ntt_fr_dev", u64(vec), C.c_uint32(ln), C.c_uint32(kind), C.c_uint32(1), sptr)
            torch.cuda.synchronize()
            t, cnt = H.profile_read("ntt")
            H.set_option("profile", 0)
            ms = t / max(cnt, 1)
            row["ntt_fft"] = {"ms": ms, "Melem_s": m / ms / 1e3}
            ms = wall_ms(lambda: L.call("mpc_cuda_ntt_fr", u64(vec_host), C.c_uint32(ln), C.c_uint32(0), C.c_uint32(1)), 2)
            row["ntt_fft_e2e_host"] = {"ms": ms, "Melem_s": m / ms / 1e3, "h2d_bytes": m * 32, "d2h_bytes": m * 32}
            sweep.append(row)
        extra["sweep"] = sweep

    # ---- the Groth16 prove sequence (x1 / a16): MySecretInputCircuit's shape, 3 parties as 3 threads on this GPU
    if world == 1 and not args.no_prove:
        extra["prove"] = bench_prove(pkg, H, S, args)
        if args.log_n >= 22:
            try:
                extra["marlin"] = bench_marlin(pkg, H, S, 20)
            except Exception as e:      # noqa: BLE001 - a failed extra must not lose the headline line
                extra["marlin"] = {"error": str(e)[:300]}

    # ---- strong scaling inside ONE process (rank 0 drives all N GPUs through the library's sharded entries)
    if world > 1:
        dist.barrier(group=host_group)
        if rank == 0 and len(devices) == world and (world & (world - 1)) == 0 and world <= 8:
            try:
                extra["strong"] = bench_strong(pkg, H, S, L, world, log_n, bases_host, scalars_host, args)
            except Exception as e:      # noqa: BLE001 - a failed extra must not lose the headline line
                extra["strong"] = {"error": str(e)[:300]}
            if not args.no_prove:
                try:
                    extra["prove_one_party_per_gpu"] = bench_prove_per_gpu(pkg, H, S, world)
                except Exception as e:      # noqa: BLE001
                    extra["prove_one_party_per_gpu"] = {"error": str(e)[:300]}
        dist.barrier(group=host_group)

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
            "config": shared_config(log_n, world, args.ref_log_n),
            "setup": {"crs": "registered on the device" + ("" if args.no_table else
                                                          " with its window table 2^(%d w) P_i, w < %d (one bucket set; "
                                                          "build time and size in extra.table)" % (table["window_bits"], table["windows"])),
                      "sharding": "point range + NCCL all-gather of Jacobian partials",
                      "l2": "inputs (%.1f GB of bases + scalars, %.1f GB of sort scratch) exceed the 126 MB L2" % (
                          n * 128 / 1e9, n * 16 * 8 / 1e9)},
            "stage_ms": stage_ms, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clock_info, "extra": extra,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


class _ThreadNet:
    """MpcSerNet::broadcast among party threads (the role LocalTestNet plays in the reference)"""

    def __init__(self, party, n_parties, barrier, slots):
        self.party, self.n_parties, self._b, self._s = party, n_parties, barrier, slots

    def exchange(self, payload):
        self._s[self.party] = payload
        self._b.wait()
        out = list(self._s)
        self._b.wait()
        return out


def _prove_instance(H, S, nc, ni, nv, spdz):
    """synthetic circuit of the given shape: matrices, proving-key arrays, opened assignment, 3 parties' shares"""
    import numpy as np
    log_n = max(nc + ni - 1, 0).bit_length()
    n = 1 << log_n
    mats = S.r1cs_matrices(0xB10 + log_n, nc, nv)

    def gen(fn, limbs):
        def g(seed, count):
            b = fn(seed, count)
            a = b.download().reshape(count, limbs)
            b.free()
            return a
        return g

    pkarr = S.proving_key_arrays(gen(H.g1_generate, 12), gen(H.g2_generate, 24), 0xB20 + log_n, nv, ni, n)
    z_open = S.fr_uniform(0xB30, nv)
    z_open[0] = S.FR_R_LIMBS
    sub = lambda a, b: H.vec_op("sub", a, b)
    shares = S.additive_shares(0xB40, z_open, 3, ni, sub)
    if spdz:            # MAC key 1 shared as (1, 0, 0): the mac planes are additive shares of the same values
        macs = S.additive_shares(0xB60, z_open, 3, ni, sub)
        shares = [np.stack([sh, mc]) for sh, mc in zip(shares, macs)]
    return mats, pkarr, z_open, shares, log_n


def _prove_parties(pkg, H, mats, pkarr, ni, nv, shares, devices, spdz, iters=4):
    """3 party threads, party p on device devices[p]; a proving key / R1CS registration per distinct device.
    Returns (per-iteration max-over-parties seconds, last results, setup seconds)."""
    G = pkg.groth16
    parties = len(shares)
    bar, slots = threading.Barrier(parties), [None] * parties
    sync = threading.Barrier(parties)
    res, errs, per_party = [None] * parties, [], [[] for _ in range(parties)]
    keys, setup = {}, [0.0]
    owners = {d: min(p for p in range(parties) if devices[p] == d) for d in set(devices)}

    def party(p):
        try:
            H.set_party(p, parties)
            H.set_device(devices[p])
            if owners[devices[p]] == p:
                t0 = time.perf_counter()
                keys[devices[p]] = (G.ProvingKey(**pkarr, table_bits_g1=int(os.environ.get('PK_BITS_G1', '0')),
                                                 table_bits_g2=int(os.environ.get('PK_BITS_G2', '0'))),
                                    G.R1CS(*mats, num_inputs=ni, num_vars=nv))
                setup[0] = max(setup[0], time.perf_counter() - t0)
            sync.wait()
            pk, r1cs = keys[devices[p]]
            session = G.ProverSession(pk, r1cs)       # per-party working set, reused across proofs
            net = _ThreadNet(p, parties, bar, slots)
            for _ in range(iters):
                sync.wait()
                t0 = time.perf_counter()
                res[p] = session.prove(shares[p], net, spdz=spdz)
                per_party[p].append(time.perf_counter() - t0)
            session.close()
            sync.wait()
            if owners[devices[p]] == p:
                pk.release()
                r1cs.release()
        except Exception as e:      # noqa: BLE001
            errs.append(e)
            bar.abort()
            sync.abort()

    ts = [threading.Thread(target=party, args=(p,)) for p in range(parties)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if errs:
        raise errs[0]
    H.set_device(0)
    H.set_party(0, 1)
    return [max(per_party[p][it] for p in range(parties)) for it in range(iters)], res, setup[0]


PROVE_WHAT = ("per proof, all 3 parties concurrently: A z, B z, C z, 3 iFFT, 3 coset FFT, Beaver batch product with 2 opens "
              "(wire bytes exchanged between the party threads; SPDZ: both planes + the MAC check of each open), coset iFFT, "
              "4 G1 + 1 G2 MSM")


def bench_prove(pkg, H, S, args):
    """prove_hot_path_s: create_proof between synthesis and reveal (src/groth16.rs:100-171) for 3 parties on one GPU, on
    the MySecretInputCircuit shape (6574 constraints + 5 inputs -> domain 2^13), the werewolf DivinationCircuit shape
    under the malicious backend (22 249 constraints, SPDZ planes, BASELINE config 5) and a synthetic 2^20 shape"""
    import numpy as np
    out = {}
    for name, nc, ni, nv, spdz in (("my_secret_input_circuit_2p13", 6574, 5, 6600, False),
                                   ("werewolf_divination_2p15_spdz", 22249, 4, 22300, True),
                                   ("synthetic_2p20", (1 << 20) - 8, 5, 1 << 20, False)):
        if name == "synthetic_2p20" and args.log_n < 22:
            continue
        mats, pkarr, z_open, shares, log_n = _prove_instance(H, S, nc, ni, nv, spdz)
        times, res, setup_s = _prove_parties(pkg, H, mats, pkarr, ni, nv, shares, [0, 0, 0], spdz)
        entry = {"prove_hot_path_s": min(times[1:]), "first_call_s": times[0], "parties": 3, "backend": "spdz" if spdz else "additive",
                 "constraints": nc, "variables": nv, "domain_log2": log_n,
                 "setup_s": setup_s, "setup": "register pk.*_query with window tables + the CSR matrices (once per circuit)",
                 "what": PROVE_WHAT}
        if not args.no_cpu and name != "synthetic_2p20":
            from oracle import oracle
            zero = np.zeros(4, dtype=np.uint64)
            t0 = time.perf_counter()
            exp = oracle.groth16_prove(pkarr, mats, ni, z_open, log_n, zero, zero, threads=1)
            entry["cpu_port_1_thread_s"] = time.perf_counter() - t0
            t0 = time.perf_counter()
            oracle.groth16_prove(pkarr, mats, ni, z_open, log_n, zero, zero, threads=os.cpu_count() or 1)
            entry["cpu_port_all_cores_s"] = time.perf_counter() - t0
            entry["cpu_port_note"] = ("the same sequence on plain values through oracle/ (ONE prover, one plane; the reference "
                                      "runs it per party, single-threaded); excludes networking on both sides")
            # the opened GPU proof equals the CPU one
            acc = {k: res[0][k] for k in ("a", "b", "c")}
            for q in res[1:]:
                for k, add in (("a", oracle.g1_add), ("b", oracle.g2_add), ("c", oracle.g1_add)):
                    acc[k] = add(acc[k][0], q[k][0], acc[k][1], q[k][1])
            entry["matches_cpu_proof"] = all(bool(np.array_equal(acc[k][0], exp[k][0])) and acc[k][1] == exp[k][1]
                                             for k in ("a", "b", "c"))
        out[name] = entry
    return out


def bench_prove_per_gpu(pkg, H, S, world):
    """BASELINE config 5: the malicious backend with one party per GPU (party p on device p mod world), all in this
    process; against the same three parties sharing device 0"""
    nc, ni, nv = 22249, 4, 22300
    mats, pkarr, z_open, shares, log_n = _prove_instance(H, S, nc, ni, nv, True)
    devices = [p % world for p in range(3)]
    t_split, res_split, _ = _prove_parties(pkg, H, mats, pkarr, ni, nv, shares, devices, True)
    t_one, res_one, _ = _prove_parties(pkg, H, mats, pkarr, ni, nv, shares, [0, 0, 0], True)
    import numpy as np
    same = all(np.array_equal(res_split[p][k][0], res_one[p][k][0]) for p in range(3) for k in ("a", "b", "c"))
    return {"shape": "werewolf DivinationCircuit (22 249 constraints, domain 2^15), SPDZ planes, 3 parties",
            "party_devices": devices, "prove_hot_path_s": min(t_split[1:]), "prove_hot_path_s_one_gpu": min(t_one[1:]),
            "same_shares_as_one_gpu": bool(same), "what": PROVE_WHAT}


def bench_marlin(pkg, H, S, log_h):
    """BASELINE config 4's shape: the share-side work of Marlin's AHP rounds 1-2 plus the seven KZG commitment MSMs over
    powers_of_g (marlin_pc without the hiding terms) on a synthetic R1CS of 2^log_h constraints, 3 parties as 3 threads
    on this GPU, every vector resident (marlin.ResidentProver); Fiat-Shamir, the third round (public index
    polynomials) and the openings are not part of this number"""
    import numpy as np
    M, K = pkg.marlin, pkg.kzg
    nh, ni, parties = 1 << log_h, 4, 3
    nc = nh
    mats = S.r1cs_matrices(0xC10 + log_h, nc, nc)
    x = S.fr_uniform(0xC20, ni)
    x[0] = S.FR_R_LIMBS
    shares = [dict(w=S.fr_uniform(0xC30 + p, nc - ni), bl=S.fr_uniform(0xC40 + p, 3), mask=S.fr_uniform(0xC50 + p, 3 * nh))
              for p in range(parties)]
    rnd = S.fr_uniform(0xC60, 4)
    alpha, etas = rnd[0], rnd[1:4]
    t0 = time.perf_counter()
    index = M.Index(mats, nc, ni)
    g = H.g1_generate(0xC70, 7 * nh)
    powers = K.Powers.__new__(K.Powers)
    powers.g = H.register_bases_dev(g, 7 * nh)
    powers.g.precompute(0)
    powers.gamma_g = powers.g
    setup_s = time.perf_counter() - t0
    bar, slots = threading.Barrier(parties), [None] * parties
    sync = threading.Barrier(parties)
    errs, per_party, coms = [], [[] for _ in range(parties)], [None] * parties
    per_party_rounds = [[] for _ in range(parties)]
    iters = 3

    def party(p):
        try:
            H.set_party(p, parties)
            H.set_device(0)
            leader = p == 0
            rp = M.ResidentProver(index, parties)
            # the party's inputs in page-locked memory (triple shares from preprocessing, mask and witness shares)
            pin = H.PinnedBuffer((4 * nh + 3 * nh + nc) * 32)
            v = pin.array(np.uint64, 16 * nh).reshape(4 * nh, 4)
            v[...] = S.FR_R_LIMBS if leader else 0
            mask = pin.array(np.uint64, 12 * nh, 4 * nh * 32).reshape(3 * nh, 4)
            mask[...] = shares[p]["mask"]
            w = pin.array(np.uint64, 4 * (nc - ni), 7 * nh * 32).reshape(nc - ni, 4)
            w[...] = shares[p]["w"]
            net = _ThreadNet(p, parties, bar, slots)
            for _ in range(iters):
                sync.wait()
                t0 = time.perf_counter()
                out = rp.rounds(x, w, shares[p]["bl"], mask, alpha, etas, net, (v, v, v), leader,
                                powers=powers, download=False)
                per_party_rounds[p].append(time.perf_counter() - t0)
                rp.open_combination([(name, etas[k % 3]) for k, name in enumerate(("w", "z_a", "z_b", "mask", "t", "g_1", "h_1"))],
                                    alpha, powers)
                per_party[p].append(time.perf_counter() - t0)
            coms[p] = out["commitments"]
            rp.close()
            pin.free()
        except Exception as e:      # noqa: BLE001
            errs.append(e)
            bar.abort()
            sync.abort()

    ts = [threading.Thread(target=party, args=(p,)) for p in range(parties)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    H.set_party(0, 1)
    index.release()
    powers.g.release()
    g.free()
    if errs:
        raise errs[0]
    times = [max(per_party[p][it] for p in range(parties)) for it in range(iters)]
    times_rounds = [max(per_party_rounds[p][it] for p in range(parties)) for it in range(iters)]
    return {"rounds_commitments_opening_s": min(times[1:]), "rounds_and_commitments_s": min(times_rounds[1:]),
            "first_call_s": times[0], "parties": parties, "constraints": nc,
            "domain_h_log2": log_h, "mul_domain_log2": log_h + 3, "msm_points_per_party": int(15 * nh - ni + 1),
            "setup_s": setup_s, "t_is_public_and_equal": bool(all(np.array_equal(coms[0]["t"][0], c["t"][0]) for c in coms)),
            "what": "per party: A z, B z; 3 iFFT |H| with v_H blinding, division by v_X and v_H; z_A z_B as 2 FFT + Beaver batch "
                    "product (2 opens of 4|H| elements as wire payloads) + iFFT on 4|H|; r(alpha,.) with 2^log_h inversions; t "
                    "through 3 transposed SpMVs; 4 FFT + 1 iFFT on 8|H| (the shared prover cannot truncate); division by v_H; "
                    "7 commitment MSMs (|H|, |H|+1, |H|+1, 3|H|, |H|, |H|-1, 7|H| points); then one opening: the linear "
                    "combination of the seven oracles, its witness polynomial / (x - alpha), evaluation and witness commitment "
                    "(7|H| points); no CPU arm at this size "
                    "(parity: tests/test_gpu_prover.py against oracle.marlin_rounds at 2^6..2^10)"}


def bench_strong(pkg, H, S, L, world, log_n, bases_host, scalars_host, args):
    """ONE party's MSM / NTT sharded over all `world` GPUs by the library (single process): strong scaling"""
    import numpy as np
    import torch
    out = {"devices": world, "how": "single process; mpc_cuda_msm_g1_register_bases_sharded + mpc_cuda_msm_g1_handle "
                                    "(Jacobian partials gathered by NVLink peer copies) and mpc_cuda_ntt_fr_sharded_dev "
                                    "(cross-device stages read/write the peers' blocks inside the butterfly kernel)"}
    n = 1 << log_n
    log_g = world.bit_length() - 1
    bases = bases_host.numpy().view(np.uint64).reshape(n, 12)
    scal = scalars_host.numpy().view(np.uint64).reshape(n, 4)

    def timed(fn, reps=3):
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) / reps * 1e3

    # MSM: 2^log_n points in total, scalars from the host (n/g x 32 B per PCIe link), CRS resident per part
    h1 = H.register_bases(bases, parts=1)
    hg = H.register_bases(bases, parts=world)
    one = H.msm_handle(h1, scal)
    t1 = timed(lambda: H.msm_handle(h1, scal))
    got = H.msm_handle(hg, scal)
    tg = timed(lambda: H.msm_handle(hg, scal))
    out["msm_2p%d" % log_n] = {"ms_1gpu": t1, "ms": tg, "Mpts_s": n / tg / 1e3, "speedup": t1 / tg, "efficiency": t1 / tg / world,
                              "same_result_as_1gpu": bool(np.array_equal(one[0], got[0])) and one[1] == got[1],
                              "api": "mpc_cuda_msm_g1_handle (host scalars), no window table"}
    h1.precompute(0)
    hg.precompute(0)
    t1 = timed(lambda: H.msm_handle(h1, scal))
    tg = timed(lambda: H.msm_handle(hg, scal))
    out["msm_2p%d_table" % log_n] = {"ms_1gpu": t1, "ms": tg, "Mpts_s": n / tg / 1e3, "speedup": t1 / tg,
                                    "efficiency": t1 / tg / world}
    h1.release()
    hg.release()

    # NTT: blocks resident on their devices
    for ln in (log_n, log_n + 2):
        m_total = 1 << ln
        m = m_total // world
        v = S.fr_uniform(0x5EED + ln, min(m_total, 1 << 22))
        bufs = []
        for q in range(world):
            H.set_device(q)
            b = H.DeviceBuffer(m * 32)
            for off in range(0, m, v.shape[0]):                 # fill with repeated random data
                k = min(v.shape[0], m - off)
                L.call("mpc_cuda_memcpy_h2d", C.c_void_p(b.ptr.value + off * 32), v.ctypes.data_as(C.c_void_p), C.c_size_t(k * 32), None)
            L.call("mpc_cuda_stream_sync", None)
            bufs.append(b)
        H.set_device(0)
        ptrs = [b.ptr.value for b in bufs]

        def run(kind):
            H.ntt_sharded_dev(ptrs, ln, kind)
            L.call("mpc_cuda_stream_sync", None)

        def run_many(kind, reps):
            for _ in range(reps):
                H.ntt_sharded_dev(ptrs, ln, kind)
            L.call("mpc_cuda_stream_sync", None)

        for kind in ("fft", "coset_ifft"):             # direct call, graph capture, first replay: all untimed
            for _ in range(3):
                run(kind)
        reps = 5
        t0 = time.perf_counter()
        run_many("fft", reps)
        tf = (time.perf_counter() - t0) / reps * 1e3
        t0 = time.perf_counter()
        run_many("coset_ifft", reps)
        ti = (time.perf_counter() - t0) / reps * 1e3
        for b in bufs:
            b.free()
        # single-GPU time of the same transform, when it fits comfortably
        single = None
        if ln <= 26:
            one = H.DeviceBuffer(m_total * 32)
            H.set_option("profile", 1)
            for _ in range(4):
                H.ntt_dev(one.ptr.value, ln, "fft")
            L.call("mpc_cuda_stream_sync", None)
            H.profile_read("ntt")
            for _ in range(3):
                H.ntt_dev(one.ptr.value, ln, "fft")
            L.call("mpc_cuda_stream_sync", None)
            t, cnt = H.profile_read("ntt")
            H.set_option("profile", 0)
            single = t / max(cnt, 1)
            one.free()
        out["ntt_2p%d" % ln] = {"ms_fft": tf, "ms_coset_ifft": ti, "Melem_s_fft": m_total / tf / 1e3, "ms_1gpu_fft": single,
                                "speedup": (single / tf) if single else None,
                                "efficiency": (single / tf / world) if single else None,
                                "nvlink_bytes_per_gpu": 2 * (world - 1) * (m // world) * 32,
                                "timing": "host clock around 5 back-to-back asynchronous transforms + one stream sync"}
    return out


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
    ap.add_argument("--no-sweep", action="store_true", help="skip extra.sweep")
    ap.add_argument("--no-prove", action="store_true", help="skip extra.prove")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    return run_reference(args) if args.impl == "reference" else run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
