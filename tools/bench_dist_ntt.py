#!/usr/bin/env python3
"""torchrun tool: time the sharded share NTT (sharding.dist_ntt over NCCL) for one vector of 2^log_n
elements block-distributed over WORLD_SIZE GPUs.  Prints one JSON line on rank 0 (max over ranks)."""
import json, os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import __graft_entry__ as ge

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ["NCCL_DEBUG"] = "WARN"
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pkg = ge.load_package(); H, S, sh = pkg.host, pkg.synth, pkg.sharding
H.init([local]); H.set_party(0, 1)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); sp = stream.cuda_stream
log_g = world.bit_length() - 1
for log_n in [int(x) for x in sys.argv[1:]] or [24]:
    n = 1 << log_n; m = n // world
    mine = torch.from_numpy(S.fr_uniform(0xA00 + rank, m).view(np.int64)).cuda()
    cross = lambda data, l0, kind: H.ntt_cross_stage_dev(data.data_ptr(), log_n, log_g, l0, data.shape[1], kind, sp)
    local_ntt = lambda block, kind: H.ntt_dev(block.data_ptr(), log_n - log_g, kind, 1, sp)
    res = {}
    for kind in ("fft", "coset_ifft"):
        for it in range(8):
            if it == 3:
                dist.barrier(); torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
            out = sh.dist_ntt(dist, rank, world, mine, log_n, kind, cross, local_ntt)
        e1.record(stream); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 5], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[kind] = round(float(t.item()), 4)
    if rank == 0:
        print(json.dumps({"tool": "dist_ntt", "n_gpus": world, "log_n": log_n, "ms": res,
                          "Melem_s_fft": round(n / res["fft"] / 1e3, 1)}), flush=True)
dist.barrier(); dist.destroy_process_group()
