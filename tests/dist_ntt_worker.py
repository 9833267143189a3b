"""torchrun worker: multi-GPU share NTT (sharding.dist_ntt over NCCL + mpc_cuda_ntt_cross_stage_dev) against
the oracle's full-size transform.  Launched by tests/test_gpu_multi.py; exits non-zero on mismatch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = ge.load_package()
    H, S, sh = pkg.host, pkg.synth, pkg.sharding
    H.init([local])
    H.set_party(0, 1)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sp = stream.cuda_stream
    log_g = world.bit_length() - 1
    for log_n in [int(x) for x in sys.argv[1:]] or [2 * log_g, 8, 13, 16]:
        n = 1 << log_n
        m = n // world
        full = S.fr_uniform(0x900 + log_n, n)

        def cross(data, l0, kind):
            H.ntt_cross_stage_dev(data.data_ptr(), log_n, log_g, l0, data.shape[1], kind, sp)

        def local_ntt(block, kind):
            H.ntt_dev(block.data_ptr(), log_n - log_g, kind, 1, sp)

        for fwd, inv in (("fft", "ifft"), ("coset_fft", "coset_ifft")):
            mine = torch.from_numpy(full[rank * m:(rank + 1) * m].view(np.int64).copy()).cuda()
            out = sh.dist_ntt(dist, rank, world, mine, log_n, fwd, cross, local_ntt)
            torch.cuda.synchronize()
            expect = orc.ntt(full, fwd)
            idx = np.arange(m) * world + sh.bitrev(rank, log_g)          # X[m*g + bitrev(r)] lives at (r, m)
            got = out.cpu().numpy().view(np.uint64)
            if not np.array_equal(got, expect[idx]):
                print("MISMATCH forward", fwd, log_n, rank, flush=True)
                sys.exit(3)
            back = sh.dist_ntt(dist, rank, world, out.clone(), log_n, inv, cross, local_ntt)
            torch.cuda.synchronize()
            if not np.array_equal(back.cpu().numpy().view(np.uint64), full[rank * m:(rank + 1) * m]):
                print("MISMATCH inverse", inv, log_n, rank, flush=True)
                sys.exit(4)
    dist.barrier()
    if rank == 0:
        print("dist_ntt OK world=%d" % world, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
