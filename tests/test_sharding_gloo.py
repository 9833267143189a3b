"""N > 1 host logic on CPU: world_size-2 (and 3) gloo processes shard one MSM by point range, exchange
Jacobian partials with an all-gather and fold them, exactly as bench.py does over NCCL.  The per-rank
MSM is played by the oracle here (no GPU in this container); the GPU path is covered by -m gpu tests."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import __graft_entry__ as ge
    from oracle import oracle as orc
    pkg = ge.load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bases = orc.g1_generate(0x77, n)
    scalars = pkg.synth.fr_witness_like(0x78, n)
    one = orc.constants()["fq_r"]

    def local_partial(lo, hi):
        xy, inf = orc.g1_msm(bases[lo:hi], scalars[lo:hi])
        return pkg.sharding.affine_to_jacobian(xy, inf, one)

    res = pkg.sharding.sharded_msm(dist, rank, world, n, local_partial, lambda parts: orc.g1_sum_jac(parts))
    if rank == 0:
        full = orc.g1_msm(bases, scalars)
        q.put((res[0].tolist(), int(res[1]), full[0].tolist(), int(full[1])))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 257), (3, 100), (2, 1)])
def test_point_range_sharding_over_gloo(world, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + world * 7 + n) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert got[0] == got[2] and got[1] == got[3]


def test_shard_range_partitions():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    sh = ge.load_package().sharding
    for n in (0, 1, 7, 8, 1 << 20, (1 << 20) + 3):
        for world in (1, 2, 3, 4, 8):
            edges = [sh.shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
