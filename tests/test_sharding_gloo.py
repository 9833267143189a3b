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


# ----------------------------------------------------------------------------- sharded NTT over gloo
def _ntt_worker(rank, world, port, log_n, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import __graft_entry__ as ge
    import pyref as P
    from oracle import oracle as orc
    pkg = ge.load_package()
    sh = pkg.sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, R = 1 << log_n, P.R_MOD
    m, k = n // world, world.bit_length() - 1
    full = pkg.synth.fr_uniform(0x990 + log_n, n)

    def to_t(vals):
        return torch.from_numpy(P.fr_to_mont_arr(vals).view(np.int64).copy())

    def from_t(t):
        return P.fr_from_mont_arr(t.numpy().view(np.uint64).reshape(-1, 4))

    def cross(data, l0, kind):
        """big-int model of mpc_cuda_ntt_cross_stage_dev (stages that pair different devices' blocks)"""
        inverse = kind in ("ifft", "coset_ifft")
        g, slen = data.shape[0], data.shape[1]
        x = [from_t(data[qq]) for qq in range(g)]
        w = P.fr_root_of_unity(log_n)
        for t in range(slen):
            l = l0 + t
            v = [x[qq][t] for qq in range(g)]
            if kind == "coset_fft":
                v = [v[qq] * pow(22, qq * m + l, R) % R for qq in range(g)]
            stages = range(k) if not inverse else range(k - 1, -1, -1)
            for s in stages:
                d = g >> (s + 1)
                for qq in range(g):
                    if qq & d:
                        continue
                    e = ((qq * m + l) % (n >> (s + 1))) << s
                    if not inverse:
                        lo, hi = v[qq], v[qq + d]
                        v[qq], v[qq + d] = (lo + hi) % R, (lo - hi) * pow(w, e, R) % R
                    else:
                        tv = v[qq + d] * pow(w, -e, R) % R
                        v[qq], v[qq + d] = (v[qq] + tv) % R, (v[qq] - tv) % R
            if inverse:
                v = [vv * pow(g, -1, R) % R for vv in v]
                if kind == "coset_ifft":
                    v = [v[qq] * pow(22, -(qq * m + l), R) % R for qq in range(g)]
            for qq in range(g):
                x[qq][t] = v[qq]
        for qq in range(g):
            data[qq] = to_t(x[qq])

    def local_ntt(block, kind):
        block.copy_(torch.from_numpy(orc.ntt(block.numpy().view(np.uint64).reshape(-1, 4), kind).view(np.int64)))

    ok = True
    for fwd, inv in (("fft", "ifft"), ("coset_fft", "coset_ifft")):
        mine = torch.from_numpy(full[rank * m:(rank + 1) * m].view(np.int64).copy())
        out = sh.dist_ntt(dist, rank, world, mine, log_n, fwd, cross, local_ntt)
        expect = orc.ntt(full, fwd)
        idx = np.arange(m) * world + sh.bitrev(rank, k)
        ok &= bool(np.array_equal(out.numpy().view(np.uint64), expect[idx]))
        assert all(sh.ntt_output_owner(int(i), world) == (rank, j) for j, i in enumerate(idx[:4]))
        back = sh.dist_ntt(dist, rank, world, out.clone(), log_n, inv, cross, local_ntt)
        ok &= bool(np.array_equal(back.numpy().view(np.uint64), full[rank * m:(rank + 1) * m]))
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,log_n", [(2, 5), (4, 6)])
def test_sharded_ntt_over_gloo(world, log_n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() + world * 11 + log_n) % 2000
    procs = [ctx.Process(target=_ntt_worker, args=(r, world, port, log_n, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(ok for _, ok in results), results
