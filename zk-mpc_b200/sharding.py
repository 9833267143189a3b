"""Point-range sharding of one party's share MSM across the GPUs of a box.

Rank k owns bases/scalars [k*n/g, (k+1)*n/g), runs the full single-GPU Pippenger pipeline and yields one
Jacobian partial (18 u64 limbs, z = 0 for infinity).  There is no NCCL reduction for elliptic-curve
addition, so the exchange is one all-gather of g x 144 bytes followed by g-1 group additions and one
normalisation on the gathering rank.  The functions here are backend-agnostic: they take the
torch.distributed module (nccl on the GPU box, gloo in the CPU tests) and callables for the two local
steps, so the same code is exercised by tests/test_sharding_gloo.py without a GPU.
"""
import numpy as np


def shard_range(n, rank, world):
    """[lo, hi) of rank's points: contiguous, sizes differ by at most one, union = [0, n)"""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def affine_to_jacobian(xy, inf, one_mont):
    """(x, y, 1) in Montgomery limbs, or z = 0 for infinity; pure limb packing, no field arithmetic"""
    xy = np.asarray(xy, dtype=np.uint64)
    limbs = xy.size // 2
    z = np.zeros(limbs, dtype=np.uint64) if inf else np.asarray(one_mont, dtype=np.uint64)
    return np.concatenate([xy, z])


def sharded_msm(dist, rank, world, n, local_partial, sum_partials, device=None):
    """local_partial(lo, hi) -> (limbs,) uint64 Jacobian partial of this rank's range;
    sum_partials((world, limbs) uint64) -> result on rank 0 (other ranks return None)."""
    import torch
    lo, hi = shard_range(n, rank, world)
    part = np.ascontiguousarray(local_partial(lo, hi), dtype=np.uint64)
    mine = torch.from_numpy(part.view(np.int64).copy())
    if device is not None:
        mine = mine.to(device)
    if world > 1:
        gathered = torch.empty(world * mine.numel(), dtype=torch.int64, device=mine.device)
        dist.all_gather_into_tensor(gathered, mine) if device is not None else \
            dist.all_gather(list(gathered.view(world, -1).unbind(0)), mine)
    else:
        gathered = mine
    if rank != 0:
        return None
    return sum_partials(gathered.cpu().numpy().view(np.uint64).reshape(world, -1))


# ----------------------------------------------------------------------------- multi-GPU share NTT
def bitrev(x, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


def ntt_output_owner(i, world):
    """(rank, local index) of forward output X[i] when the transform ran over `world` = 2^k devices"""
    k = world.bit_length() - 1
    return bitrev(i % world, k), i // world


def _all_to_all(dist, send, world):
    """send: (world, slice, 4) tensor, chunk q goes to rank q; returns (world, slice, 4) with chunk q from rank q"""
    import torch
    recv = torch.empty_like(send)
    if dist.get_backend() == "gloo":          # CPU tests: gloo has no all_to_all_single on every build
        everything = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(everything, send.contiguous())
        rank = dist.get_rank()
        for q in range(world):
            recv[q] = everything[q][rank]
        return recv
    dist.all_to_all_single(recv, send.contiguous())
    return recv


def dist_ntt(dist, rank, world, block, log_n, kind, cross_stage, local_ntt):
    """One party's size-2^log_n NTT over `world` = 2^k ranks (include/mpc_cuda.h
    mpc_cuda_ntt_cross_stage_dev).  `block`: (n/world, 4) int64 tensor, this rank's part of the vector —
    natural block order for the forward kinds (fft, coset_fft), the transposed order the forward kinds
    produce (ntt_output_owner) for the inverse kinds, which return natural block order.
    cross_stage(data (world, slice, 4) tensor, l0, kind) and local_ntt(block tensor, kind) run in place
    (C ABI on the GPU; models in the CPU test).  Two all-to-alls of (world-1)/world of the block each."""
    inverse = kind in ("ifft", "coset_ifft")
    if world == 1:
        local_ntt(block, kind)
        return block
    m = block.shape[0]
    slice_len = m // world
    assert slice_len * world == m and m * world == 1 << log_n
    if inverse:
        local_ntt(block, "ifft")
    data = _all_to_all(dist, block.view(world, slice_len, 4), world)
    cross_stage(data, rank * slice_len, kind)
    block = _all_to_all(dist, data, world).view(m, 4)
    if not inverse:
        local_ntt(block, "fft")
    return block
