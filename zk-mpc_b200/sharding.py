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
