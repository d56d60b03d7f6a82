"""Multi-GPU sharding of a candidate batch: candidates are independent, so the pair
list is dealt over the ranks by estimated work (longest-processing-time first) with no data-path
collective; the only exchange is one MIN all-reduce of a packed (cost << 32 | candidate) int64 per
round -- NCCL over NVLink on GPUs, gloo in the CPU tests."""
import numpy as np


def lpt_partition(work, world):
    """Deal items (by descending work) to the currently lightest rank.  -> list of index arrays."""
    work = np.asarray(work, np.int64)
    order = np.argsort(-work, kind="stable")
    loads = np.zeros(world, np.int64)
    parts = [[] for _ in range(world)]
    for i in order:
        r = int(np.argmin(loads))
        parts[r].append(int(i))
        loads[r] += work[i]
    return [np.array(sorted(p), np.int64) for p in parts]


def pack_best(costs, ids):
    """(cost << 32 | id) of the cheapest candidate of a shard; ties resolve to the smallest id."""
    costs = np.asarray(costs, np.int64); ids = np.asarray(ids, np.int64)
    if len(costs) == 0:
        return np.int64(np.iinfo(np.int64).max)
    return np.int64(((costs << 32) | ids).min())


def unpack_best(packed):
    packed = int(packed)
    return packed >> 32, packed & 0xFFFFFFFF


def allreduce_best(packed, device=None):
    """MIN all-reduce of the packed best candidate over torch.distributed (NCCL or gloo)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([int(packed)], dtype=torch.int64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return int(t.item())
