"""Data-parallel helpers (src/theia/scripts/train/train_rvfm.py:213,258 uses torch DDP + NCCL).

The distillation step shards naturally over the batch (no cross-sample statistic on the path), so the only
collective is the gradient all-reduce.  `RobotVisionFM` works under the reference's own
`DistributedDataParallel` wrapper unchanged; `allreduce_flat_grads` is the one-call alternative for callers
that drive the step themselves: every gradient lives in ONE flat fp32 buffer, so a single NCCL all-reduce
(sum) followed by a 1/world scale replaces DDP's bucket copies."""
from __future__ import annotations

import torch
import torch.distributed as dist


def allreduce_flat_grads(flat_grads: torch.Tensor, group=None, average: bool = True) -> torch.Tensor:
    """In-place all-reduce of a flat gradient buffer (NCCL on GPU tensors, gloo on CPU tensors)."""
    if not (dist.is_available() and dist.is_initialized()):
        return flat_grads
    world = dist.get_world_size(group)
    if world == 1:
        return flat_grads
    dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat_grads.mul_(1.0 / world)
    return flat_grads


def shard_batch(global_batch: int, rank: int, world: int) -> tuple[int, int]:
    """[start, stop) of this rank's slice of a global batch (wds.split_by_node equivalent for tensors)."""
    if global_batch % world != 0:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world}")
    per = global_batch // world
    return rank * per, (rank + 1) * per
