"""Data-parallel plumbing of the cone step: the step's positives (each with its 2N negatives)
are split in contiguous near-equal slices over the ranks, the label table is replicated, and the only
exchange per step is the sum of the table gradient (+ the scalar loss).  Backend-agnostic: NCCL on the
GPU box, gloo in the CPU tests."""
import torch
import torch.distributed as dist


def shard_bounds(n_items, rank, world):
    """[lo, hi) of rank's contiguous slice; slices differ by at most one item and cover [0, n_items)."""
    base, rem = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_groups(pos_from, pos_to, neg_to, neg_from, rank, world):
    """Slice one step's batch (compact training layout) for this rank; negatives stay with their positive."""
    lo, hi = shard_bounds(len(pos_from), rank, world)
    return pos_from[lo:hi], pos_to[lo:hi], neg_to[lo:hi], neg_from[lo:hi]


def allreduce_grad_and_loss(grad_table, loss, group=None):
    """Sum the dense table gradient [n, D] and the scalar loss over the ranks, in place."""
    if group is None and not dist.is_initialized():
        return grad_table, loss
    if dist.get_world_size(group) > 1:
        dist.all_reduce(grad_table, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=group)
    return grad_table, loss
