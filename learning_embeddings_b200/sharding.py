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


class PeerExchange:
    """Exchange buffers for the fused all-reduce + RSGD update (include/lec_b200.h: lec_p2p_publish,
    lec_rsgd_update_p2p): one buffer per rank, mapped by every rank through torch symmetric memory
    (CUDA IPC over NVLink / NVSwitch).  Two regions per buffer, one per protocol of include/lec_b200.h:
    push (lec_p2p_push / lec_rsgd_update_rows_p2p)   float slot[2][world][slot_floats]; uint32 flag[2][world]
    pull (lec_p2p_publish / lec_rsgd_update_p2p)     float slot[2][slot_floats];        uint32 flag[2][world]
    `peer_ptrs` points at the push regions, `peer_ptrs_pull` at the pull regions."""

    def __init__(self, n, D, device, group):
        import ctypes
        import torch.distributed._symmetric_memory as symm_mem
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.slot_floats = (n * D + 2 + 3) // 4 * 4
        push = (2 * self.world * self.slot_floats + 2 * self.world + 3) // 4 * 4
        pull = (2 * self.slot_floats + 2 * self.world + 3) // 4 * 4
        self.pull_offset = push      # floats
        total = push + pull
        self.buf = symm_mem.empty(total, dtype=torch.float32, device=device)
        name = getattr(group, "group_name", None)
        try:
            self.handle = symm_mem.rendezvous(self.buf, group=name if name is not None else group)
        except TypeError:
            self.handle = symm_mem.rendezvous(self.buf, name)
        self.buf.zero_()
        torch.cuda.synchronize(device)
        dist.barrier(group)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        if len(ptrs) != self.world or ptrs[self.rank] != self.buf.data_ptr():
            raise RuntimeError("symmetric memory rendezvous returned unexpected peer pointers")
        self.peer_ptrs = (ctypes.c_void_p * self.world)(*ptrs)
        self.peer_ptrs_pull = (ctypes.c_void_p * self.world)(*[q + 4 * self.pull_offset for q in ptrs])
        self.counter = torch.zeros(1, dtype=torch.int32, device=device)   # last-block counter of lec_p2p_push
        self.step = 0
        self.error = torch.zeros(1, dtype=torch.int32, device=device)

    def slot_and_tag(self):
        return self.step % 2, self.step + 1

    def my_slot_ptr(self, slot):
        return self.buf.data_ptr() + 4 * (self.pull_offset + slot * self.slot_floats)
