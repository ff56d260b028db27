"""Data-parallel plumbing of the cone step: the step's positives (each with its 2N negatives)
are split in contiguous near-equal slices over the ranks, the label table is replicated, and the only
exchange per step is the sum of the table gradient (+ the scalar loss).  The slicing helpers are backend-agnostic
(NCCL on the GPU box, gloo in the CPU tests); PeerExchange / LocalExchange own the peer-memory buffers of the exchange
that is fused into the update kernel."""
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


def exchange_packets(n, ld):
    """16-byte packets of one source region (lec_exchange_packets): two floats per packet + one packet for the loss."""
    return int(n) * int(ld) // 2 + 1


def exchange_bytes(n, ld, world, mode):
    """Bytes of one rank's exchange buffer (lec_exchange_bytes): mode 0 = one-shot packets, 1 = two-shot."""
    from . import _native as N
    return int(N.lib().lec_exchange_bytes(int(n), int(ld), int(world), int(mode)))


ONE_SHOT, TWO_SHOT = 0, 1
# one-shot sends the whole gradient to every peer as 8 bytes per float; above this table size the owner-computes
# two-shot exchange (2 (W-1)/W table volumes per rank, three launches) is used instead
ONE_SHOT_MAX_TABLE_BYTES = 256 << 10   # 8 GPUs: 8192 x 12 floats (393 KB) one-shot 27.5 us, two-shot 23.9 us


def pick_mode(n, ld):
    return ONE_SHOT if int(n) * int(ld) * 4 <= ONE_SHOT_MAX_TABLE_BYTES else TWO_SHOT


class _ExchangeBase:
    """State of the low-latency exchange of include/lec_b200.h (lec_exchange_t): slot / tag bookkeeping, the device error
    flag and the ctypes view of the peer pointers."""

    def _finish(self, ptrs, device, timeout_ms):
        import ctypes
        self.peer_ptrs = (ctypes.c_void_p * self.world)(*ptrs)
        self.step = 0
        self.timeout_ms = int(timeout_ms)
        self.error = torch.zeros(1, dtype=torch.int32, device=device)
        self.loss_global = torch.zeros(1, dtype=torch.float64, device=device)

    def slot_and_tag(self):
        return self.step % 2, self.step + 1

    def fill(self, x):
        """Write this step's lec_exchange_t into the ctypes struct x."""
        import ctypes
        x.peer_bufs = ctypes.cast(self.peer_ptrs, ctypes.c_void_p)
        x.slot_packets, x.world, x.rank = self.slot_packets, self.world, self.rank
        x.slot, x.tag = self.slot_and_tag()
        x.loss_global, x.error, x.timeout_ms = self.loss_global.data_ptr(), self.error.data_ptr(), self.timeout_ms
        x.mode, x.phases = self.mode, 0


class PeerExchange(_ExchangeBase):
    """Exchange buffers of the fused update (include/lec_b200.h: lec_update_rows with a lec_exchange_t): one buffer per
    rank, `packet slot[2][world][slot_packets]`, mapped by every rank through torch symmetric memory (CUDA IPC over
    NVLink / NVSwitch)."""

    def __init__(self, n, ld, device, group, timeout_ms=30000, mode=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.mode = pick_mode(n, ld) if mode is None else int(mode)
        self.slot_packets = exchange_packets(n, ld)
        total = (exchange_bytes(n, ld, self.world, self.mode) + 3) // 4      # floats
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
        self._finish(ptrs, device, timeout_ms)


class LocalExchange(_ExchangeBase):
    """The same exchange between `world` step engines that live in ONE process on ONE device (each on its own stream):
    the "peer" buffers are plain device tensors.  What the single-GPU tests use to run the multi-rank protocol of
    lec_update_rows for real; also a way to drive several model replicas per GPU."""

    def __init__(self, bufs, rank, slot_packets, timeout_ms=5000, mode=ONE_SHOT):
        self.world, self.rank, self.slot_packets, self.mode = len(bufs), int(rank), int(slot_packets), int(mode)
        self.bufs = bufs
        self._finish([int(b.data_ptr()) for b in bufs], bufs[0].device, timeout_ms)

    @staticmethod
    def make(world, n, ld, device, timeout_ms=5000, mode=ONE_SHOT):
        sp = exchange_packets(n, ld)
        floats = (exchange_bytes(n, ld, world, mode) + 3) // 4
        bufs = [torch.zeros(floats, dtype=torch.float32, device=device) for _ in range(world)]
        return [LocalExchange(bufs, r, sp, timeout_ms, mode) for r in range(world)]
