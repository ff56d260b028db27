"""Step engine: one fused training step of a label-only cone model on preallocated buffers.

This is the public fast path (what bench.py drives).  It does what one iteration of the reference's
`pass_samples('train')` loop does between `optimizer.zero_grad()` and the weight update
(order_embeddings_h.py:752-775 for the hyperbolic trainer, order_embeddings.py:619-635 for the
Euclidean one) with the negatives already drawn:

    rows, aux = transform(table)                    lec_rows_fwd   (also clears grad_rows; aux = per-row
                                                    aperture terms in fp64)
    loss, dL/drows = cone loss over B*(1+2N) pairs   lec_pairs_grouped
    [sum dL/drows and loss over ranks]               NCCL all-reduce, only when world_size > 1
    dL/dtable = J^T dL/drows                         lec_rows_bwd
    table     = RSGD(table, dL/dtable)               lec_rsgd_update   (hyperbolic)  | plain SGD (Euclidean)

Index batches use the compact training layout of include/lec_b200.h (lec_pairs_grouped) as one int32
block [pos_from | pos_to | neg_to | neg_from] so a step's indices travel host->device in one copy.
"""
import numpy as np
import torch

from . import _native as N
from . import ops
from . import sharding
from .criterion import inner_radius


def pack_index_block(pos_from, pos_to, neg_to, neg_from, pin=True):
    """Host int32 block [B | B | B*N | B*N] for ConeStep.step_host."""
    parts = [np.ascontiguousarray(a, dtype=np.int32).reshape(-1) for a in (pos_from, pos_to, neg_to, neg_from)]
    blk = torch.from_numpy(np.concatenate(parts))
    return blk.pin_memory() if (pin and torch.cuda.is_available()) else blk


class ConeStep:
    def __init__(self, table, geom, n_neg, max_groups, K=None, alpha=1.0, lr=1e-3, row_mode=None, update="auto",
                 precision=ops.PREC_F64CORE, process_group=None, replicas=None, comm="auto"):
        N.require_cuda(table)
        if table.dtype != torch.float32 or not table.is_contiguous():
            raise N.LecError("ConeStep: table must be a contiguous float32 CUDA tensor (updated in place)")
        self.table = table
        self.geom = geom
        self.n, self.D = table.shape
        self.ld = ops.padded_dim(self.D)
        self.n_neg = int(n_neg)
        self.K = {"euc": 3.0, "hyp": 0.1, "oe": 0.0}[geom] if K is None else float(K)
        self.alpha = float(alpha)
        self.lr = float(lr)
        self.row_mode = {"euc": N.ROWS_EUC_SOFTCLIP, "hyp": N.ROWS_HYP_SHELL, "oe": N.ROWS_NONE}[geom] \
            if row_mode is None else int(row_mode)
        self.update = ("rsgd" if geom == "hyp" else "sgd") if update == "auto" else update
        self.r_in = float(inner_radius(self.K)) if geom == "hyp" else 0.0
        self.precision = int(precision)
        self.pg = process_group
        self.max_groups = int(max_groups)
        dev = table.device
        self.replicas = ops.default_replicas(self.n, self.ld) if replicas is None else int(replicas)
        self.rows = torch.empty((self.n, self.ld), device=dev, dtype=torch.float32)
        self.aux = torch.empty((self.n, 4), device=dev, dtype=torch.float64)
        self.grad_rows = torch.empty((self.replicas, self.n, self.ld), device=dev, dtype=torch.float32)
        self.grad_table = torch.empty((self.n, self.D), device=dev, dtype=torch.float32)
        self.E_pos = torch.empty(self.max_groups, device=dev, dtype=torch.float32)
        self.E_neg = torch.empty((self.max_groups, 2 * self.n_neg), device=dev, dtype=torch.float32)
        self.loss = torch.zeros(1, device=dev, dtype=torch.float64)
        self.idx_dev = torch.empty(self.max_groups * (2 + 2 * self.n_neg), device=dev, dtype=torch.int32)
        self.loss_host = torch.zeros(1, dtype=torch.float64).pin_memory()
        self.kernel_events = None  # optional (start, stop) pairs around the pair kernel, set by bench
        self._struct = None
        # multi-GPU exchange: "p2p" = one-shot all-reduce over NVLink peer memory fused into the RSGD kernel,
        # "nccl" = all_reduce of the table gradient; "auto" tries p2p for the RSGD update and falls back
        self.comm, self.comm_note, self.px = "none", "", None
        if self.pg is not None and torch.distributed.get_world_size(self.pg) > 1:
            self.comm = "nccl"
            if comm in ("auto", "p2p") and self.update == "rsgd":
                try:
                    self.px = sharding.PeerExchange(self.n, self.D, dev, self.pg)
                    self.comm = "p2p"
                except Exception as e:  # noqa: BLE001 -- any failure of the symmetric-memory setup
                    if comm == "p2p":
                        raise
                    self.comm_note = "p2p unavailable (%s: %s)" % (type(e).__name__, str(e)[:120])
            self.loss_global = torch.zeros(1, device=dev, dtype=torch.float64)

    # -- pieces ---------------------------------------------------------------------------------
    def _split(self, blk, B):
        Nn = self.n_neg
        return blk[:B], blk[B:2 * B], blk[2 * B:2 * B + B * Nn], blk[2 * B + B * Nn:2 * B + 2 * B * Nn]

    def forward_backward(self, pos_from, pos_to, neg_to, neg_from, w_pos=None, w_neg=None):
        """rows, loss and d loss / d rows for one batch (no collective, no update)."""
        lib, st = N.lib(), N.stream_ptr(self.table.device)
        B = int(pos_from.numel())
        if B > self.max_groups:
            raise N.LecError("batch of %d positives exceeds max_groups=%d" % (B, self.max_groups))
        N.check(lib.lec_rows_fwd(N._p(self.table), self.n, self.D, self.row_mode, N.GEOM[self.geom], self.K,
                                 N._p(self.rows), self.ld, N._p(self.aux), N._p(self.grad_rows), self.replicas,
                                 N._p(self.loss), st), "lec_rows_fwd")
        ev = self.kernel_events
        if ev is not None:
            ev[0].record()
        N.check(lib.lec_pairs_grouped(
            N.GEOM[self.geom], self.precision, N._p(self.rows), N._p(self.aux), self.n, self.D, self.ld,
            N._p(pos_from), N._p(pos_to),
            N._p(neg_to), N._p(neg_from), pos_from.element_size(), B, self.n_neg, N._p(w_pos), N._p(w_neg), self.K,
            self.alpha, N._p(self.E_pos), N._p(self.E_neg), N._p(self.loss), N._p(self.grad_rows), self.replicas, st),
            "lec_pairs_grouped")
        if ev is not None:
            ev[1].record()
        return B

    def reduce_and_update(self):
        lib, st = N.lib(), N.stream_ptr(self.table.device)
        multi = self.pg is not None and torch.distributed.get_world_size(self.pg) > 1
        if not multi and self.update == "rsgd" and self.row_mode == N.ROWS_HYP_SHELL:
            # straight-through rows: d/dtable == d/drows; the update sums the replicas itself
            N.check(lib.lec_rsgd_update(N._p(self.table), N._p(self.grad_rows), self.replicas, self.n, self.D, self.ld,
                                        self.lr, self.r_in, 0, N._p(self.grad_table), st), "lec_rsgd_update")
            return
        if multi and self.comm == "p2p":
            import ctypes
            px = self.px
            slot, tag = px.slot_and_tag()
            # partial d/dtable of this rank straight into its exchange slot, publish, fused reduce + update
            N.check(lib.lec_rows_bwd(N._p(self.table), N._p(self.grad_rows), self.replicas, self.n, self.D, self.ld,
                                     self.row_mode, self.K, ctypes.c_void_p(px.my_slot_ptr(slot)), 0, st),
                    "lec_rows_bwd")
            N.check(lib.lec_p2p_publish(N._p(self.loss), px.peer_ptrs, px.slot_floats, px.world, px.rank, slot, tag, st),
                    "lec_p2p_publish")
            N.check(lib.lec_rsgd_update_p2p(N._p(self.table), px.peer_ptrs, px.slot_floats, px.world, px.rank, slot,
                                            tag, self.n, self.D, self.lr, self.r_in, 0, N._p(self.loss_global),
                                            N._p(px.error), st), "lec_rsgd_update_p2p")
            px.step += 1
            return
        # d/dtable = J^T (sum of replicas); it is linear, so ranks can be summed after it
        N.check(lib.lec_rows_bwd(N._p(self.table), N._p(self.grad_rows), self.replicas, self.n, self.D, self.ld,
                                 self.row_mode, self.K, N._p(self.grad_table), 0, st), "lec_rows_bwd")
        if multi:
            # one collective per step: the table gradient.  The scalar loss stays rank-local until someone
            # asks for it (global_loss), so logging does not put a second latency-bound all-reduce on the
            # critical path.
            torch.distributed.all_reduce(self.grad_table, group=self.pg)
        if self.update == "rsgd":
            N.check(lib.lec_rsgd_update(N._p(self.table), N._p(self.grad_table), 1, self.n, self.D, self.D, self.lr,
                                        self.r_in, 0, N._p(self.grad_table), st), "lec_rsgd_update")
        elif self.update == "sgd":
            self.table.add_(self.grad_table, alpha=-self.lr)
        elif self.update != "none":
            raise N.LecError("unknown update rule %r" % (self.update,))

    def global_loss(self):
        """Loss of the latest step summed over ranks (float64 tensor on the device)."""
        if self.comm == "p2p":
            if int(self.px.error.item()) != 0:
                raise N.LecError("peer exchange timed out: a rank did not publish its gradient")
            return self.loss_global.clone()
        out = self.loss.clone()
        if self.pg is not None and torch.distributed.get_world_size(self.pg) > 1:
            torch.distributed.all_reduce(out, group=self.pg)
        return out

    # -- whole steps ----------------------------------------------------------------------------
    def _step_struct(self):
        """lec_step_t with everything that does not change from step to step filled in."""
        import ctypes
        s = N.LecStep()
        s.geom, s.precision, s.row_mode = N.GEOM[self.geom], self.precision, self.row_mode
        s.update = {"none": 0, "rsgd": 1}[self.update]
        s.lambda_mode = 0
        s.K, s.alpha, s.lr, s.r_in = self.K, self.alpha, self.lr, self.r_in
        s.table, s.n, s.D, s.ld = self.table.data_ptr(), self.n, self.D, self.ld
        s.rows, s.aux, s.grad_rows = self.rows.data_ptr(), self.aux.data_ptr(), self.grad_rows.data_ptr()
        s.grad_replicas, s.grad_table = self.replicas, self.grad_table.data_ptr()
        s.N = self.n_neg
        s.E_pos, s.E_neg, s.loss = self.E_pos.data_ptr(), self.E_neg.data_ptr(), self.loss.data_ptr()
        s.world = 0
        if self.comm == "p2p":
            px = self.px
            s.peer_bufs = ctypes.cast(px.peer_ptrs, ctypes.c_void_p)
            s.slot_floats, s.world, s.rank = px.slot_floats, px.world, px.rank
            s.loss_global, s.error = self.loss_global.data_ptr(), px.error.data_ptr()
        return s

    def step_device(self, pos_from, pos_to, neg_to, neg_from, w_pos=None, w_neg=None):
        """Indices already on the device (int32 or int64).  Returns the device loss (float64[1])."""
        if self.update in ("none", "rsgd") and self.comm in ("none", "p2p"):
            # the whole step as ONE call into the library (lec_cone_step): one FFI crossing, 3-5 launches
            B = int(pos_from.numel())
            if B > self.max_groups:
                raise N.LecError("batch of %d positives exceeds max_groups=%d" % (B, self.max_groups))
            s = self._struct if self._struct is not None else self._step_struct()
            self._struct = s
            s.pos_from, s.pos_to = pos_from.data_ptr(), pos_to.data_ptr()
            s.neg_to, s.neg_from = neg_to.data_ptr(), neg_from.data_ptr()
            s.idx_bytes, s.B = pos_from.element_size(), B
            s.w_pos = w_pos.data_ptr() if w_pos is not None else None
            s.w_neg = w_neg.data_ptr() if w_neg is not None else None
            if self.comm == "p2p":
                s.slot, s.tag = self.px.slot_and_tag()
            ev = self.kernel_events
            if ev is not None:
                for e in ev:
                    if not e.cuda_event:
                        e.record()  # torch creates the cudaEvent lazily
                s.ev_pairs_start, s.ev_pairs_stop = ev[0].cuda_event, ev[1].cuda_event
            else:
                s.ev_pairs_start, s.ev_pairs_stop = None, None
            import ctypes
            N.check(N.lib().lec_cone_step(ctypes.byref(s), N.stream_ptr(self.table.device)), "lec_cone_step")
            if self.comm == "p2p":
                self.px.step += 1
            return self.loss
        self.forward_backward(pos_from, pos_to, neg_to, neg_from, w_pos, w_neg)
        self.reduce_and_update()
        return self.loss

    def step_host(self, index_block, B):
        """index_block: pinned host int32 block from pack_index_block.  Copies it in, runs the step and
        reads the scalar loss back (one H2D, one D2H, one sync) -- the end-to-end path."""
        n = B * (2 + 2 * self.n_neg)
        dst = self.idx_dev[:n]
        dst.copy_(index_block[:n], non_blocking=True)
        self.step_device(*self._split(dst, B))
        self.loss_host.copy_(self.loss, non_blocking=True)
        torch.cuda.current_stream(self.table.device).synchronize()
        return float(self.loss_host[0])
