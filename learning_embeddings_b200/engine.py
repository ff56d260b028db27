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
import os

import numpy as np
import torch

from . import _native as N
from . import ops
from . import sharding
from .criterion import inner_radius


def index_dtype_for(n_rows):
    """Narrowest index type the kernels take for a table of n_rows rows: uint16 up to 65 536 rows (ETHEC has
    723), else int32.  The index block is the only per-step host->device traffic, so its width is the
    end-to-end cost of a step."""
    return np.uint16 if n_rows <= 65536 else np.int32


def pack_index_block(pos_from, pos_to, neg_to, neg_from, pin=True, dtype=np.int32):
    """Host index block [B | B | B*N | B*N] (int32, or uint16 for tables of <= 65 536 rows) for ConeStep.step_host."""
    parts = [np.ascontiguousarray(a).reshape(-1) for a in (pos_from, pos_to, neg_to, neg_from)]
    cat = np.concatenate(parts)
    if cat.size and (cat.min() < 0 or cat.max() > np.iinfo(dtype).max):
        raise ValueError("index out of range for %s" % np.dtype(dtype).name)
    blk = torch.from_numpy(cat.astype(dtype))
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
        # Fused step (lec_cone_step with fused = 1): RSGD on straight-through rows is followed, in the same launch, by
        # the row transform the next step starts with -- two launches per step (pairs, update+rows) instead of three.
        # The pair kernel then adds into loss_acc and the update moves it to `loss`.
        self.loss_acc = torch.zeros(1, device=dev, dtype=torch.float64)
        self.fused = (self.update == "rsgd" and self.row_mode == N.ROWS_HYP_SHELL
                      and os.environ.get("LEC_FUSED_STEP", "1") != "0")
        self._rows_valid = False
        # host->device staging: `depth` slots so that the copy of step i+1 overlaps the kernels of step i
        self.depth = 2
        self._idx_bytes_dev = [torch.empty(self.max_groups * (2 + 2 * self.n_neg) * 4, device=dev, dtype=torch.uint8)
                               for _ in range(self.depth)]
        self.loss_host = torch.zeros(self.depth, dtype=torch.float64).pin_memory()
        self._copy_stream = None
        self._ev = None          # per slot: (indices copied, kernels done with the slot, loss read back)
        self._inflight = [False] * self.depth
        self._submitted = 0
        self._losses = []
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
        self._rows_valid = False   # the replicas now hold a gradient the fused step did not put there
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

    def invalidate_rows(self):
        """Call after changing `table` from outside the engine (loading weights, a manual update): the fused step
        otherwise reuses the transformed rows its previous update left behind."""
        self._rows_valid = False

    def reduce_and_update(self):
        lib, st = N.lib(), N.stream_ptr(self.table.device)
        self._rows_valid = False
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
            N.check(lib.lec_p2p_publish(N._p(self.loss), px.peer_ptrs_pull, px.slot_floats, px.world, px.rank, slot, tag, st),
                    "lec_p2p_publish")
            N.check(lib.lec_rsgd_update_p2p(N._p(self.table), px.peer_ptrs_pull, px.slot_floats, px.world, px.rank, slot,
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
        elif self.update in ("sgd", "adam"):
            self._apply_torch_update()
        elif self.update != "none":
            raise N.LecError("unknown update rule %r" % (self.update,))

    def _apply_torch_update(self):
        """Euclidean trainers hand the table to a stock torch optimiser (order_embeddings.py:563-565: SGD or Adam); so
        does the engine -- plain SGD, or torch's fused Adam on the table.  Not the product."""
        if self.update == "sgd":
            self.table.add_(self.grad_table, alpha=-self.lr)
            return
        if getattr(self, "_adam", None) is None:
            self._adam_param = torch.nn.Parameter(self.table, requires_grad=False)
            self._adam = torch.optim.Adam([self._adam_param], lr=self.lr, fused=True)
        self._adam_param.grad = self.grad_table
        self._adam.step()

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
        s.update = 1 if self.update == "rsgd" else 0   # sgd / adam: the library leaves d loss / d table in grad_table
        s.lambda_mode = 0
        s.K, s.alpha, s.lr, s.r_in = self.K, self.alpha, self.lr, self.r_in
        s.table, s.n, s.D, s.ld = self.table.data_ptr(), self.n, self.D, self.ld
        s.rows, s.aux, s.grad_rows = self.rows.data_ptr(), self.aux.data_ptr(), self.grad_rows.data_ptr()
        s.grad_replicas, s.grad_table = self.replicas, self.grad_table.data_ptr()
        s.N = self.n_neg
        s.E_pos, s.E_neg, s.loss = self.E_pos.data_ptr(), self.E_neg.data_ptr(), self.loss.data_ptr()
        s.world = 0
        s.fused = 1 if self.fused else 0
        s.loss_acc = self.loss_acc.data_ptr()
        if self.comm == "p2p":
            px = self.px
            s.counter = px.counter.data_ptr()
            s.peer_bufs = ctypes.cast(px.peer_ptrs if self.fused else px.peer_ptrs_pull, ctypes.c_void_p)
            s.slot_floats, s.world, s.rank = px.slot_floats, px.world, px.rank
            s.loss_global, s.error = self.loss_global.data_ptr(), px.error.data_ptr()
        return s

    def step_device(self, pos_from, pos_to, neg_to, neg_from, w_pos=None, w_neg=None):
        """Indices already on the device (int32 or int64).  Returns the device loss (float64[1])."""
        if self.update in ("none", "rsgd") and self.comm in ("none", "p2p") or \
                self.update in ("sgd", "adam") and self.comm == "none":
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
            if self.fused and not self._rows_valid:
                # first step (or the table changed behind the engine's back): rows, aux, cleared replicas and loss
                N.check(N.lib().lec_rows_fwd(N._p(self.table), self.n, self.D, self.row_mode, N.GEOM[self.geom], self.K,
                                             N._p(self.rows), self.ld, N._p(self.aux), N._p(self.grad_rows), self.replicas,
                                             N._p(self.loss_acc), N.stream_ptr(self.table.device)), "lec_rows_fwd")
                self._rows_valid = True
            N.check(N.lib().lec_cone_step(ctypes.byref(s), N.stream_ptr(self.table.device)), "lec_cone_step")
            if not self.fused:
                self._rows_valid = False
            if self.update in ("sgd", "adam"):
                self._apply_torch_update()
            if self.comm == "p2p":
                self.px.step += 1
            return self.loss
        self.forward_backward(pos_from, pos_to, neg_to, neg_from, w_pos, w_neg)
        self.reduce_and_update()
        return self.loss

    def step_sampled(self, graph, pos_from, pos_to, seed, step):
        """Device-resident training step: negatives are drawn on the GPU (sampler.SamplerGraph.draw_philox: the
        reference's candidate sets and uniform law, Philox stream keyed by (seed, step)) and consumed by the step in the
        same stream -- only the positive edges ever come from the host.  The fast mode: not the `random.choice` stream."""
        B = int(pos_from.numel())
        key = (pos_from.dtype, B)
        if getattr(self, "_neg_key", None) != key:
            dev = self.table.device
            self._neg_bufs = (torch.empty((B, self.n_neg), dtype=pos_from.dtype, device=dev),
                              torch.empty((B, self.n_neg), dtype=pos_from.dtype, device=dev))
            self._neg_key = key
        neg_to, neg_from = graph.draw_philox(pos_from, pos_to, self.n_neg, seed, step, out=self._neg_bufs, check=False)
        return self.step_device(pos_from, pos_to, neg_to.view(-1), neg_from.view(-1))

    def _slot_view(self, slot, n, dtype):
        return self._idx_bytes_dev[slot][:n * dtype.itemsize].view(dtype)

    def step_host(self, index_block, B):
        """index_block: pinned host block from pack_index_block (uint16 / int32).  Copies it in, runs the
        step and reads the scalar loss back (one H2D, one D2H, one sync) -- the synchronous end-to-end path."""
        self._collect()
        n = B * (2 + 2 * self.n_neg)
        dst = self._slot_view(0, n, index_block.dtype)
        dst.copy_(index_block[:n], non_blocking=True)
        self.step_device(*self._split(dst, B))
        self.loss_host[:1].copy_(self.loss, non_blocking=True)
        torch.cuda.current_stream(self.table.device).synchronize()
        return float(self.loss_host[0])

    def submit_host(self, index_block, B):
        """Pipelined end-to-end step: the index block goes host->device on a side stream into one of `depth`
        staging slots while the previous step's kernels run; the step is enqueued behind its copy; its loss
        comes back device->host asynchronously.  Blocks only when the slot it needs is still in flight
        (i.e. on the loss of step i - depth).  Losses are returned, in order, by drain()."""
        dev = self.table.device
        main = torch.cuda.current_stream(dev)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(dev)
            self._ev = [tuple(torch.cuda.Event() for _ in range(3)) for _ in range(self.depth)]
        slot = self._submitted % self.depth
        ev_copied, ev_free, ev_loss = self._ev[slot]
        if self._inflight[slot]:
            ev_loss.synchronize()
            self._losses.append(float(self.loss_host[slot]))
        n = B * (2 + 2 * self.n_neg)
        dst = self._slot_view(slot, n, index_block.dtype)
        with torch.cuda.stream(self._copy_stream):
            if self._inflight[slot]:
                self._copy_stream.wait_event(ev_free)
            dst.copy_(index_block[:n], non_blocking=True)
            ev_copied.record(self._copy_stream)
        main.wait_event(ev_copied)
        self.step_device(*self._split(dst, B))
        ev_free.record(main)
        self.loss_host[slot:slot + 1].copy_(self.loss, non_blocking=True)
        ev_loss.record(main)
        self._inflight[slot] = True
        self._submitted += 1

    def drain(self):
        """Wait for every submitted step; returns their losses in submission order (and forgets them)."""
        self._collect()
        out, self._losses = self._losses, []
        return out

    def _collect(self):
        if self._ev is not None:
            for k in range(self.depth):
                slot = (self._submitted + k) % self.depth  # oldest first
                if self._inflight[slot]:
                    self._ev[slot][2].synchronize()
                    self._losses.append(float(self.loss_host[slot]))
                    self._inflight[slot] = False


class JointConeStep:
    """One fused training step of a JOINT image+label cone model (the reference's oe.py / oe_h.py trainers,
    `JointEmbeddings.pass_samples('train')`, oe.py:1489-1560) on preallocated buffers:

        X        = features[img_sel]                         gather of the step's distinct images  (torch)
        Y        = X @ fc1.weight^T + fc1.bias               FeatNet.fc1, oe.py:97,113            (cuBLAS)
        rows,aux = [transform(table) ; transform(Y)]         lec_rows_fwd x2  (Embedder / FeatNet tail)
        loss, dL/drows over B*(1+2N) pairs                   lec_pairs_grouped on the concatenated row table
        dL/dtable, dL/dY                                     lec_reduce_replicas + lec_rows_bwd x2
        dL/dfc1.weight = dL/dY^T @ X, dL/dfc1.bias           cuBLAS / torch
        [one all-reduce of the flat gradient buffer]         NCCL, only when world_size > 1
        Adam (fused) on table, fc1.weight, fc1.bias          torch.optim.Adam, the reference's default optimizer

    Endpoints are row numbers of the concatenated table: < n_labels a label, >= n_labels image
    (index - n_labels) of this step's img_sel.  The FeatNet GEMM stays cuBLAS (not the product)."""

    def __init__(self, table, fc_weight, fc_bias, features, geom, n_neg, max_groups, max_images, K=None, alpha=1.0,
                 lr=1e-3, precision=ops.PREC_F64CORE, process_group=None):
        N.require_cuda(table, fc_weight, fc_bias, features)
        self.table, self.fc_w, self.fc_b, self.features = table, fc_weight, fc_bias, features
        self.geom = geom
        self.n, self.D = table.shape
        self.ld = ops.padded_dim(self.D)
        self.n_neg, self.max_groups, self.max_images = int(n_neg), int(max_groups), int(max_images)
        self.K = {"euc": 3.0, "hyp": 0.1, "oe": 0.0}[geom] if K is None else float(K)
        self.alpha, self.lr, self.precision = float(alpha), float(lr), int(precision)
        self.lab_mode = {"euc": N.ROWS_EUC_SOFTCLIP, "hyp": N.ROWS_HYP_TANH, "oe": N.ROWS_NONE}[geom]
        self.img_mode = {"euc": N.ROWS_EUC_SOFTCLIP, "hyp": N.ROWS_HYP_TANH_FEAT, "oe": N.ROWS_NONE}[geom]
        self.pg = process_group
        dev = table.device
        nt = self.n + self.max_images
        self.n_total = nt
        F = features.shape[1]
        self.replicas = ops.default_replicas(nt, self.ld)
        self.X = torch.empty((self.max_images, F), device=dev, dtype=torch.float32)
        self.Y = torch.empty((self.max_images, self.D), device=dev, dtype=torch.float32)
        self.rows = torch.empty((nt, self.ld), device=dev, dtype=torch.float32)
        self.aux = torch.empty((nt, 4), device=dev, dtype=torch.float64)
        self.grad_rows = torch.empty((self.replicas, nt, self.ld), device=dev, dtype=torch.float32)
        self.grad_sum = torch.empty((nt, self.ld), device=dev, dtype=torch.float32)
        self.gY = torch.empty((self.max_images, self.D), device=dev, dtype=torch.float32)
        # flat gradient buffer [table | fc1.weight | fc1.bias]: one all-reduce, and the optimizer's .grad views
        sizes = [self.n * self.D, fc_weight.numel(), fc_bias.numel()]
        self.gflat = torch.zeros(sum(sizes), device=dev, dtype=torch.float32)
        o = np.cumsum([0] + sizes)
        self.g_table = self.gflat[o[0]:o[1]].view(self.n, self.D)
        self.g_w = self.gflat[o[1]:o[2]].view_as(fc_weight)
        self.g_b = self.gflat[o[2]:o[3]].view_as(fc_bias)
        self.E_pos = torch.empty(self.max_groups, device=dev, dtype=torch.float32)
        self.E_neg = torch.empty((self.max_groups, 2 * self.n_neg), device=dev, dtype=torch.float32)
        self.loss = torch.zeros(1, device=dev, dtype=torch.float64)
        self.params = [torch.nn.Parameter(t, requires_grad=False) for t in (table, fc_weight, fc_bias)]
        for p, g in zip(self.params, (self.g_table, self.g_w, self.g_b)):
            p.grad = g
        self.opt = torch.optim.Adam(self.params, lr=self.lr, fused=True)
        self.kernel_events = None
        self._sel_dev = torch.empty(self.max_images, device=dev, dtype=torch.int64)
        self._idx_bytes_dev = torch.empty(self.max_groups * (2 + 2 * self.n_neg) * 4, device=dev, dtype=torch.uint8)
        self.loss_host = torch.zeros(1, dtype=torch.float64).pin_memory()

    def _split(self, blk, B):
        Nn = self.n_neg
        return blk[:B], blk[B:2 * B], blk[2 * B:2 * B + B * Nn], blk[2 * B + B * Nn:2 * B + 2 * B * Nn]

    def step_device(self, img_sel, pos_from, pos_to, neg_to, neg_from):
        lib, st = N.lib(), N.stream_ptr(self.table.device)
        m, B, n, D, ld = int(img_sel.numel()), int(pos_from.numel()), self.n, self.D, self.ld
        if m > self.max_images or B > self.max_groups:
            raise N.LecError("step of %d images / %d positives exceeds the engine's buffers" % (m, B))
        X, Y = self.X[:m], self.Y[:m]
        torch.index_select(self.features, 0, img_sel, out=X)
        torch.addmm(self.fc_b, X, self.fc_w.t(), out=Y)
        geom = N.GEOM[self.geom]
        rows_img, aux_img = self.rows[n:n + m], self.aux[n:n + m]
        # labels (clears the loss accumulator), then the projected images; the gradient accumulator spans both
        N.check(lib.lec_rows_fwd(N._p(self.table), n, D, self.lab_mode, geom, self.K, N._p(self.rows), ld, N._p(self.aux),
                                 N._p(None), 0, N._p(self.loss), st), "lec_rows_fwd")
        self.grad_rows.zero_()
        N.check(lib.lec_rows_fwd(N._p(Y), m, D, self.img_mode, geom, self.K, N._p(rows_img), ld, N._p(aux_img),
                                 N._p(None), 0, N._p(None), st), "lec_rows_fwd")
        ev = self.kernel_events
        if ev is not None:
            ev[0].record()
        N.check(lib.lec_pairs_grouped(
            geom, self.precision, N._p(self.rows), N._p(self.aux), self.n_total, D, ld, N._p(pos_from), N._p(pos_to),
            N._p(neg_to), N._p(neg_from), pos_from.element_size(), B, self.n_neg, N._p(None), N._p(None), self.K,
            self.alpha, N._p(self.E_pos), N._p(self.E_neg), N._p(self.loss), N._p(self.grad_rows), self.replicas, st),
            "lec_pairs_grouped")
        if ev is not None:
            ev[1].record()
        N.check(lib.lec_reduce_replicas(N._p(self.grad_rows), self.replicas, self.n_total * ld, N._p(self.grad_sum), st),
                "lec_reduce_replicas")
        N.check(lib.lec_rows_bwd(N._p(self.table), N._p(self.grad_sum), 1, n, D, ld, self.lab_mode, self.K,
                                 N._p(self.g_table), 0, st), "lec_rows_bwd")
        gY = self.gY[:m]
        N.check(lib.lec_rows_bwd(N._p(Y), N._p(self.grad_sum[n:n + m]), 1, m, D, ld, self.img_mode, self.K, N._p(gY), 0, st),
                "lec_rows_bwd")
        torch.mm(gY.t(), X, out=self.g_w)
        torch.sum(gY, dim=0, out=self.g_b)
        if self.pg is not None and torch.distributed.get_world_size(self.pg) > 1:
            torch.distributed.all_reduce(self.gflat, group=self.pg)
        self.opt.step()
        return self.loss

    def step_host(self, img_sel_host, index_block, B):
        """Pinned host inputs (int64 image selection, uint16/int32 index block) -> step -> loss on the host."""
        m = int(img_sel_host.numel())
        n = B * (2 + 2 * self.n_neg)
        sel = self._sel_dev[:m]
        sel.copy_(img_sel_host, non_blocking=True)
        dst = self._idx_bytes_dev[:n * index_block.dtype.itemsize].view(index_block.dtype)
        dst.copy_(index_block[:n], non_blocking=True)
        self.step_device(sel, *self._split(dst, B))
        self.loss_host.copy_(self.loss, non_blocking=True)
        torch.cuda.current_stream(self.table.device).synchronize()
        return float(self.loss_host[0])
