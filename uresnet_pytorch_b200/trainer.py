"""Training / inference step of the sparse U-ResNet on one GPU per process.

Host-side mirror of the reference's trainer for the hot path only
(/root/reference/uresnet/trainval.py:21-30 backward = zero_grad -> backward -> Adam.step,
:59-69 several forwards accumulated before one backward, :94-100 forward = blob to device -> net,
:127-130 inference outputs; loss per uresnet_sparse.py:40-82), re-designed for B200:

  * one process per GPU.  Every gradient kernel writes into ONE flat fp32 buffer (functional.GradSink; the
    parameters' .grad are views of it), so the data-parallel exchange is NCCL all-reduce(SUM) straight over that
    buffer -- SUM, not mean: the reference's DataParallel reduce-adds replica gradients and never averages
     -- with no flatten / unflatten passes.  The inner U-Net levels hold ~97 % of the gradient
    bytes and finish their backward first; their slice is all-reduced on a side stream while the outer level's
    backward still runs, only the small remainder is exchanged at the end;
  * `graph=True` runs the whole step -- rulebook build, forward, loss, backward, all-reduce, Adam, weight
    re-packing -- as one CUDA graph over capacity-sized buffers; the number of input points and of active sites
    per level live in device memory, so events of any size replay the same graph with no host synchronisation
    inside the step;
  * inputs: the reference's wire format (float64 blob [N, D+2], float32 labels; iotools_sparse.py:138-141,160) or
    the compact one (int32 coords [N, D+1] + float32 values [N]: 40 % of the bytes, no float64 anywhere), staged
    through two pinned slots and uploaded on a copy stream, so the host-side staging and the H2D copy of step i+1
    run under step i (`submit` / `result`; `train_step` is the synchronous pair);
  * bad input is reported, not trained on: out-of-range coordinates, sample ids >= events_per_step and labels
    outside [0, NUM_CLASS) are counted on the device and raise from `result` / `train_step`.
"""
import os

import numpy as np
import torch

from . import inference as _inference
from . import sparseconvnet as scn
from .sparseconvnet import functional as scn_functional
from .model import SegmentationLoss, UResNet


def _dist():
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized()) else None


def allreduce_sum_(tensors):
    """In-place SUM all-reduce of a list of dense tensors through ONE flat buffer (one collective
    per step).  SUM, not mean: the reference's DataParallel reduce-adds replica gradients
    (R/uresnet/ops.py:8 -> Broadcast.backward) and divides only by the number of accumulated
    forwards (R/uresnet/trainval.py:23-25).  No-op without an initialised process group.
    (Generic helper for callers without a GradSink; the Trainer all-reduces its flat buffer in place.)"""
    dist = _dist()
    if dist is None or dist.get_world_size() == 1 or not tensors:
        return
    flat = torch._utils._flatten_dense_tensors(tensors)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    torch._foreach_copy_(tensors, list(torch._utils._unflatten_dense_tensors(flat, tensors)))


def shard_events(n_events, world, rank):
    """Contiguous block of events owned by `rank`, as the reference's IO assigns them:
    new_id = data_id // batch_per_gpu (R/uresnet/iotools/iotools_sparse.py:140)."""
    if n_events % world != 0:
        raise ValueError("BATCH_SIZE must be a multiple of the number of GPUs (R/uresnet/flags.py:176-183)")
    per = n_events // world
    return list(range(rank * per, (rank + 1) * per))


class BadInputError(ValueError):
    """A step saw out-of-range coordinates, sample ids or labels (see Trainer docstring)."""


class Trainer(object):
    N_SLOTS = 2  # pinned staging slots (step i+1 is staged while step i runs)

    def __init__(self, flags, lr=1e-3, graph=False, capacity=None, events_per_step=1, train=True, device=None,
                 math_mode=None, seed=0, overlap_wgrad=True, fused_loss=True, input_format="blob64",
                 overlap_allreduce=True, distributed=None):
        if input_format not in ("blob64", "compact"):
            raise ValueError("input_format must be 'blob64' (reference wire format) or 'compact'")
        self.flags = flags
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.graph_mode = graph
        self.train = train
        self.events_per_step = max(int(events_per_step), 1)
        self.overlap_wgrad = overlap_wgrad
        self.overlap_allreduce = overlap_allreduce and not os.environ.get("B200SCN_NO_OVERLAP_ALLREDUCE")  # (A/B runs)
        self.fused_loss = fused_loss  # head + loss + their backward in three kernels (csrc/headloss.cu)
        self.input_format = input_format
        self.math_mode = scn.config.math_mode if math_mode is None else math_mode
        torch.manual_seed(seed)
        self.net = UResNet(flags).to(self.device)
        self.net.train(train)
        self.criterion = SegmentationLoss(flags)
        self.criterion.num_events = self.events_per_step
        dist = _dist()
        # distributed=False keeps a trainer local even inside an initialised process group (tests, evaluation)
        self.world = dist.get_world_size() if (dist is not None and distributed is not False) else 1
        if self.world > 1:
            for t in list(self.net.parameters()) + list(self.net.buffers()):
                dist.broadcast(t.data, 0)
        self.params = [p for p in self.net.parameters()]
        self.n_params = sum(p.numel() for p in self.params)
        self.opt = self.sink = None
        self._n_inner = 0
        if train:
            # flat gradient buffer, inner U-Net levels first: that prefix is complete (and all-reduced) early
            inner = [m for m in self.net.modules() if getattr(m, "_mark_tag", None) == "unet.inner0"]
            first = [p for p in inner[0].parameters()] if inner else []
            ids = set(id(p) for p in first)
            order = first + [p for p in self.params if id(p) not in ids]
            self.sink = scn_functional.GradSink(order)
            self._n_inner = (self.sink.offsets[len(first)] if len(first) < len(order) else self.sink.numel) if first else 0
            # one launch for all ~140 parameter tensors (csrc/optim.cu); B200SCN_TORCH_ADAM=1 keeps torch's
            if os.environ.get("B200SCN_TORCH_ADAM"):
                self.opt = torch.optim.Adam(self.params, lr=lr, capturable=graph, fused=True)
            else:
                from .optim import FusedAdam
                self.opt = FusedAdam(self.params, lr=lr)
        self.ncol = flags.DATA_DIM + 2
        self._graph = self._graph_micro = self._graph_opt = self._graph_pred = None
        self._pack_jobs = None
        self._cap = 0
        self._comm = torch.cuda.Stream(device=self.device)
        self._copy = torch.cuda.Stream(device=self.device)
        self._comm_pending = False
        self._slot = 0
        self._tickets = {}
        self._next_ticket = 0
        self._pred_out = {}
        if graph:
            if capacity is None:
                raise ValueError("graph mode needs a row capacity")
            self._alloc_static(int(capacity))

    # ------------------------------------------------------------------------------------
    def _alloc_static(self, cap):
        dev, D = self.device, self.flags.DATA_DIM
        self._cap = cap
        pin = lambda *shape, dtype: torch.zeros(shape, dtype=dtype).pin_memory()
        if self.input_format == "blob64":
            self.h_data = [pin(cap, self.ncol, dtype=torch.float64) for _ in range(self.N_SLOTS)]
            self.s_data = [torch.zeros((cap, self.ncol), dtype=torch.float64, device=dev) for _ in range(self.N_SLOTS)]
            self.d_data = torch.zeros((cap, self.ncol), dtype=torch.float64, device=dev)
        else:
            self.h_coords = [pin(cap, D + 1, dtype=torch.int32) for _ in range(self.N_SLOTS)]
            self.h_value = [pin(cap, 1, dtype=torch.float32) for _ in range(self.N_SLOTS)]
            self.s_coords = [torch.zeros((cap, D + 1), dtype=torch.int32, device=dev) for _ in range(self.N_SLOTS)]
            self.s_value = [torch.zeros((cap, 1), dtype=torch.float32, device=dev) for _ in range(self.N_SLOTS)]
            self.d_coords = torch.zeros((cap, D + 1), dtype=torch.int32, device=dev)
            self.d_value = torch.zeros((cap, 1), dtype=torch.float32, device=dev)
        self.h_label = [pin(cap, 1, dtype=torch.float32) for _ in range(self.N_SLOTS)]
        self.s_label = [torch.zeros((cap, 1), dtype=torch.float32, device=dev) for _ in range(self.N_SLOTS)]
        self.d_label = torch.zeros((cap, 1), dtype=torch.float32, device=dev)
        self.h_n = [pin(1, dtype=torch.int32) for _ in range(self.N_SLOTS)]
        self.s_n = [torch.zeros(1, dtype=torch.int32, device=dev) for _ in range(self.N_SLOTS)]
        self.d_n = torch.zeros(1, dtype=torch.int32, device=dev)
        # loss, accuracy, rows with a bad sample id, rows with a bad label, out-of-range coordinate flag, spare
        self.d_loss = torch.zeros(6, dtype=torch.float32, device=dev)
        self.h_loss = [pin(6, dtype=torch.float32) for _ in range(self.N_SLOTS)]
        self._graph = None

    # -- inputs ----------------------------------------------------------------------------
    def _split(self, data):
        """(coords [N, D+1], value [N, 1] float32, ids column) views of a device-resident input."""
        if isinstance(data, (tuple, list)):
            coords, value = data
            return coords, value.reshape(-1, 1).float(), coords[:, -1]
        ncol = data.shape[1]
        return data[:, :ncol - 1], data[:, ncol - 1:].float(), data[:, ncol - 2]

    def _static_input(self):
        return (self.d_coords, self.d_value) if self.input_format == "compact" else self.d_data

    # -- one forward (+ backward) ------------------------------------------------------------
    def _on_backward_mark(self, tag):
        """The inner U-Net levels finished their backward: all-reduce their gradient slice now, on the comm
        stream, under the rest of the backward pass (capturable: forked from / joined to the main stream)."""
        if tag != "unet.inner0" or self.world == 1 or not self._allreduce_armed or self._n_inner == 0:
            return
        dist = _dist()
        main = torch.cuda.current_stream(self.device)
        self._comm.wait_stream(main)                                   # BatchNorm gradients (main stream)
        self._comm.wait_stream(scn_functional._side_stream(self.device))  # weight gradients (side stream)
        with torch.cuda.stream(self._comm):
            dist.all_reduce(self.sink.flat[:self._n_inner], op=dist.ReduceOp.SUM)
        self._comm_pending = True
        self._allreduce_armed = False

    def _allreduce_grads(self):
        if self.world == 1:
            return
        dist = _dist()
        main = torch.cuda.current_stream(self.device)
        lo = self._n_inner if self._comm_pending else 0
        if lo < self.sink.numel:
            dist.all_reduce(self.sink.flat[lo:self.sink.numel], op=dist.ReduceOp.SUM)
        if self._comm_pending:
            main.wait_stream(self._comm)
            self._comm_pending = False

    def _fwd_bwd(self, data, label, n_dev=None, zero=True, step=True, grad_scale=None):
        """One forward (+ backward, + all-reduce and optimiser step when `step`) on device tensors.
        Returns a float32 tensor [5]: loss, accuracy, bad sample ids, bad labels, bad coordinates."""
        coords, value, ids = self._split(data)
        nb = self.events_per_step
        with scn.config.math(self.math_mode):
            with torch.set_grad_enabled(self.train):
                if self.train and zero:
                    self.sink.zero()
                feats = self.net.sparseModel((coords, value))
                if self.fused_loss:
                    out4 = scn_functional.head_loss(feats, self.net.linear, label, ids, nb, rows=feats.shape[0],
                                                    n_dev=n_dev, unit_grad=True, full=True)
                    loss = out4[0]
                    stats = out4.detach()
                else:
                    logits = self.net.linear(feats)
                    loss, stats = self._torch_loss(logits, ids, label, n_dev)
                flag = self.net.sparseModel[0].last_error_flag
                flag = flag.float() if flag is not None else torch.zeros(1, device=self.device)
                stats = torch.cat([stats, flag])
                if self.train:
                    prev = scn.config.overlap_wgrad, scn.config.backward_mark_callback
                    scn.config.overlap_wgrad = self.overlap_wgrad
                    scn.config.backward_mark_callback = self._on_backward_mark
                    self._allreduce_armed = step and grad_scale is None and self.overlap_allreduce
                    try:
                        with self.sink.active():
                            loss.backward()
                    finally:
                        scn.config.overlap_wgrad, scn.config.backward_mark_callback = prev
                        scn_functional.join_side_streams()
                    if step:
                        self._opt_step(grad_scale)
        return stats

    def _opt_step(self, grad_scale=None):
        if grad_scale is not None:   # mean over the accumulated forwards (R/uresnet/trainval.py:23-25)
            self.sink.flat.mul_(grad_scale)
        self._allreduce_grads()
        self.opt.step()
        # fragment-ordered weight copies follow the new weights
        self._pack_jobs = scn_functional.repack_params(self.params, self._pack_jobs)

    def _torch_loss(self, logits, ids, label, n_dev):
        """uresnet_sparse.py:57-82 in torch ops (the unfused reference path of the fused kernel), with a validity
        mask for the padded rows of the static buffers; same [4] result as functional.head_loss(full=True)."""
        nb, nc = self.events_per_step, logits.shape[1]
        n = logits.shape[0]
        w = torch.ones(n, device=logits.device)
        if n_dev is not None:
            w = (torch.arange(n, device=logits.device, dtype=torch.int32) < self.d_n).float()
        ids_l, lab = ids.long(), label.reshape(-1).long()
        ok_id, ok_lab = (ids_l >= 0) & (ids_l < nb), (lab >= 0) & (lab < nc)
        bad_id, bad_lab = ((~ok_id).float() * w).sum(), ((~ok_lab).float() * w).sum()
        ids_l = torch.where(ok_id, ids_l, torch.zeros_like(ids_l))
        w_id = w * ok_id.float()
        w_ce = w_id * ok_lab.float()
        lab_c = torch.where(ok_lab, lab, torch.zeros_like(lab))
        ce = torch.nn.functional.cross_entropy(logits, lab_c, reduction="none") * w_ce
        cnt = torch.zeros(nb, device=ce.device).index_add_(0, ids_l, w_id)
        s = torch.zeros(nb, device=ce.device).index_add_(0, ids_l, ce)
        hit = (logits.argmax(-1) == lab).float() * w_id
        h = torch.zeros(nb, device=ce.device).index_add_(0, ids_l, hit)
        safe = torch.where(cnt > 0, cnt, torch.ones_like(cnt))
        loss = (s / safe).sum()
        return loss, torch.stack([loss.detach(), (h / safe).sum(), bad_id, bad_lab])

    # ------------------------------------------------------------------------------------
    def _calibrate_tiles(self, n):
        """Row counts of the U-Net levels of the event in the static buffers (eager rulebook build, one
        host read-back per level) -> tile rows of every level's grouped pair lists in the captured graph."""
        from . import _lib
        from .sparseconvnet.metadata import Metadata
        if n <= 0:
            return
        D = self.flags.DATA_DIM
        L = _lib.load()
        md = Metadata(D)
        size = (int(self.flags.SPATIAL_SIZE),) * D
        coords = self._split(self._static_input())[0]
        lv = md.set_input(coords[:n, :D + 1], size)
        hints = {size: L.b200scn_gpl_tile_rows(max(lv.n, 1))}
        rows = {size: lv.n}
        for _ in range(int(self.flags.URESNET_NUM_STRIDES) - 1):
            out = tuple((v - 2) // 2 + 1 for v in size)
            md.down_rules(size, out, 2)
            hints[out] = L.b200scn_gpl_tile_rows(max(md.level(out).n, 1))
            rows[out] = md.level(out).n
            size = out
        scn.config.gpl_tile_rows_by_size = hints
        scn.config.gpl_rows_by_size = rows

    def _capture_fn(self, fn, warm=3):
        """Warm `fn` up on a side stream, restore the model / optimiser state it touched, capture it."""
        state = {k: v.clone() for k, v in self.net.state_dict().items()}
        opt_snap = None
        if self.train and self.opt is not None:
            opt_snap = self.opt.snapshot() if hasattr(self.opt, "snapshot") else __import__("copy").deepcopy(
                self.opt.state_dict())
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for _ in range(warm):
                fn()
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        # put back the pre-warm-up weights / BN statistics / optimiser state
        self.net.load_state_dict(state)
        if opt_snap is not None:
            self._pack_jobs = scn_functional.repack_params(self.params, self._pack_jobs)
            if hasattr(self.opt, "restore"):
                self.opt.restore(opt_snap)
            else:
                self.opt.load_state_dict(opt_snap)
        torch.cuda.synchronize(self.device)
        return g

    def _static_step(self, zero=True, step=True, grad_scale=None):
        scn.config.static_input_rows = self.d_n
        try:
            with scn.static_mode(True):
                stats = self._fwd_bwd(self._static_input(), self.d_label, n_dev=self.d_n.data_ptr(), zero=zero,
                                      step=step, grad_scale=grad_scale)
        finally:
            scn.config.static_input_rows = None
        self.d_loss[:5] = stats

    def _capture(self):
        """Capture the whole step into one CUDA graph (calibrated on the event in the static buffers)."""
        torch.cuda.synchronize(self.device)
        self._calibrate_tiles(int(self.d_n.item()))
        scn.config.md_plans.clear()  # the rule-prefetch plan is re-recorded by the first warm-up step below
        self._graph = self._capture_fn(lambda: self._static_step())

    # -- staging ---------------------------------------------------------------------------
    def _stage(self, data_np, label_np):
        """Host copy into the next pinned slot + H2D on the copy stream; returns (slot, event, n)."""
        k = self._slot
        self._slot = (k + 1) % self.N_SLOTS
        if self.input_format == "compact":
            coords_np, value_np = data_np
            n = int(coords_np.shape[0])
        else:
            n = int(data_np.shape[0])
        if n > self._cap:
            raise ValueError("step has %d rows but the captured capacity is %d" % (n, self._cap))
        if self.input_format == "compact":
            self.h_coords[k][:n].numpy()[...] = coords_np
            self.h_value[k][:n].numpy()[...] = np.asarray(value_np).reshape(-1, 1)
        else:
            self.h_data[k][:n].numpy()[...] = data_np
        if label_np is not None:
            self.h_label[k][:n].numpy()[...] = np.asarray(label_np).reshape(-1, 1)
        self.h_n[k][0] = n
        with torch.cuda.stream(self._copy):
            if self.input_format == "compact":
                self.s_coords[k][:n].copy_(self.h_coords[k][:n], non_blocking=True)
                self.s_value[k][:n].copy_(self.h_value[k][:n], non_blocking=True)
            else:
                self.s_data[k][:n].copy_(self.h_data[k][:n], non_blocking=True)
            if label_np is not None:
                self.s_label[k][:n].copy_(self.h_label[k][:n], non_blocking=True)
            self.s_n[k].copy_(self.h_n[k], non_blocking=True)
            ev = self._copy.record_event()
        return k, ev, n

    def _consume(self, k, ev, n):
        """Main stream: wait for slot k's upload, move it into the graph's static buffers."""
        main = torch.cuda.current_stream(self.device)
        main.wait_event(ev)
        if self.input_format == "compact":
            self.d_coords[:n].copy_(self.s_coords[k][:n], non_blocking=True)
            self.d_value[:n].copy_(self.s_value[k][:n], non_blocking=True)
        else:
            self.d_data[:n].copy_(self.s_data[k][:n], non_blocking=True)
        self.d_label[:n].copy_(self.s_label[k][:n], non_blocking=True)
        self.d_n.copy_(self.s_n[k], non_blocking=True)

    def upload(self, data_np, label_np=None):
        """Synchronous-order upload of one step's input: returns device tensors (eager) or fills the static
        buffers of the captured graph (graph mode)."""
        if not self.graph_mode:
            def dev(a, dtype=None):
                t = torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).pin_memory()
                return t.to(self.device, non_blocking=True)
            if self.input_format == "compact" or isinstance(data_np, (tuple, list)):
                d = (dev(data_np[0], np.int32), dev(np.asarray(data_np[1]).reshape(-1, 1), np.float32))
            else:
                d = dev(data_np)
            l = dev(np.asarray(label_np).reshape(-1, 1), np.float32) if label_np is not None else None
            return d, l
        k, ev, n = self._stage(data_np, label_np)
        self._consume(k, ev, n)
        return self._static_input(), self.d_label

    # -- public step API ---------------------------------------------------------------------
    def step_resident(self, data=None, label=None):
        """Run one training step on inputs already resident in HBM.  No host synchronisation.  Returns the device
        tensor (loss, acc, bad ids, bad labels, bad coords[, spare])."""
        if self.graph_mode:
            if self._graph is None:
                self._capture()
            self._graph.replay()
            return self.d_loss
        return self._fwd_bwd(data, label)

    def submit(self, data_np, label_np):
        """Stage + upload + launch one training step without waiting for it; `result(ticket)` returns its loss.
        Submitting step i+1 before asking for step i's result overlaps its staging and H2D copy with step i."""
        t = self._next_ticket
        self._next_ticket += 1
        if not self.graph_mode:
            d, l = self.upload(data_np, label_np)
            self._tickets[t] = (self.step_resident(d, l), None)
            return t
        if self._graph is None:
            self.upload(data_np, label_np)
            self._capture()
        k, ev, n = self._stage(data_np, label_np)
        self._consume(k, ev, n)
        self._graph.replay()
        self.h_loss[k].copy_(self.d_loss, non_blocking=True)
        self._tickets[t] = (self.h_loss[k], torch.cuda.current_stream(self.device).record_event())
        return t

    def result(self, ticket):
        out, ev = self._tickets.pop(ticket)
        if ev is not None:
            ev.synchronize()
        vals = out.tolist()
        self._check(vals)
        return {"loss": vals[0], "acc": vals[1]}

    @staticmethod
    def _check(vals):
        if vals[4] != 0:
            raise BadInputError("coordinates / sample ids must be integers in [0, 65535)")
        if vals[2] != 0:
            raise BadInputError("%d rows carry a sample id outside [0, events_per_step)" % int(vals[2]))
        if vals[3] != 0:
            raise BadInputError("%d rows carry a label outside [0, NUM_CLASS)" % int(vals[3]))

    def train_step(self, data_np, label_np):
        """Public end-to-end call: host blob in, (loss, accuracy) floats out."""
        return self.result(self.submit(data_np, label_np))

    def train_step_accum(self, blobs):
        """Several forwards, ONE optimiser step on the mean of their losses -- the reference's
        BATCH_SIZE > MINIBATCH_SIZE x GPUs regime (R/uresnet/trainval.py:21-30,59-69).  `blobs` = list of
        (data, label) host arrays.  Returns the summed loss / accuracy of the forwards."""
        n_acc = len(blobs)
        if n_acc == 0:
            raise ValueError("train_step_accum: no input")
        tot = [0.0, 0.0]
        for i, (data_np, label_np) in enumerate(blobs):
            last = i == n_acc - 1
            if self.graph_mode:
                if self._graph is None:
                    self.upload(data_np, label_np)
                    self._capture()
                if self._graph_micro is None:
                    self.upload(data_np, label_np)
                    torch.cuda.synchronize(self.device)
                    self._graph_micro = self._capture_fn(lambda: self._static_step(zero=False, step=False))
                self.upload(data_np, label_np)
                if i == 0:
                    self.sink.zero()
                self._graph_micro.replay()
                if last:
                    self._opt_step(1.0 / n_acc)
                vals = self.d_loss.tolist()
            else:
                d, l = self.upload(data_np, label_np)
                vals = self._fwd_bwd(d, l, zero=(i == 0), step=last, grad_scale=1.0 / n_acc).tolist()
            self._check(vals)
            tot[0] += vals[0]
            tot[1] += vals[1]
        return {"loss": tot[0], "acc": tot[1]}

    # -- inference -------------------------------------------------------
    @torch.no_grad()
    def _predict_dev(self, data, label, n_dev=None, want_logits=False):
        coords, value, ids = self._split(data)
        with scn.config.math(self.math_mode):
            feats = self.net.sparseModel((coords, value))
        lin = self.net.linear
        return _inference.head_predict(feats, lin.weight, lin.bias, ids=ids, label=label, energy=value,
                                       nb=self.events_per_step, rows=feats.shape[0], n_dev=n_dev,
                                       want_logits=want_logits, out=self._pred_out if n_dev is not None else None)

    def predict_resident(self):
        """Graph mode: forward + softmax / argmax / metric sums on the static buffers.  No host sync."""
        if self._graph_pred is None:
            was = self.net.training
            self.net.eval()
            try:
                torch.cuda.synchronize(self.device)
                self._calibrate_tiles(int(self.d_n.item()))
                scn.config.md_plans.clear()

                def fn():
                    scn.config.static_input_rows = self.d_n
                    try:
                        with scn.static_mode(True):
                            self._predict_dev(self._static_input(), self.d_label, n_dev=self.d_n.data_ptr())
                    finally:
                        scn.config.static_input_rows = None
                self._graph_pred = self._capture_fn(fn)
            finally:
                self.net.train(was)
        self._graph_pred.replay()
        return self._pred_out

    def predict_step(self, data_np, label_np=None, to_host=True):
        """Inference end to end (R/uresnet/trainval.py:92-130 with TRAIN=False): host blob in ->
        dict(softmax [N, nc], pred [N], and, with labels, `metrics` = the reference's per-event accuracy /
        confusion block (inference.metrics_from_sums))."""
        nc = self.flags.NUM_CLASS
        if self.graph_mode:
            if label_np is None:
                label_np = np.full((len(data_np[0]) if self.input_format == "compact" else len(data_np), 1), -1.0,
                                   np.float32)
            k, ev, n = self._stage(data_np, label_np)
            self._consume(k, ev, n)
            out = self.predict_resident()
            res = {key: out[key][:n] for key in ("softmax", "pred") if key in out}
            sums = out.get("sums")
        else:
            was = self.net.training
            self.net.eval()
            try:
                d, l = self.upload(data_np, label_np)
                out = self._predict_dev(d, l)
            finally:
                self.net.train(was)
            res = {key: out[key] for key in ("softmax", "pred") if key in out}
            sums = out.get("sums")
        if to_host:
            res = {key: v.cpu().numpy() for key, v in res.items()}
            if sums is not None:
                res["metrics"] = _inference.metrics_from_sums(sums, nc)
        elif sums is not None:
            res["sums"] = sums
        return res

    # ------------------------------------------------------------------------------------
    def save_state(self, path, iteration):
        """The reference's checkpoint (R/uresnet/trainval.py:32-40): {'global_step', 'state_dict', 'optimizer'};
        state-dict keys carry the 'module.' prefix of the DataParallel wrapper the reference saves through
        (R/uresnet/trainval.py:159-163), so its tools read these files and vice versa."""
        torch.cuda.synchronize(self.device)
        sd = {"module." + k: v.detach().clone() for k, v in self.net.state_dict().items()}
        torch.save({"global_step": int(iteration), "state_dict": sd,
                    "optimizer": self.opt.state_dict() if self.opt is not None else None}, path)

    def load_state(self, path, lr=None):
        """R/uresnet/trainval.py:171-196: weights (strict=False), optimizer state when training, learning rate
        reset to the current flag; returns the next iteration.  Works on a captured graph too: everything is
        copied in place."""
        ck = torch.load(path, map_location=self.device, weights_only=False)
        sd = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in ck["state_dict"].items()}
        own = self.net.state_dict()
        with torch.no_grad():
            for k, v in sd.items():
                if k in own:
                    if v.dim() == own[k].dim() + 1 and v.shape[1] == 1:  # newer scn: [K, 1, nIn, nOut]
                        v = v.squeeze(1)
                    own[k].copy_(v)
        if self.train and self.opt is not None and ck.get("optimizer") is not None:
            keep = self.opt.lr if lr is None else lr
            self.opt.load_state_dict(ck["optimizer"])
            if hasattr(self.opt, "param_groups"):
                for g in self.opt.param_groups:
                    g["lr"] = keep
            else:
                self.opt.lr = float(keep)
        # fragment-ordered weight copies follow the restored weights
        self._pack_jobs = scn_functional.repack_params(self.params, self._pack_jobs)
        torch.cuda.synchronize(self.device)
        return int(ck["global_step"]) + 1

    @torch.no_grad()
    def infer(self, data_np):
        """Forward only: host blob in, logits [N, nc] on the device out (eager regime)."""
        was = self.net.training
        self.net.eval()
        try:
            d, _ = (self.upload(data_np) if not self.graph_mode else
                    (torch.from_numpy(np.ascontiguousarray(data_np)).to(self.device), None))
            coords, value, _ = self._split(d)
            with scn.config.math(self.math_mode):
                return self.net.linear(self.net.sparseModel((coords, value)))
        finally:
            self.net.train(was)


def init_distributed():
    """torchrun-style rendezvous (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the env)."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        # The gradient all-reduce runs next to the outer level's backward kernels, which fill the SMs: with NCCL's
        # default channel count its kernel gets few CTAs scheduled and outlives the backward pass.  Measured on 2
        # GPUs (profiles/README.md section 0.1): 4.39 ms per step by default, 4.29 with NCCL_MIN_CTAS=32, 4.51 / 4.64
        # with NCCL_MAX_CTAS=4 / 2.
        os.environ.setdefault("NCCL_MIN_CTAS", "32")
        dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo", rank=rank, world_size=world)
    return rank, local, world
