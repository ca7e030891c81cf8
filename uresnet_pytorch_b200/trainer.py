"""Training / inference step of the sparse U-ResNet on one GPU per process.

Host-side mirror of the reference's trainer for the hot path only
(/root/reference/uresnet/trainval.py:21-30 backward = zero_grad -> backward -> Adam.step,
:94-100 forward = blob to device -> net; loss per uresnet_sparse.py:40-82), re-designed for
B200:

  * one process per GPU; gradients are combined with ONE NCCL all-reduce (SUM -- the
    reference's DataParallel reduce-adds replica gradients and never averages)
    over a single flat buffer, instead of a per-forward parameter broadcast + gather;
  * `graph=True` runs the whole step -- rulebook build, forward, loss, backward, all-reduce
    pack/unpack, Adam -- as one CUDA graph over capacity-sized buffers; the number of input
    points and of active sites per level live in device memory, so events of any size replay
    the same graph with no host synchronisation inside the step;
  * the blob crosses PCIe exactly as the reference ships it (float64 [N, D+2]) from pinned
    memory; labels as float32 [N, 1].
"""
import os

import numpy as np
import torch

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
    forwards (R/uresnet/trainval.py:23-25).  No-op without an initialised process group."""
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


class Trainer(object):
    def __init__(self, flags, lr=1e-3, graph=False, capacity=None, events_per_step=1, train=True, device=None,
                 math_mode=None, seed=0, overlap_wgrad=True, fused_loss=True):
        self.flags = flags
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.graph_mode = graph
        self.train = train
        self.events_per_step = events_per_step
        self.overlap_wgrad = overlap_wgrad
        self.fused_loss = fused_loss  # head + loss + their backward in three kernels (csrc/headloss.cu)
        self.math_mode = scn.config.math_mode if math_mode is None else math_mode
        torch.manual_seed(seed)
        self.net = UResNet(flags).to(self.device)
        self.net.train(train)
        self.criterion = SegmentationLoss(flags)
        self.criterion.num_events = events_per_step
        dist = _dist()
        if dist is not None and dist.get_world_size() > 1:
            for t in list(self.net.parameters()) + list(self.net.buffers()):
                dist.broadcast(t.data, 0)
        self.params = [p for p in self.net.parameters()]
        self.n_params = sum(p.numel() for p in self.params)
        self.opt = None
        if train:
            # one launch for all ~190 parameter tensors (csrc/optim.cu); B200SCN_TORCH_ADAM=1 keeps torch's
            if os.environ.get("B200SCN_TORCH_ADAM"):
                self.opt = torch.optim.Adam(self.params, lr=lr, capturable=graph, fused=True)
            else:
                from .optim import FusedAdam
                self.opt = FusedAdam(self.params, lr=lr)
        self.ncol = flags.DATA_DIM + 2
        self._graph = None
        self._cap = 0
        if graph:
            if capacity is None:
                raise ValueError("graph mode needs a row capacity")
            self._alloc_static(int(capacity))

    # ------------------------------------------------------------------------------------
    def _alloc_static(self, cap):
        dev = self.device
        self._cap = cap
        self.h_data = torch.zeros((cap, self.ncol), dtype=torch.float64).pin_memory()
        self.h_label = torch.zeros((cap, 1), dtype=torch.float32).pin_memory()
        self.h_n = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.d_data = torch.zeros((cap, self.ncol), dtype=torch.float64, device=dev)
        self.d_label = torch.zeros((cap, 1), dtype=torch.float32, device=dev)
        self.d_n = torch.zeros(1, dtype=torch.int32, device=dev)
        self.d_loss = torch.zeros(2, dtype=torch.float32, device=dev)
        self.h_loss = torch.zeros(2, dtype=torch.float32).pin_memory()
        self._graph = None

    def _allreduce_grads(self):
        allreduce_sum_([p.grad for p in self.params if p.grad is not None])

    def _fwd_bwd(self, data, label, valid=None):
        """One forward (+ backward + optimiser step when training) on device tensors."""
        with scn.config.math(self.math_mode):
            with torch.set_grad_enabled(self.train):
                logits = None
                if self.fused_loss:
                    ncol = data.shape[1]
                    feats = self.net.sparseModel((data[:, :ncol - 1], data[:, ncol - 1:].float()))
                    loss, acc = scn_functional.head_loss(
                        feats, self.net.linear, label, data[:, ncol - 2], max(self.events_per_step, 1),
                        rows=feats.shape[0], n_dev=self.d_n.data_ptr() if valid is not None else None,
                        unit_grad=True)
                else:
                    logits = self.net(data)[0]
                if self.fused_loss:
                    pass
                elif valid is not None:
                    lab = torch.where(valid, label.reshape(-1), torch.zeros_like(label.reshape(-1)))
                    w = valid.float()
                    loss, acc = self._masked_loss(logits, data, lab, w)
                else:
                    loss, acc = self.criterion([logits], [data], [label], None)
                if self.train:
                    self.opt.zero_grad(set_to_none=True)
                    prev = scn.config.overlap_wgrad
                    scn.config.overlap_wgrad = self.overlap_wgrad
                    try:
                        loss.backward()
                    finally:
                        scn.config.overlap_wgrad = prev
                        scn_functional.join_side_streams()
                    self._allreduce_grads()
                    self.opt.step()
                    scn_functional.repack_all()  # fragment-ordered weight copies follow the new weights
        return loss, acc, logits

    def _masked_loss(self, logits, data, lab, w):
        """uresnet_sparse.py:57-82 with a validity mask for the padded rows of the static buffers."""
        ids = torch.where(w > 0, data[:, -2], torch.zeros_like(data[:, -2])).long()
        ce = torch.nn.functional.cross_entropy(logits, lab.long(), reduction="none") * w
        nb = self.events_per_step
        cnt = torch.zeros(nb, device=ce.device).index_add_(0, ids, w)
        s = torch.zeros(nb, device=ce.device).index_add_(0, ids, ce)
        hit = (logits.argmax(-1) == lab.long()).float() * w
        h = torch.zeros(nb, device=ce.device).index_add_(0, ids, hit)
        safe = torch.where(cnt > 0, cnt, torch.ones_like(cnt))
        return (s / safe).sum(), (h / safe).sum()

    # ------------------------------------------------------------------------------------
    def _calibrate_tiles(self):
        """Row counts of the U-Net levels of the event in the staging buffer (eager rulebook build, one
        host read-back per level) -> tile rows of every level's grouped pair lists in the captured graph."""
        from . import _lib
        from .sparseconvnet.metadata import Metadata
        n = int(self.h_n[0])
        if n <= 0:
            return
        D = self.flags.DATA_DIM
        L = _lib.load()
        md = Metadata(D)
        size = (int(self.flags.SPATIAL_SIZE),) * D
        lv = md.set_input(self.d_data[:n, :D + 1], size)
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

    def _capture(self):
        """Warm up on a side stream, then capture the whole step into one CUDA graph."""
        valid = None
        self._calibrate_tiles()
        scn.config.md_plans.clear()  # the rule-prefetch plan is re-recorded by the first warm-up step below

        def step():
            nonlocal valid
            valid = torch.arange(self._cap, device=self.device, dtype=torch.int32) < self.d_n
            scn.config.static_input_rows = self.d_n
            try:
                with scn.static_mode(True):
                    loss, acc, _ = self._fwd_bwd(self.d_data, self.d_label, valid)
            finally:
                scn.config.static_input_rows = None
            self.d_loss[0] = loss.detach()
            self.d_loss[1] = acc.detach() if torch.is_tensor(acc) else float(acc)

        # snapshot state so that warm-up iterations do not perturb training
        state = {k: v.clone() for k, v in self.net.state_dict().items()}
        opt_state = None
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for _ in range(3):
                step()
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        if self.train:
            self.opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(g):
            step()
        self._graph = g
        # restore the pre-warm-up weights / BN stats and reset Adam
        self.net.load_state_dict(state)
        scn_functional.repack_all()
        if self.train:
            if hasattr(self.opt, "reset_state"):
                self.opt.reset_state()
            else:
                for st in self.opt.state.values():
                    for k, v in st.items():
                        if torch.is_tensor(v):
                            v.zero_()
        del opt_state
        torch.cuda.synchronize(self.device)

    # ------------------------------------------------------------------------------------
    def upload(self, data_np, label_np=None):
        """Pinned host staging + async H2D of one step's blob (float64 [N, D+2]) and labels."""
        n = int(data_np.shape[0])
        if not self.graph_mode:
            hd = torch.from_numpy(np.ascontiguousarray(data_np)).pin_memory()
            d = hd.to(self.device, non_blocking=True)
            l = None
            if label_np is not None:
                l = torch.from_numpy(np.ascontiguousarray(label_np, dtype=np.float32)).pin_memory().to(
                    self.device, non_blocking=True)
            return d, l
        if n > self._cap:
            raise ValueError("step has %d rows but the captured capacity is %d" % (n, self._cap))
        self.h_data[:n].numpy()[...] = data_np
        if label_np is not None:
            self.h_label[:n].numpy()[...] = label_np
        self.h_n[0] = n
        self.d_data[:n].copy_(self.h_data[:n], non_blocking=True)
        if label_np is not None:
            self.d_label[:n].copy_(self.h_label[:n], non_blocking=True)
        self.d_n.copy_(self.h_n, non_blocking=True)
        return self.d_data, self.d_label

    def step_resident(self, data=None, label=None):
        """Run one step on inputs already resident in HBM.  No host synchronisation."""
        if self.graph_mode:
            if self._graph is None:
                self._capture()
            self._graph.replay()
            return self.d_loss
        loss, acc, _ = self._fwd_bwd(data, label)
        return torch.stack([loss.detach(), torch.as_tensor(acc, device=loss.device, dtype=loss.dtype)])

    def train_step(self, data_np, label_np):
        """Public end-to-end call: host blob in, (loss, accuracy) floats out."""
        if self.graph_mode and self._graph is None:
            self.upload(data_np, label_np)
            self._capture()
        d, l = self.upload(data_np, label_np)
        out = self.step_resident(d, l)
        if self.graph_mode:
            self.h_loss.copy_(out, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
            vals = self.h_loss.tolist()
        else:
            vals = out.tolist()
        return {"loss": vals[0], "acc": vals[1]}

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
        scn_functional.repack_all()  # fragment-ordered weight copies follow the restored weights
        torch.cuda.synchronize(self.device)
        return int(ck["global_step"]) + 1

    @torch.no_grad()
    def infer(self, data_np):
        """Forward only: host blob in, logits [N, nc] on the device out (eager regime)."""
        was = self.net.training
        self.net.eval()
        try:
            d = torch.from_numpy(np.ascontiguousarray(data_np)).to(self.device)
            with scn.config.math(self.math_mode):
                return self.net(d)[0]
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
        dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo", rank=rank, world_size=world)
    return rank, local, world
