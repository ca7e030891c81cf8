"""Multi-tensor Adam over the C-ABI (csrc/optim.cu): torch.optim.Adam's default update
(/root/reference/uresnet/trainval.py:159-167 builds `torch.optim.Adam(net.parameters(), lr=...)`) for every
parameter in ONE kernel launch.  The step counter lives on the device, so the launch can be captured in a
CUDA graph; gradient pointers are re-read on every call (autograd allocates fresh gradient tensors), through
a pinned staging buffer so that the refresh is a capturable async copy.  One step counter for all tensors
(torch keeps one per tensor; they agree whenever every parameter receives a gradient every step, as here)."""
import torch

from . import _lib


class FusedAdam(object):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        self.params = [p for p in params]
        if not self.params:
            raise ValueError("FusedAdam: no parameters")
        dev = self.params[0].device
        if any((not p.is_cuda) or p.dtype != torch.float32 or not p.is_contiguous() for p in self.params):
            raise RuntimeError("FusedAdam: parameters must be contiguous float32 CUDA tensors (there is no CPU path)")
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.exp_avg = [torch.zeros_like(p) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p) for p in self.params]
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.ticket = torch.zeros(1, dtype=torch.int32, device=dev)
        n = len(self.params)
        self._host = torch.zeros((n, 5), dtype=torch.int64).pin_memory()
        self._jobs = torch.zeros((n, 5), dtype=torch.int64, device=dev)
        self._max = max(p.numel() for p in self.params)
        for i, p in enumerate(self.params):
            self._host[i, 0] = p.data_ptr()
            self._host[i, 2] = self.exp_avg[i].data_ptr()
            self._host[i, 3] = self.exp_avg_sq[i].data_ptr()
            self._host[i, 4] = p.numel()

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def reset_state(self):
        """Forget the moments and the step count (used after warm-up iterations before graph capture)."""
        for t in self.exp_avg + self.exp_avg_sq:
            t.zero_()
        self.step_dev.zero_()

    def snapshot(self):
        """Copy of the moments and the step count (graph capture warms up on real steps, then puts this back)."""
        return ([t.clone() for t in self.exp_avg], [t.clone() for t in self.exp_avg_sq], self.step_dev.clone())

    @torch.no_grad()
    def restore(self, snap):
        for dst, src in zip(self.exp_avg + self.exp_avg_sq, snap[0] + snap[1]):
            dst.copy_(src)
        self.step_dev.copy_(snap[2])

    @torch.no_grad()
    def step(self):
        live = [i for i, p in enumerate(self.params) if p.grad is not None]
        if not live:
            return
        for i in live:
            g = self.params[i].grad
            if g.dtype != torch.float32 or not g.is_contiguous():
                self.params[i].grad = g = g.float().contiguous()
            self._host[i, 1] = g.data_ptr()
        if len(live) == len(self.params):
            host, jobs, n = self._host, self._jobs, len(self.params)
        else:  # parameters without a gradient are skipped, as torch.optim does
            host = self._host[live].contiguous().pin_memory()
            jobs, n = self._jobs[:len(live)], len(live)
        jobs.copy_(host, non_blocking=True)
        self._keep = host
        _lib.call("b200scn_adam_step", jobs.data_ptr(), n, self._max, self.step_dev.data_ptr(), self.ticket.data_ptr(),
                  self.lr, self.betas[0], self.betas[1], self.eps, torch.cuda.current_stream().cuda_stream)
        # the kernel wrote through raw pointers: tell autograd (and the packed-weight caches keyed on the version
        # counter, functional.packed_weights) that the parameters changed
        torch.autograd.graph.increment_version([self.params[i] for i in live])

    # ------------------------------------------------------------------------------------
    def state_dict(self):
        """torch.optim.Adam's state-dict layout (what the reference stores under 'optimizer',
        /root/reference/uresnet/trainval.py:32-40), so checkpoints move both ways."""
        step = self.step_dev.to(torch.float32).reshape(()).cpu()
        state = {i: {"step": step.clone(), "exp_avg": self.exp_avg[i].detach().clone(),
                     "exp_avg_sq": self.exp_avg_sq[i].detach().clone()} for i in range(len(self.params))}
        if int(step.item()) == 0:
            state = {}
        group = {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": 0, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group]}

    @torch.no_grad()
    def load_state_dict(self, sd):
        groups = sd.get("param_groups", [])
        order = [i for g in groups for i in g["params"]]
        if len(order) != len(self.params):
            raise ValueError("FusedAdam.load_state_dict: %d parameters in the checkpoint, %d here"
                             % (len(order), len(self.params)))
        if groups:
            g = groups[0]
            if g.get("weight_decay", 0) != 0 or g.get("amsgrad", False):
                raise ValueError("FusedAdam: weight_decay / amsgrad checkpoints are not supported")
            self.lr, self.eps = float(g["lr"]), float(g["eps"])
            self.betas = (float(g["betas"][0]), float(g["betas"][1]))
        steps = set()
        for pos, key in enumerate(order):
            st = sd["state"].get(key)
            if st is None:
                self.exp_avg[pos].zero_()
                self.exp_avg_sq[pos].zero_()
                continue
            self.exp_avg[pos].copy_(st["exp_avg"])
            self.exp_avg_sq[pos].copy_(st["exp_avg_sq"])
            steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError("FusedAdam keeps one step counter; the checkpoint holds %s" % sorted(steps))
        self.step_dev.fill_(steps.pop() if steps else 0)

