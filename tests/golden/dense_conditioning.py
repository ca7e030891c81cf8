"""How well conditioned is the uf = 16 dense fixture?  Runs the UNMODIFIED reference model (torch-CPU, fp32) on
tests/golden/dense_uf16_3d.npz with its convolution weights perturbed by a relative 1e-7 / 1e-6 / 1e-5 and prints
how far logits and gradients move.  Measured in the build container (two seeds each):

    eps 1e-07: logits 2.3e-06  gradients global 1.4e-06  worst significant tensor 3.7e-06
    eps 1e-06: logits 5.6e-06  gradients global 2.0e-03  worst significant tensor 4.7e-03
    eps 1e-05: logits 5.1e-05 .. 6.0e-05  gradients global 7.6e-03 .. 1.6e-02  worst tensor 1.3e-02 .. 3.2e-02

A forward perturbation of a few 1e-6 already flips isolated ReLU masks (8^3 x 2 voxels per channel at the inner
level), and one flipped mask moves a BatchNorm bias gradient by a full |dy|: the gradient tolerance of
tests/test_gpu_dense.py::test_dense_model_uf16_tensor_core_modes_match_reference_golden follows this curve, not the
kernels' own error (each tcgen05 layer matches the FFMA layer to 1e-5, profiles/dbg_dense_layers.py).
Only runs where /root/reference exists.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
from make_dense_golden import Flags  # noqa: E402
from oracle import dense_ref  # noqa: E402


def run(ref, g, eps, seed):
    net = ref.UResNet(Flags(3, 16, 16, 2)).train()
    sd = {k[3:]: torch.from_numpy(g[k].astype(np.float32) if g[k].dtype == np.float16 else g[k])
          for k in g.files if k.startswith("sd.")}
    net.load_state_dict(sd)
    torch.manual_seed(seed)
    if eps:
        with torch.no_grad():
            for _, p in net.named_parameters():
                if p.dim() > 1:
                    p.mul_(1 + eps * torch.randn_like(p))
    out = net(torch.from_numpy(g["x"].astype(np.float32)))
    (out * torch.from_numpy(g["w"].astype(np.float32))).sum().backward()
    return out.detach(), {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}


def main():
    g = np.load(os.path.join(HERE, "dense_uf16_3d.npz"))
    ref = dense_ref.load()
    o0, g0 = run(ref, g, 0, 0)
    top = max(v.norm() for v in g0.values())
    sig = [n for n in g0 if g0[n].norm() > 1e-3 * top]
    for eps in (1e-7, 1e-6, 1e-5):
        for seed in (1, 2):
            o1, g1 = run(ref, g, eps, seed)
            num = sum(((g1[n] - g0[n]).double() ** 2).sum().item() for n in g0)
            den = sum((g0[n].double() ** 2).sum().item() for n in g0)
            worst = max(((g1[n] - g0[n]).norm() / g0[n].norm()).item() for n in sig)
            print("eps %.0e seed %d: logits %.2e  gradients global %.2e  worst significant tensor %.2e"
                  % (eps, seed, ((o1 - o0).abs().max() / o0.abs().max()).item(), (num / den) ** 0.5, worst))


if __name__ == "__main__":
    main()
