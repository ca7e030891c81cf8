"""Layer-level tf32x3-vs-fp32 check of the dense convolutions at the shapes of the uf = 16 model."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from uresnet_pytorch_b200 import sparseconvnet as scn  # noqa: E402
from uresnet_pytorch_b200.dense import Conv, ConvTranspose  # noqa: E402

torch.manual_seed(0)
for kind, cin, cout, k, stride, size in (("conv", 64, 32, 3, 1, 8), ("conv", 64, 32, 1, 1, 8), ("conv", 32, 16, 3, 1, 16),
                                         ("conv", 32, 16, 1, 1, 16), ("conv", 64, 64, 3, 1, 4), ("conv", 32, 64, 3, 1, 4),
                                         ("conv", 32, 64, 2, 2, 8), ("conv", 16, 32, 2, 2, 16), ("convT", 64, 32, 3, 2, 4),
                                         ("convT", 32, 16, 3, 2, 8), ("conv", 32, 32, 3, 1, 8), ("conv", 64, 32, 3, 1, 9)):
    mod = (Conv(3, cin, cout, k, stride) if kind == "conv" else ConvTranspose(3, cin, cout)).cuda()
    for xkind in ("randn", "relu"):
        x0 = torch.randn((2, size, size, size, cin), device="cuda")
        if xkind == "relu":
            x0 = torch.relu(x0 + 0.5)
        res = []
        for mode in (scn.MATH_FP32, scn.MATH_TF32X3):
            x = x0.clone().requires_grad_(True)
            mod.zero_grad(set_to_none=True)
            with scn.math(mode):
                y = mod(x)
                w = torch.sin(torch.arange(y.numel(), device="cuda", dtype=torch.float32)).reshape(y.shape) + 0.3
                (y * w).sum().backward()
            res.append((y.detach(), x.grad.detach(), mod.weight.grad.detach().clone(), mod.bias.grad.detach().clone()))
        errs = [((a - b).abs().max() / a.abs().max()).item() for a, b in zip(*res)]
        mean_off = ((res[0][1] - res[1][1]).mean() / res[0][1].abs().mean()).item()
        print("%-5s %3d->%3d k%d s%d size %2d %-5s  y %.1e  dx %.1e (mean offset %.1e)  dW %.1e  db %.1e"
              % (kind, cin, cout, k, stride, size, xkind, errs[0], errs[1], mean_off, errs[2], errs[3]))
