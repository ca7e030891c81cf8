"""Generate tests/golden/dense_tiny_{2d,3d}.npz by running the UNMODIFIED reference dense model
(/root/reference/uresnet/models/uresnet_dense.py, loaded by path through oracle/dense_ref.py)
on torch-CPU.  Only runs where /root/reference exists (the build container).
    python tests/golden/make_dense_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import dense_ref  # noqa: E402


class Flags(object):
    def __init__(self, dim, ss, uf, uns):
        self.DATA_DIM, self.SPATIAL_SIZE, self.URESNET_FILTERS, self.URESNET_NUM_STRIDES = dim, ss, uf, uns
        self.NUM_CLASS, self.BN_MOMENTUM = 5, 0.9


def main():
    ref = dense_ref.load()
    for dim, ss, uf, uns, bs in ((3, 16, 4, 2, 2), (2, 32, 4, 3, 2)):
        torch.manual_seed(1000 + dim)
        net = ref.UResNet(Flags(dim, ss, uf, uns)).train()
        with torch.no_grad():
            for m in net.modules():
                if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm3d)):
                    m.weight.uniform_(0.7, 1.3)
                    m.bias.uniform_(-0.2, 0.2)
        x = torch.zeros((bs, 1) + (ss,) * dim)
        mask = torch.rand_like(x) < 0.05                      # sparse "LArTPC-like" occupancy
        x[mask] = torch.rand(int(mask.sum())) + 0.1
        out = net(x)
        w = torch.randn_like(out)
        (out * w).sum().backward()
        blob = {"dim": dim, "ss": ss, "uf": uf, "uns": uns, "x": x.numpy(), "w": w.numpy(), "out": out.detach().numpy()}
        for k, v in net.state_dict().items():
            blob["sd." + k] = v.numpy()
        for k, p in net.named_parameters():
            if p.grad is not None:
                blob["grad." + k] = p.grad.numpy()
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dense_tiny_%dd.npz" % dim)
        np.savez_compressed(path, **blob)
        print("wrote", path, os.path.getsize(path), "bytes", tuple(out.shape))


if __name__ == "__main__":
    main()
