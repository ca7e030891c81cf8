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


def grad_probes(name, n):
    """Three seeded probe vectors per parameter tensor (shared with tests/test_gpu_dense.py)."""
    seed = sum((i + 1) * ord(c) for i, c in enumerate(name)) % (2 ** 31)
    return np.random.RandomState(seed).standard_normal((3, n)).astype(np.float64)


def main_uf16():
    """uf = 16 (3D, two levels: 16- and 32-channel layers, the widths the tcgen05 dense kernel takes).  Compact
    fixture: weights rounded to fp16-representable values and stored as fp16, gradients as their norm and three
    seeded projections per parameter tensor."""
    ref = dense_ref.load()
    dim, ss, uf, uns, bs = 3, 16, 16, 2, 2
    torch.manual_seed(4242)
    net = ref.UResNet(Flags(dim, ss, uf, uns)).train()
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm3d)):
                m.weight.uniform_(0.7, 1.3)
                m.bias.uniform_(-0.2, 0.2)
        for p_ in net.parameters():
            p_.copy_(p_.half().float())
    x = torch.zeros((bs, 1) + (ss,) * dim)
    mask = torch.rand_like(x) < 0.05
    x[mask] = (torch.rand(int(mask.sum())) + 0.1).half().float()
    out = net(x)
    w = torch.randn_like(out).half().float()
    (out * w).sum().backward()
    blob = {"dim": dim, "ss": ss, "uf": uf, "uns": uns, "x": x.numpy().astype(np.float16), "w": w.numpy().astype(np.float16),
            "out": out.detach().numpy()}
    for k, v in net.state_dict().items():
        a = v.numpy()
        blob["sd." + k] = a.astype(np.float16) if (a.dtype == np.float32 and "running" not in k) else a
    for k, p_ in net.named_parameters():
        if p_.grad is not None:
            g = p_.grad.numpy().astype(np.float64).ravel()
            blob["gnorm." + k] = np.float64(np.sqrt((g * g).sum()))
            blob["gproj." + k] = grad_probes(k, g.size) @ g
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dense_uf16_3d.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path), "bytes", tuple(out.shape))


if __name__ == "__main__":
    main()
    main_uf16()
