"""The bar of the dense path (BASELINE configs[3], row a-D): the reference's dense U-ResNet architecture
(/root/reference/uresnet/models/uresnet_dense.py:28-226 -- replicate pre-padding, Conv3d k3 s1/s2 and k1, BatchNorm3d on
batch statistics, ConvTranspose3d k3 s2 p1 op1, skip concat) on torch.nn / cuDNN on the same B200, one training step =
forward, per-voxel cross entropy, backward, Adam.  /root/reference does not exist on the GPU box, so the architecture is
restated here with plain torch.nn layers; the parameter count is asserted against the reference's (50 320 383 at
uf=16, uns=5).  Usage: python profiles/dense_cudnn_bar.py [--ss 192] [--bs 4] [--no-tf32]"""
import argparse
import json
import sys
import time

import torch
import torch.nn as nn
import torch.nn.functional as F


def rpad(x, k, s):
    n = x.shape[-1]
    p = max(k - s, 0) if n % s == 0 else max(k - n % s, 0)
    return F.pad(x, (p // 2, p - p // 2) * 3, mode="replicate") if p else x


def cbn(ci, co, k, s, m):
    return nn.Sequential(nn.Conv3d(ci, co, k, s, 0), nn.BatchNorm3d(co, momentum=m, track_running_stats=False))


class Res(nn.Module):
    def __init__(self, ci, co, s, m):
        super().__init__()
        self.use_shortcut = ci != co or s != 1
        self.shortcut, self.residual1, self.residual2 = cbn(ci, co, 1, s, m), cbn(ci, co, 3, s, m), cbn(co, co, 3, 1, m)
        self.s = s

    def forward(self, x):
        sc = self.shortcut(rpad(x, 1, self.s)) if self.use_shortcut else x
        r = self.residual1(rpad(x, 3, self.s))
        r = self.residual2(rpad(r, 3, 1))
        return F.relu(sc + r)


class Dense(nn.Module):
    def __init__(self, uf=16, uns=5, nc=5, m=0.9):
        super().__init__()
        self.conv1 = nn.Sequential(*cbn(1, uf, 3, 1, m), nn.ReLU())
        self.enc, self.dec, self.up = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        c = uf
        for _ in range(uns):
            self.enc.append(nn.Sequential(Res(c, 2 * c, 2, m), Res(2 * c, 2 * c, 1, m)))
            c *= 2
        for _ in range(uns):
            self.dec.append(nn.Sequential(Res(c, c // 2, 1, m), Res(c // 2, c // 2, 1, m)))
            self.up.append(nn.Sequential(nn.ConvTranspose3d(c, c // 2, 3, 2, 1, 1),
                                         nn.BatchNorm3d(c // 2, momentum=m, track_running_stats=False), nn.ReLU()))
            c //= 2
        self.conv2 = nn.Sequential(*cbn(c, uf, 3, 1, m), nn.ReLU())
        self.conv3 = cbn(uf, nc, 3, 1, m)

    def forward(self, x):
        net = self.conv1(rpad(x, 3, 1))
        skip = {net.shape[1]: net}
        for e in self.enc:
            net = e(net)
            skip[net.shape[1]] = net
        for u, d in zip(self.up, self.dec):
            net = u(net)
            net = d(torch.cat((net, skip[net.shape[1]]), 1))
        return self.conv3(rpad(self.conv2(rpad(net, 3, 1)), 3, 1))


ap = argparse.ArgumentParser()
ap.add_argument("--ss", type=int, default=192)
ap.add_argument("--bs", type=int, default=4)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--no-tf32", action="store_true")
a = ap.parse_args()
torch.backends.cudnn.allow_tf32 = not a.no_tf32
torch.backends.cudnn.benchmark = True
net = Dense().cuda().train()
n_par = sum(p.numel() for p in net.parameters())
assert n_par == 50320383, n_par
opt = torch.optim.Adam(net.parameters(), lr=1e-3, fused=True)
torch.manual_seed(0)
img = (torch.rand(a.bs, 1, a.ss, a.ss, a.ss, device="cuda") < 1e-3).float() * torch.rand(a.bs, 1, a.ss, a.ss, a.ss, device="cuda")
lab = torch.randint(0, 5, (a.bs, a.ss, a.ss, a.ss), device="cuda")


def step():
    out = net(img)
    loss = F.cross_entropy(out, lab)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
ts = []
for _ in range(a.steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step()
    e1.record()
    e1.synchronize()
    ts.append(e0.elapsed_time(e1))
print(json.dumps({"what": "uresnet_dense architecture on torch.nn / cuDNN (%s), train step" % ("TF32 convs allowed: torch default" if not a.no_tf32 else "fp32 convs"),
                  "ss": a.ss, "bs": a.bs, "ms_per_step": sum(ts) / len(ts), "grid_voxels_per_s": a.bs * a.ss ** 3 / (sum(ts) / len(ts) * 1e-3),
                  "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30, "cudnn": torch.backends.cudnn.version(), "params": n_par}))
