"""One U-Net level's SubmanifoldConvolution forward + backward on the tcgen05 kernels, a few launches (short command
for ncu).  Usage: python profiles/tc_micro.py [--level 3] [--config c2|c5] [--reps 3]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from uresnet_pytorch_b200 import sparseconvnet as scn, synthetic  # noqa: E402
from uresnet_pytorch_b200.sparseconvnet.metadata import Metadata  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c2")
ap.add_argument("--level", type=int, default=3)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
ss, uf, ev = (512, 16, 1) if a.config == "c2" else (768, 32, 8)
data, _, _ = synthetic.make_batch(ev, ss, 3, seed=1234)
coords = torch.from_numpy(data[:, :4]).cuda()
md = Metadata(3)
size = (ss,) * 3
md.set_input(coords, size)
for _ in range(a.level):
    out = tuple((v - 2) // 2 + 1 for v in size)
    md.down_rules(size, out, 2)
    size = out
C = uf * (a.level + 1)
n = md.level(size).n
torch.manual_seed(0)
conv = scn.SubmanifoldConvolution(3, C, C, 3, False).cuda()
feats = torch.randn(n, C, device="cuda", requires_grad=True)
with scn.math(scn.MATH_TF32X3):
    for _ in range(a.reps):
        y = conv(scn.SparseConvNetTensor(feats, md, size)).features
        y.sum().backward()
torch.cuda.synchronize()
print("level %d rows %d C %d ok" % (a.level, n, C))
