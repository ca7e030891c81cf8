"""BASELINE configs[3] (uresnet_dense 3D train, ss=192, uns=5, uf=16, nc=5, bs=4) on this repo's dense path:
one training step = forward, per-voxel cross entropy, backward, Adam.  Not the metric's line (bench.py is);
recorded in profiles/README.md section 3.  --bs 1 first when memory is in doubt."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from uresnet_pytorch_b200 import sparseconvnet as scn  # noqa: E402
from uresnet_pytorch_b200 import synthetic  # noqa: E402
from uresnet_pytorch_b200.dense import DenseUResNet  # noqa: E402
from uresnet_pytorch_b200.model import Flags  # noqa: E402
from uresnet_pytorch_b200.optim import FusedAdam  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ss", type=int, default=192)
ap.add_argument("--bs", type=int, default=4)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--math", default="tf32x3", choices=["fp32", "tf32x3", "tf32"])
ap.add_argument("--profile", action="store_true", help="per-kernel device time of one step (torch.profiler)")
a = ap.parse_args()
mode = {"fp32": scn.MATH_FP32, "tf32x3": scn.MATH_TF32X3, "tf32": scn.MATH_TF32}[a.math]
dev = torch.device("cuda")
flags = Flags(DATA_DIM=3, SPATIAL_SIZE=a.ss, URESNET_FILTERS=16, URESNET_NUM_STRIDES=5, NUM_CLASS=5)
torch.manual_seed(0)
net = DenseUResNet(flags).to(dev).train()
opt = FusedAdam([p for p in net.parameters()], lr=1e-3)

# the sparse generator's events, densified
data, label, _ = synthetic.make_batch(a.bs, a.ss, 3, seed=99)
img = torch.zeros((a.bs, 1, a.ss, a.ss, a.ss), dtype=torch.float32)
lab = torch.zeros((a.bs, a.ss, a.ss, a.ss), dtype=torch.int64)
c = data[:, :3].astype(np.int64)
b = data[:, 3].astype(np.int64)
img[b, 0, c[:, 0], c[:, 1], c[:, 2]] = torch.from_numpy(data[:, 4].astype(np.float32))
lab[b, c[:, 0], c[:, 1], c[:, 2]] = torch.from_numpy(label[:, 0].astype(np.int64))
img, lab = img.to(dev), lab.to(dev)


def step():
    with scn.config.math(mode):
        out = net(img)
        loss = torch.nn.functional.cross_entropy(out, lab)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
    return loss


step()
torch.cuda.synchronize()
times = []
for _ in range(a.steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loss = step()
    e1.record()
    e1.synchronize()
    times.append(e0.elapsed_time(e1))
ms = sum(times) / len(times)
if a.profile:
    import collections
    import re
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            name = re.sub(r"^void ", "", ev.name)
            name = re.sub(r"(b200scn::|\(anonymous namespace\)::)", "", name)[:70]
            agg[name][0] += 1
            agg[name][1] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
    tot = sum(v[1] for v in agg.values())
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
        print("%-70s n=%4d  ms=%9.2f  %5.1f%%" % (k, c, t / 1e3, 100 * t / tot))
print(json.dumps({"config": "uresnet_dense 3D train ss=%d uns=5 uf=16 nc=5 bs=%d" % (a.ss, a.bs), "math": a.math,
                  "ms_per_step": ms, "grid_voxels_per_s": a.bs * a.ss ** 3 / ms * 1e3, "loss": float(loss),
                  "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30, "steps": a.steps}))
