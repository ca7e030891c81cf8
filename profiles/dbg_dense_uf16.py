"""Per-parameter gradient error of the dense model (uf = 16 fixture) in each math mode, in model order."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from conftest import small_flags  # noqa: E402
from make_dense_golden import grad_probes  # noqa: E402
from uresnet_pytorch_b200 import sparseconvnet as scn  # noqa: E402
from uresnet_pytorch_b200.dense import DenseUResNet  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "dense_uf16_3d.npz"))
flags = small_flags(DATA_DIM=3, SPATIAL_SIZE=int(g["ss"]), URESNET_FILTERS=int(g["uf"]), URESNET_NUM_STRIDES=int(g["uns"]))
sd = {k[3:]: torch.from_numpy(g[k].astype(np.float32) if g[k].dtype == np.float16 else g[k]) for k in g.files if k.startswith("sd.")}
res = {}
for mode in ("fp32", "tf32x3", "tf32"):
    net = DenseUResNet(flags)
    net.load_state_dict(sd)
    net.cuda().train()
    with scn.math({"fp32": scn.MATH_FP32, "tf32x3": scn.MATH_TF32X3, "tf32": scn.MATH_TF32}[mode]):
        out = net(torch.from_numpy(g["x"].astype(np.float32)).cuda())
        (out * torch.from_numpy(g["w"].astype(np.float32)).cuda()).sum().backward()
    for name, p in net.named_parameters():
        if "gnorm." + name not in g.files:
            continue
        a = p.grad.detach().cpu().double().numpy().ravel()
        norm = float(g["gnorm." + name])
        d = float(np.sqrt(np.mean((grad_probes(name, a.size) @ a - g["gproj." + name]) ** 2)))
        res.setdefault(name, {})[mode] = (d / max(norm, 1e-30), norm, tuple(p.shape))
for name, r in res.items():
    print("%-55s %-22s norm %.2e  fp32 %.1e  tf32x3 %.1e  tf32 %.1e" % (name, r["fp32"][2], r["fp32"][1], r["fp32"][0], r["tf32x3"][0], r["tf32"][0]))
