"""Per-level forward time of SubmanifoldConvolution at the U-Net levels of one BASELINE event: the tcgen05 rule-table
kernel (sparse_tc.cu) against the mma.sync pair-list kernel (conv_gpl.cu), warm L2 (the in-step regime), CUDA events
around 50 back-to-back launches.  Usage: python profiles/conv_levels_bench.py [--config c2|c5] [--events N]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from uresnet_pytorch_b200 import sparseconvnet as scn, synthetic  # noqa: E402
from uresnet_pytorch_b200.sparseconvnet.metadata import Metadata  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c2")
ap.add_argument("--events", type=int, default=0)
a = ap.parse_args()
ss, uf, uns, ev = (512, 16, 5, 1) if a.config == "c2" else (768, 32, 6, 8)
ev = a.events or ev
data, _, _ = synthetic.make_batch(ev, ss, 3, seed=1234)
coords = torch.from_numpy(data[:, :4]).cuda()
md = Metadata(3)
size = (ss,) * 3
md.set_input(coords, size)
torch.manual_seed(0)
print("config %s, %d event(s)" % (a.config, ev))
for lvl in range(uns):
    C = uf * (lvl + 1)
    lv = md.level(size)
    n = lv.n
    x = scn.SparseConvNetTensor(torch.randn(n, C, device="cuda"), md, size)
    conv = scn.SubmanifoldConvolution(3, C, C, 3, False).cuda()
    res = {}
    with torch.no_grad(), scn.math(scn.MATH_TF32X3):
        for name, tc in (("tc", True), ("gpl", False)):
            scn.config.use_tc = tc
            for _ in range(5):
                y = conv(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50):
                y = conv(x)
            e1.record()
            e1.synchronize()
            res[name] = (e0.elapsed_time(e1) / 50 * 1e3, y.features)
        scn.config.use_tc = True
    err = (res["tc"][1] - res["gpl"][1]).abs().max().item() / res["gpl"][1].abs().max().item()
    pairs = int((md.subm_table(size, 3)[:, :n] >= 0).sum().item())
    print("level %d: rows %6d pairs %7d C %3d | tcgen05 %7.1f us | mma.sync gpl %7.1f us | max rel diff %.1e"
          % (lvl, n, pairs, C, res["tc"][0], res["gpl"][0], err))
    if lvl + 1 < uns:
        out = tuple((v - 2) // 2 + 1 for v in size)
        md.down_rules(size, out, 2)
        size = out
