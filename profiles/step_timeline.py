"""Per-kernel device time of the captured training step under real (warm, back-to-back) execution:
torch.profiler (CUPTI) over a few CUDA-graph replays of bench.py's workload.  Complements the ncu
launch lists, which serialise and cold-start every kernel."""
import argparse
import collections
import re
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402
from uresnet_pytorch_b200 import sparseconvnet as scn  # noqa: E402
from uresnet_pytorch_b200.model import Flags  # noqa: E402
from uresnet_pytorch_b200.trainer import Trainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--math", default="tf32x3")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--no-overlap", action="store_true")
ap.add_argument("--config", default="c2", choices=["c2", "c3", "c5"])
a = ap.parse_args()
mode = {"fp32": scn.MATH_FP32, "tf32x3": scn.MATH_TF32X3, "tf32": scn.MATH_TF32}[a.math]
events = 1
if a.config != "c2":
    cfg, events, _ = bench.EXTRA_CONFIGS[a.config]
    bench.CONFIG.update(cfg)
C = bench.CONFIG
flags = Flags(DATA_DIM=C["dim"], SPATIAL_SIZE=C["ss"], URESNET_FILTERS=C["uf"], URESNET_NUM_STRIDES=C["uns"],
              NUM_CLASS=C["nc"])


class A(object):
    events_per_gpu, pool = events, 4


pool = bench.make_pool(A, 0)
cap = ((max(len(d) for d, _ in pool) * 9 // 8 + 4095) // 4096) * 4096 * int(os.environ.get("B200SCN_BENCH_CAP_MULT", "1"))
tr = Trainer(flags, graph=True, capacity=cap, events_per_step=events, math_mode=mode, overlap_wgrad=not a.no_overlap)
for i in range(4):
    tr.train_step(*pool[i % 4])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(a.steps):
        tr.train_step(*pool[i % 4])
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
t_min, t_max = None, None
for ev in prof.events():
    if ev.device_type.name != "CUDA":
        continue
    m = re.search(r"(k_\w+(<[^>]*>)?)", ev.name)
    name = m.group(1) if m else ev.name.split("(")[0][-70:]
    agg[name][0] += 1
    agg[name][1] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
    s, e = ev.time_range.start, ev.time_range.end
    t_min = s if t_min is None else min(t_min, s)
    t_max = e if t_max is None else max(t_max, e)
total = sum(v[1] for v in agg.values())
print("steps %d  span %.3f ms/step  sum of kernel time %.3f ms/step  kernels+memops/step %.0f" % (
    a.steps, (t_max - t_min) / 1e3 / a.steps, total / 1e3 / a.steps, sum(v[0] for v in agg.values()) / a.steps))
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:32]:
    print("%-70s n/step=%6.1f  us/step=%8.1f  avg=%6.1f us  %4.1f%%" % (k, c / a.steps, t / a.steps, t / c, 100 * t / total))
