"""Dense-path convolution microbenchmark: the tcgen05 implicit-GEMM kernel (csrc/dense_tc.cu, single-pass
TF32) against the exact-fp32 FFMA kernel (csrc/dense.cu) on the layer shapes of BASELINE config 4
(uresnet_dense 3D ss=192, bs=4: 16->32->...->512 filters, levels 192^3 ... 12^3), forward only.
CUDA events on the launching stream, L2 flushed between launches."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from uresnet_pytorch_b200 import sparseconvnet as scn  # noqa: E402
from uresnet_pytorch_b200.dense import Conv  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--only-tc", action="store_true")
a = ap.parse_args()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
shapes = [(96, 32, 32), (48, 64, 64), (24, 128, 128), (12, 256, 256), (12, 512, 512), (96, 32, 64)]
for size, cin, cout in shapes:
    torch.manual_seed(0)
    conv = Conv(3, cin, cout, 3, 1).cuda()
    x = torch.randn(a.batch, size, size, size, cin, device="cuda")
    flops = 2.0 * a.batch * size ** 3 * 27 * cin * cout
    row = {"shape": "B%d %d^3 %d->%d k3s1" % (a.batch, size, cin, cout), "gflop": flops / 1e9}
    for name, mode in (("ffma_fp32", scn.MATH_FP32), ("tcgen05_tf32", scn.MATH_TF32), ("tcgen05_tf32x3", scn.MATH_TF32X3)):
        if a.only_tc and mode == scn.MATH_FP32:
            continue
        ts = []
        with torch.no_grad(), scn.math(mode):
            for i in range(a.reps + 2):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                y = conv(x)
                e1.record()
                e1.synchronize()
                if i >= 2:
                    ts.append(e0.elapsed_time(e1) * 1e-3)
        t = sum(ts) / len(ts)
        row[name + "_ms"] = t * 1e3
        row[name + "_tflops"] = flops / t / 1e12
    print(json.dumps(row))
