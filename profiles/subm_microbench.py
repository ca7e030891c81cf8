"""Large-batch submanifold-conv microbenchmark on its own (the kernel bench.py reports in `roofline`).
Used as the short command for `ncu --set full -k regex:k_conv` captures."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--roofline-events", type=int, default=128)
ap.add_argument("--math", default="tf32x3")
a = ap.parse_args()
from uresnet_pytorch_b200 import sparseconvnet as scn  # noqa: E402
mode = {"fp32": scn.MATH_FP32, "tf32x3": scn.MATH_TF32X3, "tf32": scn.MATH_TF32}[a.math]
print(json.dumps(bench.subm_roofline(a, mode)))
