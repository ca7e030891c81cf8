"""uresnet_pytorch_b200 -- B200-native (sm_100a) sparse U-ResNet hot path.

Layout:
    csrc/            hand-written CUDA kernels + the C-ABI (libb200scn.so, include/b200scn.h)
    sparseconvnet/   the scn-compatible Python module surface the reference model builds on
    model.py         the sparse U-ResNet + segmentation loss (mirror of the reference's model file)
    trainer.py       train step: eager or whole-step CUDA graph, data-parallel SUM all-reduce
    synthetic.py     seeded LArTPC-like event generator in the reference's blob layout
"""
from . import synthetic  # noqa: F401

__all__ = ["synthetic"]
