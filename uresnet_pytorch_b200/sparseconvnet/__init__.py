"""A `sparseconvnet`-compatible module surface backed by hand-written sm_100a CUDA.

`/root/reference/uresnet/models/uresnet_sparse.py:9` does `import sparseconvnet as scn`; call
`uresnet_pytorch_b200.sparseconvnet.install()` once (or put this package on the path under
that name) and the reference model builds and runs unchanged on top of libb200scn.
"""
import sys

from . import config
from .config import MATH_FP32, MATH_TF32, MATH_TF32X3, math, static_mode
from .metadata import Metadata
from .modules import (AddTable, BatchNormalization, BatchNormLeakyReLU, BatchNormReLU, ConcatTable, Convolution,
                      Deconvolution, Identity, InputLayer, JoinTable, NetworkInNetwork, OutputLayer, Sequential,
                      SparseConvNetTensor, SubmanifoldConvolution, ValidConvolution)
from .unet import UNet


def install(name="sparseconvnet"):
    """Register this package under `name` so that `import sparseconvnet as scn` resolves to it."""
    sys.modules[name] = sys.modules[__name__]
    return sys.modules[__name__]
