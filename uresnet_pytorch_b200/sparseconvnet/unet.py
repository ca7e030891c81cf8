"""scn.UNet(dimension, reps, nPlanes, residual_blocks, downsample) -- uresnet_sparse.py:22.

Graph per SURVEY.md 3.5 / Appendix A.8: `reps` pre-activation residual blocks per level,
BN+ReLU -> strided Convolution down, recursive inner U, BN+ReLU -> Deconvolution up, skip
concat, `reps` more blocks (the first maps 2P -> P through a NetworkInNetwork shortcut).
Child indices are what scn produces, so state-dict keys match reference checkpoints.
"""
from .modules import (AddTable, BatchNormLeakyReLU, ConcatTable, Convolution, Deconvolution, Identity, JoinTable,
                      MarkedSequential, NetworkInNetwork, Sequential, SubmanifoldConvolution)


def _residual_block(dimension, a, b, leakiness):
    shortcut = Identity() if a == b else NetworkInNetwork(a, b, False)
    body = Sequential()
    body.add(BatchNormLeakyReLU(a, leakiness=leakiness))
    body.add(SubmanifoldConvolution(dimension, a, b, 3, False))
    body.add(BatchNormLeakyReLU(b, leakiness=leakiness))
    body.add(SubmanifoldConvolution(dimension, b, b, 3, False))
    return ConcatTable().add(shortcut).add(body)


def _plain_block(dimension, a, b, leakiness):
    body = Sequential()
    body.add(BatchNormLeakyReLU(a, leakiness=leakiness))
    body.add(SubmanifoldConvolution(dimension, a, b, 3, False))
    return body


def UNet(dimension, reps, nPlanes, residual_blocks=False, downsample=[2, 2], leakiness=0, n_input_planes=-1):
    def add_block(seq, a, b):
        if residual_blocks:
            seq.add(_residual_block(dimension, a, b, leakiness))
            seq.add(AddTable())
        else:
            seq.add(_plain_block(dimension, a, b, leakiness))

    def level(planes, n_in, depth=0):
        seq = Sequential()
        width = planes[0]
        for r in range(reps):
            add_block(seq, n_in if (r == 0 and n_in != -1) else width, width)
        if len(planes) > 1:
            inner = MarkedSequential("unet.inner%d" % depth)  # a Sequential; tells a trainer when its backward ended
            inner.add(BatchNormLeakyReLU(width, leakiness=leakiness))
            inner.add(Convolution(dimension, width, planes[1], downsample[0], downsample[1], False))
            inner.add(level(planes[1:], -1, depth + 1))
            inner.add(BatchNormLeakyReLU(planes[1], leakiness=leakiness))
            inner.add(Deconvolution(dimension, planes[1], width, downsample[0], downsample[1], False))
            seq.add(ConcatTable().add(Identity()).add(inner))
            seq.add(JoinTable())
            for r in range(reps):
                add_block(seq, 2 * width if r == 0 else width, width)
        return seq

    return level(list(nPlanes), n_input_planes)
