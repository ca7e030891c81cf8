"""ORACLE -- TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see oracle/README.md).

"Oracle A": a torch-CPU restatement of the SparseConvNet (scn) module surface that
/root/reference/uresnet/models/uresnet_sparse.py:19-24 builds on.  scn is an un-vendored,
unpinned dependency absent from /root/reference; the semantics restated here are the ones
written out in SURVEY.md section 8(a) rows a2-a11 and Appendix A, anchored on the
reference's call sites:

    uresnet_sparse.py:20  scn.InputLayer(dimension, spatial_size, mode=3)
    uresnet_sparse.py:21  scn.SubmanifoldConvolution(dimension, nIn, nOut, 3, False)
    uresnet_sparse.py:22  scn.UNet(dimension, reps, nPlanes, residual_blocks=True, downsample=[2, 2])
    uresnet_sparse.py:23  scn.BatchNormReLU(m)
    uresnet_sparse.py:24  scn.OutputLayer(dimension)
    uresnet_sparse.py:35  self.sparseModel((coords, features))

The arithmetic follows scn's CPU backend shape: per kernel offset
`index_select -> mm -> index_add_` over (input row, output row) rule lists; autograd is by
construction.  Works in float32 or float64 (`module.double()`).

Nothing under uresnet_pytorch_b200/ imports this file.
"""
import numpy as np
import torch

from . import rules as _rules


class Metadata(object):
    """Per-forward cache of active sites and rulebooks, keyed by spatial size (a11)."""

    def __init__(self, dimension):
        self.dimension = dimension
        self.active = {}      # spatial size tuple -> int64 [n, D+1]
        self.in2row = None
        self.subm = {}        # (size, k) -> rules
        self.down = {}        # (in size, s) -> (parent, kofs)

    def subm_rules(self, size, k):
        key = (size, k)
        if key not in self.subm:
            self.subm[key] = [(torch.from_numpy(a).long(), torch.from_numpy(b).long())
                              for a, b in _rules.subm_rules(self.active[size], k)]
        return self.subm[key]

    def down_rules(self, size, out_size, s):
        key = (size, s)
        if key not in self.down:
            parent, kofs, coarse = _rules.down_rules(self.active[size], s)
            self.active[out_size] = coarse
            self.down[key] = (torch.from_numpy(parent).long(), torch.from_numpy(kofs).long())
        return self.down[key]


class SparseConvNetTensor(object):
    def __init__(self, features=None, metadata=None, spatial_size=None):
        self.features, self.metadata, self.spatial_size = features, metadata, spatial_size


def _size_key(spatial_size):
    return tuple(int(v) for v in spatial_size)


class Sequential(torch.nn.Sequential):
    def add(self, module):
        self._modules[str(len(self._modules))] = module
        return self


class ConcatTable(torch.nn.Module):
    def add(self, module):
        self._modules[str(len(self._modules))] = module
        return self

    def forward(self, x):
        return [m(x) for m in self._modules.values()]


class Identity(torch.nn.Module):
    def forward(self, x):
        return x


class AddTable(torch.nn.Module):
    def forward(self, xs):
        out = SparseConvNetTensor(None, xs[0].metadata, xs[0].spatial_size)
        out.features = sum(x.features for x in xs)
        return out


class JoinTable(torch.nn.Module):
    def forward(self, xs):
        out = SparseConvNetTensor(None, xs[0].metadata, xs[0].spatial_size)
        out.features = torch.cat([x.features for x in xs], 1)
        return out


class InputLayer(torch.nn.Module):
    """a2: rows in first-occurrence order, duplicates summed in input order (mode 3)."""

    def __init__(self, dimension, spatial_size, mode=3):
        super().__init__()
        self.dimension = dimension
        if isinstance(spatial_size, int):
            spatial_size = [spatial_size] * dimension
        self.spatial_size = _size_key(spatial_size)
        assert mode == 3, "the reference only uses mode=3 (uresnet_sparse.py:20)"

    def forward(self, inp):
        coords, feats = inp[0], inp[1]
        c = coords.detach().cpu().long().numpy()
        if c.shape[1] == self.dimension:  # no sample column -> sample 0
            c = np.concatenate([c, np.zeros((len(c), 1), np.int64)], 1)
        md = Metadata(self.dimension)
        in2row, active = _rules.input_rows(c)
        md.active[self.spatial_size] = active
        md.in2row = torch.from_numpy(in2row).long()
        out = torch.zeros(len(active), feats.shape[1], dtype=feats.dtype)
        out = out.index_add(0, md.in2row, feats)  # sequential over input rows on CPU
        return SparseConvNetTensor(out, md, self.spatial_size)


class OutputLayer(torch.nn.Module):
    """a10: inverse of the InputLayer rules; duplicates receive the same row."""

    def __init__(self, dimension):
        super().__init__()

    def forward(self, x):
        return x.features.index_select(0, x.metadata.in2row)


def _rule_conv(feats, weight, rules, n_out):
    out = torch.zeros(n_out, weight.shape[2], dtype=feats.dtype)
    for f, (ri, ro) in enumerate(rules):
        if len(ri):
            out = out.index_add(0, ro, feats.index_select(0, ri) @ weight[f])
    return out


class SubmanifoldConvolution(torch.nn.Module):
    """a3: out[o] = sum_k in[nbr_k(o)] @ W[k]."""

    def __init__(self, dimension, nIn, nOut, filter_size, bias):
        super().__init__()
        self.dimension, self.nIn, self.nOut, self.filter_size = dimension, nIn, nOut, filter_size
        K = filter_size ** dimension
        self.weight = torch.nn.Parameter(torch.empty(K, nIn, nOut).normal_(0, (2.0 / nIn / K) ** 0.5))
        self.bias = torch.nn.Parameter(torch.zeros(nOut)) if bias else None

    def forward(self, x):
        assert x.features.shape[1] == self.nIn
        rules = x.metadata.subm_rules(x.spatial_size, self.filter_size)
        out = _rule_conv(x.features, self.weight, rules, x.features.shape[0])
        if self.bias is not None:
            out = out + self.bias
        return SparseConvNetTensor(out, x.metadata, x.spatial_size)


class Convolution(torch.nn.Module):
    """a4: filter == stride; out[p // s] += in[p] @ W[k(p % s)]."""

    def __init__(self, dimension, nIn, nOut, filter_size, filter_stride, bias):
        super().__init__()
        assert filter_size == filter_stride, "oracle restates the k == s case the reference uses"
        self.dimension, self.nIn, self.nOut, self.s = dimension, nIn, nOut, filter_stride
        K = filter_size ** dimension
        self.weight = torch.nn.Parameter(torch.empty(K, nIn, nOut).normal_(0, (2.0 / nIn / K) ** 0.5))
        self.bias = torch.nn.Parameter(torch.zeros(nOut)) if bias else None

    def forward(self, x):
        out_size = tuple((v - self.s) // self.s + 1 for v in x.spatial_size)
        assert all((o - 1) * self.s + self.s == v for o, v in zip(out_size, x.spatial_size))
        parent, kofs = x.metadata.down_rules(x.spatial_size, out_size, self.s)
        n_out = len(x.metadata.active[out_size])
        rules = []
        for f in range(self.weight.shape[0]):
            sel = (kofs == f).nonzero()[:, 0]
            rules.append((sel, parent[sel]))
        out = _rule_conv(x.features, self.weight, rules, n_out)
        if self.bias is not None:
            out = out + self.bias
        return SparseConvNetTensor(out, x.metadata, out_size)


class Deconvolution(torch.nn.Module):
    """a5: the Convolution rulebook with roles swapped; output set = existing finer level."""

    def __init__(self, dimension, nIn, nOut, filter_size, filter_stride, bias):
        super().__init__()
        assert filter_size == filter_stride
        self.dimension, self.nIn, self.nOut, self.s = dimension, nIn, nOut, filter_stride
        K = filter_size ** dimension
        self.weight = torch.nn.Parameter(torch.empty(K, nIn, nOut).normal_(0, (2.0 / nIn / K) ** 0.5))
        self.bias = torch.nn.Parameter(torch.zeros(nOut)) if bias else None

    def forward(self, x):
        out_size = tuple((v - 1) * self.s + self.s for v in x.spatial_size)
        parent, kofs = x.metadata.down_rules(out_size, x.spatial_size, self.s)
        n_out = len(x.metadata.active[out_size])
        rules = []
        for f in range(self.weight.shape[0]):
            sel = (kofs == f).nonzero()[:, 0]
            rules.append((parent[sel], sel))
        out = _rule_conv(x.features, self.weight, rules, n_out)
        if self.bias is not None:
            out = out + self.bias
        return SparseConvNetTensor(out, x.metadata, out_size)


class NetworkInNetwork(torch.nn.Module):
    """a7: [N, nIn] @ W[nIn, nOut]."""

    def __init__(self, nIn, nOut, bias):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.empty(nIn, nOut).normal_(0, (2.0 / nIn) ** 0.5))
        self.bias = torch.nn.Parameter(torch.zeros(nOut)) if bias else None

    def forward(self, x):
        out = x.features @ self.weight
        if self.bias is not None:
            out = out + self.bias
        return SparseConvNetTensor(out, x.metadata, x.spatial_size)


class BatchNormalization(torch.nn.Module):
    """a6 / Appendix A.7: two-pass variance, eps inside the sqrt, `momentum` = retention."""

    def __init__(self, nPlanes, eps=1e-4, momentum=0.9, affine=True, leakiness=1):
        super().__init__()
        self.nPlanes, self.eps, self.momentum, self.leakiness = nPlanes, eps, momentum, leakiness
        self.register_buffer("running_mean", torch.zeros(nPlanes))
        self.register_buffer("running_var", torch.ones(nPlanes))
        self.weight = torch.nn.Parameter(torch.ones(nPlanes)) if affine else None
        self.bias = torch.nn.Parameter(torch.zeros(nPlanes)) if affine else None

    def forward(self, x):
        f = x.features
        assert f.shape[1] == self.nPlanes
        if self.training:
            n = f.shape[0]
            mean = f.mean(0)
            s = ((f - mean) ** 2).sum(0)
            with torch.no_grad():
                self.running_mean.mul_(self.momentum).add_((1 - self.momentum) * mean.detach())
                self.running_var.mul_(self.momentum).add_((1 - self.momentum) * s.detach() / max(n - 1, 1))
            invstd = (s / n + self.eps) ** -0.5
        else:
            mean = self.running_mean
            invstd = (self.running_var + self.eps) ** -0.5
        y = (f - mean) * invstd
        if self.weight is not None:
            y = y * self.weight + self.bias
        if self.leakiness != 1:
            y = torch.where(y > 0, y, y * self.leakiness)
        return SparseConvNetTensor(y, x.metadata, x.spatial_size)


class BatchNormReLU(BatchNormalization):
    def __init__(self, nPlanes, eps=1e-4, momentum=0.9):
        super().__init__(nPlanes, eps, momentum, True, 0)


class BatchNormLeakyReLU(BatchNormalization):
    def __init__(self, nPlanes, eps=1e-4, momentum=0.9, leakiness=0.333):
        super().__init__(nPlanes, eps, momentum, True, leakiness)


def UNet(dimension, reps, nPlanes, residual_blocks=False, downsample=[2, 2], leakiness=0, n_input_planes=-1):
    """a9: the recursive encoder/decoder of SURVEY.md 3.5."""

    def block(m, a, b):
        if residual_blocks:
            m.add(ConcatTable()
                  .add(Identity() if a == b else NetworkInNetwork(a, b, False))
                  .add(Sequential()
                       .add(BatchNormLeakyReLU(a, leakiness=leakiness))
                       .add(SubmanifoldConvolution(dimension, a, b, 3, False))
                       .add(BatchNormLeakyReLU(b, leakiness=leakiness))
                       .add(SubmanifoldConvolution(dimension, b, b, 3, False)))).add(AddTable())
        else:
            m.add(Sequential()
                  .add(BatchNormLeakyReLU(a, leakiness=leakiness))
                  .add(SubmanifoldConvolution(dimension, a, b, 3, False)))

    def U(planes, n_in=-1):
        m = Sequential()
        for _ in range(reps):
            block(m, n_in if n_in != -1 else planes[0], planes[0])
            n_in = -1
        if len(planes) > 1:
            m.add(ConcatTable()
                  .add(Identity())
                  .add(Sequential()
                       .add(BatchNormLeakyReLU(planes[0], leakiness=leakiness))
                       .add(Convolution(dimension, planes[0], planes[1], downsample[0], downsample[1], False))
                       .add(U(planes[1:]))
                       .add(BatchNormLeakyReLU(planes[1], leakiness=leakiness))
                       .add(Deconvolution(dimension, planes[1], planes[0], downsample[0], downsample[1], False))))
            m.add(JoinTable())
            for i in range(reps):
                block(m, planes[0] * (2 if i == 0 else 1), planes[0])
        return m

    return U(nPlanes, n_input_planes)
