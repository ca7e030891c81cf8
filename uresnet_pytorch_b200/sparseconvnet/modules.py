"""The scn module surface /root/reference/uresnet/models/uresnet_sparse.py:19-24 builds on,
re-implemented over libb200scn.  Constructor signatures, child naming ('0', '1', ...),
parameter names/shapes and initialisation follow scn, so the reference
model file runs unchanged with this package imported as `sparseconvnet`.
"""
import torch

from . import config
from . import functional as F
from .metadata import Metadata


class SparseConvNetTensor(object):
    """{features [rows, C] float32, metadata, spatial_size}."""

    def __init__(self, features=None, metadata=None, spatial_size=None):
        self.features = features
        self.metadata = metadata
        self.spatial_size = spatial_size

    def _like(self, features, spatial_size=None):
        return SparseConvNetTensor(features, self.metadata,
                                   self.spatial_size if spatial_size is None else spatial_size)

    def get_spatial_locations(self):
        """int64 [n, D+1] coordinates (+ sample index) of the rows, like scn's method of the same name."""
        return self.metadata.active_coords(self.spatial_size)

    def cuda(self):
        return self

    def __repr__(self):
        f = self.features
        return "SparseConvNetTensor<features=%s, spatial_size=%s>" % (
            None if f is None else tuple(f.shape), self.spatial_size)


def _size_tuple(dimension, spatial_size):
    if isinstance(spatial_size, int):
        return (spatial_size,) * dimension
    if torch.is_tensor(spatial_size):
        spatial_size = spatial_size.tolist()
    return tuple(int(v) for v in spatial_size)


class Sequential(torch.nn.Sequential):
    def add(self, module):
        self._modules[str(len(self._modules))] = module
        return self

    def input_spatial_size(self, out_size):
        for m in reversed(list(self._modules.values())):
            if hasattr(m, "input_spatial_size"):
                out_size = m.input_spatial_size(out_size)
        return out_size


class _FusedSum(list):
    """[shortcut, body] of a residual block whose sum was already formed inside the body's last convolution."""
    fused = None


class ConcatTable(torch.nn.Module):
    def add(self, module):
        self._modules[str(len(self._modules))] = module
        return self

    def forward(self, x):
        mods = list(self._modules.values())
        # residual block of scn.UNet (shortcut | BN-conv-BN-conv) followed by AddTable: hand the shortcut to the last
        # convolution, which adds it in its epilogue -- the AddTable after it then has nothing left to do
        if (config.fuse_residual and len(mods) == 2 and isinstance(mods[0], (Identity, NetworkInNetwork))
                and isinstance(mods[1], Sequential) and len(mods[1]) > 0
                and isinstance(list(mods[1]._modules.values())[-1], SubmanifoldConvolution)
                and list(mods[1]._modules.values())[-1].bias is None):
            s = mods[0](x)
            body = list(mods[1]._modules.values())
            y = x
            for m in body[:-1]:
                y = m(y)
            if tuple(y.spatial_size) == tuple(s.spatial_size) and s.features.shape[1] == body[-1].nOut:
                out = _FusedSum([s, None])
                out.fused = body[-1](y, res=s.features)
                return out
            return [s, body[-1](y)]
        return [m(x) for m in mods]

    def input_spatial_size(self, out_size):
        return list(self._modules.values())[0].input_spatial_size(out_size) if len(self._modules) else out_size


class Identity(torch.nn.Module):
    def forward(self, x):
        return x


class _Mark(torch.autograd.Function):
    """Identity whose backward tells `config.backward_marks` that everything downstream of this point has run its
    backward (used to start the gradient all-reduce of the inner U-Net levels while the outer level still runs)."""

    @staticmethod
    def forward(ctx, x, tag):
        ctx.tag = tag
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        cb = config.backward_mark_callback
        if cb is not None:
            cb(ctx.tag)
        return g, None


class MarkedSequential(Sequential):
    """scn.Sequential that announces the end of its own backward pass (no parameters, same child names)."""

    def __init__(self, tag):
        super().__init__()
        self._mark_tag = tag

    def forward(self, x):
        if config.backward_mark_callback is not None and torch.is_grad_enabled() and x.features.requires_grad:
            x = x._like(_Mark.apply(x.features, self._mark_tag))
        return super().forward(x)

    def input_spatial_size(self, out_size):
        return out_size


class AddTable(torch.nn.Module):
    def forward(self, xs):
        if getattr(xs, "fused", None) is not None:
            return xs.fused
        feats = xs[0].features
        for x in xs[1:]:
            feats = feats + x.features
        return xs[0]._like(feats)

    def input_spatial_size(self, out_size):
        return out_size


class JoinTable(torch.nn.Module):
    def forward(self, xs):
        return xs[0]._like(torch.cat([x.features for x in xs], 1))

    def input_spatial_size(self, out_size):
        return out_size


class InputLayer(torch.nn.Module):
    """scn.InputLayer(dimension, spatial_size, mode=3) -- uresnet_sparse.py:20.

    forward((coords [N, D+1], features [N, C])) -> SparseConvNetTensor; rows in order of first
    occurrence, duplicate points summed (mode 3).  Modes 1/2/4 of scn are not on the
    reference's path and raise."""

    def __init__(self, dimension, spatial_size, mode=3):
        super().__init__()
        if mode != 3:
            raise NotImplementedError("only InputLayer mode=3 (sum duplicates) is implemented")
        self.dimension = dimension
        self.spatial_size = _size_tuple(dimension, spatial_size)
        self.mode = mode
        self.last_error_flag = None

    def forward(self, inp):
        coords, feats = inp[0], inp[1]
        if not feats.is_cuda:
            raise RuntimeError("b200 sparseconvnet: features must be on a CUDA device (there is no CPU path)")
        if coords.device != feats.device:
            coords = coords.to(feats.device)
        md = Metadata(self.dimension)
        lv = md.set_input(coords, self.spatial_size)
        # static regime: nothing is read back, so the out-of-range flag of this forward stays reachable for the
        # owner of the step (Trainer folds it into its loss read-back)
        self.last_error_flag = lv.n_dev[1:2] if md.static else None
        out = F.InputGather.apply(feats.float(), md, lv.rows, md.n_dev_ptr(lv))
        return SparseConvNetTensor(out, md, self.spatial_size)

    def input_spatial_size(self, out_size):
        return out_size


class OutputLayer(torch.nn.Module):
    """scn.OutputLayer(dimension) -- uresnet_sparse.py:24: back to the original point order."""

    def __init__(self, dimension):
        super().__init__()
        self.dimension = dimension

    def forward(self, x):
        md = x.metadata
        lv = md.level(x.spatial_size)
        return F.OutputGather.apply(x.features, md, lv.rows, md.n_dev_ptr(lv))

    def input_spatial_size(self, out_size):
        return out_size


def _conv_weight(K, nIn, nOut):
    return torch.nn.Parameter(torch.empty(K, nIn, nOut).normal_(0, (2.0 / nIn / K) ** 0.5))


def _load_compat(module, state_dict, prefix):
    # newer scn checkpoints carry a groups dim: [K, 1, nIn, nOut]
    key = prefix + "weight"
    if key in state_dict and state_dict[key].dim() == module.weight.dim() + 1 and state_dict[key].shape[1] == 1:
        state_dict[key] = state_dict[key].squeeze(1)


class SubmanifoldConvolution(torch.nn.Module):
    """scn.SubmanifoldConvolution(dimension, nIn, nOut, filter_size, bias) -- uresnet_sparse.py:21."""

    def __init__(self, dimension, nIn, nOut, filter_size, bias, groups=1):
        super().__init__()
        if groups != 1:
            raise NotImplementedError("groups != 1")
        if not isinstance(filter_size, int):
            fs = _size_tuple(dimension, filter_size)
            if len(set(fs)) != 1:
                raise NotImplementedError("anisotropic filters")
            filter_size = fs[0]
        self.dimension, self.nIn, self.nOut, self.filter_size = dimension, nIn, nOut, filter_size
        self.filter_volume = filter_size ** dimension
        self.weight = _conv_weight(self.filter_volume, nIn, nOut)
        self.bias = torch.nn.Parameter(torch.zeros(nOut)) if bias else None

    def _load_from_state_dict(self, state_dict, prefix, *args, **kw):
        _load_compat(self, state_dict, prefix)
        super()._load_from_state_dict(state_dict, prefix, *args, **kw)

    def forward(self, x, res=None):
        if x.features.shape[1] != self.nIn:
            raise AssertionError("SubmanifoldConvolution: expected %d input planes, got %d"
                                 % (self.nIn, x.features.shape[1]))
        md = x.metadata
        lv = md.level(x.spatial_size)
        nbr = md.subm_table(x.spatial_size, self.filter_size)
        nd = md.n_dev_ptr(lv)
        gpl = None
        if (config.use_gpl and config.math_mode != 0 and self.nIn % 16 == 0 and self.nOut % 16 == 0
                and lv.rows >= config.gpl_min_rows):
            gpl = md.subm_gpl(x.spatial_size, self.filter_size)
        hint = config.gpl_rows_by_size.get(tuple(x.spatial_size), 0) if md.static else 0
        out = F.TableConv.apply(x.features, self.weight, nbr, lv.rows, nd, lv.rows, nd, True, None, None, gpl, gpl,
                                int(hint), res)
        if self.bias is not None:
            out = out + self.bias
        return x._like(out)

    def input_spatial_size(self, out_size):
        return out_size

    def __repr__(self):
        return "SubmanifoldConvolution %d->%d C%d" % (self.nIn, self.nOut, self.filter_size)


class ValidConvolution(SubmanifoldConvolution):
    """old scn name of SubmanifoldConvolution"""


def _strided_sizes(dimension, filter_size, filter_stride):
    fs = _size_tuple(dimension, filter_size)
    st = _size_tuple(dimension, filter_stride)
    if len(set(fs)) != 1 or len(set(st)) != 1 or fs[0] != st[0]:
        raise NotImplementedError(
            "strided sparse convolutions are implemented for filter_size == filter_stride "
            "(the reference uses downsample=[2, 2], uresnet_sparse.py:22)")
    return fs[0]


class Convolution(torch.nn.Module):
    """scn.Convolution(dimension, nIn, nOut, filter_size, filter_stride, bias) -- inside scn.UNet."""

    def __init__(self, dimension, nIn, nOut, filter_size, filter_stride, bias, groups=1):
        super().__init__()
        if groups != 1:
            raise NotImplementedError("groups != 1")
        self.s = _strided_sizes(dimension, filter_size, filter_stride)
        self.dimension, self.nIn, self.nOut = dimension, nIn, nOut
        self.filter_volume = self.s ** dimension
        self.weight = _conv_weight(self.filter_volume, nIn, nOut)
        self.bias = torch.nn.Parameter(torch.zeros(nOut)) if bias else None

    def _load_from_state_dict(self, state_dict, prefix, *args, **kw):
        _load_compat(self, state_dict, prefix)
        super()._load_from_state_dict(state_dict, prefix, *args, **kw)

    def forward(self, x):
        if x.features.shape[1] != self.nIn:
            raise AssertionError("Convolution: expected %d input planes, got %d" % (self.nIn, x.features.shape[1]))
        out_size = tuple((v - self.s) // self.s + 1 for v in x.spatial_size)
        if any((o - 1) * self.s + self.s != v for o, v in zip(out_size, x.spatial_size)):
            raise AssertionError("Convolution: spatial size %s is not divisible by stride %d"
                                 % (x.spatial_size, self.s))
        md = x.metadata
        lv = md.level(x.spatial_size)
        rules = md.down_rules(x.spatial_size, out_size, self.s)
        nxt = md.level(out_size)
        g_child = g_parent = None
        if (config.use_gpl and config.math_mode != 0 and self.nIn % 16 == 0 and self.nOut % 16 == 0
                and nxt.rows >= config.gpl_min_rows):
            g_child, g_parent = md.strided_gpl(x.spatial_size, out_size, self.s)
        hint = config.gpl_rows_by_size.get(tuple(out_size), 0) if md.static else 0
        out = F.TableConv.apply(x.features, self.weight, rules["child"], nxt.rows, md.n_dev_ptr(nxt), lv.rows,
                                md.n_dev_ptr(lv), False, rules["parent"], rules["kofs"], g_child, g_parent, int(hint))
        if self.bias is not None:
            out = out + self.bias
        return x._like(out, out_size)

    def input_spatial_size(self, out_size):
        return tuple((v - 1) * self.s + self.s for v in out_size)

    def __repr__(self):
        return "Convolution %d->%d C%d/%d" % (self.nIn, self.nOut, self.s, self.s)


class Deconvolution(torch.nn.Module):
    """scn.Deconvolution(dimension, nIn, nOut, filter_size, filter_stride, bias) -- inside scn.UNet.

    The output active set is the *existing* finer level (the one the matching Convolution
    consumed), so skip connections line up row for row."""

    def __init__(self, dimension, nIn, nOut, filter_size, filter_stride, bias, groups=1):
        super().__init__()
        if groups != 1:
            raise NotImplementedError("groups != 1")
        self.s = _strided_sizes(dimension, filter_size, filter_stride)
        self.dimension, self.nIn, self.nOut = dimension, nIn, nOut
        self.filter_volume = self.s ** dimension
        self.weight = _conv_weight(self.filter_volume, nIn, nOut)
        self.bias = torch.nn.Parameter(torch.zeros(nOut)) if bias else None

    def _load_from_state_dict(self, state_dict, prefix, *args, **kw):
        _load_compat(self, state_dict, prefix)
        super()._load_from_state_dict(state_dict, prefix, *args, **kw)

    def forward(self, x):
        if x.features.shape[1] != self.nIn:
            raise AssertionError("Deconvolution: expected %d input planes, got %d" % (self.nIn, x.features.shape[1]))
        out_size = tuple((v - 1) * self.s + self.s for v in x.spatial_size)
        md = x.metadata
        if out_size not in md.levels:
            raise RuntimeError("Deconvolution: no finer active set of spatial size %s exists" % (out_size,))
        fine = md.level(out_size)
        rules = md.down_rules(out_size, x.spatial_size, self.s)
        coarse = md.level(x.spatial_size)
        g_child = g_parent = None
        if (config.use_gpl and config.math_mode != 0 and self.nIn % 16 == 0 and self.nOut % 16 == 0
                and coarse.rows >= config.gpl_min_rows):
            g_child, g_parent = md.strided_gpl(out_size, x.spatial_size, self.s)
        out = F.ParentConv.apply(x.features, self.weight, rules["parent"], rules["kofs"], rules["child"], fine.rows,
                                 md.n_dev_ptr(fine), coarse.rows, md.n_dev_ptr(coarse), g_parent, g_child)
        if self.bias is not None:
            out = out + self.bias
        return x._like(out, out_size)

    def input_spatial_size(self, out_size):
        return tuple((v - self.s) // self.s + 1 for v in out_size)

    def __repr__(self):
        return "Deconvolution %d->%d C%d/%d" % (self.nIn, self.nOut, self.s, self.s)


class NetworkInNetwork(torch.nn.Module):
    """scn.NetworkInNetwork(nIn, nOut, bias): per-site linear map (residual shortcut in scn.UNet)."""

    def __init__(self, nIn, nOut, bias):
        super().__init__()
        self.nIn, self.nOut = nIn, nOut
        self.weight = torch.nn.Parameter(torch.empty(nIn, nOut).normal_(0, (2.0 / nIn) ** 0.5))
        self.bias = torch.nn.Parameter(torch.zeros(nOut)) if bias else None

    def forward(self, x):
        if x.features.shape[1] != self.nIn:
            raise AssertionError("NetworkInNetwork: expected %d input planes, got %d"
                                 % (self.nIn, x.features.shape[1]))
        md = x.metadata
        lv = md.level(x.spatial_size)
        nd = md.n_dev_ptr(lv)
        hint = config.gpl_rows_by_size.get(tuple(x.spatial_size), 0) if md.static else 0
        out = F.TableConv.apply(x.features, self.weight, None, lv.rows, nd, lv.rows, nd, True, None, None, None, None,
                                int(hint))
        if self.bias is not None:
            out = out + self.bias
        return x._like(out)

    def input_spatial_size(self, out_size):
        return out_size

    def __repr__(self):
        return "NetworkInNetwork %d->%d" % (self.nIn, self.nOut)


class BatchNormalization(torch.nn.Module):
    """scn.BatchNormalization(nPlanes, eps=1e-4, momentum=0.9, affine=True, leakiness=1).

    `momentum` is scn's retention factor (0.9 == torch momentum 0.1); leakiness 1 = no
    activation, 0 = ReLU."""

    def __init__(self, nPlanes, eps=1e-4, momentum=0.9, affine=True, leakiness=1):
        super().__init__()
        self.nPlanes, self.eps, self.momentum, self.leakiness = nPlanes, eps, momentum, leakiness
        self.register_buffer("running_mean", torch.zeros(nPlanes))
        self.register_buffer("running_var", torch.ones(nPlanes))
        if affine:
            self.weight = torch.nn.Parameter(torch.ones(nPlanes))
            self.bias = torch.nn.Parameter(torch.zeros(nPlanes))
        else:
            self.weight = self.bias = None
        self._sync = None

    def forward(self, x):
        if x.features.shape[1] != self.nPlanes:
            raise AssertionError("BatchNormalization: expected %d planes, got %d"
                                 % (self.nPlanes, x.features.shape[1]))
        md = x.metadata
        lv = md.level(x.spatial_size)
        sync = self._sync
        if sync is None or sync.device != x.features.device:
            # barrier state of the single-launch kernels: owned by the module (never shared between calls that
            # could overlap), zeroed once, not part of the state dict
            sync = self._sync = torch.zeros(2, dtype=torch.int32, device=x.features.device)
        hint = config.gpl_rows_by_size.get(tuple(x.spatial_size), 0) if md.static else 0
        out = F.BatchNormAct.apply(x.features, self.weight, self.bias, self.running_mean, self.running_var,
                                   float(self.eps), float(self.momentum), float(self.leakiness), self.training,
                                   lv.rows, md.n_dev_ptr(lv), sync, int(hint))
        return x._like(out)

    def input_spatial_size(self, out_size):
        return out_size

    def __repr__(self):
        return "BatchNorm(%d, eps=%g, momentum=%g, leakiness=%g)" % (self.nPlanes, self.eps, self.momentum,
                                                                   self.leakiness)


class BatchNormReLU(BatchNormalization):
    """scn.BatchNormReLU(nPlanes) -- uresnet_sparse.py:23."""

    def __init__(self, nPlanes, eps=1e-4, momentum=0.9):
        super().__init__(nPlanes, eps, momentum, True, 0)


class BatchNormLeakyReLU(BatchNormalization):
    def __init__(self, nPlanes, eps=1e-4, momentum=0.9, leakiness=0.333):
        super().__init__(nPlanes, eps, momentum, True, leakiness)
