"""Dense U-ResNet path on hand-written CUDA (north-star job 4).

Behavioural mirror of /root/reference/uresnet/models/uresnet_dense.py:113-226: the same layer
graph, the same parameter names and shapes (so a reference state dict loads unchanged), the
same "replicate-pad then zero-padding-free convolution" arithmetic (:19-26, :75-80), and
BatchNorm that always uses batch statistics (track_running_stats=False, :45).  The layers run
on channels-last [B, D, H, W, C] fp32 tensors through libb200scn (csrc/dense.cu: implicit-index
gather-GEMM convolutions, replicate pad, bias column sums; csrc/bnrelu.cu for BatchNorm).
`forward` accepts and returns the reference's NC(D)HW layout.
"""
import ctypes

import torch

from . import _lib
from .sparseconvnet import functional as F_sparse


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _i32(vals):
    return (ctypes.c_int32 * len(vals))(*[int(v) for v in vals])


def _same_pad(kernel, stride, size):
    """Total padding that makes ceil(size / stride) outputs, split low/high (the rule the
    reference applies before every convolution, uresnet_dense.py:19-26)."""
    rem = size % stride
    total = max(kernel - (stride if rem == 0 else rem), 0)
    lo = total // 2
    return lo, total - lo


def _geom(mode, dims, in_sz, out_sz, k, stride, pad):
    return _i32([mode, dims] + list(in_sz) + list(out_sz) + [k, stride, pad])


def _sz3(x):
    """(z, y, x) spatial sizes of a channels-last [B, D, H, W, C] tensor."""
    return [x.shape[1], x.shape[2], x.shape[3]]


class _PadReplicate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, lo, hi, dims):
        x = x.contiguous()
        B, C = x.shape[0], x.shape[4]
        in_sz = _sz3(x)
        pl = [0 if (dims == 2 and d == 0) else lo for d in range(3)]
        pr = [0 if (dims == 2 and d == 0) else hi for d in range(3)]
        out = torch.empty([B] + [s + a + b for s, a, b in zip(in_sz, pl, pr)] + [C], dtype=torch.float32,
                          device=x.device)
        _lib.call("b200scn_pad_replicate", x.data_ptr(), out.data_ptr(), B, C, _i32(in_sz), _i32(pl), _i32(pr), 0,
                  _stream())
        ctx.geom = (B, C, in_sz, pl, pr)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, C, in_sz, pl, pr = ctx.geom
        dout = dout.contiguous()
        din = torch.empty([B] + in_sz + [C], dtype=torch.float32, device=dout.device)
        _lib.call("b200scn_pad_replicate", dout.data_ptr(), din.data_ptr(), B, C, _i32(in_sz), _i32(pl), _i32(pr), 1,
                  _stream())
        return din, None, None, None


def _conv_call(x_ptr, Cin, w_ptr, w_si, w_so, bias, g, B, Cout, out, T, device):
    """One dense convolution launch: the tcgen05 implicit-GEMM kernel (csrc/dense_tc.cu) in the tensor-core
    math modes (tf32x3 = fp32-equivalent, tf32 = single pass) when the channel counts allow, else the
    exact-fp32 FFMA kernel."""
    from .sparseconvnet import config
    if config.math_mode != config.MATH_FP32 and (Cin % 32 == 0 or Cin == 16) and Cout % 16 == 0:
        L = _lib.load()
        nbytes = L.b200scn_dense_conv_tc_scratch_bytes(T, Cin, Cout)
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _lib.call("b200scn_dense_conv_tc", x_ptr, Cin, w_ptr, 1, w_si, w_so, _lib.ptr(bias), g, B, Cin, Cout,
                  out.data_ptr(), Cout, config.math_mode, scratch.data_ptr(), nbytes, _stream())
        return
    _lib.call("b200scn_dense_conv", x_ptr, Cin, w_ptr, 1, w_si, w_so, _lib.ptr(bias), g, B, Cin, Cout, out.data_ptr(),
              Cout, _stream())


class _DenseConv(torch.autograd.Function):
    """nn.Conv{2,3}d(padding=0) / nn.ConvTranspose{2,3}d(padding, output_padding) on channels-last data."""

    @staticmethod
    def forward(ctx, x, weight, bias, k, stride, transposed, pad, out_pad, dims):
        if not x.is_cuda:
            raise RuntimeError("b200 dense path: tensors must be on a CUDA device (there is no CPU path)")
        x = x.contiguous()
        w = weight.contiguous()
        B, Cin = x.shape[0], x.shape[4]
        T = k ** dims
        in_sz = _sz3(x)
        if transposed:
            Cout = w.shape[1]
            out_sz = [1 if (dims == 2 and d == 0) else (s - 1) * stride - 2 * pad + k + out_pad
                      for d, s in enumerate(in_sz)]
            mode, w_si, w_so = 2, Cout * T, T
        else:
            Cout = w.shape[0]
            out_sz = [1 if (dims == 2 and d == 0) else (s - k) // stride + 1 for d, s in enumerate(in_sz)]
            mode, w_si, w_so = 1, T, Cin * T
        out = torch.empty([B] + out_sz + [Cout], dtype=torch.float32, device=x.device)
        g = _geom(mode, dims, in_sz, out_sz, k, stride, pad)
        _conv_call(x.data_ptr(), Cin, w.data_ptr(), w_si, w_so, bias, g, B, Cout, out, T, x.device)
        ctx.save_for_backward(x, w)
        ctx.cfg = (k, stride, transposed, pad, dims, in_sz, out_sz, Cin, Cout, bias is not None, tuple(weight.shape))
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w = ctx.saved_tensors
        k, stride, transposed, pad, dims, in_sz, out_sz, Cin, Cout, has_bias, w_shape = ctx.cfg
        dout = dout.contiguous()
        B, T = x.shape[0], k ** dims
        L = _lib.load()
        dx = dw = db = None
        if transposed:
            w_si, w_so = Cout * T, T
        else:
            w_si, w_so = T, Cin * T
        if ctx.needs_input_grad[0]:
            # the transpose of the forward index map, over the *input* grid, with W^T
            dx = torch.empty([B] + in_sz + [Cin], dtype=torch.float32, device=x.device)
            g = _geom(1 if transposed else 2, dims, out_sz, in_sz, k, stride, pad)
            _conv_call(dout.data_ptr(), Cout, w.data_ptr(), w_so, w_si, None, g, B, Cin, dx, T, x.device)
        if ctx.needs_input_grad[1]:
            g = _geom(2 if transposed else 1, dims, in_sz, out_sz, k, stride, pad)
            nbytes = L.b200scn_dense_wgrad_scratch_bytes(g, B, Cin, Cout)
            scratch = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
            dwk = torch.empty((T, Cin, Cout), dtype=torch.float32, device=x.device)
            from .sparseconvnet import config
            _lib.call("b200scn_dense_wgrad", x.data_ptr(), Cin, dout.data_ptr(), Cout, g, B, 0, Cin, Cout,
                      dwk.data_ptr(), config.math_mode, scratch.data_ptr(), nbytes, _stream())
            dw = (dwk.permute(1, 2, 0) if transposed else dwk.permute(2, 1, 0)).reshape(w_shape).contiguous()
        if has_bias and ctx.needs_input_grad[2]:
            nb = L.b200scn_colsum_scratch_bytes(Cout)
            scratch = torch.empty(nb, dtype=torch.uint8, device=x.device)
            db = torch.empty(Cout, dtype=torch.float32, device=x.device)
            rows = dout.numel() // Cout
            _lib.call("b200scn_colsum", dout.data_ptr(), Cout, rows, Cout, db.data_ptr(), scratch.data_ptr(), nb,
                      _stream())
        return dx, dw, db, None, None, None, None, None, None


class Conv(torch.nn.Module):
    """nn.Conv{2,3}d(in, out, kernel_size, stride, padding=0) with the reference's replicate
    pre-padding folded into the module (uresnet_dense.py:75-80)."""

    def __init__(self, dims, cin, cout, kernel, stride):
        super().__init__()
        self.dims, self.kernel, self.stride = dims, kernel, stride
        ref = (torch.nn.Conv3d if dims == 3 else torch.nn.Conv2d)(cin, cout, kernel, stride)  # same init as torch
        self.weight = torch.nn.Parameter(ref.weight.detach().clone())
        self.bias = torch.nn.Parameter(ref.bias.detach().clone())

    def forward(self, x):
        lo, hi = _same_pad(self.kernel, self.stride, x.shape[3])
        if lo or hi:
            x = _PadReplicate.apply(x, lo, hi, self.dims)
        return _DenseConv.apply(x, self.weight, self.bias, self.kernel, self.stride, False, 0, 0, self.dims)


class ConvTranspose(torch.nn.Module):
    """nn.ConvTranspose{2,3}d(in, out, 3, stride=2, padding=1, output_padding=1) (uresnet_dense.py:165-172)."""

    def __init__(self, dims, cin, cout, kernel=3, stride=2, padding=1, output_padding=1):
        super().__init__()
        self.dims, self.kernel, self.stride, self.padding, self.output_padding = dims, kernel, stride, padding, \
            output_padding
        ref = (torch.nn.ConvTranspose3d if dims == 3 else torch.nn.ConvTranspose2d)(
            cin, cout, kernel, stride, padding, output_padding)
        self.weight = torch.nn.Parameter(ref.weight.detach().clone())
        self.bias = torch.nn.Parameter(ref.bias.detach().clone())

    def forward(self, x):
        return _DenseConv.apply(x, self.weight, self.bias, self.kernel, self.stride, True, self.padding,
                                self.output_padding, self.dims)


class BatchNorm(torch.nn.Module):
    """nn.BatchNorm{2,3}d(C, track_running_stats=False): always batch statistics, eps 1e-5;
    optionally with the following ReLU fused in."""

    def __init__(self, channels, relu=False):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.ones(channels))
        self.bias = torch.nn.Parameter(torch.zeros(channels))
        self.relu = relu

    def forward(self, x):
        C = x.shape[-1]
        flat = x.reshape(-1, C)
        y = F_sparse.BatchNormAct.apply(flat, self.weight, self.bias, None, None, 1e-5, 0.0, 0.0 if self.relu else 1.0,
                                        True, flat.shape[0], None)
        return y.reshape(x.shape)


class _Id(torch.nn.Module):
    """placeholder keeping nn.Sequential indices aligned with the reference (ReLU is fused into BatchNorm)"""

    def forward(self, x):
        return x


def _conv_bn(dims, cin, cout, kernel, stride, relu=False, tail=False):
    layers = [Conv(dims, cin, cout, kernel, stride), BatchNorm(cout, relu=relu)]
    if tail:
        layers.append(_Id())
    return torch.nn.Sequential(*layers)


class ResBlock(torch.nn.Module):
    """uresnet_dense.py:29-83: relu(shortcut(x) + bn(conv(bn(conv(x)))))"""

    def __init__(self, dims, cin, cout, kernel=3, stride=1):
        super().__init__()
        self.use_shortcut = (cin != cout) or stride != 1
        self.shortcut = _conv_bn(dims, cin, cout, 1, stride)
        self.residual1 = _conv_bn(dims, cin, cout, kernel, stride)
        self.residual2 = _conv_bn(dims, cout, cout, kernel, 1)

    def forward(self, x):
        sc = self.shortcut(x) if self.use_shortcut else x
        return torch.relu(sc + self.residual2(self.residual1(x)))


class DoubleBlock(torch.nn.Module):
    def __init__(self, dims, cin, cout, kernel=3, stride=1):
        super().__init__()
        self.resnet1 = ResBlock(dims, cin, cout, kernel, stride)
        self.resnet2 = ResBlock(dims, cout, cout, kernel, 1)

    def forward(self, x):
        return self.resnet2(self.resnet1(x))


class DenseUResNet(torch.nn.Module):
    def __init__(self, flags):
        super().__init__()
        self._flags = flags
        dims = flags.DATA_DIM
        self.dims = dims
        base, depth = flags.URESNET_FILTERS, flags.URESNET_NUM_STRIDES
        self.conv1 = _conv_bn(dims, 1, base, 3, 1, relu=True, tail=True)
        self.double_resnet = torch.nn.ModuleList()
        width = base
        for _ in range(depth):
            self.double_resnet.append(DoubleBlock(dims, width, 2 * width, 3, 2))
            width *= 2
        self.decode_conv = torch.nn.ModuleList()
        self.decode_double_resnet = torch.nn.ModuleList()
        for _ in range(depth):
            self.decode_double_resnet.append(DoubleBlock(dims, width, width // 2, 3, 1))
            self.decode_conv.append(torch.nn.Sequential(ConvTranspose(dims, width, width // 2),
                                                        BatchNorm(width // 2, relu=True), _Id()))
            width //= 2
        self.conv2 = _conv_bn(dims, width, base, 3, 1, relu=True, tail=True)
        self.conv3 = _conv_bn(dims, base, flags.NUM_CLASS, 3, 1)

    def forward(self, image):
        """image [B, 1, (D,) H, W] -> scores [B, num_class, (D,) H, W]"""
        if self.dims == 3:
            x = image.permute(0, 2, 3, 4, 1)
        else:
            x = image.permute(0, 2, 3, 1).unsqueeze(1)
        x = self.conv1(x.float().contiguous())
        skips = {x.shape[-1]: x}
        for blk in self.double_resnet:
            x = blk(x)
            skips[x.shape[-1]] = x
        for up, blk in zip(self.decode_conv, self.decode_double_resnet):
            x = up(x)
            x = blk(torch.cat((x, skips[x.shape[-1]]), dim=-1))
        x = self.conv3(self.conv2(x))
        if self.dims == 3:
            return x.permute(0, 4, 1, 2, 3).contiguous()
        return x.squeeze(1).permute(0, 3, 1, 2).contiguous()
