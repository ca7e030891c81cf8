"""Each scn layer of the hot path, CUDA (through the C-ABI) vs the CPU oracle, forward and
backward.  Tolerances: exact-fp32 kernels vs the fp32 oracle differ only by summation order:
max|d| <= 2e-5 * scale for outputs, rel-L2 <= 1e-4 for gradients."""
import numpy as np
import pytest
import torch

from helpers import event_blob, rel_l2
from oracle import scn_cpu as ocn

pytestmark = pytest.mark.gpu

OUT_TOL, GRAD_TOL = 2e-5, 1e-4


def _pair(dim, ss, n_events=2, dup=0, nfeat=1, seed=3):
    """Input layer on both sides -> (oracle tensor, cuda tensor)."""
    from uresnet_pytorch_b200 import sparseconvnet as scn
    data, _ = event_blob(n_events, ss, dim=dim, dup=dup, seed=seed)
    coords = torch.from_numpy(data[:, : dim + 1])
    torch.manual_seed(seed)
    feats = torch.randn(len(data), nfeat)
    f_o = feats.clone().requires_grad_(True)
    f_g = feats.clone().cuda().requires_grad_(True)
    x_o = ocn.InputLayer(dim, ss, mode=3)((coords, f_o))
    x_g = scn.InputLayer(dim, ss, mode=3)((coords.cuda(), f_g))
    return x_o, x_g, f_o, f_g


def _check(layer_o, layer_g, x_o, x_g, f_o, f_g, out_tol=OUT_TOL, grad_tol=GRAD_TOL):
    layer_g.load_state_dict(layer_o.state_dict())
    layer_g.cuda()
    y_o, y_g = layer_o(x_o), layer_g(x_g)
    yo = y_o.features if hasattr(y_o, "features") else y_o
    yg = y_g.features if hasattr(y_g, "features") else y_g
    assert yo.shape == yg.shape
    scale = max(yo.abs().max().item(), 1e-6)
    err = (yg.detach().cpu() - yo.detach()).abs().max().item()
    assert err <= out_tol * scale, "forward: %g vs scale %g" % (err, scale)
    torch.manual_seed(1)
    go = torch.randn_like(yo)
    yo.backward(go)
    yg.backward(go.cuda())
    torch.cuda.synchronize()
    assert rel_l2(f_g.grad, f_o.grad) <= grad_tol, "dgrad %g" % rel_l2(f_g.grad, f_o.grad)
    for (n, p_o), (_, p_g) in zip(layer_o.named_parameters(), layer_g.named_parameters()):
        assert rel_l2(p_g.grad, p_o.grad) <= grad_tol, "wgrad %s %g" % (n, rel_l2(p_g.grad, p_o.grad))
    return y_o, y_g


def _lift(x_o, x_g, f_o, f_g, c_in):
    """Widen the single input feature to c_in channels with a shared NiN so later layers see real widths."""
    from uresnet_pytorch_b200 import sparseconvnet as scn
    if c_in == f_o.shape[1]:
        return x_o, x_g
    torch.manual_seed(5)
    a = ocn.NetworkInNetwork(f_o.shape[1], c_in, False)
    b = scn.NetworkInNetwork(f_o.shape[1], c_in, False)
    b.load_state_dict(a.state_dict())
    b.cuda()
    return a(x_o), b(x_g)


@pytest.mark.parametrize("dim,cin,cout", [(3, 1, 16), (3, 16, 16), (3, 32, 16), (3, 48, 48), (3, 80, 80),
                                          (3, 7, 5), (2, 16, 32), (3, 160, 136)])
def test_submanifold_convolution(dim, cin, cout):
    from uresnet_pytorch_b200 import sparseconvnet as scn
    x_o, x_g, f_o, f_g = _pair(dim, 24, dup=15)
    x_o, x_g = _lift(x_o, x_g, f_o, f_g, cin)
    torch.manual_seed(2)
    _check(ocn.SubmanifoldConvolution(dim, cin, cout, 3, False), scn.SubmanifoldConvolution(dim, cin, cout, 3, False),
           x_o, x_g, f_o, f_g)


@pytest.mark.parametrize("dim,cin,cout", [(3, 16, 32), (3, 64, 80), (2, 16, 32), (3, 5, 3)])
def test_convolution_and_deconvolution(dim, cin, cout):
    from uresnet_pytorch_b200 import sparseconvnet as scn
    x_o, x_g, f_o, f_g = _pair(dim, 24, dup=15)
    x_o, x_g = _lift(x_o, x_g, f_o, f_g, cin)
    torch.manual_seed(2)
    down_o, down_g = ocn.Convolution(dim, cin, cout, 2, 2, False), scn.Convolution(dim, cin, cout, 2, 2, False)
    up_o, up_g = ocn.Deconvolution(dim, cout, cin, 2, 2, False), scn.Deconvolution(dim, cout, cin, 2, 2, False)
    seq_o = ocn.Sequential().add(down_o).add(up_o)
    seq_g = scn.Sequential().add(down_g).add(up_g)
    y_o, y_g = _check(seq_o, seq_g, x_o, x_g, f_o, f_g)
    assert y_g.spatial_size == x_g.spatial_size


def test_strided_conv_rejects_odd_spatial_size():
    from uresnet_pytorch_b200 import sparseconvnet as scn
    x_o, x_g, f_o, f_g = _pair(3, 24)
    x_g.spatial_size = (25, 25, 25)
    x_g.metadata.levels[(25, 25, 25)] = x_g.metadata.levels[(24, 24, 24)]
    with pytest.raises(AssertionError):
        scn.Convolution(3, 1, 4, 2, 2, False).cuda()(x_g)
    with pytest.raises(AssertionError):
        scn.SubmanifoldConvolution(3, 3, 4, 3, False).cuda()(x_g)  # channel mismatch


@pytest.mark.parametrize("c,leak", [(16, 0.0), (48, 0.0), (80, 0.333), (5, 1.0), (320, 0.0)])
def test_batchnorm_relu_train(c, leak):
    from uresnet_pytorch_b200 import sparseconvnet as scn
    x_o, x_g, f_o, f_g = _pair(3, 24, dup=10)
    x_o, x_g = _lift(x_o, x_g, f_o, f_g, c)
    torch.manual_seed(4)
    bn_o = ocn.BatchNormalization(c, leakiness=leak).train()
    with torch.no_grad():
        bn_o.weight.uniform_(0.5, 1.5)
        bn_o.bias.uniform_(-0.5, 0.5)
    bn_g = scn.BatchNormalization(c, leakiness=leak).train()
    _check(bn_o, bn_g, x_o, x_g, f_o, f_g, grad_tol=3e-4)
    assert (bn_g.running_mean.cpu() - bn_o.running_mean).abs().max().item() <= 1e-6
    assert rel_l2(bn_g.running_var, bn_o.running_var) <= 1e-5


def test_batchnorm_eval_uses_running_stats():
    from uresnet_pytorch_b200 import sparseconvnet as scn
    x_o, x_g, f_o, f_g = _pair(3, 24)
    x_o, x_g = _lift(x_o, x_g, f_o, f_g, 16)
    bn_o = ocn.BatchNormReLU(16)
    with torch.no_grad():
        bn_o.running_mean.uniform_(-0.3, 0.3)
        bn_o.running_var.uniform_(0.5, 2.0)
        bn_o.weight.uniform_(0.5, 1.5)
    bn_o.eval()
    bn_g = scn.BatchNormReLU(16).eval()
    _check(bn_o, bn_g, x_o, x_g, f_o, f_g)
    assert torch.equal(bn_g.running_mean.cpu(), bn_o.running_mean)


def test_batchnorm_large_mean_is_stable():
    """Shifted / merged variance must survive mean >> std (catastrophic cancellation check)."""
    from uresnet_pytorch_b200 import sparseconvnet as scn
    torch.manual_seed(0)
    x = (torch.randn(5000, 16) * 0.01 + 100.0)
    md_holder = _pair(3, 24)[1]
    lv = md_holder.metadata.level(md_holder.spatial_size)
    n = lv.n
    xt = scn.SparseConvNetTensor(x[:n].cuda(), md_holder.metadata, md_holder.spatial_size)
    bn = scn.BatchNormalization(16).cuda().train()
    y = bn(xt).features
    ref = (x[:n].double() - x[:n].double().mean(0)) / (x[:n].double().var(0, unbiased=False) + 1e-4).sqrt()
    assert (y.cpu().double() - ref).abs().max().item() < 5e-3


@pytest.mark.parametrize("n,c,live", [(1, 16, None), (37, 1, None), (1362, 128, None), (23000, 32, None),
                                      (32768, 16, 9405), (32768, 64, 3987), (40000, 48, None), (6000, 320, 1017),
                                      (300000, 32, None)])
def test_batchnorm_single_launch_matches_two_kernel(n, c, live):
    """The single-launch kernels (grid barrier, csrc/bnrelu.cu) against the two-kernel path and an fp64 torch
    restatement; the same barrier words are reused by every call (count must return to zero each time)."""
    from uresnet_pytorch_b200.sparseconvnet import functional as F
    torch.manual_seed(n + c)
    x = (torch.randn(n, c, device="cuda") * 1.5 + 0.7)
    dy = torch.randn(n, c, device="cuda")
    w = torch.rand(c, device="cuda") + 0.5
    b = torch.rand(c, device="cuda") - 0.5
    n_dev = torch.tensor([live], dtype=torch.int32, device="cuda") if live is not None else None
    m = live if live is not None else n
    sync = torch.zeros(2, dtype=torch.int32, device="cuda")
    res = []
    for sy in (None, sync, sync, sync):
        rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
        xx = x.clone().requires_grad_(True)
        ww, bb = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
        y = F.BatchNormAct.apply(xx, ww, bb, rm, rv, 1e-4, 0.9, 0.0, True, n,
                                 n_dev.data_ptr() if n_dev is not None else None, sy, m if live is not None else 0)
        y[:m].backward(dy[:m])
        torch.cuda.synchronize()
        res.append((y.detach()[:m], xx.grad[:m], ww.grad, bb.grad, rm, rv))
    assert int(sync[0].item()) == 0  # the wrapping arrival counter is back at zero after every call
    xd = x[:m].double()
    mu, var = xd.mean(0), xd.var(0, unbiased=False)
    ref = torch.relu((xd - mu) / (var + 1e-4).sqrt() * w.double() + b.double())
    for r in res:
        assert (r[0].double() - ref).abs().max().item() <= 2e-5 * max(ref.abs().max().item(), 1.0)
    for r in res[1:]:
        for a, o in zip(r, res[0]):
            assert rel_l2(a, o) <= 2e-5
    # run-to-run determinism of the single-launch path
    for a, o in zip(res[1], res[2]):
        assert torch.equal(a, o)


def test_input_output_layers_with_duplicates():
    from uresnet_pytorch_b200 import sparseconvnet as scn
    x_o, x_g, f_o, f_g = _pair(3, 16, dup=120, nfeat=3)
    assert torch.equal(x_g.features.detach().cpu(), x_o.features.detach())  # in-order sums: bit-exact
    y_o = ocn.OutputLayer(3)(x_o)
    y_g = scn.OutputLayer(3)(x_g)
    assert torch.equal(y_g.detach().cpu(), y_o.detach())
    g = torch.randn_like(y_o)
    y_o.backward(g)
    y_g.backward(g.cuda())
    assert rel_l2(f_g.grad, f_o.grad) <= 1e-6


def test_unet_block_tables():
    """ConcatTable / AddTable / JoinTable / NetworkInNetwork plumbing inside a 2-level UNet."""
    from uresnet_pytorch_b200 import sparseconvnet as scn
    x_o, x_g, f_o, f_g = _pair(3, 24, dup=10)
    torch.manual_seed(7)
    net_o = ocn.Sequential().add(ocn.SubmanifoldConvolution(3, 1, 8, 3, False)).add(
        ocn.UNet(3, 2, [8, 16], residual_blocks=True, downsample=[2, 2]))
    net_g = scn.Sequential().add(scn.SubmanifoldConvolution(3, 1, 8, 3, False)).add(
        scn.UNet(3, 2, [8, 16], residual_blocks=True, downsample=[2, 2]))
    net_o.train()
    net_g.train()
    _check(net_o, net_g, x_o, x_g, f_o, f_g, out_tol=1e-4, grad_tol=1e-3)
