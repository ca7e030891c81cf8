"""Whole sparse U-ResNet: CUDA vs oracle (logits, loss, every gradient), golden fixture,
static/graph regime vs eager, permutation invariance, trainer."""
import os

import numpy as np
import pytest
import torch

from conftest import small_flags
from helpers import check_grads, event_blob, rel_l2
from oracle.model import OracleUResNet, segmentation_loss

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

LOGIT_TOL = 1e-4   # max|d| <= LOGIT_TOL * max|logit|   (fp32 kernels vs fp32 oracle)
GRAD_TOL = 1e-3    # rel-L2 of the whole flattened gradient
TENSOR_TOL = 2e-2  # rel-L2 per parameter tensor (guard; see helpers.check_grads)


def _models(flags, seed=0):
    from uresnet_pytorch_b200.model import UResNet
    torch.manual_seed(seed)
    ref = OracleUResNet(flags).train()
    with torch.no_grad():
        for mod in ref.modules():
            if hasattr(mod, "running_mean"):
                mod.weight.uniform_(0.7, 1.3)
                mod.bias.uniform_(-0.2, 0.2)
    net = UResNet(flags)
    net.load_state_dict(ref.state_dict())
    return ref, net.cuda().train()


def _loss_gpu(net, d, l, n_events):
    from uresnet_pytorch_b200.model import SegmentationLoss
    crit = SegmentationLoss()
    crit.num_events = n_events
    out = net(d)[0]
    loss, acc = crit([out], [d], [l], None)
    return out, loss, acc


@pytest.mark.parametrize("dim,ss,uf,uns", [(3, 32, 8, 3), (3, 64, 16, 4), (2, 64, 16, 4)])
def test_model_forward_backward_matches_oracle(dim, ss, uf, uns):
    flags = small_flags(DATA_DIM=dim, SPATIAL_SIZE=ss, URESNET_FILTERS=uf, URESNET_NUM_STRIDES=uns)
    ref, net = _models(flags)
    data, label = event_blob(2, ss, dim=dim, dup=25)
    d, l = torch.from_numpy(data), torch.from_numpy(label)
    out_ref = ref(d)[0]
    loss_ref, acc_ref = segmentation_loss(out_ref, d, l)
    loss_ref.backward()
    out, loss, acc = _loss_gpu(net, d.cuda(), l.cuda(), 2)
    loss.backward()
    torch.cuda.synchronize()
    scale = out_ref.abs().max().item()
    assert (out.detach().cpu() - out_ref.detach()).abs().max().item() <= LOGIT_TOL * scale
    assert abs(loss.item() - loss_ref.item()) <= 1e-4 * abs(loss_ref.item())
    assert abs(float(acc) - acc_ref) < 1e-3
    check_grads(net, ref, GRAD_TOL, TENSOR_TOL)
    b_ref = dict(ref.named_buffers())
    for name, b in net.named_buffers():
        assert rel_l2(b, b_ref[name]) <= 1e-4, name


def test_model_matches_golden_fixture():
    from uresnet_pytorch_b200.model import UResNet
    g = np.load(os.path.join(GOLDEN, "sparse_tiny.npz"))
    flags = small_flags(SPATIAL_SIZE=int(g["ss"]), URESNET_FILTERS=int(g["uf"]), URESNET_NUM_STRIDES=int(g["uns"]))
    net = UResNet(flags)
    net.load_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")})
    net.cuda().train()
    d, l = torch.from_numpy(g["data"]).cuda(), torch.from_numpy(g["label"]).cuda()
    out, loss, _ = _loss_gpu(net, d, l, 2)
    loss.backward()
    assert np.abs(out.detach().cpu().numpy() - g["logits"]).max() <= LOGIT_TOL * np.abs(g["logits"]).max()
    assert abs(loss.item() - float(g["loss"])) <= 1e-4 * float(g["loss"])
    assert rel_l2(net.sparseModel[1].weight.grad, torch.from_numpy(g["grad_stem"])) <= GRAD_TOL
    assert rel_l2(net.linear.weight.grad, torch.from_numpy(g["grad_linear"])) <= GRAD_TOL
    for k in g.files:
        if k.startswith("after."):
            assert rel_l2(net.state_dict()[k[6:]], torch.from_numpy(g[k])) <= 1e-4, k


def test_permutation_invariance_on_gpu():
    flags = small_flags()
    _, net = _models(flags)
    data, _ = event_blob(2, 32, dup=20)
    d = torch.from_numpy(data).cuda()
    with torch.no_grad():
        out = net(d)[0]
        perm = torch.randperm(len(d), device="cuda")
        out_p = net(d[perm])[0]
    assert (out[perm] - out_p).abs().max().item() <= 1e-4 * out.abs().max().item()


def test_eval_mode_and_inference():
    flags = small_flags()
    ref, net = _models(flags)
    data, label = event_blob(1, 32)
    d = torch.from_numpy(data)
    ref.eval()
    net.eval()
    with torch.no_grad():
        a = ref(d)[0]
        b = net(d.cuda())[0]
    assert (b.cpu() - a).abs().max().item() <= LOGIT_TOL * a.abs().max().item()


def test_static_regime_matches_eager_forward_backward():
    from uresnet_pytorch_b200 import sparseconvnet as scn
    flags = small_flags()
    _, net = _models(flags)
    data, label = event_blob(2, 32, dup=10)
    d, l = torch.from_numpy(data).cuda(), torch.from_numpy(label).cuda()
    out_e, loss_e, _ = _loss_gpu(net, d, l, 2)
    loss_e.backward()
    g_e = {n: p.grad.clone() for n, p in net.named_parameters()}
    net.zero_grad(set_to_none=True)
    with scn.static_mode(True):
        out_s, loss_s, _ = _loss_gpu(net, d, l, 2)
        loss_s.backward()
    # same kernels, but capacity-sized grids change the (fixed) reduction split: round-off only
    assert (out_s - out_e).abs().max().item() <= 1e-5 * out_e.abs().max().item()
    for n, p in net.named_parameters():
        assert rel_l2(p.grad, g_e[n]) <= 1e-4, n


@pytest.mark.parametrize("mode,uf", [(0, 8), (1, 16)])
def test_graph_trainer_matches_eager_trainer(mode, uf):
    """Whole-step CUDA graph (static regime, padded capacity) == eager step, over several
    steps with different event sizes; FP32 FFMA kernels and the tensor-core pair-list path."""
    from uresnet_pytorch_b200.trainer import Trainer
    flags = small_flags(URESNET_FILTERS=uf)
    blobs = [event_blob(2, 32, seed=s) for s in (1, 2, 3, 4)]
    cap = max(len(b[0]) for b in blobs) + 37
    te = Trainer(flags, lr=1e-3, graph=False, events_per_step=2, seed=5, math_mode=mode)
    tg = Trainer(flags, lr=1e-3, graph=True, capacity=cap, events_per_step=2, seed=5, math_mode=mode)
    for data, label in blobs:
        re_ = te.train_step(data, label)
        rg = tg.train_step(data, label)
        assert abs(re_["loss"] - rg["loss"]) <= 2e-4 * abs(re_["loss"]), (re_, rg)
        assert abs(re_["acc"] - rg["acc"]) <= 1e-3
    for (n, a), (_, b) in zip(te.net.named_parameters(), tg.net.named_parameters()):
        assert rel_l2(b, a) <= 1e-3, n


def test_training_reduces_loss():
    from uresnet_pytorch_b200.trainer import Trainer
    flags = small_flags()
    data, label = event_blob(2, 32, seed=8)
    t = Trainer(flags, lr=3e-3, graph=False, events_per_step=2, seed=1)
    first = t.train_step(data, label)["loss"]
    for _ in range(15):
        last = t.train_step(data, label)["loss"]
    assert last < 0.7 * first, (first, last)


@pytest.mark.parametrize("with_weight", [False, True])
def test_fused_head_loss_matches_torch(with_weight):
    """csrc/headloss.cu against torch.nn.Linear + the reference-semantics SegmentationLoss
    (uresnet_sparse.py:25,36,44-82): loss, accuracy and every gradient."""
    from uresnet_pytorch_b200.model import SegmentationLoss
    from uresnet_pytorch_b200.sparseconvnet import functional as F
    torch.manual_seed(5)
    n, m, nc, nb = 3000, 16, 5, 4
    dev = torch.device("cuda")
    feat = torch.randn(n, m, device=dev, requires_grad=True)
    lin = torch.nn.Linear(m, nc).to(dev)
    label = torch.randint(0, nc, (n, 1), device=dev).float()
    ids = torch.randint(0, nb - 1, (n,), device=dev)          # event nb-1 is absent on purpose
    data = torch.zeros(n, 5, dtype=torch.float64, device=dev)
    data[:, 3] = ids.double()
    pw = (torch.rand(n, 1, device=dev) + 0.5) if with_weight else None

    crit = SegmentationLoss()
    crit.num_events = nb
    loss_ref, acc_ref = crit([lin(feat)], [data], [label], [pw] if with_weight else None)
    loss_ref.backward()
    g_ref = (feat.grad.clone(), lin.weight.grad.clone(), lin.bias.grad.clone())
    feat.grad = None
    lin.zero_grad()

    loss, acc = F.head_loss(feat, lin, label, data[:, 3], nb, pix_w=pw)
    (2.0 * loss).backward()   # exercises the upstream-gradient scaling
    assert abs(loss.item() - loss_ref.item()) <= 1e-5 * abs(loss_ref.item())
    assert abs(acc.item() - float(acc_ref)) <= 1e-5
    for a, b in zip((feat.grad, lin.weight.grad, lin.bias.grad), g_ref):
        assert (a - 2.0 * b).abs().max().item() <= 2e-5 * (2.0 * b).abs().max().item() + 1e-9


def test_fused_and_unfused_trainers_agree():
    """Trainer(fused_loss=True) and the torch loss path give the same loss trajectory."""
    from uresnet_pytorch_b200.model import Flags
    from uresnet_pytorch_b200.trainer import Trainer
    flags = Flags(DATA_DIM=3, SPATIAL_SIZE=64, URESNET_FILTERS=16, URESNET_NUM_STRIDES=3, NUM_CLASS=5)
    data, label = event_blob(2, 64, seed=3)
    out = []
    for fused in (False, True):
        tr = Trainer(flags, graph=False, events_per_step=2, seed=1, fused_loss=fused)
        out.append([tr.train_step(data, label) for _ in range(3)])
    for a, b in zip(*out):
        assert abs(a["loss"] - b["loss"]) <= 2e-4 * abs(a["loss"])
        assert abs(a["acc"] - b["acc"]) <= 1e-3


def test_fused_adam_matches_torch_adam():
    """csrc/optim.cu (one launch for all tensors, device-side step counter) against torch.optim.Adam."""
    from uresnet_pytorch_b200.optim import FusedAdam
    torch.manual_seed(11)
    shapes = [(27, 16, 16), (16,), (8, 32, 48), (5, 16), (1,)]
    pa = [torch.randn(s, device="cuda") for s in shapes]
    pb = [p.clone() for p in pa]
    ref = torch.optim.Adam(pa, lr=3e-3)
    mine = FusedAdam(pb, lr=3e-3)
    for step in range(6):
        gs = [torch.randn_like(p) * (0.1 + step) for p in pa]
        for p, q, g in zip(pa, pb, gs):
            p.grad = g.clone()
            q.grad = g.clone()
        ref.step()
        mine.step()
    for p, q in zip(pa, pb):
        assert (p - q).abs().max().item() <= 2e-6 * max(1.0, p.abs().max().item())
    assert int(mine.step_dev.item()) == 6
    # a parameter without a gradient is left untouched (the step counter is global, unlike torch's per-tensor one)
    before = pb[1].clone()
    for q in pb:
        q.grad = torch.ones_like(q)
    pb[1].grad = None
    mine.step()
    assert torch.equal(pb[1], before) and int(mine.step_dev.item()) == 7


def test_checkpoint_round_trip_in_reference_format(tmp_path):
    """save_state / load_state use the reference's checkpoint layout (R/uresnet/trainval.py:32-40, 171-196):
    'module.'-prefixed state dict + torch.optim.Adam-format optimizer state; a restored trainer (eager or
    captured graph) continues bit-identically, and torch.optim.Adam accepts the optimizer state."""
    from uresnet_pytorch_b200.trainer import Trainer
    flags = small_flags()
    blobs = [event_blob(2, 32, seed=s) for s in (1, 2, 3, 4)]
    cap = max(len(b[0]) for b in blobs) + 5
    a = Trainer(flags, lr=1e-3, graph=False, events_per_step=2, seed=5)
    for data, label in blobs[:2]:
        a.train_step(data, label)
    path = str(tmp_path / "weights-1.ckpt")
    a.save_state(path, 1)
    ck = torch.load(path, map_location="cpu", weights_only=False)
    assert set(ck) == {"global_step", "state_dict", "optimizer"} and ck["global_step"] == 1
    assert all(k.startswith("module.") for k in ck["state_dict"])
    assert "module.sparseModel.1.weight" in ck["state_dict"] and "module.linear.bias" in ck["state_dict"]
    ref_opt = torch.optim.Adam([torch.nn.Parameter(p.detach().clone()) for p in a.net.parameters()], lr=1e-3)
    ref_opt.load_state_dict(ck["optimizer"])  # torch's own loader validates the layout
    assert int(ref_opt.state_dict()["state"][0]["step"]) == 2
    nxt = [a.train_step(d, l)["loss"] for d, l in blobs[2:]]
    for graph in (False, True):
        b = Trainer(flags, lr=1e-3, graph=graph, capacity=cap if graph else None, events_per_step=2, seed=99)
        if graph:
            b.train_step(*blobs[0])  # capture first: restoring must work on a live graph
        assert b.load_state(path) == 2
        got = [b.train_step(d, l)["loss"] for d, l in blobs[2:]]
        for x, y in zip(got, nxt):
            assert abs(x - y) <= (2e-4 if graph else 1e-6) * abs(y), (graph, got, nxt)
