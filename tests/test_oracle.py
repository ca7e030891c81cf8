"""CPU tests of the oracle itself: the C++ rulebook restatement against the pure-Python one,
Oracle A (scn-style rule lists) against Oracle B (dense convolutions), golden fixtures."""
import os

import numpy as np
import pytest
import torch

from helpers import event_blob
from oracle import dense_equiv as de
from oracle import rules
from oracle import scn_cpu as scn
from oracle.model import OracleUResNet, segmentation_loss
from conftest import small_flags

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("dim", [2, 3])
def test_rules_cpp_matches_python(dim):
    data, _ = event_blob(2, 24, dim=dim, dup=40)
    c = data[:, : dim + 1].astype(np.int64)
    i2r, act = rules.input_rows(c)
    i2r_py, act_py = rules.input_rows_py(c)
    assert (i2r == i2r_py).all() and (act == act_py).all()
    assert len(act) < len(c)  # duplicates were merged
    for a, b in zip(rules.subm_rules(act), rules.subm_rules_py(act)):
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all()
    for a, b in zip(rules.down_rules(act), rules.down_rules_py(act)):
        assert (a == b).all()


def test_rules_properties():
    data, _ = event_blob(3, 32)
    c = data[:, :4].astype(np.int64)
    i2r, act = rules.input_rows(c)
    # no duplicates in the generator: identity map, rows in input order
    assert (i2r == np.arange(len(c))).all() and (act == c).all()
    r = rules.subm_rules(act)
    K = len(r)
    assert K == 27 and (r[13][0] == r[13][1]).all() and len(r[13][0]) == len(act)
    # symmetry: (k, i, o) present  <=>  (K-1-k, o, i) present
    for k in range(K):
        fwd = set(zip(r[k][0].tolist(), r[k][1].tolist()))
        bwd = set(zip(r[K - 1 - k][1].tolist(), r[K - 1 - k][0].tolist()))
        assert fwd == bwd
    parent, kofs, coarse = rules.down_rules(act)
    assert (coarse[parent][:, :3] == act[:, :3] // 2).all() and (coarse[parent][:, 3] == act[:, 3]).all()
    assert len(np.unique(coarse, axis=0)) == len(coarse)
    off = (act[:, 0] % 2) * 4 + (act[:, 1] % 2) * 2 + act[:, 2] % 2
    assert (kofs == off).all()


@pytest.mark.parametrize("dim", [2, 3])
def test_oracle_a_equals_dense_b(dim):
    torch.manual_seed(1)
    ss, nb = 16, 2
    data, _ = event_blob(nb, ss, dim=dim)
    _, act = rules.input_rows(data[:, : dim + 1].astype(np.int64))
    size = (ss,) * dim
    feats = torch.randn(len(act), 3, dtype=torch.float64)
    x = scn.SparseConvNetTensor(feats, scn.Metadata(dim), size)
    x.metadata.active[size] = act
    coords = torch.from_numpy(act)
    conv = scn.SubmanifoldConvolution(dim, 3, 5, 3, False).double()
    assert torch.allclose(conv(x).features, de.subm_dense(coords, feats, conv.weight, ss, nb), atol=1e-12)
    down = scn.Convolution(dim, 3, 4, 2, 2, False).double()
    z = down(x)
    coarse = torch.from_numpy(x.metadata.active[(ss // 2,) * dim])
    assert torch.allclose(z.features, de.down_dense(coords, feats, down.weight, ss, nb, coarse), atol=1e-12)
    up = scn.Deconvolution(dim, 4, 6, 2, 2, False).double()
    u = up(z)
    assert u.features.shape[0] == len(act)
    assert torch.allclose(u.features, de.up_dense(coarse, z.features, up.weight, ss // 2, nb, coords), atol=1e-12)
    bn = scn.BatchNormReLU(3).double().train()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.uniform_(-0.5, 0.5)
    y = bn(x).features
    yb, rm, rv = de.bnrelu_dense(feats, bn.weight, bn.bias)
    assert torch.allclose(y, yb, atol=1e-12)
    assert torch.allclose(bn.running_mean, rm, atol=1e-12) and torch.allclose(bn.running_var, rv, atol=1e-12)


def test_oracle_model_shape_and_params():
    from uresnet_pytorch_b200.model import Flags
    m = OracleUResNet(Flags(SPATIAL_SIZE=192))
    assert sum(p.numel() for p in m.parameters()) == 2741477
    keys = list(m.state_dict().keys())
    assert keys[0] == "sparseModel.1.weight" and keys[-2:] == ["linear.weight", "linear.bias"]


def test_oracle_permutation_invariance():
    torch.manual_seed(0)
    flags = small_flags()
    m = OracleUResNet(flags).double().train()
    data, _ = event_blob(2, 32, dup=20)
    d = torch.from_numpy(data)
    out = m(d)[0]
    perm = torch.randperm(len(d))
    m2 = OracleUResNet(flags).double().train()
    m2.load_state_dict(m.state_dict())
    out_p = m2(d[perm])[0]
    assert torch.allclose(out[perm], out_p, atol=1e-9)


def test_oracle_matches_golden():
    """The committed fixture was produced by tests/golden/make_golden.py from this oracle."""
    g = np.load(os.path.join(GOLDEN, "sparse_tiny.npz"))
    flags = small_flags(SPATIAL_SIZE=int(g["ss"]), URESNET_FILTERS=int(g["uf"]), URESNET_NUM_STRIDES=int(g["uns"]))
    m = OracleUResNet(flags).train()
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    m.load_state_dict(sd)
    d, l = torch.from_numpy(g["data"]), torch.from_numpy(g["label"])
    out = m(d)[0]
    loss, _ = segmentation_loss(out, d, l)
    loss.backward()
    assert np.allclose(out.detach().numpy(), g["logits"], atol=2e-5, rtol=1e-4)
    assert abs(loss.item() - float(g["loss"])) < 1e-5
    assert np.allclose(m.sparseModel[1].weight.grad.numpy(), g["grad_stem"], atol=1e-5, rtol=1e-3)


def test_metrics_oracle_matches_reference_golden():
    """oracle/metrics.py against outputs of the UNMODIFIED reference function (utils.py:78), generated by
    tests/golden/make_metrics_golden.py."""
    import os
    import numpy as np
    from oracle import metrics as om
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "metrics_small.npz"))
    with np.errstate(all="ignore"):
        res = om.compute_metrics_sparse([g["data"]], [g["label"]], [g["softmax"]])
    for k in ("acc", "correct_softmax", "id", "nonzero_pixels", "class_acc", "class_pixel", "class_mean_softmax",
              "confusion_matrix", "energy_confusion_matrix", "loss_seg"):
        got = np.array([np.asarray(v, dtype=np.float64).reshape(-1) if np.ndim(v) else float(v) for v in res[k]])
        ref = g["ref." + k]
        assert got.shape == ref.shape, k
        assert np.allclose(got, ref, rtol=1e-5, atol=1e-6, equal_nan=True), (k, got, ref)
