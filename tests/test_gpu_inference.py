"""Inference outputs and per-event metrics: csrc/headloss.cu k_hp_main through
inference.head_predict / Trainer.predict_step, against (a) outputs of the UNMODIFIED reference function
/root/reference/uresnet/utils.py:78 committed as tests/golden/metrics_small.npz, (b) the numpy oracle
(oracle/metrics.py), (c) the CPU oracle model in eval mode (trainval.py:92-130 with TRAIN=False)."""
import os

import numpy as np
import pytest
import torch

from conftest import small_flags
from helpers import event_blob

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
KEYS = ("acc", "correct_softmax", "id", "nonzero_pixels", "class_acc", "class_pixel", "class_mean_softmax",
        "confusion_matrix", "energy_confusion_matrix", "loss_seg")


def _flat(res, k):
    return np.array([np.asarray(v, dtype=np.float64).reshape(-1) if np.ndim(v) else float(v) for v in res[k]])


def test_head_predict_matches_reference_golden():
    from uresnet_pytorch_b200 import inference
    g = np.load(os.path.join(GOLDEN, "metrics_small.npz"))
    data, label, softmax = g["data"], g["label"], g["softmax"]
    nc = softmax.shape[1]
    # identity head on log-probabilities reproduces the golden softmax (up to fp32 round-off)
    feat = torch.from_numpy(np.log(softmax)).cuda()
    d = torch.from_numpy(data).cuda()
    out = inference.head_predict(feat, torch.eye(nc, device="cuda"), torch.zeros(nc, device="cuda"), ids=d[:, 3],
                                 label=torch.from_numpy(label).cuda(), energy=d[:, 4], nb=3, want_logits=True)
    torch.cuda.synchronize()
    assert np.abs(out["softmax"].cpu().numpy() - softmax).max() <= 1e-6
    assert np.array_equal(out["pred"].cpu().numpy(), softmax.argmax(1))
    res = inference.metrics_from_sums(out["sums"], nc)
    for k in KEYS:
        got, ref = _flat(res, k), g["ref." + k]
        assert got.shape == ref.shape, k
        assert np.allclose(got, ref, rtol=2e-5, atol=1e-6, equal_nan=True), (k, got, ref)


@pytest.mark.parametrize("graph", [False, True])
def test_predict_step_matches_oracle_model(graph):
    """Trainer.predict_step (eager and captured graph, eval-mode BatchNorm, tensor-core convs) == the CPU oracle
    model's softmax / argmax, and its metrics == oracle/metrics.py on the oracle's softmax."""
    from oracle import metrics as om
    from oracle.model import OracleUResNet
    from uresnet_pytorch_b200 import sparseconvnet as scn
    from uresnet_pytorch_b200.model import Flags
    from uresnet_pytorch_b200.trainer import Trainer
    flags = Flags(DATA_DIM=3, SPATIAL_SIZE=64, URESNET_FILTERS=16, URESNET_NUM_STRIDES=3, NUM_CLASS=5)
    blobs = [event_blob(2, 64, seed=s) for s in (5, 6)]
    cap = max(len(b[0]) for b in blobs) + 100
    t = Trainer(flags, graph=graph, capacity=cap if graph else None, events_per_step=2, train=False,
                math_mode=scn.MATH_TF32X3, seed=3)
    with torch.no_grad():   # non-trivial running statistics
        for m in t.net.modules():
            if hasattr(m, "running_mean"):
                m.running_mean.uniform_(-0.1, 0.1)
                m.running_var.uniform_(0.5, 1.5)
    ref = OracleUResNet(flags)
    ref.load_state_dict({k: v.detach().cpu().clone() for k, v in t.net.state_dict().items()})
    ref.eval()
    for data, label in blobs:
        with torch.no_grad():
            sm_ref = torch.softmax(ref(torch.from_numpy(data))[0], 1).numpy()
        res = t.predict_step(data, label)
        assert res["softmax"].shape == sm_ref.shape
        assert np.abs(res["softmax"] - sm_ref).max() <= 1e-4
        assert (res["pred"] == sm_ref.argmax(1)).mean() >= 0.999
        with np.errstate(all="ignore"):
            want = om.compute_metrics_sparse([data], [label], [sm_ref])
        for k in KEYS:
            got, exp = _flat(res["metrics"], k), _flat(want, k)
            assert got.shape == exp.shape, k
            tol = 3e-3 if k in ("acc", "class_acc", "confusion_matrix", "energy_confusion_matrix") else 2e-4
            assert np.allclose(got, exp, rtol=tol, atol=3 if k.endswith("confusion_matrix") else tol,
                               equal_nan=True), (k, got, exp)


def test_collate_kernel():
    """b200scn_collate: events back to back + offsets -> [N, D+1] with the sample-id column (replaces the host
    padding of iotools_sparse.py:131-141)."""
    from uresnet_pytorch_b200 import _lib
    rng = np.random.default_rng(0)
    counts = [100, 0, 257, 31]
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    coords = rng.integers(0, 512, size=(off[-1], 3)).astype(np.int32)
    c = torch.from_numpy(coords).cuda()
    o = torch.from_numpy(off).cuda()
    out = torch.full((off[-1] + 5, 4), -1, dtype=torch.int32, device="cuda")
    n = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.call("b200scn_collate", c.data_ptr(), 3, o.data_ptr(), len(counts), 3, 7, out.shape[0], out.data_ptr(), 4,
              n.data_ptr(), torch.cuda.current_stream().cuda_stream)
    got = out.cpu().numpy()
    assert int(n.item()) == off[-1]
    want_id = np.repeat(np.arange(len(counts)), counts) + 7
    assert np.array_equal(got[:off[-1], :3], coords) and np.array_equal(got[:off[-1], 3], want_id)
    assert (got[off[-1]:] == -1).all()
