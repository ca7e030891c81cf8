"""Inference outputs and per-event metrics on the GPU.

The reference copies logits and softmax of every forward pass to the host
(/root/reference/uresnet/trainval.py:127-130) and, for `inference --full`, recomputes per-event accuracy, mean
softmax of the true class, per-class accuracy / pixel counts and the (energy-weighted) confusion matrices in numpy,
one boolean-mask pass per event and class (/root/reference/uresnet/utils.py:127-200).  Here one kernel
(csrc/headloss.cu: k_hp_main) applies the Linear head, writes softmax / argmax and accumulates all of those sums per
event; the host only divides.  The physics add-ons of compute_metrics_sparse (DBSCAN Michel clustering, distance
histograms, misclassified-pixel dumps) are analysis scripts outside the hot path and are not reproduced.
"""
import numpy as np
import torch

from . import _lib

_CODES = {torch.int32: 0, torch.int64: 1, torch.float32: 2, torch.float64: 3}


def head_predict(feat, weight, bias, ids=None, label=None, energy=None, nb=1, rows=None, n_dev=None, want_logits=False,
                 want_softmax=True, want_pred=True, out=None):
    """feat [rows, m] (OutputLayer rows) -> dict(logits?, softmax?, pred?, sums? [nb, E] float64).

    `out`: optional dict of preallocated output tensors with the same keys (static buffers of a captured graph)."""
    if not feat.is_cuda or feat.dtype != torch.float32:
        raise RuntimeError("head_predict: float32 CUDA features only (there is no CPU path)")
    if feat.stride(1) != 1:
        feat = feat.contiguous()
    rows = feat.shape[0] if rows is None else rows
    W = weight.detach().contiguous()
    nc, m = W.shape
    dev = feat.device
    out = {} if out is None else out
    L = _lib.load()

    def buf(key, want, shape, dtype):
        if not want:
            return None
        t = out.get(key)
        if t is None:
            t = out[key] = torch.empty(shape, dtype=dtype, device=dev)
        return t

    logits = buf("logits", want_logits, (rows, nc), torch.float32)
    softmax = buf("softmax", want_softmax, (rows, nc), torch.float32)
    pred = buf("pred", want_pred, (rows,), torch.int32)
    E = 4 + nc + 2 * nc * nc
    sums = scratch = None
    nbytes = 0
    lab = None
    if label is not None:
        if ids is None:
            raise ValueError("head_predict: metrics need the sample-id column")
        lab = label.reshape(-1)
        if lab.dtype != torch.float32:
            lab = lab.float()
        sums = buf("sums", True, (nb, E), torch.float64)
        nbytes = L.b200scn_head_predict_scratch_bytes(rows, nc, nb)
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    for t in (ids, energy):
        if t is not None and t.dtype not in _CODES:
            raise TypeError("head_predict: ids / energy must be int32, int64, float32 or float64")
    _lib.call("b200scn_head_predict", feat.data_ptr(), feat.stride(0), W.data_ptr(), _lib.ptr(bias),
              _lib.ptr(lab), lab.stride(0) if lab is not None else 0,
              _lib.ptr(ids), _CODES[ids.dtype] if ids is not None else 0, ids.stride(0) if ids is not None else 0,
              _lib.ptr(energy), _CODES[energy.dtype] if energy is not None else 0,
              energy.stride(0) if energy is not None else 0, rows, n_dev, m, nc, nb,
              _lib.ptr(logits), nc, _lib.ptr(softmax), nc, _lib.ptr(pred), _lib.ptr(sums), _lib.ptr(scratch), nbytes,
              torch.cuda.current_stream().cuda_stream)
    return out


def metrics_from_sums(sums, nc):
    """Host side of compute_metrics_sparse's accuracy / confusion block (R/uresnet/utils.py:127-200): `sums` is the
    [nb, E] float64 table of head_predict (numpy or tensor).  Returns the reference's keys, one entry per event
    PRESENT in the batch (events with zero pixels are skipped, like np.unique over the batch-id column)."""
    s = sums.detach().cpu().numpy() if torch.is_tensor(sums) else np.asarray(sums)
    res = {k: [] for k in ("acc", "correct_softmax", "id", "nonzero_pixels", "class_acc", "class_pixel",
                           "class_mean_softmax", "confusion_matrix", "energy_confusion_matrix", "loss_seg")}
    for b in range(s.shape[0]):
        n = s[b, 0]
        if n <= 0:
            continue
        conf = s[b, 4 + nc:4 + nc + nc * nc].reshape(nc, nc)
        econf = s[b, 4 + nc + nc * nc:4 + nc + 2 * nc * nc].reshape(nc, nc)
        pix = conf.sum(1)
        with np.errstate(divide="ignore", invalid="ignore"):   # the reference divides by zero for absent classes too
            class_acc = np.diag(conf) / pix
            class_sm = s[b, 4:4 + nc] / pix
        res["acc"].append(s[b, 1] / n)
        res["loss_seg"].append(s[b, 2] / n)
        res["correct_softmax"].append(s[b, 3] / n)
        res["id"].append(float(b))
        res["nonzero_pixels"].append(int(round(n)))
        res["class_acc"].append(class_acc)
        res["class_pixel"].append(pix.astype(np.int64))
        res["class_mean_softmax"].append(class_sm)
        res["confusion_matrix"].append(conf.astype(np.int32))
        res["energy_confusion_matrix"].append(econf.astype(np.float32))
    return res
