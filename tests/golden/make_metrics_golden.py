"""Generates tests/golden/metrics_small.npz by running the UNMODIFIED reference function
/root/reference/uresnet/utils.py:78 (`compute_metrics_sparse`, imported by file path; matplotlib, which that file
imports but never uses on this path, is stubbed) on a seeded random batch.  Run in the build container only:

    python tests/golden/make_metrics_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference_utils():
    sys.modules.setdefault("matplotlib", types.ModuleType("matplotlib"))
    spec = importlib.util.spec_from_file_location("ref_utils", "/root/reference/uresnet/utils.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def make_inputs(seed=11, n=3000, nb=3, nc=5, ss=192):
    rng = np.random.default_rng(seed)
    data = np.zeros((n, 5), np.float64)
    data[:, :3] = rng.integers(0, ss, size=(n, 3))
    data[:, 3] = rng.integers(0, nb, size=n)
    data[:, 4] = rng.lognormal(-1.0, 0.8, size=n)
    label = rng.integers(0, nc - 1, size=(n, 1)).astype(np.float32)   # class nc-1 absent on purpose
    logits = rng.normal(size=(n, nc)).astype(np.float32) * 2.0
    logits[np.arange(n), label[:, 0].astype(int)] += 1.5              # a better-than-random classifier
    e = np.exp(logits - logits.max(1, keepdims=True))
    softmax = (e / e.sum(1, keepdims=True)).astype(np.float32)
    return data, label, softmax


if __name__ == "__main__":
    ref = load_reference_utils()
    data, label, softmax = make_inputs()
    with np.errstate(all="ignore"):
        out = ref.compute_metrics_sparse([data], [label], [softmax], [0], N=192, particles=None)
    res = out[0] if isinstance(out, tuple) else out
    keep = ("acc", "correct_softmax", "id", "nonzero_pixels", "class_acc", "class_pixel", "class_mean_softmax",
            "confusion_matrix", "energy_confusion_matrix", "loss_seg")
    arrays = {"data": data, "label": label, "softmax": softmax}
    for k in keep:
        arrays["ref." + k] = np.array([np.asarray(v, dtype=np.float64).reshape(-1) if np.ndim(v) else float(v)
                                       for v in res[k]], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "metrics_small.npz"), **arrays)
    print({k: arrays["ref." + k].shape for k in keep})
